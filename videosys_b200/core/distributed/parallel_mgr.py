"""Process mesh for inference -- mirror of videosys/core/distributed/parallel_mgr.py without colossalai.

``ParallelManager(dp, cp, sp)`` lays ranks out as a dp x cp x sp grid with sp innermost (reference :14-39, which
builds the same groups through colossalai's ProcessGroupMesh) and exposes {dp,cp,sp}_{size,rank,group}.
``initialize()`` is the reference's ``videosys.initialize`` (:103-117): one process per GPU, NCCL on CUDA
(gloo when no GPU is visible, for the host-logic tests), ``cuda.set_device(rank % device_count)``.
"""
import os
from typing import Optional

import torch
import torch.distributed as dist


class ParallelManager:
    def __init__(self, dp_size: int, cp_size: int, sp_size: int):
        world = dist.get_world_size() if dist.is_initialized() else 1
        assert dp_size * cp_size * sp_size == world, f"dp*cp*sp = {dp_size * cp_size * sp_size} != world {world}"
        self.dp_size, self.cp_size, self.sp_size = dp_size, cp_size, sp_size
        rank = dist.get_rank() if dist.is_initialized() else 0
        self.dp_rank = rank // (cp_size * sp_size)
        self.cp_rank = (rank // sp_size) % cp_size
        self.sp_rank = rank % sp_size
        self.dp_group = self._axis_group(lambda d, c, s: (c, s), dp_size, cp_size, sp_size)
        self.cp_group = self._axis_group(lambda d, c, s: (d, s), dp_size, cp_size, sp_size)
        self.sp_group = self._axis_group(lambda d, c, s: (d, c), dp_size, cp_size, sp_size)
        self.enable_sp = sp_size > 1

    @staticmethod
    def _axis_group(key, dp, cp, sp):
        """Groups of ranks that share ``key`` (i.e. differ only along one mesh axis); returns mine."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return None
        buckets = {}
        for d in range(dp):
            for c in range(cp):
                for s in range(sp):
                    buckets.setdefault(key(d, c, s), []).append((d * cp + c) * sp + s)
        mine = None
        me = dist.get_rank()
        for ranks in buckets.values():  # every rank must create every group, in the same order
            g = dist.new_group(ranks)
            if me in ranks:
                mine = g
        return mine


def set_distributed_state(distributed_profile=None):
    """Training-time DCP helper in the reference (:120-148); inference never calls it."""
    raise NotImplementedError("DCP profiling is training-only and out of scope (SURVEY.md section 2.1 row 10)")


def initialize(
    rank: Optional[int] = 0,
    world_size: Optional[int] = 1,
    init_method: Optional[str] = None,
):
    if dist.is_initialized():
        return
    has_gpu = torch.cuda.is_available()
    if init_method is None and "MASTER_ADDR" not in os.environ:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
    if rank is None:
        rank = int(os.environ.get("RANK", 0))
    if world_size is None:
        world_size = int(os.environ.get("WORLD_SIZE", 1))
    if has_gpu:
        torch.cuda.set_device(rank % torch.cuda.device_count())
    kw = dict(backend="nccl" if has_gpu else "gloo", rank=rank, world_size=world_size)
    if init_method is not None:
        kw["init_method"] = init_method
    if has_gpu:
        kw["device_id"] = torch.device("cuda", rank % torch.cuda.device_count())
    dist.init_process_group(**kw)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
