"""What the CogVideoX and Latte pipelines share with the reference's: the process-mesh set-up (reference
pipeline_cogvideox.py:195-209 = pipeline_latte.py:261-275) and the seeding rule (utils/utils.py:19-34)."""
import random
from typing import Optional

import torch
import torch.distributed as dist


class ParallelPipelineMixin:
    """Needs ``self.transformer`` (with ``enable_parallel`` / ``parallel_manager``) and ``self._device``."""

    def _set_parallel(self, dp_size: Optional[int] = None, sp_size: Optional[int] = None, enable_cp: Optional[bool] = False):
        """Every rank joins one sequence-parallel group unless told otherwise; ``enable_cp`` turns a factor 2 of it into
        CFG parallelism (the transformer's ``enable_parallel``)."""
        world = dist.get_world_size() if dist.is_initialized() else 1
        if sp_size is None:
            sp_size, dp_size = world, 1
        else:
            assert world % sp_size == 0, f"world_size {world} must be divisible by sp_size"
            dp_size = world // sp_size
        self.transformer.enable_parallel(dp_size, sp_size, enable_cp)

    def _set_seed(self, seed):
        """One seed per dp replica.  seed < 0 draws one; rank 0's draw is broadcast so that the ranks of a sequence- /
        CFG-parallel group denoise the same latent (every rank draws the latent itself)."""
        if seed is None or seed < 0:
            seed = random.randint(0, 1000000)
        if dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.tensor([seed], dtype=torch.int64, device=self._device)
            dist.broadcast(t, src=0)
            seed = int(t.item())
        pm = self.transformer.parallel_manager
        seed = int(seed) + (pm.dp_rank if pm is not None else 0)
        random.seed(seed)
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        return seed

    def _maybe_seed(self, seed):
        """Seed when asked to, and always under multi-GPU (unseeded ranks would draw different latents)."""
        if (seed is not None and seed >= 0) or (dist.is_initialized() and dist.get_world_size() > 1):
            self._set_seed(seed)
