"""VideoSysEngine -- process-per-GPU driver with the reference's surface (videosys/core/engine/engine.py:13-128):
``VideoSysEngine(config).generate(...)``, ``.save_video``, ``.shutdown``.

One process per GPU (torch.multiprocessing 'spawn'); rank 0 is the caller's process; every rank builds the pipeline
named by ``config.pipeline_cls`` after ``initialize()`` (NCCL).  The workers run ``generate`` on request over a pipe
and rank 0's result is returned (the reference returns the driver's output too, :74-95).
"""
import os
import socket
import traceback

import torch
import torch.multiprocessing as mp

from ..distributed.parallel_mgr import initialize


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_main(rank, world, init_method, config, conn):
    try:
        initialize(rank=rank, world_size=world, init_method=init_method)
        pipe = config.pipeline_cls(config)
        conn.send(("ready", None))
        while True:
            msg = conn.recv()
            if msg is None:
                break
            method, args, kwargs = msg
            try:
                getattr(pipe, method)(*args, **kwargs)
                conn.send(("ok", None))
            except Exception:  # pragma: no cover
                conn.send(("error", traceback.format_exc()))
    except Exception:  # pragma: no cover
        conn.send(("error", traceback.format_exc()))


class VideoSysEngine:
    def __init__(self, config):
        self.config = config
        self.workers = []
        self._init_worker(config.pipeline_cls)

    def _init_worker(self, pipeline_cls):
        world = self.config.num_gpus
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        assert world <= torch.cuda.device_count(), "num_gpus exceeds visible devices"
        init_method = f"tcp://127.0.0.1:{_free_port()}"
        ctx = mp.get_context("spawn")
        for rank in range(1, world):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker_main, args=(rank, world, init_method, self.config, child), daemon=True)
            p.start()
            self.workers.append((p, parent))
        if world > 1:
            initialize(rank=0, world_size=world, init_method=init_method)
        self.driver_worker = pipeline_cls(self.config)
        for p, conn in self.workers:
            kind, payload = conn.recv()
            if kind != "ready":
                raise ChildProcessError(payload)

    def _run_workers(self, method, *args, **kwargs):
        for _, conn in self.workers:
            conn.send((method, args, kwargs))
        out = getattr(self.driver_worker, method)(*args, **kwargs)
        for p, conn in self.workers:
            if not p.is_alive():
                raise ChildProcessError("worker died")
            kind, payload = conn.recv()
            if kind == "error":
                raise ChildProcessError(payload)
        return out

    def generate(self, *args, **kwargs):
        return self._run_workers("generate", *args, **kwargs)

    def save_video(self, video, output_path):
        os.makedirs(os.path.dirname(os.path.abspath(output_path)), exist_ok=True)
        torch.save(video, output_path)  # encoders (imageio/mp4) are outside the hot-path scope
        return output_path

    def shutdown(self):
        for p, conn in self.workers:
            try:
                conn.send(None)
            except Exception:
                pass
        for p, _ in self.workers:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
        self.workers = []

    def __del__(self):
        try:
            self.shutdown()
        except Exception:
            pass
