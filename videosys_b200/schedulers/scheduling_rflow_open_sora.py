"""Rectified-flow sampler for OpenSora -- the caller of the hot path (mirror of
videosys/schedulers/scheduling_rflow_open_sora.py: timestep_transform :47-70, RFLOW.sample :188-257).

Inference only (no training losses).  The loop body is what bench.py times as one "step": CFG batch of 2 through
STDiT3.forward, guidance combine, Euler update.  With ``use_cuda_graph`` (default on CUDA) the step runs as a replayed
CUDA graph keyed by the PAB skip pattern (core/graph_step.py): the per-step bookkeeping that stays on the host is the
integer timestep list, the mask update and the graph lookup.
"""
from typing import Callable, List, Optional

import torch


def timestep_transform(t, model_kwargs, base_resolution=512 * 512, base_num_frames=1, scale=1.0, num_timesteps=1):
    t = t / num_timesteps
    resolution = model_kwargs["height"] * model_kwargs["width"]
    ratio_space = (resolution / base_resolution).sqrt()
    if model_kwargs["num_frames"][0] == 1:
        num_frames = torch.ones_like(model_kwargs["num_frames"])
    else:
        num_frames = model_kwargs["num_frames"] // 17 * 5  # temporal reduction of the OpenSora VAE
    ratio = ratio_space * (num_frames / base_num_frames).sqrt() * scale
    return ratio * t / (1 + (ratio - 1) * t) * num_timesteps


class RFLOW:
    def __init__(self, num_sampling_steps=10, num_timesteps=1000, cfg_scale=4.0, use_discrete_timesteps=False,
                 use_timestep_transform=False, **kwargs):
        self.num_sampling_steps = num_sampling_steps
        self.num_timesteps = num_timesteps
        self.cfg_scale = cfg_scale
        self.use_discrete_timesteps = use_discrete_timesteps
        self.use_timestep_transform = use_timestep_transform

    def prepare_timesteps(self, batch: int, device, model_args) -> List[torch.Tensor]:
        ts = [(1.0 - i / self.num_sampling_steps) * self.num_timesteps for i in range(self.num_sampling_steps)]
        if self.use_discrete_timesteps:
            ts = [int(round(t)) for t in ts]
        ts = [torch.tensor([t] * batch, device=device) for t in ts]
        if self.use_timestep_transform:
            ts = [timestep_transform(t, model_args, num_timesteps=self.num_timesteps) for t in ts]
        return ts

    @staticmethod
    def add_noise(x0, noise, t, num_timesteps=1000):
        tp = (1 - t.float() / num_timesteps)[:, None, None, None, None]
        return tp * x0 + (1 - tp) * noise

    def step(self, model: Callable, z, t, dt, model_args, guidance_scale):
        """One sampling step: CFG pair through the denoiser, guidance, Euler update (reference :238-250)."""
        z_in = torch.cat([z, z], 0)
        tt = torch.cat([t, t], 0)
        out = model(z_in, tt, **model_args)
        pred = out.chunk(2, dim=1)[0]
        cond, uncond = pred.chunk(2, dim=0)
        v = uncond + guidance_scale * (cond - uncond)
        return z + v * dt[:, None, None, None, None]

    def sample(self, model, z, model_args, y_null, device, mask=None, guidance_scale=None, progress=True, verbose=False,
               use_cuda_graph=None):
        guidance_scale = self.cfg_scale if guidance_scale is None else guidance_scale
        if use_cuda_graph is None:
            use_cuda_graph = z.is_cuda
        stepper = None
        if use_cuda_graph:
            from ..core.graph_step import StepGraph

            stepper = StepGraph(model, guidance_scale)
        model_args = dict(model_args)
        model_args["y"] = torch.cat([model_args["y"], y_null], 0)
        timesteps = self.prepare_timesteps(z.shape[0], device, model_args)
        noise_added = None
        if mask is not None:
            noise_added = torch.zeros_like(mask, dtype=torch.bool) | (mask == 1)
        dtype = model.x_embedder.proj.weight.dtype
        model_args["all_timesteps"] = [int(t.to(dtype).item()) for t in timesteps]
        fwd_args = {k: v for k, v in model_args.items() if k not in ("num_frames", "ar")}
        for i, t in enumerate(timesteps):
            x0 = mask_t_upper = None
            if mask is not None:
                mask_t = mask * self.num_timesteps
                x0 = z.clone()
                x_noise = self.add_noise(x0, torch.randn_like(x0), t, self.num_timesteps)
                mask_t_upper = mask_t >= t.unsqueeze(1)
                fwd_args["x_mask"] = mask_t_upper.repeat(2, 1)
                z = torch.where((mask_t_upper & ~noise_added)[:, None, :, None, None], x_noise, x0)
                noise_added = mask_t_upper
            dt = (timesteps[i] - timesteps[i + 1] if i < len(timesteps) - 1 else timesteps[i]) / self.num_timesteps
            if stepper is not None:
                z = stepper.step(z, t, dt, fwd_args, ts_int=model_args["all_timesteps"][i])
            else:
                z = self.step(model, z, t, dt, fwd_args, guidance_scale)
            if mask is not None:
                z = torch.where(mask_t_upper[:, None, :, None, None], z, x0)
        return z
