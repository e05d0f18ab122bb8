"""DDIM sampler as the Latte pipeline configures it -- restatement of ``diffusers.DDIMScheduler`` (diffusers==0.30.0,
third party: not in the reference tree) for the arguments pipeline_latte.py:229-237 passes: linear betas
(1e-4 .. 0.02), epsilon prediction, clip_sample=False, set_alpha_to_one, "leading" timestep spacing, eta = 0
(``variance_type="learned_range"`` only matters for eta > 0).  Host arithmetic + a few elementwise device ops."""
from typing import Optional

import numpy as np
import torch


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=False,
                 set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon", timestep_spacing="leading",
                 variance_type="learned_range"):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        if clip_sample or prediction_type != "epsilon":
            raise NotImplementedError("the Latte configuration is epsilon prediction without sample clipping")
        self.num_train_timesteps, self.steps_offset, self.timestep_spacing = num_train_timesteps, steps_offset, timestep_spacing
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        N = self.num_train_timesteps
        if num_inference_steps > N:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        if self.timestep_spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (N // num_inference_steps)).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.timestep_spacing == "trailing":
            ts = np.round(np.arange(N, 0, -N / num_inference_steps)).astype(np.int64) - 1
        elif self.timestep_spacing == "linspace":
            ts = np.linspace(0, N - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = torch.from_numpy(ts).to(device)

    def step(self, model_output, timestep, sample, eta: float = 0.0, return_dict: bool = False, **kwargs):
        if eta != 0.0:
            raise NotImplementedError("eta > 0 (stochastic DDIM) is not built")
        t = int(timestep)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.final_alpha_cumprod)
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t**0.5
        out = a_p**0.5 * x0 + (1 - a_p) ** 0.5 * model_output
        return (out,) if not return_dict else type("Out", (), {"prev_sample": out, "pred_original_sample": x0})()
