"""DDIM sampler of CogVideoX -- mirror of videosys/schedulers/scheduling_ddim_cogvideox.py (CogVideoXDDIMScheduler
:172-394): scaled-linear betas, SD3-style SNR shift (:208), zero-terminal-SNR rescale (:87-113), "trailing" timestep
spacing (:286-291), v-prediction step (:355-389).  Host arithmetic + a handful of elementwise device ops; the defaults
are the THUDM/CogVideoX-2b scheduler config (HF file, not in the reference tree: SURVEY.md 8d)."""
from typing import Optional

import numpy as np
import torch


def rescale_zero_terminal_snr(alphas_cumprod: torch.Tensor) -> torch.Tensor:
    """Reference :87-113 (arXiv 2305.08891 algorithm 1 on alphas_cumprod)."""
    s = alphas_cumprod.sqrt()
    s0, sT = s[0].clone(), s[-1].clone()
    s = (s - sT) * (s0 / (s0 - sT))
    return s**2


class CogVideoXDDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.0120, beta_schedule="scaled_linear",
                 clip_sample=False, set_alpha_to_one=True, steps_offset=0, prediction_type="v_prediction",
                 timestep_spacing="trailing", rescale_betas_zero_snr=True, snr_shift_scale=3.0):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float64) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type, self.timestep_spacing, self.steps_offset = prediction_type, timestep_spacing, steps_offset
        ac = torch.cumprod(1.0 - betas, dim=0)
        ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
        if rescale_betas_zero_snr:
            ac = rescale_zero_terminal_snr(ac)
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else ac[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        N = self.num_train_timesteps
        if self.timestep_spacing == "linspace":
            ts = np.linspace(0, N - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (N // num_inference_steps)).round()[::-1].copy().astype(np.int64)
            ts += self.steps_offset
        elif self.timestep_spacing == "trailing":
            ts = np.round(np.arange(N, 0, -N / num_inference_steps)).astype(np.int64) - 1
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = torch.from_numpy(ts).to(device)

    def coefficients(self, timestep: int):
        """(sqrt(alpha_t), sqrt(1 - alpha_t), a_t, b_t) of the step at `timestep` as python floats (reference :359-386)."""
        prev = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t_ = self.alphas_cumprod[timestep]
        a_p_ = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        a = ((1 - a_p_) / (1 - a_t_)) ** 0.5
        b = a_p_**0.5 - a_t_**0.5 * a
        return float(a_t_**0.5), float((1 - a_t_) ** 0.5), float(a), float(b)

    def step(self, model_output, timestep, sample, eta: float = 0.0, return_dict: bool = False, **kwargs):
        if self.num_inference_steps is None:
            raise ValueError("run set_timesteps first")
        sa, sb, a, b = self.coefficients(int(timestep))
        if self.prediction_type == "v_prediction":
            x0 = sa * sample - sb * model_output
        elif self.prediction_type == "epsilon":
            x0 = (sample - sb * model_output) / sa
        elif self.prediction_type == "sample":
            x0 = model_output
        else:
            raise ValueError(self.prediction_type)
        prev = a * sample + b * x0
        return (prev,) if not return_dict else type("Out", (), {"prev_sample": prev, "pred_original_sample": x0})()
