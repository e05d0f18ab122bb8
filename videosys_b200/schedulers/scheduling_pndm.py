"""PNDMScheduler as the Open-Sora-Plan v1.1.0 pipeline constructs it (``PNDMScheduler()``: reference
pipelines/open_sora_plan/pipeline_open_sora_plan.py:304; linear betas 1e-4 .. 0.02, epsilon prediction, leading spacing,
Runge-Kutta warm-up NOT skipped).  The class is diffusers' (==0.30.0, not installed here): restated from its published
algorithm (pseudo numerical methods for diffusion models, arXiv:2202.09778: F-PNDM), parity unpinned.
"""
import numpy as np
import torch


class PNDMScheduler:
    order = 1
    init_noise_sigma = 1.0
    pndm_order = 4

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 skip_prk_steps: bool = False, set_alpha_to_one: bool = False, steps_offset: int = 0, **unused):
        self.n_train, self.skip_prk, self.steps_offset = num_train_timesteps, skip_prk_steps, steps_offset
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = None
        self._reset()

    def _reset(self):
        self.cur_model_output, self.counter, self.cur_sample, self.ets = 0, 0, None, []

    def scale_model_input(self, sample, *a, **k):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.n_train // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round() + self.steps_offset
        if self.skip_prk:
            self.prk_timesteps = np.array([])
            self.plms_timesteps = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()
        else:
            prk = np.array(ts[-self.pndm_order:]).repeat(2) + np.tile(np.array([0, ratio // 2]), self.pndm_order)
            self.prk_timesteps = (prk[:-1].repeat(2)[1:-1])[::-1].copy()
            self.plms_timesteps = ts[:-3][::-1].copy()
        self.timesteps = torch.from_numpy(np.concatenate([self.prk_timesteps, self.plms_timesteps]).astype(np.int64)).to(device)
        self._reset()

    def _prev_sample(self, sample, t, t_prev, eps):
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t_prev] if t_prev >= 0 else self.final_alpha_cumprod
        b_t, b_prev = 1 - a_t, 1 - a_prev
        sample_coeff = (a_prev / a_t) ** 0.5
        denom = a_t * b_prev**0.5 + (a_t * b_t * a_prev) ** 0.5
        return sample_coeff * sample - (a_prev - a_t) * eps / denom

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = False, **unused):
        t = int(timestep)
        ratio = self.n_train // self.num_inference_steps
        if self.counter < len(self.prk_timesteps) and not self.skip_prk:  # Runge-Kutta warm-up (4 evaluations per step)
            t_prev = t - (0 if self.counter % 2 else ratio // 2)
            t = int(self.prk_timesteps[self.counter // 4 * 4])
            k = self.counter % 4
            if k == 0:
                self.cur_model_output = self.cur_model_output + 1 / 6 * model_output
                self.ets.append(model_output)
                self.cur_sample = sample
            elif k in (1, 2):
                self.cur_model_output = self.cur_model_output + 1 / 3 * model_output
            else:
                model_output = self.cur_model_output + 1 / 6 * model_output
                self.cur_model_output = 0
            cur = self.cur_sample if self.cur_sample is not None else sample
            prev = self._prev_sample(cur, t, t_prev, model_output)
        else:  # linear multi-step
            t_prev = t - ratio
            if self.counter != 1:
                self.ets = self.ets[-3:]
                self.ets.append(model_output)
            else:
                t_prev, t = t, t + ratio
            e = self.ets
            if len(e) == 1 and self.counter == 0:
                self.cur_sample = sample
            elif len(e) == 1 and self.counter == 1:
                model_output = (model_output + e[-1]) / 2
                sample, self.cur_sample = self.cur_sample, None
            elif len(e) == 2:
                model_output = (3 * e[-1] - e[-2]) / 2
            elif len(e) == 3:
                model_output = (23 * e[-1] - 16 * e[-2] + 5 * e[-3]) / 12
            else:
                model_output = (1 / 24) * (55 * e[-1] - 59 * e[-2] + 37 * e[-3] - 9 * e[-4])
            prev = self._prev_sample(sample, t, t_prev, model_output)
        self.counter += 1
        return (prev,)
