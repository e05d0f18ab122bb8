"""EulerAncestralDiscreteScheduler as the Open-Sora-Plan v1.2.0 pipeline constructs it (``EulerAncestralDiscreteScheduler()``:
reference pipelines/open_sora_plan/pipeline_open_sora_plan.py:306; linear betas 1e-4 .. 0.02, epsilon prediction, linspace
timesteps).  The class is diffusers' (==0.30.0, not installed here): restated from its published algorithm (ancestral Euler
sampling of Karras et al. / k-diffusion), parity unpinned.
"""
import numpy as np
import torch


class EulerAncestralDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02, **unused):
        self.n_train = num_train_timesteps
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        self._train_sigmas = sig
        self.sigmas = torch.from_numpy(np.concatenate([sig[::-1], [0.0]]).astype(np.float32))
        self.init_noise_sigma = float(self.sigmas.max())  # linspace spacing
        self.timesteps = None
        self._step_index = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        ts = np.linspace(0, self.n_train - 1, num_inference_steps, dtype=np.float32)[::-1].copy()
        sig = np.interp(ts, np.arange(0, len(self._train_sigmas)), self._train_sigmas)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32)).to(device)
        self.timesteps = torch.from_numpy(ts).to(device)
        self._step_index = None

    def _index(self, timestep):
        if self._step_index is None:
            self._step_index = int((self.timesteps == float(timestep)).nonzero()[0].item())
        return self._step_index

    def scale_model_input(self, sample: torch.Tensor, timestep) -> torch.Tensor:
        sigma = self.sigmas[self._index(timestep)]
        return sample / ((sigma**2 + 1) ** 0.5)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None, return_dict: bool = False, **unused):
        i = self._index(timestep)
        sigma, sigma_to = self.sigmas[i], self.sigmas[i + 1]
        sample = sample.to(torch.float32)
        pred_original = sample - sigma * model_output  # epsilon prediction
        sigma_up = (sigma_to**2 * (sigma**2 - sigma_to**2) / sigma**2) ** 0.5
        sigma_down = (sigma_to**2 - sigma_up**2) ** 0.5
        derivative = (sample - pred_original) / sigma
        prev = sample + derivative * (sigma_down - sigma)
        noise = torch.randn(model_output.shape, dtype=model_output.dtype, device=model_output.device, generator=generator)
        prev = (prev + noise * sigma_up).to(model_output.dtype)
        self._step_index += 1
        return (prev,)
