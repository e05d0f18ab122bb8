"""FlowMatchEulerDiscreteScheduler as the Vchitect pipeline uses it (reference pipelines/vchitect/pipeline_vchitect.py:
``retrieve_timesteps`` :1035-1057, ``scheduler.step`` :954).  The class itself is diffusers' (==0.30.0, not installed
here): restated from its published semantics, parity unpinned -- including its double shift (``sigma_max`` / ``sigma_min``
are taken from the already shifted training table, and ``set_timesteps`` shifts the interpolated values again).
"""
import numpy as np
import torch


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, **unused):
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        ts = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        sigmas = torch.from_numpy(ts) / num_train_timesteps
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = sigmas * num_train_timesteps
        self.sigmas = sigmas
        self.sigma_min, self.sigma_max = sigmas[-1].item(), sigmas[0].item()
        self._step_index = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        n = self.num_train_timesteps
        ts = np.linspace(self.sigma_max * n, self.sigma_min * n, num_inference_steps)
        sigmas = ts / n
        sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        self.timesteps = sigmas * n
        self.sigmas = torch.cat([sigmas, torch.zeros(1, device=sigmas.device)])
        self._step_index = None

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = False):
        """x_{t-1} = x_t + (sigma_next - sigma) * v, in fp32, cast to the model output's dtype."""
        if self._step_index is None:
            t = float(timestep)
            self._step_index = int((self.timesteps == t).nonzero()[0].item()) if (self.timesteps == t).any() else 0
        sigma, sigma_next = self.sigmas[self._step_index], self.sigmas[self._step_index + 1]
        prev = (sample.to(torch.float32) + (sigma_next - sigma) * model_output).to(model_output.dtype)
        self._step_index += 1
        return (prev,)
