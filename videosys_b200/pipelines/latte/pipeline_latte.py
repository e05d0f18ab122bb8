"""Latte pipeline surface (mirror of videosys/pipelines/latte/pipeline_latte.py: LattePABConfig :35-76, LatteConfig
:127-160, LattePipeline.generate :675-905) around the B200 LatteT2V.

In scope: the config classes, ``generate()``'s signature, the CFG + DDIM loop (:815-873, learned-sigma split :864-868)
and the denoiser.  Out of scope as for the other pipelines: T5 encoder, caption cleaning, VAE decode --
pass ``prompt_embeds`` / ``negative_prompt_embeds`` or a ``text_encoder_fn``; without a ``vae_decode_fn`` the LATENTS are
returned.  dtype fp16 as the reference (:201).  Latte generates 16 frames at 512 x 512 (:764-766).
"""
import zlib
from typing import Callable, Optional

import torch

from ...core.pab.pab_mgr import PABConfig, enable_pab, set_pab_manager, update_steps
from ...models.transformers.latte_transformer_3d import LatteT2V
from ...schedulers.scheduling_ddim import DDIMScheduler
from .._common import ParallelPipelineMixin
from ..open_sora.pipeline_open_sora import VideoSysPipelineOutput

_MLP_CFG = {k: {"block": [0, 1, 2, 3, 4], "skip_count": 2} for k in (720, 640, 560, 480, 400)}


class LattePABConfig(PABConfig):
    def __init__(self, spatial_broadcast=True, spatial_threshold=(100, 800), spatial_range=2, temporal_broadcast=True,
                 temporal_threshold=(100, 800), temporal_range=3, cross_broadcast=True, cross_threshold=(100, 800),
                 cross_range=6, mlp_broadcast=True, mlp_spatial_broadcast_config=None, mlp_temporal_broadcast_config=None):
        super().__init__(
            spatial_broadcast=spatial_broadcast, spatial_threshold=list(spatial_threshold), spatial_range=spatial_range,
            temporal_broadcast=temporal_broadcast, temporal_threshold=list(temporal_threshold), temporal_range=temporal_range,
            cross_broadcast=cross_broadcast, cross_threshold=list(cross_threshold), cross_range=cross_range,
            mlp_broadcast=mlp_broadcast,
            mlp_spatial_broadcast_config=dict(_MLP_CFG) if mlp_spatial_broadcast_config is None else mlp_spatial_broadcast_config,
            mlp_temporal_broadcast_config=dict(_MLP_CFG) if mlp_temporal_broadcast_config is None else mlp_temporal_broadcast_config,
        )


class LatteConfig:
    def __init__(self, model_path: str = "maxin-cn/Latte-1", num_gpus: int = 1, enable_vae_temporal_decoder: bool = True,
                 beta_start: float = 0.0001, beta_end: float = 0.02, beta_schedule: str = "linear",
                 variance_type: str = "learned_range", enable_pab: bool = False, pab_config=None,
                 transformer_config: Optional[dict] = None, state_dict=None, text_encoder_fn: Optional[Callable] = None,
                 vae_decode_fn: Optional[Callable] = None):
        self.model_path = model_path
        self.pipeline_cls = LattePipeline
        self.num_gpus = num_gpus
        self.enable_vae_temporal_decoder = enable_vae_temporal_decoder
        self.beta_start, self.beta_end, self.beta_schedule, self.variance_type = beta_start, beta_end, beta_schedule, variance_type
        self.enable_pab = enable_pab
        self.pab_config = pab_config if pab_config is not None else LattePABConfig()
        # B200 build extras: architecture / weights / out-of-scope stages supplied by the caller
        self.transformer_config = transformer_config
        self.state_dict = state_dict
        self.text_encoder_fn = text_encoder_fn
        self.vae_decode_fn = vae_decode_fn


class LattePipeline(ParallelPipelineMixin):
    vae_scale_factor = 8

    def __init__(self, config: LatteConfig, device=None, dtype: torch.dtype = torch.float16):
        if not torch.cuda.is_available():
            raise RuntimeError("videosys_b200 pipelines need an sm_100a GPU (no CPU path)")
        self._config = config
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._dtype = dtype
        import os

        if config.transformer_config is None and config.state_dict is None and os.path.isdir(str(config.model_path)):
            self.transformer = LatteT2V.from_pretrained(config.model_path, subfolder="transformer", video_length=16).to(dtype)
        else:
            self.transformer = LatteT2V(**(config.transformer_config or {})).to(dtype)
        if config.state_dict is not None:
            self.transformer.load_state_dict(config.state_dict)
        self.transformer = self.transformer.to(self._device).eval()
        self.scheduler = DDIMScheduler(beta_start=config.beta_start, beta_end=config.beta_end, beta_schedule=config.beta_schedule,
                                       variance_type=config.variance_type, clip_sample=False)
        if config.enable_pab:
            set_pab_manager(config.pab_config)
        self._set_parallel()

    def _embeds(self, prompt, negative_prompt, L=120):
        cfg = self.transformer.config
        if self._config.text_encoder_fn is not None:
            return self._config.text_encoder_fn(prompt, negative_prompt)
        g = torch.Generator(device="cpu").manual_seed(zlib.crc32(str(prompt).encode()))
        return torch.randn(1, L, cfg.caption_channels, generator=g), torch.randn(1, L, cfg.caption_channels, generator=g)

    @torch.no_grad()
    def generate(self, prompt=None, negative_prompt: str = "", num_inference_steps: int = 50, guidance_scale: float = 7.5,
                 num_images_per_prompt: int = 1, eta: float = 0.0, seed: int = -1, generator=None, latents=None,
                 prompt_embeds=None, negative_prompt_embeds=None, output_type: str = "pil", return_dict: bool = True,
                 callback=None, callback_steps: int = 1, clean_caption: bool = True, mask_feature: bool = True,
                 enable_temporal_attentions: bool = True, verbose: bool = True, video_length: int = 16, height: int = 512,
                 width: int = 512):
        """video_length / height / width are fixed to 16 / 512 / 512 by the reference (:764-766); they are arguments here so
        that the tests can run a small model."""
        update_steps(num_inference_steps)
        self.transformer.reset_pab_state()
        self._maybe_seed(seed)
        dev, dt = self._device, self._dtype
        if prompt_embeds is None:
            prompt_embeds, negative_prompt_embeds = self._embeds(prompt, negative_prompt)
        do_cfg = guidance_scale > 1.0
        pe = prompt_embeds.to(dev, dt)
        if do_cfg:
            pe = torch.cat([negative_prompt_embeds.to(dev, dt), pe], dim=0)
        self.scheduler.set_timesteps(num_inference_steps, dev)
        ts = [int(v) for v in self.scheduler.timesteps.tolist()]
        cin = self.transformer.config.in_channels
        if latents is None:
            latents = torch.randn(prompt_embeds.shape[0], cin, video_length, height // self.vae_scale_factor,
                                  width // self.vae_scale_factor, device=dev, dtype=dt)
        lat = latents.to(dev, dt) * self.scheduler.init_noise_sigma
        for t in ts:
            inp = torch.cat([lat] * 2) if do_cfg else lat
            tt = torch.full((inp.shape[0],), t, device=dev, dtype=torch.int64)
            noise = self.transformer(inp, timestep=tt, all_timesteps=ts, encoder_hidden_states=pe,
                                     added_cond_kwargs={"resolution": None, "aspect_ratio": None},
                                     enable_temporal_attentions=enable_temporal_attentions, return_dict=False,
                                     ts_int=t if enable_pab() else None)[0]
            if do_cfg:
                un, tx = noise.chunk(2)
                noise = un + guidance_scale * (tx - un)
            if self.transformer.config.out_channels // 2 == cin:  # learned sigma: keep the mean prediction (:864-868)
                noise = noise.chunk(2, dim=1)[0]
            lat = self.scheduler.step(noise, t, lat, eta=eta)[0]
        if self._config.vae_decode_fn is not None and output_type != "latents":
            video = self._config.vae_decode_fn(lat)
        else:
            video = lat.float().cpu()  # latents: the VAE is out of scope
        return VideoSysPipelineOutput(video=video) if return_dict else (video,)
