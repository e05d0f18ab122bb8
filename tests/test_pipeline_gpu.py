"""Public API smoke on the GPU: VideoSysEngine(OpenSoraConfig(...)).generate(...) through RFLOW + STDiT3 on the kernels
(tiny transformer, 144p, 17 frames, 4 steps; the reference's own pipeline tests are smoke tests too)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    from videosys_b200 import OpenSoraConfig
    from videosys_b200.models.transformers.open_sora_transformer_3d import STDiT3Config

    tcfg = STDiT3Config(hidden_size=288, num_heads=4, depth=2, caption_channels=64, model_max_length=24)
    return OpenSoraConfig(num_sampling_steps=4, cfg_scale=7.0, transformer_config=tcfg, **kw)


def test_engine_generate_latents_and_pab_speedup_path():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200 import OpenSoraPABConfig, VideoSysEngine
    from videosys_b200 import kernels
    from videosys_b200.core.pab import pab_mgr

    eng = VideoSysEngine(_cfg())
    # un-zero the temporal projections so every kernel contributes (reference init zeros them, :508-511)
    for blk in eng.driver_worker.transformer.temporal_blocks:
        for lin in (blk.attn.proj, blk.cross_attn.proj, blk.mlp.fc2):
            torch.nn.init.normal_(lin.weight, std=0.02)
    n0 = kernels.executed_launch_count()
    out = eng.generate("Sunset over the sea.", resolution="144p", aspect_ratio="9:16", num_frames=17, seed=0, verbose=False)
    n_plain = kernels.executed_launch_count() - n0
    lat = out.video
    assert lat.shape == (1, 4, 5, 18, 32) and torch.isfinite(lat).all()
    out2 = eng.generate("Sunset over the sea.", resolution="144p", aspect_ratio="9:16", num_frames=17, seed=0, verbose=False)
    assert torch.equal(out2.video, lat), "same seed, same prompt -> same latents"
    eng.shutdown()

    # PAB on: wide thresholds so that the 4-step schedule actually skips; fewer kernels must launch
    pab = OpenSoraPABConfig(spatial_threshold=(0, 1001), temporal_threshold=(0, 1001), cross_threshold=(0, 1001))
    eng = VideoSysEngine(_cfg(enable_pab=True, pab_config=pab))
    try:
        n0 = kernels.executed_launch_count()
        out3 = eng.generate("Sunset over the sea.", resolution="144p", aspect_ratio="9:16", num_frames=17, seed=0, verbose=False)
        n_pab = kernels.executed_launch_count() - n0
        assert torch.isfinite(out3.video).all() and out3.video.shape == lat.shape
        print(f"[pipeline] kernels launched: plain {n_plain}, PAB {n_pab}")
        assert n_pab < n_plain
    finally:
        pab_mgr.set_pab_manager(None)
        eng.shutdown()
