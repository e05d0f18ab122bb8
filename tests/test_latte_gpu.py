"""Latte on the kernels vs the Latte oracle: block stack (bf16 / fp16), the whole LatteT2V forward at config 1's shape,
PAB gates, and the pipeline surface (the oracle is a restatement: parity unpinned, see oracle/latte_oracle.py).  Config 1 of BASELINE.json: Latte 16x256x256 -> latent 16 x (16x16 patches), hidden 1152."""
import pytest
import torch

from oracle import latte_oracle as LO, synth

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("BF", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C,H,depth,B,Fr,S,L", [(288, 4, 2, 2, 6, 36, 20), (1152, 16, 1, 2, 16, 256, 120)])
def test_latte_block_stack(C, H, depth, B, Fr, S, L, BF):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200.models.transformers.latte_transformer_3d import LatteBlockStack

    dev = torch.device("cuda:0")
    net = LatteBlockStack(C, H, depth).to(BF)
    sd = synth.fill_state_dict(net.state_dict(), "latte.")
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    x = synth.normalish("latte.x", (B, Fr, S, C)).to(BF)
    enc = synth.normalish("latte.enc", (B, L, C)).to(BF)
    t6 = synth.normalish("latte.t", (B, 6 * C), std=0.5).to(BF)
    tpe = synth.normalish("latte.tpe", (1, Fr, C), std=0.3).to(BF)
    out = net(x.to(dev), enc.to(dev), t6.to(dev), tpe.to(dev)).cpu()
    with torch.no_grad():
        ref16 = LO.block_stack(sd, x, enc, t6, H, depth, tpe)
        sd32 = {k: v.float() for k, v in sd.items()}
        ref32 = LO.block_stack(sd32, x.float(), enc.float(), t6.float(), H, depth, tpe.float())
    e_ours, e_ref = _rel(out, ref32), _rel(ref16, ref32)
    print(f"[parity] latte stack C={C} depth={depth}: ours-vs-fp32 {e_ours:.3e}, oracle bf16-vs-fp32 {e_ref:.3e}, "
          f"bit-equal to oracle bf16 {(out == ref16).float().mean().item()*100:.1f} %")
    assert e_ours <= 1.25 * e_ref + 1e-4


SMALL = dict(num_attention_heads=4, attention_head_dim=72, in_channels=4, out_channels=8, num_layers=2, sample_size=16,
             caption_channels=64, video_length=6)
SMALL_O = dict(heads=4, head_dim=72, layers=2, patch=2, sample_size=16, out_channels=8, video_length=6)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("cfg,ocfg,shape,L", [(SMALL, SMALL_O, (2, 4, 6, 16, 16), 20),
                                             # BASELINE.json configs[0]: Latte 16 x 256 x 256 -> latent [2, 4, 16, 32, 32], 120 text
                                             # tokens, hidden 1152 = 16 heads x 72 (2 of the 28 layers)
                                             (dict(num_layers=2, sample_size=32, caption_channels=4096, video_length=16),
                                              dict(heads=16, head_dim=72, layers=2, patch=2, sample_size=32, out_channels=8, video_length=16),
                                              (2, 4, 16, 32, 32), 120)])
def test_latte_transformer_forward(cfg, ocfg, shape, L, dt):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200.models.transformers.latte_transformer_3d import LatteT2V

    dev = torch.device("cuda:0")
    net = LatteT2V(**cfg).to(dt)
    sd = synth.fill_state_dict(net.state_dict(), "lattef.")
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    lat = synth.normalish("lattef.lat", shape).to(dt)
    txt = synth.normalish("lattef.txt", (shape[0], L, net.config.caption_channels)).to(dt)
    ts = torch.tensor([999] * shape[0], dtype=torch.int64)
    out = net(lat.to(dev), timestep=ts.to(dev), encoder_hidden_states=txt.to(dev), return_dict=False)[0].cpu()
    with torch.no_grad():
        r16 = LO.transformer_forward(sd, ocfg, lat, ts, txt)
        r32 = LO.transformer_forward({k: v.float() for k, v in sd.items()}, ocfg, lat.float(), ts, txt.float())
    e_ours, e_ref = _rel(out, r32), _rel(r16, r32)
    print(f"[parity] latte forward {dt} latent {shape}: ours-vs-fp32 {e_ours:.3e}, oracle 16-bit-vs-fp32 {e_ref:.3e}")
    assert out.shape == r32.shape
    assert e_ours <= 1.3 * e_ref + 1e-4


def test_latte_pipeline_generate_and_pab():
    """Public surface: LatteConfig -> VideoSysEngine.generate(...) (tiny transformer, 10 DDIM steps), PAB (attention
    broadcast + MLP skip) on and off; the PAB run launches fewer kernels and stays finite."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200 import LatteConfig, LattePABConfig, VideoSysEngine, kernels
    from videosys_b200.core.pab import pab_mgr

    kw = dict(num_inference_steps=10, guidance_scale=7.5, seed=0, video_length=6, height=128, width=128)
    eng = VideoSysEngine(LatteConfig(transformer_config=SMALL))
    assert eng.driver_worker._dtype == torch.float16  # reference pipeline_latte.py:201
    n0 = kernels.launch_count()
    out = eng.generate("Sunset over the sea.", **kw).video
    n_plain = kernels.launch_count() - n0
    assert out.shape == (1, 4, 6, 16, 16) and torch.isfinite(out).all()
    assert torch.equal(eng.generate("Sunset over the sea.", **kw).video, out), "same seed, same prompt -> same latents"
    eng.shutdown()
    mlp = {900: {"block": [0, 1], "skip_count": 2}, 500: {"block": [0], "skip_count": 2}}
    pab = LattePABConfig(spatial_threshold=(0, 1001), temporal_threshold=(0, 1001), cross_threshold=(0, 1001),
                         mlp_spatial_broadcast_config=mlp, mlp_temporal_broadcast_config=mlp)
    eng = VideoSysEngine(LatteConfig(transformer_config=SMALL, enable_pab=True, pab_config=pab))
    try:
        n0 = kernels.launch_count()
        out2 = eng.generate("Sunset over the sea.", **kw).video
        n_pab = kernels.launch_count() - n0
        assert torch.isfinite(out2).all() and out2.shape == out.shape
        assert not pab_mgr.PAB_MANAGER.get_spatial_mlp_outputs() and not pab_mgr.PAB_MANAGER.get_temporal_mlp_outputs(), \
            "every stored MLP output must have been consumed and released (pab_mgr.py:140-143)"
        print(f"[pipeline] latte kernels launched: plain {n_plain}, PAB {n_pab}")
        assert n_pab < n_plain
    finally:
        pab_mgr.set_pab_manager(None)
        eng.shutdown()
