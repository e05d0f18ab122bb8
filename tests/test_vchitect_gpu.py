"""Vchitect-2.0 on the kernels vs the Vchitect oracle (attention pinned bit for bit against the reference's processor,
tests/test_oracle_vs_reference.py; the diffusers norm / embedder pieces restated: parity unpinned): the transformer
forward at a small and at the 2B model's width, temporal attention on both of its paths (<= 32 frames: vsb_attn_short;
more: RoPE pre-pass + vsb_attn_flash on strided views), PAB over eight steps, the pipeline surface."""
import pytest
import torch

from oracle import pab_oracle, synth, vchitect_oracle as VO

pytestmark = pytest.mark.gpu

SMALL = dict(sample_size=8, patch_size=2, in_channels=4, num_layers=3, attention_head_dim=64, num_attention_heads=2,
             joint_attention_dim=48, caption_projection_dim=128, pooled_projection_dim=40, out_channels=4, pos_embed_max_size=12)
SMALL_O = dict(heads=2, head_dim=64, layers=3, patch=2, sample_size=8, pos_embed_max_size=12, out_channels=4)
# Vchitect-2.0-2B's width (24 heads x 64 = 1536, 16 latent channels, 4096-wide captions), 2 of its 24 layers
WIDE = dict(sample_size=128, patch_size=2, in_channels=16, num_layers=2, attention_head_dim=64, num_attention_heads=24,
            joint_attention_dim=4096, caption_projection_dim=1536, pooled_projection_dim=2048, out_channels=16,
            pos_embed_max_size=96)
WIDE_O = dict(heads=24, head_dim=64, layers=2, patch=2, sample_size=128, pos_embed_max_size=96, out_channels=16)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _net(cfg, dt, tag):
    from videosys_b200.models.transformers.vchitect_transformer_3d import VchitectXLTransformerModel

    net = VchitectXLTransformerModel(**cfg)
    sd = synth.fill_state_dict(net.state_dict(), tag)
    sd["pos_embed.pos_embed"] = net.state_dict()["pos_embed.pos_embed"]
    sd = {k: v.to(dt) for k, v in sd.items()}
    net = net.to(dt)
    net.load_state_dict(sd)
    return net.to("cuda:0").eval(), sd


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cfg,ocfg,shape,L", [(SMALL, SMALL_O, (1, 5, 4, 12, 16), 9),       # 5 frames: vsb_attn_short
                                             (SMALL, SMALL_O, (1, 34, 4, 12, 16), 9),      # 34 frames: vsb_attn_short's 64-token instantiation
                                             (SMALL, SMALL_O, (1, 66, 4, 12, 16), 9),      # 66 frames: RoPE pre-pass + flash on strided views
                                             (SMALL, SMALL_O, (1, 1, 4, 12, 16), 9),       # one frame: temporal branch * 0
                                             (WIDE, WIDE_O, (1, 8, 16, 36, 60), 333)])     # 288 x 480, 77 + 256 text tokens
def test_vchitect_forward(cfg, ocfg, shape, L, dt):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    net, sd = _net(cfg, dt, "vchf.")
    lat = synth.normalish("vchf.lat", shape).to(dt)
    enc = synth.normalish("vchf.enc", (1, L, cfg["joint_attention_dim"])).to(dt)
    pooled = synth.normalish("vchf.pool", (1, cfg["pooled_projection_dim"])).to(dt)
    ts = torch.tensor([500.0])
    out = net(lat.cuda(), enc.cuda(), pooled.cuda(), ts.cuda(), return_dict=False)[0].cpu()
    with torch.no_grad():
        r16 = VO.transformer_forward(sd, ocfg, lat, enc, pooled, ts)
        r32 = VO.transformer_forward({k: v.float() for k, v in sd.items()}, ocfg, lat.float(), enc.float(), pooled.float(), ts)
    e_ours, e_ref = _rel(out, r32), _rel(r16, r32)
    print(f"[parity] vchitect forward {dt} latent {shape}: ours-vs-fp32 {e_ours:.3e}, oracle 16-bit-vs-fp32 {e_ref:.3e}")
    assert out.shape == r32.shape
    assert e_ours <= 1.3 * e_ref + 1e-4


def test_vchitect_pab_steps():
    """Eight steps with the temporal / cross / spatial gates on: same hits as the oracle's gates, per-step error within the
    oracle's own 16-bit error, fewer kernels on the steps that reuse."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200 import kernels
    from videosys_b200.core.pab import pab_mgr

    dt = torch.bfloat16
    net, sd = _net(SMALL, dt, "vchp.")
    sd32 = {k: v.float() for k, v in sd.items()}
    cfg = pab_mgr.PABConfig(spatial_broadcast=True, spatial_threshold=[100, 800], spatial_range=2, temporal_broadcast=True,
                            temporal_threshold=[100, 800], temporal_range=3, cross_broadcast=True, cross_threshold=[100, 800],
                            cross_range=4)
    pab_mgr.set_pab_manager(cfg)
    pab_mgr.update_steps(8)
    net.reset_pab_state()
    try:
        G = pab_oracle.PABGate((True, (100, 800), 2), (True, (100, 800), 3), (True, (100, 800), 4), 8)
        Ls = SMALL["num_layers"]
        enc = synth.normalish("vchp.enc", (1, 9, 48)).to(dt)
        pooled = synth.normalish("vchp.pool", (1, 40)).to(dt)
        state = {k: ([{"spatial": 0, "temporal": 0, "cross": 0} for _ in range(Ls)], [{} for _ in range(Ls)]) for k in (16, 32)}
        launches = []
        for step, t in enumerate([900, 700, 650, 600, 550, 500, 450, 50]):
            lat = synth.normalish(f"vchp.lat{step}", (1, 4, 4, 12, 16)).to(dt)
            ts = torch.tensor([float(t)])
            n0 = kernels.launch_count()
            out = net(lat.cuda(), enc.cuda(), pooled.cuda(), ts.cuda(), return_dict=False)[0].cpu()
            launches.append(kernels.launch_count() - n0)
            refs = {}
            for bits, (s, x) in ((16, (sd, lambda v: v)), (32, (sd32, lambda v: v.float()))):
                counts, caches = state[bits]

                def gate(i, kind, counts=counts, t=t):
                    hit, counts[i][kind] = G.gate(kind, t, counts[i][kind])
                    return hit

                with torch.no_grad():
                    refs[bits] = VO.transformer_forward(s, SMALL_O, x(lat), x(enc), x(pooled), ts, gate, caches)
            e_ours, e_ref = _rel(out, refs[32]), _rel(refs[16], refs[32])
            print(f"[parity] vchitect PAB step {step} t={t}: ours-vs-fp32 {e_ours:.3e}, oracle bf16-vs-fp32 {e_ref:.3e}, "
                  f"kernels {launches[-1]}")
            assert e_ours <= 1.3 * e_ref + 1e-4, step
        assert min(launches) < launches[0], launches
    finally:
        pab_mgr.set_pab_manager(None)


def test_vchitect_pipeline_generate():
    """Public surface: VchitectConfig -> VideoSysEngine.generate(...) on a tiny transformer, PAB off and on."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200 import VchitectConfig, VchitectPABConfig, VideoSysEngine, kernels
    from videosys_b200.core.pab import pab_mgr

    kw = dict(num_inference_steps=8, guidance_scale=7.5, seed=0, frames=4, height=96, width=128)
    eng = VideoSysEngine(VchitectConfig(transformer_config=SMALL))
    assert eng.driver_worker._dtype == torch.bfloat16  # reference pipeline_vchitect.py:186
    n0 = kernels.launch_count()
    out = eng.generate("Sunset over the sea.", **kw).video
    n_plain = kernels.launch_count() - n0
    assert out.shape == (1, 4, 4, 12, 16) and torch.isfinite(out).all()
    assert torch.equal(eng.generate("Sunset over the sea.", **kw).video, out), "same seed, same prompt -> same latents"
    eng.shutdown()
    pab = VchitectPABConfig(spatial_threshold=(0, 1001), temporal_threshold=(0, 1001), cross_threshold=(0, 1001))
    eng = VideoSysEngine(VchitectConfig(transformer_config=SMALL, enable_pab=True, pab_config=pab))
    try:
        n0 = kernels.launch_count()
        out2 = eng.generate("Sunset over the sea.", **kw).video
        n_pab = kernels.launch_count() - n0
        assert torch.isfinite(out2).all() and out2.shape == out.shape
        print(f"[pipeline] vchitect kernels launched: plain {n_plain}, PAB {n_pab}")
        assert n_pab < n_plain
    finally:
        pab_mgr.set_pab_manager(None)
        eng.shutdown()
