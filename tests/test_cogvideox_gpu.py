"""CogVideoX on the kernels vs the CogVideoX oracle: block stack (bf16 and fp16), the whole transformer forward, the
joint attention at the real 226 + 17 550-token length, and the pipeline surface (see oracle/cogvideox_oracle.py for what
is pinned against the reference and what is restated)."""
import pytest
import torch

from oracle import cogvideox_oracle as CO, pab_oracle, synth

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _fill(net):
    sd = net.state_dict()
    out = synth.fill_state_dict(sd, "cogx.")
    for k in sd:  # LayerNorm weights ~1 (fill_state_dict treats them as generic matrices / vectors)
        if k.endswith("norm.weight") or k.endswith("norm_q.weight") or k.endswith("norm_k.weight"):
            out[k] = (1.0 + 0.2 * synth.uniform("cogx." + k, tuple(sd[k].shape))).to(sd[k].dtype)
    return out


@pytest.mark.parametrize("BF", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("heads,D,layers,B,Nv,Nt", [(4, 64, 2, 2, 300, 26), (30, 64, 1, 2, 1350, 226)])
def test_cogvideox_block_stack(heads, D, layers, B, Nv, Nt, BF):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200.models.transformers.cogvideox_transformer_3d import CogVideoXBlockStack

    dev = torch.device("cuda:0")
    C = heads * D
    net = CogVideoXBlockStack(heads, D, layers, 512).to(BF)
    sd = _fill(net)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    hid = synth.normalish("cogx.h", (B, Nv, C)).to(BF)
    enc = synth.normalish("cogx.e", (B, Nt, C)).to(BF)
    temb = synth.normalish("cogx.t", (B, 512)).to(BF)
    oh, oe = net(hid.to(dev), enc.to(dev), temb.to(dev))
    oh, oe = oh.cpu(), oe.cpu()
    with torch.no_grad():
        h16, e16 = hid, enc
        h32, e32 = hid.float(), enc.float()
        sd32 = {k: v.float() for k, v in sd.items()}
        for i in range(layers):
            h16, e16 = CO.block(sd, f"transformer_blocks.{i}.", h16, e16, temb, heads)
            h32, e32 = CO.block(sd32, f"transformer_blocks.{i}.", h32, e32, temb.float(), heads)
    for name, o, r16, r32 in (("video", oh, h16, h32), ("text", oe, e16, e32)):
        e_ours, e_ref = _rel(o, r32), _rel(r16, r32)
        print(f"[parity] cogvideox {name} C={C} layers={layers} {BF}: ours-vs-fp32 {e_ours:.3e}, oracle 16-bit-vs-fp32 {e_ref:.3e}, "
              f"bit-equal to oracle bf16 {(o == r16).float().mean().item()*100:.1f} %")
        assert e_ours <= 1.3 * e_ref + 1e-4


def test_cogvideox_pab_spatial_gate():
    """PAB on CogVideoX: spatial gate only, threshold (100, 850), range 2 (pipeline_cogvideox.py:33-44)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200.core.pab import pab_mgr as P
    from videosys_b200.models.transformers.cogvideox_transformer_3d import CogVideoXBlockStack

    dev = torch.device("cuda:0")
    heads, D, layers, B, Nv, Nt = 4, 64, 2, 2, 200, 26
    C = heads * D
    net = CogVideoXBlockStack(heads, D, layers, 512).to(BF)
    sd = _fill(net)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    steps = [999, 859, 839, 819, 499, 79]
    P.set_pab_manager(P.PABConfig(spatial_broadcast=True, spatial_threshold=[100, 850], spatial_range=2))
    P.update_steps(len(steps))
    gate = pab_oracle.PABGate(spatial=(True, (100, 850), 2), steps=len(steps))
    states = [CO.BlockPAB() for _ in range(layers)]
    temb = synth.normalish("cogx.t", (B, 512)).to(BF)
    try:
        for i, t in enumerate(steps):
            hid = synth.normalish(f"cogx.h{i}", (B, Nv, C)).to(BF)
            enc = synth.normalish(f"cogx.e{i}", (B, Nt, C)).to(BF)
            oh, oe = net(hid.to(dev), enc.to(dev), temb.to(dev), timestep=torch.tensor([float(t)] * B))
            h, e = hid, enc
            with torch.no_grad():
                for l in range(layers):
                    h, e = CO.block(sd, f"transformer_blocks.{l}.", h, e, temb, heads, gate, states[l], t)
            for blk, st in zip(net.transformer_blocks, states):
                assert blk.attn_count == st.attn_count
            assert _rel(oh.cpu(), h) < 2e-2 and _rel(oe.cpu(), e) < 2e-2, f"step {i}"
    finally:
        P.set_pab_manager(None)
        net.reset_pab_state()


SMALL = dict(num_attention_heads=4, attention_head_dim=64, in_channels=4, out_channels=4, time_embed_dim=64, text_embed_dim=48,
             num_layers=2, sample_width=16, sample_height=12, sample_frames=9, max_text_seq_length=16)
SMALL_O = dict(heads=4, head_dim=64, layers=2, patch=2, max_text=16, sample_width=16, sample_height=12, sample_frames=9, out_channels=4)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_cogvideox_transformer_forward(dt):
    """CogVideoXTransformer3DModel.forward (time / patch / position embedders, 2 blocks, norm_final, AdaLayerNorm head,
    unpatchify) in the reference's dtype for the 2b model (fp16) and in bf16."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200.models.transformers.cogvideox_transformer_3d import CogVideoXTransformer3DModel

    dev = torch.device("cuda:0")
    net = CogVideoXTransformer3DModel(**SMALL).to(dt)
    sd = _fill(net)
    for k in sd:
        if k.endswith("norm_final.weight") or k.endswith("norm_out.norm.weight"):
            sd[k] = (1.0 + 0.2 * synth.uniform("cogx." + k, tuple(sd[k].shape))).to(dt)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    B, Fr, H, W = 2, 3, 12, 16
    lat = synth.normalish("cogx.lat", (B, Fr, 4, H, W)).to(dt)
    txt = synth.normalish("cogx.txt", (B, 16, 48)).to(dt)
    ts = torch.tensor([499, 499], dtype=torch.int64)
    out = net(lat.to(dev), txt.to(dev), ts.to(dev), return_dict=False)[0].cpu()
    with torch.no_grad():
        r16 = CO.transformer_forward(sd, SMALL_O, lat, txt, ts)
        r32 = CO.transformer_forward({k: v.float() for k, v in sd.items()}, SMALL_O, lat.float(), txt.float(), ts)
    e_ours, e_ref = _rel(out, r32), _rel(r16, r32)
    print(f"[parity] cogvideox forward {dt}: ours-vs-fp32 {e_ours:.3e}, oracle 16-bit-vs-fp32 {e_ref:.3e}")
    assert out.shape == r32.shape
    assert e_ours <= 1.3 * e_ref + 1e-4


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_cogvideox_rotary_forward(dt):
    """use_rotary_positional_embeddings=True (CogVideoX-5b; its dtype is bf16): no position table, q / k of the video tokens
    rotated by the 3-D rotary tables through vsb_qk_rmsnorm_rope's RoPE-only mode (identity rows for the text tokens); the oracle
    is pinned bit for bit against the reference model (tests/test_oracle_vs_reference.py)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200.models.transformers.cogvideox_transformer_3d import CogVideoXTransformer3DModel

    dev = torch.device("cuda:0")
    net = CogVideoXTransformer3DModel(**dict(SMALL, use_rotary_positional_embeddings=True)).to(dt)
    sd = _fill(net)
    for k in sd:
        if k.endswith("norm_final.weight") or k.endswith("norm_out.norm.weight"):
            sd[k] = (1.0 + 0.2 * synth.uniform("cogxr." + k, tuple(sd[k].shape))).to(dt)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    B, Fr, H, W = 2, 3, 12, 16
    lat = synth.normalish("cogxr.lat", (B, Fr, 4, H, W)).to(dt)
    txt = synth.normalish("cogxr.txt", (B, 16, 48)).to(dt)
    ts = torch.tensor([499, 499], dtype=torch.int64)
    rot = CO.rotary_3d(64, CO.resize_crop_region_for_grid((6, 8), 45, 30), (6, 8), Fr)
    out = net(lat.to(dev), txt.to(dev), ts.to(dev), image_rotary_emb=(rot[0].to(dev), rot[1].to(dev)), return_dict=False)[0].cpu()
    with torch.no_grad():
        r16 = CO.transformer_forward(sd, SMALL_O, lat, txt, ts, rotary=rot)
        r32 = CO.transformer_forward({k: v.float() for k, v in sd.items()}, SMALL_O, lat.float(), txt.float(), ts, rotary=rot)
        plain = CO.transformer_forward({k: v.float() for k, v in sd.items()}, SMALL_O, lat.float(), txt.float(), ts)
    e_ours, e_ref = _rel(out, r32), _rel(r16, r32)
    print(f"[parity] cogvideox rotary forward {dt}: ours-vs-fp32 {e_ours:.3e}, oracle 16-bit-vs-fp32 {e_ref:.3e}, "
          f"rotary-vs-table model {_rel(plain, r32):.3e}")
    assert e_ours <= 1.3 * e_ref + 1e-4
    assert _rel(plain, r32) > 10 * e_ref, "the rotary path must matter in this test"


def test_cogvideox_joint_attention_real_length():
    """The joint text + video attention at cfg4's real sequence length (226 + 17 550 = 17 776 tokens, head_dim 64, fp16),
    two heads of the 30: flash kernel vs fp64 attention on a subsample of query rows."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200 import kernels as K

    dev = torch.device("cuda:0")
    N, H, D = 17776, 2, 64
    C = H * D
    qkv = synth.normalish("cogx.long", (1, N, 3, H, D)).to(torch.float16)
    g = qkv.to(dev)
    got = K.attn_flash(g[:, :, 0], g[:, :, 1], g[:, :, 2], 1, N, N, H, D, 3 * C, N * 3 * C, 3 * C, N * 3 * C, D**-0.5).cpu()
    rows = torch.arange(0, N, 97)
    q, k, v = qkv[0].permute(1, 2, 0, 3).double().unbind(0)  # [H, N, D]
    s = (q[:, rows] @ k.transpose(1, 2)) * D**-0.5
    exact = (s.softmax(-1) @ v).permute(1, 0, 2).reshape(len(rows), C)
    err = (got[0, rows].double() - exact).abs()
    print(f"[parity] cogvideox joint attention N={N} fp16: max|err| vs fp64 {err.max().item():.3e}, mean {err.mean().item():.3e}")
    assert (err <= 2.0**-9 * exact.abs().clamp_min(0.02) + 1e-3).all()


def test_cogvideox_pipeline_generate_and_pab():
    """Public surface: CogVideoXConfig -> VideoSysEngine.generate(...) (tiny transformer, 8 DDIM steps), PAB on and off."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from videosys_b200 import CogVideoXConfig, CogVideoXPABConfig, VideoSysEngine, kernels
    from videosys_b200.core.pab import pab_mgr

    kw = dict(height=96, width=128, num_frames=9, num_inference_steps=8, guidance_scale=6, seed=0, max_sequence_length=16)
    eng = VideoSysEngine(CogVideoXConfig("THUDM/CogVideoX-2b", transformer_config=SMALL))
    assert eng.driver_worker._dtype == torch.float16  # reference pipeline_cogvideox.py:138-139
    n0 = kernels.launch_count()
    out = eng.generate("Sunset over the sea.", **kw).video
    n_plain = kernels.launch_count() - n0
    assert out.shape == (1, 3, 4, 12, 16) and torch.isfinite(out).all()
    assert torch.equal(eng.generate("Sunset over the sea.", **kw).video, out), "same seed, same prompt -> same latents"
    eng.shutdown()
    eng = VideoSysEngine(CogVideoXConfig("THUDM/CogVideoX-2b", transformer_config=SMALL, enable_pab=True,
                                         pab_config=CogVideoXPABConfig(spatial_threshold=[0, 1001])))
    try:
        n0 = kernels.launch_count()
        out2 = eng.generate("Sunset over the sea.", **kw).video
        n_pab = kernels.launch_count() - n0
        assert torch.isfinite(out2).all() and out2.shape == out.shape
        print(f"[pipeline] cogvideox kernels launched: plain {n_plain}, PAB {n_pab}")
        assert n_pab < n_plain
    finally:
        pab_mgr.set_pab_manager(None)
        eng.shutdown()
