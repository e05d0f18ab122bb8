"""CogVideoX blocks on the kernels vs the CogVideoX oracle (block level, bf16; see oracle/cogvideox_oracle.py for what
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


@pytest.mark.parametrize("heads,D,layers,B,Nv,Nt", [(4, 64, 2, 2, 300, 26), (30, 64, 1, 2, 1350, 226)])
def test_cogvideox_block_stack(heads, D, layers, B, Nv, Nt):
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
        print(f"[parity] cogvideox {name} C={C} layers={layers}: ours-vs-fp32 {e_ours:.3e}, oracle bf16-vs-fp32 {e_ref:.3e}, "
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
