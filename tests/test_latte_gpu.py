"""Latte blocks on the kernels vs the Latte oracle (block level; the oracle is a restatement: parity unpinned, see
oracle/latte_oracle.py).  Config 1 of BASELINE.json: Latte 16x256x256 -> latent 16 x (16x16 patches), hidden 1152."""
import pytest
import torch

from oracle import latte_oracle as LO, synth

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("C,H,depth,B,Fr,S,L", [(288, 4, 2, 2, 6, 36, 20), (1152, 16, 1, 2, 16, 256, 120)])
def test_latte_block_stack(C, H, depth, B, Fr, S, L):
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
