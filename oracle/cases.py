"""TEST INFRASTRUCTURE ONLY -- the synthetic parity cases shared by gen_golden.py and the tests.

All inputs and weights are pure functions of the case name (oracle/synth.py), so only the reference's
OUTPUTS are committed under tests/golden/.
"""
import torch

from . import synth

SMALL = dict(hidden_size=288, num_heads=4, depth=1, caption_channels=64, model_max_length=20)


def small_model_cfg(depth=1):
    c = dict(SMALL)
    c["depth"] = depth
    return c


def oracle_cfg(c):
    return dict(hidden_size=c["hidden_size"], num_heads=c["num_heads"], depth=c["depth"])


def forward_inputs(dtype, B=2, T=5, H=12, W=11, L=20, n_valid=15, cap=64):
    x = synth.normalish("fwd.x", (B, 4, T, H, W))
    ts = torch.tensor([700.0] * B)
    y = synth.normalish("fwd.y", (B, 1, L, cap))
    mask = torch.ones(1, L, dtype=torch.long)
    mask[0, n_valid:] = 0
    x_mask = torch.ones(B, T, dtype=torch.bool)
    x_mask[:, 0] = False
    fps = torch.tensor([24.0] * B).to(dtype)
    h = torch.tensor([240.0] * B).to(dtype)
    w = torch.tensor([426.0] * B).to(dtype)
    return dict(x=x, timestep=ts, y=y, mask=mask, x_mask=x_mask, fps=fps, height=h, width=w)


def block_inputs(dtype, C=288, B=2, T=5, S=36, L=15, tag="blk"):
    x = synth.normalish(tag + ".x", (B, T * S, C)).to(dtype)
    y = synth.normalish(tag + ".y", (1, B * L, C)).to(dtype)
    t = synth.normalish(tag + ".t", (B, 6 * C), std=0.5).to(dtype)
    t0 = synth.normalish(tag + ".t0", (B, 6 * C), std=0.5).to(dtype)
    x_mask = torch.ones(B, T, dtype=torch.bool)
    x_mask[0, 0] = False
    x_mask[1, T - 1] = False
    return dict(x=x, y=y, t=t, t0=t0, x_mask=x_mask, y_lens=[L] * B, T=T, S=S)
