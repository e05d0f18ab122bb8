#!/usr/bin/env python
"""Per-tile timeline of vsb_attn_flash from in-kernel clock64() stamps (CTA 0 only): where does a KV tile's time go?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosys_b200 import _lib, kernels as K  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16
C, H, D = 1152, 16, 72
nb, n = 40, 3600
qkv = torch.randn(nb, n, 3, H, D, device=dev, dtype=bf)
lib = _lib.load()
trace = torch.zeros(9 * 16 * 4, dtype=torch.int64, device=dev)
for variant, pp, poly in ((2, 1, 0), (3, 1, 0)):
    K.set_option("attn_variant", variant)
    K.set_option("attn_pingpong", pp)
    K.set_option("attn_poly_exp", poly)
    K.attn_flash(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], nb, n, n, H, D, 3 * C, n * 3 * C, 3 * C, n * 3 * C, D**-0.5)
    torch.cuda.synchronize()
    trace.zero_()
    lib.vsb_debug_attn_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K.attn_flash(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], nb, n, n, H, D, 3 * C, n * 3 * C, 3 * C, n * 3 * C, D**-0.5)
    e1.record()
    torch.cuda.synchronize()
    lib.vsb_debug_attn_trace(None)
    t = trace.cpu().view(9, 16, 4)
    base = int(t[t > 0].min())
    print(f"\n=== variant {variant} pingpong {pp} poly {poly}: kernel {e0.elapsed_time(e1):.3f} ms; clock64 deltas (cycles), CTA 0 ===")
    print("softmax warp per tile: wait->loaded, loaded->exps done, exps->arrived, arrive->next s_full | period")
    actors = (1, 2) if variant == 0 else range(1, 9)
    for a in actors:
        name = f"WG{'AB'[a-1]}" if variant == 0 else f"WG{'AB'[(a-1)//4]} warp {a+3:2d} (SMSP {(a+3)%4})"
        for j in range(2, 10):
            r, nx = t[a, j], t[a, j + 1]
            print(f"  {name} tile {j}: start@{int(r[0])-base:7d}  ld {int(r[1]-r[0]):5d}  exp {int(r[2]-r[1]):5d}  st+arrive {int(r[3]-r[2]):5d}  arrive@{int(r[3])-base:7d}  idle {int(nx[0]-r[3]):5d} | period {int(nx[0]-r[0]):5d}")
    print("MMA thread per tile: [p_full A seen, issued PV_A+S_A', p_full B seen, issued PV_B+S_B']")
    for j in range(1, 10):
        r = t[0, j]
        print(f"  tile {j}: pA@{int(r[0])-base:7d} issueA {int(r[1]-r[0]):4d}  pB@{int(r[2])-base:7d} issueB {int(r[3]-r[2]):4d}")
K.set_option('attn_variant', -1); K.set_option('attn_poly_exp', 0)
