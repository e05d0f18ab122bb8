#!/usr/bin/env python
"""Times vsb_attn_flash at Open-Sora-Plan v1.2.0's shapes (head_dim 96 -> csrc/attn_mma.cu, the warp-level kernel): self-attention
of the CFG pair over the 8 x 30 x 40 = 9600 tokens of a 29-frame 480p video, 24 heads; and its 512-key text cross attention.
CUDA events around 5 launches after 2 warm-ups; prints one JSON line (profiles/r02_attn_mma_bench.json)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosys_b200 import kernels as K  # noqa: E402


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    B, N, H, D, L = 2, 9600, 24, 96, 512
    C = H * D
    torch.manual_seed(0)
    qkv = torch.randn(B * N, 3, C, device=dev, dtype=torch.bfloat16)
    q = torch.randn(B * N, C, device=dev, dtype=torch.bfloat16)
    kv = torch.randn(B * L, 2, C, device=dev, dtype=torch.bfloat16)
    ms_self = timed(lambda: K.attn_flash(qkv[:, 0], qkv[:, 1], qkv[:, 2], B, N, N, H, D, 3 * C, N * 3 * C, 3 * C, N * 3 * C, D**-0.5))
    ms_cross = timed(lambda: K.attn_flash(q, kv[:, 0], kv[:, 1], B, N, L, H, D, C, N * C, 2 * C, L * 2 * C, D**-0.5, kv_lens=[512, 300]))
    # the tcgen05 kernel on the nearest shape it supports (head_dim 64, same token count and heads) for scale
    C64 = H * 64
    qkv64 = torch.randn(B * N, 3, C64, device=dev, dtype=torch.bfloat16)
    ms_64 = timed(lambda: K.attn_flash(qkv64[:, 0], qkv64[:, 1], qkv64[:, 2], B, N, N, H, 64, 3 * C64, N * 3 * C64, 3 * C64, N * 3 * C64, 0.125))
    fl = lambda d, nk: 4.0 * B * H * N * nk * d  # noqa: E731
    print(json.dumps({"shape": {"B": B, "tokens": N, "heads": H, "head_dim": D, "text_keys": L},
                      "attn_mma_self_ms": ms_self, "attn_mma_self_tflops": fl(D, N) / ms_self / 1e9,
                      "attn_mma_cross_ms": ms_cross, "attn_mma_cross_tflops": fl(D, L) / ms_cross / 1e9,
                      "tcgen05_head_dim64_self_ms": ms_64, "tcgen05_head_dim64_self_tflops": fl(64, N) / ms_64 / 1e9}))


if __name__ == "__main__":
    main()
