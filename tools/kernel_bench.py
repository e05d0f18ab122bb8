#!/usr/bin/env python
"""Per-shape microbenchmarks on the GPU box: our GEMM (1-CTA / CTA-pair) vs cuBLAS, our flash attention vs SDPA,
elementwise kernels vs the HBM copy peak.  Library kernels are timed only as yardsticks (never on the product path)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosys_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


out = {}
M = 144000


def elem_section():
    C, H, D = 1152, 16, 72
    x = torch.randn(2, 72000, C, device=dev, dtype=bf)
    y = torch.randn(2, 72000, C, device=dev, dtype=bf)
    mod = torch.randn(2, 2, 6, C, device=dev, dtype=bf)
    m8 = torch.ones(2, 20, dtype=torch.uint8, device=dev)
    o = torch.empty_like(x)
    nb = x.numel() * 2
    r = {}
    t = timeit(lambda: K.ln_modulate(x, mod, m8, 0, 1, 2, 20, 3600, out=o)); r["ln_modulate_GBs"] = round(2 * nb / t / 1e9)
    K.set_option("ln_occupancy", 4)
    t = timeit(lambda: K.ln_modulate(x, mod, m8, 0, 1, 2, 20, 3600, out=o)); r["ln_modulate_occ4_GBs"] = round(2 * nb / t / 1e9)
    K.set_option("ln_occupancy", 3)
    t = timeit(lambda: K.gate_residual(x, y, mod, m8, 2, 2, 20, 3600, out=o)); r["gate_residual_GBs"] = round(3 * nb / t / 1e9)
    t = timeit(lambda: K.residual_add(x, y, out=o)); r["residual_add_GBs"] = round(3 * nb / t / 1e9)
    t = timeit(lambda: o.copy_(x)); r["torch_copy_GBs"] = round(2 * nb / t / 1e9)
    qkv = torch.randn(144000, 3, H, D, device=dev, dtype=bf)
    wq = torch.ones(D, device=dev, dtype=bf)
    t = timeit(lambda: K.qk_rmsnorm_(qkv, wq, wq, H, D)); r["qk_rmsnorm_GBs"] = round(4 * 144000 * C * 2 / t / 1e9)
    cos = torch.randn(20, D, device=dev); sin = torch.randn(20, D, device=dev)
    t = timeit(lambda: K.attn_short(qkv, wq, wq, cos, sin, 2, 3600, 20 * 3600, 1, 3600, 20, H, D, D**-0.5, out=o.view(-1, C)))
    r["attn_short_GBs"] = round(4 * 144000 * C * 2 / t / 1e9); r["attn_short_ms"] = round(t * 1e3, 3)
    # patch embedding (vsb_patch_embed) against the torch / cuDNN chain it replaces, 720p CFG pair
    z = torch.randn(2, 4, 20, 90, 160, device=dev, dtype=bf)
    conv = torch.nn.Conv3d(4, C, (1, 2, 2), (1, 2, 2)).to(dev, bf)
    pos = torch.randn(1, 3600, C, device=dev, dtype=bf)
    t = timeit(lambda: K.patch_embed(z, conv.weight, conv.bias, pos[0], 2, 2)); r["patch_embed_ms"] = round(t * 1e3, 4)
    r["patch_embed_GBs"] = round(2 * 72000 * C * 2 / t / 1e9)
    with torch.no_grad():
        t = timeit(lambda: (conv(z).flatten(2).transpose(1, 2).reshape(2, 20, 3600, C) + pos).reshape(2, 72000, C).contiguous())
    r["patch_embed_torch_chain_ms"] = round(t * 1e3, 4)
    return r


if os.environ.get("KB_ONLY", "") == "elem":
    out["elementwise"] = elem_section()
    print("elementwise", out["elementwise"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/kernel_bench_elem.json", "w"), indent=1)
    sys.exit(0)
ONLY = os.environ.get("KB_ONLY", "")  # "attn": skip the GEMM and elementwise sections
SHAPES = {} if ONLY == "attn" else {"qkv": (3456, 1152, 0), "proj": (1152, 1152, 0), "fc1": (4608, 1152, 1), "fc2": (1152, 4608, 0)}
for name, (N, Kd, act) in SHAPES.items():
    a = torch.randn(M, Kd, device=dev, dtype=bf)
    w = torch.randn(N, Kd, device=dev, dtype=bf) * 0.02
    b = torch.randn(N, device=dev, dtype=bf)
    o = torch.empty(M, N, device=dev, dtype=bf)
    fl = 2.0 * M * N * Kd
    r = {}
    for opt in (0, 1):
        K.set_option("gemm_2sm", opt)
        t = timeit(lambda: K.gemm_bias_act(a, w, b, act=act, out=o))
        r["2sm" if opt else "1sm"] = round(fl / t / 1e12, 1)
    t = timeit(lambda: torch.nn.functional.linear(a, w, b))
    r["cublas_linear"] = round(fl / t / 1e12, 1)
    out["gemm_" + name] = r
    print("gemm", name, r, flush=True)
    del a, w, o
K.set_option("gemm_2sm", 1)

# gate + residual fused into the GEMM epilogue (residual tile TMA-prefetched into the staging buffer) vs the two kernels
if ONLY != "attn":
    xres = torch.randn(2, 72000, 1152, device=dev, dtype=bf)
    modg = torch.randn(2, 2, 6, 1152, device=dev, dtype=bf)
    m8g = torch.ones(2, 20, dtype=torch.uint8, device=dev)
    for name, Kd in (("proj", 1152), ("fc2", 4608)):
        a = torch.randn(M, Kd, device=dev, dtype=bf)
        w = torch.randn(1152, Kd, device=dev, dtype=bf) * 0.02
        b = torch.randn(1152, device=dev, dtype=bf)
        yb = torch.empty(M, 1152, device=dev, dtype=bf)

        def unfused():
            K.gemm_bias_act(a, w, b, out=yb)
            K.gate_residual(xres, yb.view(2, 72000, 1152), modg, m8g, 2, 2, 20, 3600, out=xres)

        def fused():
            K.gemm_bias_residual(a, w, b, xres, modg, m8g, 2, 2, 20, 3600)

        tu, tf = [], []
        for _ in range(6):
            tu.append(timeit(unfused, iters=8, warm=2))
            tf.append(timeit(fused, iters=8, warm=2))
        tg = timeit(lambda: K.gemm_bias_act(a, w, b, out=yb), iters=8, warm=2)
        out["fused_epilogue_" + name] = {"gemm_plus_gate_residual_ms": round(sorted(tu)[3] * 1e3, 4), "fused_ms": round(sorted(tf)[3] * 1e3, 4),
                                          "gemm_alone_ms": round(tg * 1e3, 4)}
        print("fused epilogue", name, out["fused_epilogue_" + name], flush=True)
        del a, w, yb
    del xres

# attention: spatial 720p (40 x 16 heads x 3600 x 72), cross (2 x 16 x 72000 x 300)
C, H, D = 1152, 16, 72
qkv = torch.randn(40, 3600, 3, H, D, device=dev, dtype=bf)
fl = 4.0 * 40 * H * 3600 * 3600 * D
# Variants are timed ROUND-ROBIN after a warm-up that brings the GPU to its sustained (power-capped) clocks: timed one
# after the other, whichever ran first on a cool GPU looked 20-40 % faster than it is inside a denoising step.
def spatial(var, poly):
    K.set_option("attn_variant", var)
    K.set_option("attn_poly_exp", poly if var != 6 else 0)
    K.attn_flash(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 40, 3600, 3600, H, D, 3 * C, 3600 * 3 * C, 3 * C, 3600 * 3 * C, D**-0.5)


qx = torch.randn(2, 72000, H, D, device=dev, dtype=bf)
kvx = torch.randn(2, 300, 2, H, D, device=dev, dtype=bf)
flx = 4.0 * 2 * H * 72000 * 300 * D


def cross(var, poly):
    K.set_option("attn_variant", var)
    K.set_option("attn_poly_exp", poly if var != 6 else 0)
    K.set_option("attn_pingpong", 0 if (var == 6 and poly == 9) else 1)  # variant 6: poly slot 9 = pipelines NOT staggered
    K.attn_flash(qx, kvx[:, :, 0], kvx[:, :, 1], 2, 72000, 300, H, D, C, 72000 * C, 2 * C, 300 * 2 * C, D**-0.5)


VARIANTS = {"kt128_pingpong": (0, 0), "kt64": (2, 0), "kt64_poly25": (2, 1), "kt64_poly37": (2, 2), "kt64p": (3, 0), "kt64p_poly25": (3, 1),
            "kt64_qtmem": (4, 0), "kt64_qtmem_poly25": (4, 1), "kt64_qtmem_poly37": (4, 2), "kt64_qtmem_poly50": (4, 3),
            "kvres_cross_only": (6, 0), "kvres_cross_only_nostagger": (6, 9), "kt64_qtsum": (5, 0), "kt64_qtsum_poly25": (5, 1), "kt64_qtsum_poly37": (5, 2), "kt64_qtsum_poly50": (5, 3)}
# cuDNN's fused attention on the same problem, inside the same round-robin (a yardstick: never on the product path)
from torch.nn.attention import SDPBackend, sdpa_kernel  # noqa: E402

q_l, k_l, v_l = qkv.permute(2, 0, 3, 1, 4).unbind(0)


def cudnn_spatial():
    with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
        torch.nn.functional.scaled_dot_product_attention(q_l, k_l, v_l)


try:
    cudnn_spatial()
    torch.cuda.synchronize()
    HAVE_CUDNN = True
except Exception as e:  # noqa: BLE001
    print("cudnn sdpa unavailable:", type(e).__name__, str(e)[:120])
    HAVE_CUDNN = False
cudnn_ts = []
for _ in range(600):  # ~2.5 s of attention: clocks settle under the power cap
    spatial(2, 0)
torch.cuda.synchronize()
acc = {k: ([], []) for k in VARIANTS}
for rnd in range(24):
    for nm, (var, poly) in VARIANTS.items():
        acc[nm][0].append(timeit(lambda: spatial(var, poly), iters=3, warm=1))
        acc[nm][1].append(timeit(lambda: cross(var, poly), iters=6, warm=1))
    if HAVE_CUDNN:
        cudnn_ts.append(timeit(cudnn_spatial, iters=3, warm=1))
r = {}
if cudnn_ts:
    cudnn_ts.sort()
    r["cudnn_sdpa_interleaved"] = round(fl / cudnn_ts[len(cudnn_ts) // 2] / 1e12, 1)
for nm, (ts, tx) in acc.items():
    ts, tx = sorted(ts), sorted(tx)
    r["ours_" + nm] = round(fl / ts[len(ts) // 2] / 1e12, 1)                       # median of 24 interleaved rounds
    r["ours_" + nm + "_spread"] = [round(fl / ts[-1] / 1e12), round(fl / ts[0] / 1e12)]
    r["cross_" + nm + "_ms"] = round(tx[len(tx) // 2] * 1e3, 3)
K.set_option("attn_variant", -1)
K.set_option("attn_poly_exp", 0)
del qx, kvx
q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
for nm, be in (() if ONLY == "attn" else (("sdpa_flash", "FLASH_ATTENTION"), ("sdpa_cudnn", "CUDNN_ATTENTION"), ("sdpa_efficient", "EFFICIENT_ATTENTION"))):
    try:
        from torch.nn.attention import SDPBackend, sdpa_kernel

        with sdpa_kernel([getattr(SDPBackend, be)]):
            t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), iters=5)
        r[nm] = round(fl / t / 1e12, 1)
    except Exception as e:
        r[nm] = f"n/a ({type(e).__name__})"
out["attn_spatial_720p"] = r
print("attn spatial", r, flush=True)
del qkv

if ONLY == "attn":
    json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "kernel_bench_attn.json"), "w"), indent=1)
    sys.exit(0)
r = elem_section()
out["elementwise"] = r
print("elementwise", r, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/kernel_bench.json", "w"), indent=1)
