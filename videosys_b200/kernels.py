"""Thin torch-tensor front end of the C-ABI (videosys_b200/_lib.py): pointer + shape marshalling only.

Every function enqueues exactly one sm_100a kernel on the current CUDA stream (capturable in a CUDA graph).  CPU tensors are rejected:
there is no fallback path.
"""
import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib

_initialised = set()

# bench.py sets this to a list to time individual launches with CUDA events on the launching stream:
# entries are (kind, start_event, end_event, algorithmic_work) -- work = FLOPs for gemm/attn, bytes otherwise.
# PROFILE_KINDS (a set, or None for every kind) limits which launches carry events: inside bench.py's timed region
# only the dominant kernel does, so the measurement itself costs < 1 % (2 events x 3472 launches cost ~1 % of a step).
PROFILE = None
PROFILE_KINDS = None


class _Timed:
    __slots__ = ("kind", "work", "e0", "cancelled")

    def __init__(self, kind, work):
        self.kind, self.work = kind, work
        self.cancelled = False

    def __enter__(self):
        self.e0 = None
        if PROFILE is not None and (PROFILE_KINDS is None or self.kind in PROFILE_KINDS):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def cancel(self):
        """Nothing was launched inside this block: record no entry."""
        self.cancelled = True

    def __exit__(self, *exc):
        if self.e0 is not None and not self.cancelled:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.kind, self.e0, e1, self.work))
        return False


def _init_dev(lib, dev):
    if dev not in _initialised:
        _lib.check(lib.vsb_init(dev), "init")
        _initialised.add(dev)


def _prep(*tensors):
    lib = _lib.load()
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.VsbError("vsb200 kernels need CUDA tensors (no CPU fallback)")
        if not t.is_contiguous():
            raise _lib.VsbError("vsb200 kernels need contiguous tensors")
        dev = t.device.index if dev is None else dev
    _init_dev(lib, dev)
    return lib, torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def require_cuda(t, what: str = "vsb200", half_only: bool = False):
    """The model front ends call this first: there is no CPU execution path (and the kernels take 16-bit activations)."""
    if not t.is_cuda:
        raise RuntimeError(f"videosys_b200 {what} runs on sm_100a CUDA devices only (no CPU path)")
    if half_only and t.dtype not in (torch.bfloat16, torch.float16):
        raise RuntimeError(f"videosys_b200 {what} runs in fp16 / bf16 only")


def _bf16(t, name):
    if t is not None and t.dtype not in (torch.bfloat16, torch.float16):
        raise _lib.VsbError(f"{name} must be bfloat16 or float16")


def _fn(lib, name, ref):
    """The entry for ref's dtype: bf16 -> `name`, IEEE fp16 -> its twin `name_f16` (include/vsb200.h)."""
    if ref.dtype == torch.float16:
        return getattr(lib, name + "_f16")
    if ref.dtype != torch.bfloat16:
        raise _lib.VsbError(f"vsb200 kernels take bfloat16 or float16 tensors, got {ref.dtype}")
    return getattr(lib, name)


def _same_dtype(*tensors):
    dt = None
    for t in tensors:
        if t is None or not t.dtype.is_floating_point or t.dtype == torch.float32:
            continue
        if dt is None:
            dt = t.dtype
        elif t.dtype != dt:
            raise _lib.VsbError(f"vsb200 kernels need one 16-bit dtype per call, got {dt} and {t.dtype}")


def set_option(name: str, value: int) -> None:
    _lib.check(_lib.load().vsb_set_option(name.encode(), int(value)), "set_option")


def launch_count() -> int:
    return int(_lib.load().vsb_launch_count())


# kernels executed by replaying captured CUDA graphs (core/graph_step.py adds a graph's launch count on every replay
# after the first; the first replay's launches were counted by vsb_launch_count at capture time)
GRAPH_REPLAYED_LAUNCHES = 0


def executed_launch_count() -> int:
    """Kernels of this library that have run: host launches + launches inside replayed CUDA graphs."""
    return launch_count() + GRAPH_REPLAYED_LAUNCHES


def tmap_cache_stats():
    """(hits, misses) of the encoded-tensor-map cache; misses = host cuTensorMapEncodeTiled calls made."""
    lib = _lib.load()
    return int(lib.vsb_tmap_cache_stats(0)), int(lib.vsb_tmap_cache_stats(1))


def modulation_table(table: torch.Tensor, t: torch.Tensor, t0: Optional[torch.Tensor]) -> torch.Tensor:
    """[2, B, rows, C] = table + t (and + t0); rows = table.shape[0] (6 for blocks, 2 for the final layer)."""
    lib, st = _prep(table, t, t0)
    rows, Cc = table.shape
    B = t.shape[0]
    _same_dtype(table, t, t0)
    mod = torch.empty(2, B, rows, Cc, dtype=table.dtype, device=table.device)
    _lib.check(_fn(lib, "vsb_modulation_table", table)(_p(table), _p(t), _p(t0), _p(mod), B, Cc, rows, st), "modulation_table")
    return mod


def patch_embed(z, weight, bias, pos, ph: int, pw: int, s0: int = 0, s_local: Optional[int] = None):
    """Tokens [B, T, S_local, C] of a latent z [B, Cin, T, H, W] (contiguous): the (1, ph, pw)-strided patch convolution
    (weight = conv.weight, any [C, Cin, (1,) ph, pw] shape), + bias, + pos [S, C], for patch columns s0 .. s0+S_local-1
    (columns beyond the grid are zero).  Returns None when the kernel does not take the shape (Cin*ph*pw != 16)."""
    lib, st = _prep(z, weight, bias, pos)
    _bf16(z, "z")
    _same_dtype(z, weight, bias, pos)
    B, Cin, T, H, W = z.shape
    Cc = weight.shape[0]
    S = -(-H // ph) * -(-W // pw)
    s_local = S if s_local is None else s_local
    out = torch.empty(B, T, s_local, Cc, dtype=z.dtype, device=z.device)
    with _Timed("patch_embed", out.numel() * 2) as tm:
        rc = _fn(lib, "vsb_patch_embed", z)(_p(z), _p(weight), _p(bias), _p(pos), _p(out), B, Cin, T, H, W, Cin * T * H * W,
                                            T * H * W, H * W, ph, pw, Cc, s0, s_local, st)
        if rc == 1:
            tm.cancel()
    if rc == 1:
        return None
    _lib.check(rc, "patch_embed")
    return out


def ln_modulate(x, mod, x_mask_u8, shift_row, scale_row, B, T, S, out=None, eps=1e-6, gamma=None, beta=None):
    """LayerNorm (optionally affine: gamma/beta) + modulate + per-frame select; x viewed as [B, T, S, C]."""
    lib, st = _prep(x, mod, x_mask_u8, out, gamma, beta)
    _bf16(x, "x")
    _same_dtype(x, mod, out, gamma, beta)
    Cc = x.shape[-1]
    out = torch.empty_like(x) if out is None else out
    with _Timed("ln_modulate", 2 * x.numel() * 2):
        if gamma is None and beta is None:  # STDiT3 / Latte: LayerNorm without affine parameters
            rc = _fn(lib, "vsb_ln_modulate", x)(_p(x), _p(out), _p(mod), _p(x_mask_u8), shift_row, scale_row, B, T, S, Cc, eps, st)
        else:  # CogVideoXLayerNormZero / norm_final / AdaLayerNorm: nn.LayerNorm with weight and bias in front
            rc = _fn(lib, "vsb_ln_modulate_affine", x)(_p(x), _p(out), _p(mod), _p(x_mask_u8), _p(gamma), _p(beta), shift_row,
                                                       scale_row, B, T, S, Cc, eps, st)
        _lib.check(rc, "ln_modulate")
    return out


def gate_residual(x, y, mod, x_mask_u8, gate_row, B, T, S, out=None, cache_out=None):
    lib, st = _prep(x, y, mod, x_mask_u8, out, cache_out)
    _same_dtype(x, y, mod, out, cache_out)
    Cc = x.shape[-1]
    out = torch.empty_like(x) if out is None else out
    with _Timed("gate_residual", (3 + (cache_out is not None)) * x.numel() * 2):
        _lib.check(
            _fn(lib, "vsb_gate_residual", x)(_p(x), _p(y), _p(out), _p(cache_out), _p(mod), _p(x_mask_u8), gate_row, B, T, S, Cc,
                                  st),
            "gate_residual",
        )
    return out


def residual_add(x, y, out=None):
    lib, st = _prep(x, y, out)
    _same_dtype(x, y, out)
    out = torch.empty_like(x) if out is None else out
    with _Timed("residual_add", 3 * x.numel() * 2):
        _lib.check(_fn(lib, "vsb_residual_add", x)(_p(x), _p(y), _p(out), x.numel(), st), "residual_add")
    return out


def qk_rmsnorm_(qkv, wq, wk, H, D, eps=1e-6, rope_cos=None, rope_sin=None, pos_div=1, pos_mod=1):
    """In-place per-head RMSNorm of q and k; with rope tables also RoPE at position (row // pos_div) % pos_mod."""
    lib, st = _prep(qkv, wq, wk, rope_cos, rope_sin)
    rows = qkv.numel() // (3 * H * D)
    with _Timed("qk_rmsnorm", 4 * rows * H * D * 2):
        _lib.check(_fn(lib, "vsb_qk_rmsnorm_rope", qkv)(_p(qkv), _p(wq), _p(wk), rows, H, D, eps, _p(rope_cos), _p(rope_sin),
                                           int(pos_div), int(pos_mod), st), "qk_rmsnorm")
    return qkv


def qk_rope_halves_(qkv, rope_cos, rope_sin_signed, H, D, half, pos_div=1, pos_mod=1):
    """In-place half-rotation RoPE of q and k (Open-Sora-Plan RoPE1D / 2D / 3D): tables [pos_mod, D] fp32, the sign of
    rotate_half folded into the sin table; token row r uses table row (r // pos_div) % pos_mod."""
    lib, st = _prep(qkv, rope_cos, rope_sin_signed)
    rows = qkv.numel() // (3 * H * D)
    with _Timed("qk_rmsnorm", 4 * rows * H * D * 2):
        _lib.check(_fn(lib, "vsb_qk_rope_halves", qkv)(_p(qkv), rows, H, D, int(half), _p(rope_cos), _p(rope_sin_signed),
                                                       int(pos_div), int(pos_mod), st), "qk_rope_halves")
    return qkv


def qk_layernorm_(qkv, wq, bq, wk, bk, H, D, eps=1e-6):
    lib, st = _prep(qkv, wq, bq, wk, bk)
    rows = qkv.numel() // (3 * H * D)
    with _Timed("qk_rmsnorm", 4 * rows * H * D * 2):
        _lib.check(_fn(lib, "vsb_qk_layernorm", qkv)(_p(qkv), _p(wq), _p(bq), _p(wk), _p(bk), rows, H, D, eps, st), "qk_layernorm")
    return qkv


def attn_short(qkv, wq, wk, rope_cos, rope_sin, n_outer, n_inner, outer_stride, inner_stride, tok_stride, n, H, D,
               scale, out=None, eps=1e-6, flags: int = 0):
    lib, st = _prep(qkv, wq, wk, rope_cos, rope_sin, out)
    rows = qkv.numel() // (3 * H * D)
    out = torch.empty(rows, H * D, dtype=qkv.dtype, device=qkv.device) if out is None else out
    with _Timed("attn_short", 4 * rows * H * D * 2):
        _lib.check(
            _fn(lib, "vsb_attn_short", qkv)(_p(qkv), _p(out), _p(wq), _p(wk), _p(rope_cos), _p(rope_sin), n_outer, n_inner,
                               outer_stride, inner_stride, tok_stride, n, H, D, eps, scale, flags, st),
            "attn_short",
        )
    return out


def gemm_bias_act(a, w, bias=None, act: int = 0, out=None):
    """out[..., N] = act(a[..., K] @ w[N, K]^T + bias)."""
    lib, st = _prep(a, w, bias, out)
    _bf16(a, "a"), _bf16(w, "w"), _bf16(bias, "bias")
    K = a.shape[-1]
    M = a.numel() // K
    N = w.shape[0]
    if w.shape[1] != K:
        raise _lib.VsbError("gemm: K mismatch")
    _same_dtype(a, w, bias, out)
    out = torch.empty(*a.shape[:-1], N, dtype=a.dtype, device=a.device) if out is None else out
    with _Timed("gemm", 2 * M * N * K):
        _lib.check(_fn(lib, "vsb_gemm_bias_act", a)(_p(a), _p(w), _p(bias), _p(out), M, N, K, act, st), "gemm_bias_act")
    return out


def gemm_bias_residual(a, w, bias, resid, mod=None, x_mask_u8=None, gate_row: int = -1, B=1, T=1, S=1, out=None):
    """out = resid + [gate *] (a @ w^T + bias) with the branch fused into the GEMM epilogue.
    Returns None (nothing launched) when the fused kernel does not take the shape: use the unfused pair then."""
    lib, st = _prep(a, w, bias, resid, mod, x_mask_u8, out)
    K = a.shape[-1]
    M = a.numel() // K
    N = w.shape[0]
    out = resid if out is None else out
    with _Timed("gemm", 2 * M * N * K) as tm:
        rc = _fn(lib, "vsb_gemm_bias_residual", a)(_p(a), _p(w), _p(bias), _p(resid), _p(out), _p(mod), _p(x_mask_u8), gate_row, M,
                                        N, K, B, T, S, st)
        if rc == 1:
            tm.cancel()
    if rc == 1:
        return None
    _lib.check(rc, "gemm_bias_residual")
    return out


def attn_flash(q, k, v, nb, nq, nk, H, D, q_row_stride, q_batch_stride, kv_row_stride, kv_batch_stride, scale,
               kv_lens: Optional[Sequence[int]] = None, out=None, out_row_stride=None, out_batch_stride=None):
    """q/k/v: tensors whose data_ptr() is the first element of the strided view (may be slices of one buffer).
    out_row_stride / out_batch_stride (elements): strided output view starting at out.data_ptr() (default: contiguous
    [nb, nq, H*D])."""
    lib = _lib.load()
    if not q.is_cuda:
        raise _lib.VsbError("vsb200 kernels need CUDA tensors (no CPU fallback)")
    _init_dev(lib, q.device.index)
    st = torch.cuda.current_stream().cuda_stream
    _same_dtype(q, k, v, out)
    out = torch.empty(nb, nq, H * D, dtype=q.dtype, device=q.device) if out is None else out
    lens = None
    if kv_lens is not None:
        lens = (C.c_int * len(kv_lens))(*[int(v_) for v_ in kv_lens])
    with _Timed("attn_flash", 4 * nb * H * nq * nk * D):
        _lib.check(
            _fn(lib, "vsb_attn_flash_strided", q)(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), nb, nq, nk, H, D,
                                       q_row_stride, q_batch_stride, kv_row_stride, kv_batch_stride,
                                       H * D if out_row_stride is None else out_row_stride,
                                       nq * H * D if out_batch_stride is None else out_batch_stride, lens, scale, st),
            "attn_flash",
        )
    return out


def pab_gate(on: bool, timestep: Optional[int], count: int, rng: int, lo: int, hi: int, steps: int):
    lib = _lib.load()
    c = C.c_int(count)
    rc = lib.vsb_pab_gate(int(bool(on)), int(timestep is not None), int(timestep or 0), C.byref(c), int(rng or 1),
                          int(lo), int(hi), int(steps))
    _lib.check(rc, "pab_gate")
    return bool(rc), c.value
