"""Latte transformer blocks on the vsb200 sm_100a kernels (block level: SURVEY.md section 8 row a16).

``LatteBlockStack`` holds the parameters of LatteT2V's ``transformer_blocks`` / ``temporal_transformer_blocks`` under
the reference's state_dict names (diffusers ``Attention``: attn1/attn2.to_q|to_k|to_v|to_out.0, ``FeedForward``:
ff.net.0.proj / ff.net.2, ``scale_shift_table``; norms carry no parameters) and runs the block loop of
models/transformers/latte_transformer_3d.py:1312-1425 (inference, use_image_num = 0) on the same kernels as STDiT3:

  * activations stay token-major [B, F, S, C]; the reference's two rearranges per block pair are strides;
  * to_q / to_k / to_v are fused into one [3C, C] GEMM (weights concatenated once, lazily);
  * spatial self-attention = vsb_attn_flash on the packed qkv; temporal (F < 30) = vsb_attn_short with flags 3 (no q/k
    norm, SDPA rounding); cross attention = one kv GEMM per sample (the reference repeats the text per frame and
    recomputes it F times: latte_transformer_3d.py:1290-1296) + vsb_attn_flash.

The embedders around the blocks (PatchEmbed, PixArtAlphaCombinedTimestepSizeEmbeddings, caption projection, output
head) are diffusers classes and are not restated this round.
"""
from typing import Optional

import torch
import torch.nn as nn

from ... import kernels


class _Attn(nn.Module):
    def __init__(self, dim, cross_dim=None):
        super().__init__()
        kv = cross_dim or dim
        self.to_q = nn.Linear(dim, dim, bias=True)
        self.to_k = nn.Linear(kv, dim, bias=True)
        self.to_v = nn.Linear(kv, dim, bias=True)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim, bias=True), nn.Identity()])


class _GELUProj(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner)


class _FF(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, dim * mult), nn.Identity(), nn.Linear(dim * mult, dim)])


class LatteBlock(nn.Module):
    def __init__(self, dim, heads, temporal):
        super().__init__()
        self.temporal = temporal
        self.attn1 = _Attn(dim)
        if not temporal:
            self.attn2 = _Attn(dim, dim)
        self.ff = _FF(dim)
        self.scale_shift_table = nn.Parameter(torch.randn(6, dim) / dim**0.5)
        self._fused = {}

    def fused(self, which: str):
        """Concatenated projection weights (built once; call ``invalidate()`` after loading new weights)."""
        if which not in self._fused:
            a = self.attn1 if which == "qkv" else self.attn2
            mods = (a.to_q, a.to_k, a.to_v) if which == "qkv" else (a.to_k, a.to_v)
            self._fused[which] = (torch.cat([m.weight for m in mods], 0).contiguous(), torch.cat([m.bias for m in mods], 0).contiguous())
        return self._fused[which]

    def invalidate(self):
        self._fused = {}


class LatteBlockStack(nn.Module):
    def __init__(self, hidden_size=1152, num_heads=16, depth=28):
        super().__init__()
        self.hidden_size, self.num_heads, self.depth = hidden_size, num_heads, depth
        self.transformer_blocks = nn.ModuleList([LatteBlock(hidden_size, num_heads, False) for _ in range(depth)])
        self.temporal_transformer_blocks = nn.ModuleList([LatteBlock(hidden_size, num_heads, True) for _ in range(depth)])

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        for b in [*self.transformer_blocks, *self.temporal_transformer_blocks]:
            b.invalidate()
        return r

    @torch.no_grad()
    def forward(self, x: torch.Tensor, enc: torch.Tensor, timestep6: torch.Tensor, temp_pos_embed: Optional[torch.Tensor] = None):
        """x [B, F, S, C] bf16 (CUDA), enc [B, L, C], timestep6 [B, 6C]; returns [B, F, S, C]."""
        if not x.is_cuda or x.dtype != torch.bfloat16:
            raise RuntimeError("videosys_b200 Latte blocks run on sm_100a CUDA devices in bf16 only (no CPU path)")
        K = kernels
        B, Fr, S, C = x.shape
        H = self.num_heads
        D = C // H
        L = enc.shape[1]
        x = x.contiguous().clone()
        enc2 = enc.reshape(B * L, C).contiguous()
        xf = x.view(B, Fr * S, C)
        for i in range(self.depth):
            for blk in (self.transformer_blocks[i], self.temporal_transformer_blocks[i]):
                if blk.temporal and i == 0 and Fr > 1 and temp_pos_embed is not None:
                    x.add_(temp_pos_embed.reshape(1, Fr, 1, C).to(x.dtype))  # glue op (once per forward)
                mod = K.modulation_table(blk.scale_shift_table, timestep6, None)
                xm = K.ln_modulate(xf, mod, None, 0, 1, B, Fr, S)
                wqkv, bqkv = blk.fused("qkv")
                qkv = K.gemm_bias_act(xm, wqkv, bqkv)
                if blk.temporal:
                    if Fr < 30:
                        o = K.attn_short(qkv.view(-1, 3, H, D), None, None, None, None, B, S, Fr * S, 1, S, Fr, H, D, D**-0.5, flags=3)
                    else:
                        raise RuntimeError("temporal sequences >= 30 frames are not supported by vsb_attn_short")
                else:
                    q3 = qkv.view(-1, 3, C)
                    if S >= 30:
                        o = K.attn_flash(q3[:, 0], q3[:, 1], q3[:, 2], B * Fr, S, S, H, D, 3 * C, S * 3 * C, 3 * C, S * 3 * C, D**-0.5)
                    else:
                        o = K.attn_short(qkv.view(-1, 3, H, D), None, None, None, None, B * Fr, 1, S, 0, 1, S, H, D, D**-0.5, flags=3)
                y = K.gemm_bias_act(o.view(-1, C), blk.attn1.to_out[0].weight, blk.attn1.to_out[0].bias)
                K.gate_residual(xf, y.view(B, Fr * S, C), mod, None, 2, B, Fr, S, out=xf)
                if not blk.temporal:
                    a2 = blk.attn2
                    q = K.gemm_bias_act(xf, a2.to_q.weight, a2.to_q.bias)
                    wkv, bkv = blk.fused("kv")
                    kv = K.gemm_bias_act(enc2, wkv, bkv).view(-1, 2, C)
                    o = K.attn_flash(q, kv[:, 0], kv[:, 1], B, Fr * S, L, H, D, C, Fr * S * C, 2 * C, L * 2 * C, D**-0.5)
                    xc = K.gemm_bias_act(o, a2.to_out[0].weight, a2.to_out[0].bias)
                    K.residual_add(xf, xc.view(B, Fr * S, C), out=xf)
                xm = K.ln_modulate(xf, mod, None, 3, 4, B, Fr, S)
                h = K.gemm_bias_act(xm, blk.ff.net[0].proj.weight, blk.ff.net[0].proj.bias, act=1)
                y = K.gemm_bias_act(h, blk.ff.net[2].weight, blk.ff.net[2].bias)
                K.gate_residual(xf, y, mod, None, 5, B, Fr, S, out=xf)
        return x
