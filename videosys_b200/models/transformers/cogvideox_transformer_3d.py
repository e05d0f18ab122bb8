"""CogVideoX transformer blocks on the vsb200 sm_100a kernels (block level: SURVEY.md section 8 row a15).

``CogVideoXBlockStack`` keeps the reference's state_dict names for ``transformer_blocks.{i}`` (norm1/norm2 =
CogVideoXLayerNormZero: linear + norm; attn1 = diffusers Attention: to_q|to_k|to_v|norm_q|norm_k|to_out.0;
ff.net.0.proj / ff.net.2) and runs CogVideoXBlock.forward (models/transformers/cogvideox_transformer_3d.py:268-312,
attention processor :88-175, sp = 1, no rotary = the 2B model) on the kernels:

  * the two residual streams stay separate ([B, Nv, C] video, [B, Nt, C] text); every LayerNormZero writes its
    modulated output straight into the concatenated [text | video] buffer the joint attention / feed-forward read
    (vsb_ln_modulate_affine on per-sample slices: no torch.cat);
  * to_q / to_k / to_v are one fused GEMM; norm_q / norm_k = vsb_qk_layernorm in place on the packed qkv; the joint
    attention over 226 + 17 550 tokens is vsb_attn_flash (head_dim 64);
  * PAB (spatial gate only, :284-295) caches the un-gated attention output, as the reference does.

Compute dtype is bf16 (the reference runs the 2B model in fp16, pipeline_cogvideox.py:138-139; an fp16 instantiation
of the kernels is future work).  The time/patch embedders and the output head are diffusers classes, not restated.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import kernels
from ...core.pab import pab_mgr


class _LayerNormZero(nn.Module):
    def __init__(self, cond_dim, dim, eps=1e-5):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 6 * dim, bias=True)
        self.norm = nn.LayerNorm(dim, eps=eps, elementwise_affine=True)
        self.eps = eps


class _Attn(nn.Module):
    def __init__(self, dim, head_dim, bias=True):
        super().__init__()
        self.to_q = nn.Linear(dim, dim, bias=bias)
        self.to_k = nn.Linear(dim, dim, bias=bias)
        self.to_v = nn.Linear(dim, dim, bias=bias)
        self.norm_q = nn.LayerNorm(head_dim, eps=1e-6)
        self.norm_k = nn.LayerNorm(head_dim, eps=1e-6)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim, bias=True), nn.Identity()])


class _GELUProj(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner)


class _FF(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, dim * mult), nn.Identity(), nn.Linear(dim * mult, dim), nn.Identity()])


class CogVideoXBlock(nn.Module):
    def __init__(self, dim, heads, head_dim, time_embed_dim, block_idx=0):
        super().__init__()
        self.norm1 = _LayerNormZero(time_embed_dim, dim)
        self.attn1 = _Attn(dim, head_dim, bias=True)
        self.norm2 = _LayerNormZero(time_embed_dim, dim)
        self.ff = _FF(dim)
        self.block_idx = block_idx
        self.attn_count = 0
        self.last_attn = None
        self._qkv = None

    def fused_qkv(self):
        if self._qkv is None:
            a = self.attn1
            self._qkv = (torch.cat([a.to_q.weight, a.to_k.weight, a.to_v.weight], 0).contiguous(),
                         torch.cat([a.to_q.bias, a.to_k.bias, a.to_v.bias], 0).contiguous())
        return self._qkv


class CogVideoXBlockStack(nn.Module):
    def __init__(self, num_attention_heads=30, attention_head_dim=64, num_layers=30, time_embed_dim=512):
        super().__init__()
        self.heads, self.head_dim = num_attention_heads, attention_head_dim
        dim = num_attention_heads * attention_head_dim
        self.dim = dim
        self.transformer_blocks = nn.ModuleList(
            [CogVideoXBlock(dim, num_attention_heads, attention_head_dim, time_embed_dim, i) for i in range(num_layers)]
        )

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        for b in self.transformer_blocks:
            b._qkv = None
        return r

    def reset_pab_state(self):
        for b in self.transformer_blocks:
            b.attn_count, b.last_attn = 0, None

    def _norm_zero(self, nz: _LayerNormZero, hid, enc, temb, ncat, B, Nv, Nt, C):
        """CogVideoXLayerNormZero for both streams, written into ncat[b] = [text | video]; returns mod [1,B,6,C]
        (rows: shift, scale, gate, enc_shift, enc_scale, enc_gate)."""
        K = kernels
        mod = F.linear(F.silu(temb), nz.linear.weight, nz.linear.bias).view(1, B, 6, C).contiguous()  # tiny (M = B)
        for b in range(B):
            mb = mod[:, b : b + 1]
            K.ln_modulate(enc[b], mb, None, 3, 4, 1, 1, Nt, out=ncat[b, :Nt], eps=nz.eps, gamma=nz.norm.weight, beta=nz.norm.bias)
            K.ln_modulate(hid[b], mb, None, 0, 1, 1, 1, Nv, out=ncat[b, Nt:], eps=nz.eps, gamma=nz.norm.weight, beta=nz.norm.bias)
        return mod

    @torch.no_grad()
    def forward(self, hidden: torch.Tensor, enc: torch.Tensor, temb: torch.Tensor, timestep=None):
        """hidden [B, Nv, C], enc [B, Nt, C] (bf16, CUDA), temb [B, time_embed_dim]; returns the two streams."""
        if not hidden.is_cuda or hidden.dtype != torch.bfloat16:
            raise RuntimeError("videosys_b200 CogVideoX blocks run on sm_100a CUDA devices in bf16 only (no CPU path)")
        K = kernels
        B, Nv, C = hidden.shape
        Nt = enc.shape[1]
        N = Nt + Nv
        H, D = self.heads, self.head_dim
        hid = hidden.contiguous().clone()
        en = enc.contiguous().clone()
        ncat = torch.empty(B, N, C, dtype=hid.dtype, device=hid.device)
        pab_on = pab_mgr.enable_pab()
        ts_int = int(timestep[0]) if (pab_on and timestep is not None) else None
        for blk in self.transformer_blocks:
            mod = self._norm_zero(blk.norm1, hid, en, temb, ncat, B, Nv, Nt, C)
            reuse = False
            if pab_on:
                reuse, blk.attn_count = pab_mgr.if_broadcast_spatial(ts_int, blk.attn_count)
            if reuse:
                a = blk.last_attn
            else:
                w, bias = blk.fused_qkv()
                qkv = K.gemm_bias_act(ncat.view(B * N, C), w, bias)
                at = blk.attn1
                K.qk_layernorm_(qkv, at.norm_q.weight, at.norm_q.bias, at.norm_k.weight, at.norm_k.bias, H, D, eps=1e-6)
                q3 = qkv.view(-1, 3, C)
                o = K.attn_flash(q3[:, 0], q3[:, 1], q3[:, 2], B, N, N, H, D, 3 * C, N * 3 * C, 3 * C, N * 3 * C, D**-0.5)
                a = K.gemm_bias_act(o, at.to_out[0].weight, at.to_out[0].bias)  # [B, N, C] = [text | video]
                if pab_on:
                    blk.last_attn = a
            for b in range(B):
                mb = mod[:, b : b + 1]
                K.gate_residual(hid[b], a[b, Nt:], mb, None, 2, 1, 1, Nv, out=hid[b])
                K.gate_residual(en[b], a[b, :Nt], mb, None, 5, 1, 1, Nt, out=en[b])
            mod = self._norm_zero(blk.norm2, hid, en, temb, ncat, B, Nv, Nt, C)
            h = K.gemm_bias_act(ncat.view(B * N, C), blk.ff.net[0].proj.weight, blk.ff.net[0].proj.bias, act=1)
            f = K.gemm_bias_act(h, blk.ff.net[2].weight, blk.ff.net[2].bias).view(B, N, C)
            for b in range(B):
                mb = mod[:, b : b + 1]
                K.gate_residual(hid[b], f[b, Nt:], mb, None, 2, 1, 1, Nv, out=hid[b])
                K.gate_residual(en[b], f[b, :Nt], mb, None, 5, 1, 1, Nt, out=en[b])
        return hid, en
