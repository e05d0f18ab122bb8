"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of the Latte transformer blocks (block level).

PARITY UNPINNED: Latte's blocks lean on ``diffusers==0.30.0`` (``Attention`` + ``AttnProcessor2_0``, ``GELU``;
requirements.txt:25), which is not installed in this image and not vendored in /root/reference, so the reference
modules cannot be executed here.  This file restates the in-tree block code
(models/transformers/latte_transformer_3d.py:357-517 spatial ``BasicTransformerBlock``, :680-824 temporal
``BasicTransformerBlock_``, local ``FeedForward`` :92-148, block loop :1312-1425) plus the published semantics of the
two diffusers classes as the reference uses them:
  * ``Attention(query_dim, heads, dim_head, bias=True, cross_attention_dim)`` with the default processor:
    to_q / to_k / to_v Linears, heads split, F.scaled_dot_product_attention (no mask: the Latte pipeline never passes
    one, pipeline_latte.py:854-862), to_out[0] Linear, dropout 0, rescale_output_factor 1;
  * ``GELU(dim, inner, approximate="tanh")`` = Linear then tanh-GELU.
Configuration: norm_type "ada_norm_single", norm_elementwise_affine False, eps 1e-6, activation "gelu-approximate".
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _ln(x: Tensor, eps: float = 1e-6) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)


def attention(sd: Dict[str, Tensor], p: str, x: Tensor, enc: Optional[Tensor], heads: int) -> Tensor:
    """diffusers Attention.forward -> AttnProcessor2_0.__call__ (self-attention when enc is None)."""
    B, N, C = x.shape
    D = C // heads
    src = x if enc is None else enc
    q = F.linear(x, sd[p + "to_q.weight"], sd.get(p + "to_q.bias"))
    k = F.linear(src, sd[p + "to_k.weight"], sd.get(p + "to_k.bias"))
    v = F.linear(src, sd[p + "to_v.weight"], sd.get(p + "to_v.bias"))
    q = q.view(B, -1, heads, D).transpose(1, 2)
    k = k.view(B, -1, heads, D).transpose(1, 2)
    v = v.view(B, -1, heads, D).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, -1, C).to(q.dtype)
    return F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def feed_forward(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """FeedForward(activation_fn='gelu-approximate'): latte_transformer_3d.py:92-148."""
    h = F.gelu(F.linear(x, sd[p + "net.0.proj.weight"], sd[p + "net.0.proj.bias"]), approximate="tanh")
    return F.linear(h, sd[p + "net.2.weight"], sd[p + "net.2.bias"])


def spatial_block(sd, p: str, x: Tensor, enc: Tensor, timestep6: Tensor, heads: int) -> Tensor:
    """BasicTransformerBlock.forward, ada_norm_single, PAB off: latte_transformer_3d.py:357-517.
    x [(b f), S, C]; enc [(b f), L, C]; timestep6 [(b f), 6C]."""
    bs = x.shape[0]
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = (sd[p + "scale_shift_table"][None] + timestep6.reshape(bs, 6, -1)).chunk(6, dim=1)
    n = _ln(x) * (1 + sc_a) + sh_a
    x = g_a * attention(sd, p + "attn1.", n, None, heads) + x
    x = attention(sd, p + "attn2.", x, enc, heads) + x  # ada_norm_single: no norm before cross attention (:447-450)
    n = _ln(x) * (1 + sc_m) + sh_m  # norm2
    return g_m * feed_forward(sd, p + "ff.", n) + x


def temporal_block(sd, p: str, x: Tensor, timestep6: Tensor, heads: int) -> Tensor:
    """BasicTransformerBlock_.forward, ada_norm_single, sp=1, PAB off: latte_transformer_3d.py:680-824.
    x [(b s), F, C]; timestep6 [(b s), 6C]."""
    bs = x.shape[0]
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = (sd[p + "scale_shift_table"][None] + timestep6.reshape(bs, 6, -1)).chunk(6, dim=1)
    n = _ln(x) * (1 + sc_a) + sh_a
    x = g_a * attention(sd, p + "attn1.", n, None, heads) + x
    n = _ln(x) * (1 + sc_m) + sh_m  # norm3 (:778)
    return g_m * feed_forward(sd, p + "ff.", n) + x


def block_stack(sd, x: Tensor, enc: Tensor, timestep6: Tensor, heads: int, depth: int, temp_pos_embed: Optional[Tensor]):
    """The LatteT2V block loop at inference (sp=1, use_image_num=0): latte_transformer_3d.py:1290-1425.
    x [B, F, S, C] -> [B, F, S, C]; enc [B, L, C]; timestep6 [B, 6C]; temp_pos_embed [1, F, C] or None."""
    B, Fr, S, C = x.shape
    enc_sp = enc[:, None].expand(B, Fr, *enc.shape[1:]).reshape(B * Fr, *enc.shape[1:]).contiguous()
    t_sp = timestep6[:, None].expand(B, Fr, -1).reshape(B * Fr, -1).contiguous()
    t_tm = timestep6[:, None].expand(B, S, -1).reshape(B * S, -1).contiguous()
    h = x.reshape(B * Fr, S, C)
    for i in range(depth):
        h = spatial_block(sd, f"transformer_blocks.{i}.", h, enc_sp, t_sp, heads)
        h = h.reshape(B, Fr, S, C).permute(0, 2, 1, 3).reshape(B * S, Fr, C).contiguous()  # (b f) t d -> (b t) f d
        if i == 0 and Fr > 1 and temp_pos_embed is not None:
            h = h + temp_pos_embed
        h = temporal_block(sd, f"temporal_transformer_blocks.{i}.", h, t_tm, heads)
        h = h.reshape(B, S, Fr, C).permute(0, 2, 1, 3).reshape(B * Fr, S, C).contiguous()
    return h.reshape(B, Fr, S, C)
