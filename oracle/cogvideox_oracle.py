"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of the CogVideoX transformer block (block level).

PARITY PARTLY PINNED: ``CogVideoXLayerNormZero`` (models/modules/normalization.py:36-57) is in the reference tree and is
executed as-is by tests/test_oracle_vs_reference.py::test_cogvideox_layernorm_zero.  The attention processor
(models/transformers/cogvideox_transformer_3d.py:88-175) is in-tree but drives a ``diffusers==0.30.0`` ``Attention``
object (requirements.txt:25, not installed here), and the feed-forward is diffusers' ``FeedForward``; both are restated
from their published semantics as the reference configures them (cogvideox_transformer_3d.py:237-261):
  * ``Attention(query_dim, heads, dim_head, qk_norm="layer_norm", eps=1e-6, bias=attention_bias, out_bias=True)``:
    to_q/to_k/to_v Linears, ``norm_q``/``norm_k`` = nn.LayerNorm(dim_head, eps=1e-6) (affine), to_out[0] Linear;
  * ``FeedForward(dim, activation_fn="gelu-approximate", final_dropout=True, bias=True)``: Linear -> tanh-GELU -> Linear.
Those two are PARITY UNPINNED.  2B configuration: no rotary embedding (use_rotary_positional_embeddings False), sp = 1.
"""
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def layer_norm_zero(sd: Dict[str, Tensor], p: str, hidden: Tensor, enc: Tensor, temb: Tensor, eps: float = 1e-5):
    """CogVideoXLayerNormZero.forward: models/modules/normalization.py:51-57."""
    C = hidden.shape[-1]
    shift, scale, gate, e_shift, e_scale, e_gate = F.linear(F.silu(temb), sd[p + "linear.weight"], sd[p + "linear.bias"]).chunk(6, dim=1)
    w, b = sd[p + "norm.weight"], sd[p + "norm.bias"]
    hs = F.layer_norm(hidden, (C,), w, b, eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    es = F.layer_norm(enc, (C,), w, b, eps) * (1 + e_scale)[:, None, :] + e_shift[:, None, :]
    return hs, es, gate[:, None, :], e_gate[:, None, :]


def joint_attention(sd, p: str, hidden: Tensor, enc: Tensor, heads: int):
    """CogVideoXAttnProcessor2_0.__call__ with sp_size == 1, no mask, no rotary: cogvideox_transformer_3d.py:88-175."""
    text_len = enc.size(1)
    x = torch.cat([enc, hidden], dim=1)
    B, N, C = x.shape
    D = C // heads
    q = F.linear(x, sd[p + "to_q.weight"], sd.get(p + "to_q.bias"))
    k = F.linear(x, sd[p + "to_k.weight"], sd.get(p + "to_k.bias"))
    v = F.linear(x, sd[p + "to_v.weight"], sd.get(p + "to_v.bias"))
    q = q.view(B, -1, heads, D).transpose(1, 2)
    k = k.view(B, -1, heads, D).transpose(1, 2)
    v = v.view(B, -1, heads, D).transpose(1, 2)
    q = F.layer_norm(q, (D,), sd[p + "norm_q.weight"], sd[p + "norm_q.bias"], 1e-6)
    k = F.layer_norm(k, (D,), sd[p + "norm_k.weight"], sd[p + "norm_k.bias"], 1e-6)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, -1, C)
    o = F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])
    e, h = o.split([text_len, o.size(1) - text_len], dim=1)
    return h, e


def feed_forward(sd, p: str, x: Tensor) -> Tensor:
    h = F.gelu(F.linear(x, sd[p + "net.0.proj.weight"], sd[p + "net.0.proj.bias"]), approximate="tanh")
    return F.linear(h, sd[p + "net.2.weight"], sd[p + "net.2.bias"])


class BlockPAB:
    def __init__(self):
        self.attn_count = 0
        self.last_attn = None


def block(sd, p: str, hidden: Tensor, enc: Tensor, temb: Tensor, heads: int, pab=None, pab_state=None, timestep_int=None):
    """CogVideoXBlock.forward: cogvideox_transformer_3d.py:268-312 (PAB: spatial gate only, :284-295)."""
    text_len = enc.size(1)
    nh, ne, gate, e_gate = layer_norm_zero(sd, p + "norm1.", hidden, enc, temb)
    reuse = False
    if pab is not None and pab.enabled():
        reuse, pab_state.attn_count = pab.gate("spatial", timestep_int, pab_state.attn_count)
    if reuse:
        ah, ae = pab_state.last_attn
    else:
        ah, ae = joint_attention(sd, p + "attn1.", nh, ne, heads)
        if pab is not None and pab.enabled():
            pab_state.last_attn = (ah, ae)
    hidden = hidden + gate * ah
    enc = enc + e_gate * ae
    nh, ne, gate_ff, e_gate_ff = layer_norm_zero(sd, p + "norm2.", hidden, enc, temb)
    ff = feed_forward(sd, p + "ff.", torch.cat([ne, nh], dim=1))
    hidden = hidden + gate_ff * ff[:, text_len:]
    enc = enc + e_gate_ff * ff[:, :text_len]
    return hidden, enc
