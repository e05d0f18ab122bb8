"""TEST INFRASTRUCTURE ONLY -- small Latte / CogVideoX / Vchitect cases shared by oracle/gen_golden_models.py (runs the
UNMODIFIED reference on them, authoring container) and tests/test_oracle_golden.py (checks the oracles against the stored
reference outputs anywhere)."""
import torch

from . import synth

LATTE = dict(num_attention_heads=2, attention_head_dim=72, in_channels=4, out_channels=8, num_layers=2, cross_attention_dim=144,
             attention_bias=True, sample_size=8, patch_size=2, activation_fn="gelu-approximate", norm_type="ada_norm_single",
             norm_elementwise_affine=False, norm_eps=1e-6, caption_channels=32, video_length=6)
LATTE_O = dict(heads=2, head_dim=72, layers=2, patch=2, sample_size=8, out_channels=8, video_length=6)
COGX = dict(num_attention_heads=4, attention_head_dim=64, in_channels=4, out_channels=4, time_embed_dim=64, text_embed_dim=48,
            num_layers=2, sample_width=16, sample_height=12, sample_frames=9, max_text_seq_length=16)
COGX_O = dict(heads=4, head_dim=64, layers=2, patch=2, max_text=16, sample_width=16, sample_height=12, sample_frames=9, out_channels=4)
VCH = dict(sample_size=8, patch_size=2, in_channels=4, num_layers=3, attention_head_dim=64, num_attention_heads=2,
           joint_attention_dim=48, caption_projection_dim=128, pooled_projection_dim=40, out_channels=4, pos_embed_max_size=12)
VCH_O = dict(heads=2, head_dim=64, layers=3, patch=2, sample_size=8, pos_embed_max_size=12, out_channels=4)


def weights(state_dict, tag, dtype, norm_ones=False, keep=()):
    """Deterministic weights for a reference state dict (by parameter name); ``keep``: entries that are tables, not weights;
    norm_ones: LayerNorm weights around 1; Vchitect-only members of the vendored attention class are filled like the rest."""
    sd0 = {k: v.float() for k, v in state_dict.items()}
    sd = synth.fill_state_dict(sd0, f"mc.{tag}.")
    if norm_ones:
        for k in sd:
            if k.endswith("norm.weight") or k.endswith("norm_final.weight") or k.endswith("norm_q.weight") or k.endswith("norm_k.weight"):
                sd[k] = 1.0 + 0.2 * synth.uniform(f"mc.{tag}." + k, tuple(sd[k].shape))
    for k in keep:
        sd[k] = sd0[k]
    return {k: v.to(dtype) for k, v in sd.items()}


def latte_inputs(dtype):
    return (synth.normalish("mc.latte.x", (2, 4, 6, 8, 8)).to(dtype), torch.tensor([500, 500]),
            synth.normalish("mc.latte.enc", (2, 7, 32)).to(dtype))


def cogx_inputs(dtype):
    return (synth.normalish("mc.cogx.lat", (2, 3, 4, 12, 16)).to(dtype), synth.normalish("mc.cogx.txt", (2, 16, 48)).to(dtype),
            torch.tensor([499, 499]))


def cogx_rotary():
    from . import cogvideox_oracle as CO

    return CO.rotary_3d(64, CO.resize_crop_region_for_grid((6, 8), 45, 30), (6, 8), 3)


def vch_inputs(dtype):
    return (synth.normalish("mc.vch.lat", (1, 5, 4, 12, 16)).to(dtype), synth.normalish("mc.vch.enc", (1, 9, 48)).to(dtype),
            synth.normalish("mc.vch.pool", (1, 40)).to(dtype), torch.tensor([500.0]))
