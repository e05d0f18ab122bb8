"""Shared test helpers (no reference, no GPU needed)."""
import torch


def stdit3_state_dict_template(cfg: dict, dtype=torch.float32) -> dict:
    """Key set / shapes of the reference STDiT3.state_dict(), zero-filled."""
    C = cfg["hidden_size"]
    H = cfg["num_heads"]
    depth = cfg["depth"]
    cap = cfg.get("caption_channels", 4096)
    L = cfg.get("model_max_length", 300)
    ps = cfg.get("patch_size", (1, 2, 2))
    inc = cfg.get("in_channels", 4)
    outc = inc * 2
    z = lambda *s: torch.zeros(*s, dtype=dtype)  # noqa: E731
    from oracle.stdit3_oracle import rope_freqs

    sd = {"rope.freqs": rope_freqs(C // H).to(dtype)}
    sd["x_embedder.proj.weight"] = z(C, inc, *ps)
    sd["x_embedder.proj.bias"] = z(C)
    for e in ("t_embedder", "fps_embedder"):
        sd[f"{e}.mlp.0.weight"] = z(C, 256)
        sd[f"{e}.mlp.0.bias"] = z(C)
        sd[f"{e}.mlp.2.weight"] = z(C, C)
        sd[f"{e}.mlp.2.bias"] = z(C)
    sd["t_block.1.weight"] = z(6 * C, C)
    sd["t_block.1.bias"] = z(6 * C)
    sd["y_embedder.y_embedding"] = z(L, cap)
    sd["y_embedder.y_proj.fc1.weight"] = z(C, cap)
    sd["y_embedder.y_proj.fc1.bias"] = z(C)
    sd["y_embedder.y_proj.fc2.weight"] = z(C, C)
    sd["y_embedder.y_proj.fc2.bias"] = z(C)
    for kind in ("spatial", "temporal"):
        for i in range(depth):
            p = f"{kind}_blocks.{i}."
            sd[p + "scale_shift_table"] = z(6, C)
            sd[p + "attn.qkv.weight"] = z(3 * C, C)
            sd[p + "attn.qkv.bias"] = z(3 * C)
            sd[p + "attn.q_norm.weight"] = z(C // H)
            sd[p + "attn.k_norm.weight"] = z(C // H)
            sd[p + "attn.proj.weight"] = z(C, C)
            sd[p + "attn.proj.bias"] = z(C)
            sd[p + "cross_attn.q_linear.weight"] = z(C, C)
            sd[p + "cross_attn.q_linear.bias"] = z(C)
            sd[p + "cross_attn.kv_linear.weight"] = z(2 * C, C)
            sd[p + "cross_attn.kv_linear.bias"] = z(2 * C)
            sd[p + "cross_attn.proj.weight"] = z(C, C)
            sd[p + "cross_attn.proj.bias"] = z(C)
            sd[p + "mlp.fc1.weight"] = z(int(C * 4), C)
            sd[p + "mlp.fc1.bias"] = z(int(C * 4))
            sd[p + "mlp.fc2.weight"] = z(C, int(C * 4))
            sd[p + "mlp.fc2.bias"] = z(C)
    sd["final_layer.scale_shift_table"] = z(2, C)
    sd["final_layer.linear.weight"] = z(ps[0] * ps[1] * ps[2] * outc, C)
    sd["final_layer.linear.bias"] = z(ps[0] * ps[1] * ps[2] * outc)
    return sd
