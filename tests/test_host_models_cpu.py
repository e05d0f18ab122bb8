"""Host logic of the model mirrors that widen the path on the CPU: every kernel entry is replaced
by its torch stand-in (tests/kernels_emul.py, test infrastructure), so what is checked here is what the front end does
around the kernels -- layouts and strides handed to the attention entries, weight fusion, PAB bookkeeping, un-patchify --
against the oracle in fp32 (tight tolerance: same arithmetic, different op order).  The GPU parity of the same models is
in tests/test_vchitect_gpu.py / tests/test_osp_gpu.py."""
import pytest
import torch

from oracle import pab_oracle, synth
from tests import kernels_emul

VCH = dict(sample_size=8, patch_size=2, in_channels=4, num_layers=3, attention_head_dim=64, num_attention_heads=2,
           joint_attention_dim=48, caption_projection_dim=128, pooled_projection_dim=40, out_channels=4, pos_embed_max_size=12)
VCH_O = dict(heads=2, head_dim=64, layers=3, patch=2, sample_size=8, pos_embed_max_size=12, out_channels=4)


def _vchitect(tag="vch."):
    from videosys_b200.models.transformers.vchitect_transformer_3d import VchitectXLTransformerModel

    net = VchitectXLTransformerModel(**VCH)
    sd = synth.fill_state_dict(net.state_dict(), tag)
    sd["pos_embed.pos_embed"] = net.state_dict()["pos_embed.pos_embed"]  # the sin-cos table is not a weight
    net.load_state_dict(sd)
    return net.eval(), sd


@pytest.mark.parametrize("Fr", [5, 1, 34])  # 34 frames: RoPE pre-pass + flash attention on strided views
def test_vchitect_forward_host_logic(monkeypatch, Fr):
    from oracle import vchitect_oracle as VO

    kernels_emul.emulate(monkeypatch)
    net, sd = _vchitect()
    lat = synth.normalish("vch.lat", (1, Fr, 4, 12, 16))
    enc = synth.normalish("vch.enc", (1, 9, 48))
    pooled = synth.normalish("vch.pool", (1, 40))
    ts = torch.tensor([500])
    out = net(lat, enc, pooled, ts, return_dict=False)[0]
    with torch.no_grad():
        ref = VO.transformer_forward(sd, VCH_O, lat, enc, pooled, ts)
    assert out.shape == ref.shape == (Fr, 4, 12, 16)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5), (out - ref).abs().max()


def test_vchitect_pab_host_logic(monkeypatch):
    """Eight steps with the three gates on: the mirror's counters / caches follow the reference's (pinned per attention
    in test_oracle_vs_reference.py::test_vchitect_attention_pab_vs_reference)."""
    from oracle import vchitect_oracle as VO
    from videosys_b200.core.pab import pab_mgr

    kernels_emul.emulate(monkeypatch)
    net, sd = _vchitect()
    cfg = pab_mgr.PABConfig(spatial_broadcast=True, spatial_threshold=[100, 800], spatial_range=2, temporal_broadcast=True,
                            temporal_threshold=[100, 800], temporal_range=3, cross_broadcast=True, cross_threshold=[100, 800],
                            cross_range=4)
    pab_mgr.set_pab_manager(cfg)
    pab_mgr.update_steps(8)
    net.reset_pab_state()
    try:
        G = pab_oracle.PABGate((True, (100, 800), 2), (True, (100, 800), 3), (True, (100, 800), 4), 8)
        L = VCH["num_layers"]
        counts = [{"spatial": 0, "temporal": 0, "cross": 0} for _ in range(L)]
        caches = [{} for _ in range(L)]
        enc = synth.normalish("vchp.enc", (1, 9, 48))
        pooled = synth.normalish("vchp.pool", (1, 40))
        hits = 0
        for step, t in enumerate([900, 700, 650, 600, 550, 500, 450, 50]):
            lat = synth.normalish(f"vchp.lat{step}", (1, 4, 4, 12, 16))

            def gate(i, kind, t=t):
                nonlocal hits
                hit, counts[i][kind] = G.gate(kind, t, counts[i][kind])
                hits += hit
                return hit

            ts = torch.tensor([t])
            out = net(lat, enc, pooled, ts, return_dict=False)[0]
            with torch.no_grad():
                ref = VO.transformer_forward(sd, VCH_O, lat, enc, pooled, ts, gate, caches)
            assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5), (step, (out - ref).abs().max())
        assert hits > 0
    finally:
        pab_mgr.set_pab_manager(None)
