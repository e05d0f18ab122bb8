"""The oracle (oracle/*.py) against the committed golden vectors produced by the unmodified reference
(oracle/gen_golden.py).  CPU only."""
import json
import os

import pytest
import torch

from oracle import cases, pab_oracle, stdit3_oracle as O, synth


def _state_dict(dtype, depth=1):
    """Same key set/shapes as the reference STDiT3, filled by synth."""
    from tests.helpers import stdit3_state_dict_template

    tmpl = stdit3_state_dict_template(cases.small_model_cfg(depth), dtype)
    return synth.fill_state_dict(tmpl, "golden.")


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "stdit3_small.pt"))


@pytest.mark.parametrize("dn,dtype,tol", [("fp32", torch.float32, 2e-5), ("bf16", torch.bfloat16, 0.0)])
def test_forward_matches_reference_golden(gold, dn, dtype, tol):
    torch.set_num_threads(1)
    sd = _state_dict(dtype)
    inp = cases.forward_inputs(dtype)
    with torch.no_grad():
        out = O.stdit3_forward(sd, cases.oracle_cfg(cases.small_model_cfg()), **inp)
    ref = gold[f"forward_{dn}"].float()
    if dtype == torch.bfloat16:
        # bf16 GEMM reduction order may differ across CPU ISAs: allow a few bf16 ulps on a few elements
        close = torch.isclose(out, ref, rtol=2e-2, atol=2e-2)
        assert close.float().mean().item() > 0.999, (out - ref).abs().max()
    else:
        assert torch.allclose(out, ref, rtol=1e-4, atol=tol), (out - ref).abs().max()


@pytest.mark.parametrize("kind", ["spatial", "temporal"])
def test_block_matches_reference_golden(gold, kind):
    torch.set_num_threads(1)
    dtype = torch.bfloat16
    sd = _state_dict(dtype)
    b = cases.block_inputs(dtype)
    temporal = kind == "temporal"
    with torch.no_grad():
        out = O.stdit3_block(sd, f"{kind}_blocks.0.", b["x"], b["y"], b["t"], b["y_lens"], b["x_mask"], b["t0"],
                             b["T"], b["S"], 4, temporal, sd["rope.freqs"] if temporal else None)
    ref = gold[f"block_{kind}_bf16"].float()
    close = torch.isclose(out.float(), ref, rtol=2e-2, atol=2e-2)
    assert close.float().mean().item() > 0.999
    # on the same ISA (the authoring container) this is bit-exact; report it
    print(kind, "bit-equal fraction", (out.float() == ref).float().mean().item())


def test_pab_schedule_known_answers(golden_dir):
    """skip bitmaps and the int(bf16(t)) timestep lists, bit-exact."""
    kat = json.load(open(os.path.join(golden_dir, "pab_schedules.json")))
    for name, (hh, ww, nf, steps) in {"240p_51f_30": (240, 426, 51, 30), "720p_68f_50": (720, 1280, 68, 50)}.items():
        _, ints = O.rflow_timesteps(steps, hh, ww, nf)
        assert ints == kat[name]["timesteps"]
        gate = pab_oracle.opensora_default(steps)
        for kind in ("spatial", "temporal", "cross"):
            assert gate.schedule(kind, ints) == kat[name][kind]
            # second generate(): the counter has wrapped to 0 again (pab_mgr.py:64)
            assert gate.schedule(kind, ints) == kat[name][kind + "_second_run"]
    assert kat["720p_68f_50"]["spatial"].count("1") == 13
    assert kat["240p_51f_30"]["cross"].count("1") == 15


def test_pab_gate_edges():
    g = pab_oracle.PABGate(spatial=(True, (450, 930), 2), steps=4)
    assert g.gate("spatial", 450, 1) == (False, 2)  # strict lower bound
    assert g.gate("spatial", 930, 1) == (False, 2)  # strict upper bound
    assert g.gate("spatial", 451, 1) == (True, 2)
    assert g.gate("spatial", 451, 2) == (False, 3)  # count % range == 0
    assert g.gate("spatial", 451, 3) == (True, 0)  # wraps modulo steps
    assert g.gate("spatial", None, 1) == (False, 2)
    assert g.gate("cross", 500, 1) == (False, 2)  # kind off, manager on: counter still advances
    off = pab_oracle.PABGate(steps=4)
    assert off.gate("spatial", 500, 3) == (False, 3)  # disabled wrapper leaves the count untouched


# ---- Latte / CogVideoX / Vchitect oracles against outputs of the UNMODIFIED reference (tests/golden/models_small.pt, written by
#      oracle/gen_golden_models.py); the reference tree is not needed here ------------------------------------------------------------
@pytest.fixture(scope="module")
def model_gold(golden_dir):
    import os

    return torch.load(os.path.join(golden_dir, "models_small.pt"))


def _check(got, gold, name, dn):
    want = gold[f"{name}.{dn}"]
    assert got.shape == want.shape
    if dn == "bf16":
        assert torch.equal(got, want), f"{name}: oracle bf16 differs from the reference's bf16 output ({(got == want).float().mean().item()*100:.2f} % equal)"
    else:
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), (name, (got - want).abs().max())


@pytest.mark.parametrize("dtype,dn", [(torch.float32, "fp32"), (torch.bfloat16, "bf16")])
def test_model_oracles_match_reference_golden(model_gold, dtype, dn):
    from oracle import cogvideox_oracle as CO, latte_oracle as LO, model_cases as MC, vchitect_oracle as VO
    from videosys_b200.models.transformers.cogvideox_transformer_3d import CogVideoXTransformer3DModel
    from videosys_b200.models.transformers.latte_transformer_3d import LatteT2V
    from videosys_b200.models.transformers.vchitect_transformer_3d import VchitectXLTransformerModel

    with torch.no_grad():
        # the state-dict templates come from the product's modules (same names / shapes as the reference's, strict-load tested)
        sd = MC.weights(LatteT2V(**MC.LATTE).state_dict(), "latte", dtype)
        x, t, enc = MC.latte_inputs(dtype)
        _check(LO.transformer_forward(sd, MC.LATTE_O, x, t, enc), model_gold, "latte", dn)
        for rot in (False, True):
            tmpl = CogVideoXTransformer3DModel(**dict(MC.COGX, use_rotary_positional_embeddings=rot)).state_dict()
            sd = MC.weights(tmpl, "cogx", dtype, norm_ones=True)
            lat, txt, ts = MC.cogx_inputs(dtype)
            got = CO.transformer_forward(sd, MC.COGX_O, lat, txt, ts, rotary=MC.cogx_rotary() if rot else None)
            _check(got, model_gold, "cogx_rotary" if rot else "cogx", dn)
        sd = MC.weights(VchitectXLTransformerModel(**MC.VCH).state_dict(), "vch", dtype, keep=("pos_embed.pos_embed",))
        lat, enc, pooled, ts = MC.vch_inputs(dtype)
        _check(VO.transformer_forward(sd, MC.VCH_O, lat, enc, pooled, ts), model_gold, "vchitect", dn)
