"""DSP on real GPUs (needs >= 2 devices; run with `gpurun --gpus 2 -- python -m pytest tests/test_dsp_gpu.py -m gpu`).

  * the P2P reshard kernel (vsb_dsp_scatter / vsb_dsp_wait) against the oracle's index math, both directions, with and
    without padding, repeated (epoch / window reuse);
  * the producer- / consumer-fused forms (vsb_ln_modulate_dsp: the modulate kernel's stores are the switch;
    vsb_gate_residual_dsp: the gate + residual kernel pulls the branch from the peers) against the standalone kernels;
  * STDiT3 forward sharded over the ranks (fused P2P, standalone P2P scatter -- row-wise and first version -- and NCCL
    transports) == the single-rank forward, bit for bit: the reshard is a permutation and every kernel is
    row-independent; the same for 4 denoising steps through the replayed step graph (device-side epochs).
"""
import os
import traceback

import pytest
import torch
import torch.multiprocessing as mp

from oracle import cases, dsp_oracle, synth
from tests.helpers import stdit3_state_dict_template

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _need(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def _worker(rank, world, port, q, transport):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        os.environ["VSB_DSP_P2P"] = "0" if transport == "nccl" else "1"
        os.environ["VSB_DSP_FUSED"] = "1" if transport == "p2p" else "0"
        import torch.distributed as dist

        from videosys_b200.core.distributed import comm
        from videosys_b200.core.distributed.parallel_mgr import ParallelManager, initialize
        from videosys_b200.models.transformers.open_sora_transformer_3d import STDiT3, STDiT3Config

        initialize(rank, world)
        dev = torch.device("cuda", rank)
        pm = ParallelManager(1, 1, world)
        res = {}
        if transport == "p2p-scatter-v1":
            from videosys_b200 import kernels as K0

            K0.set_option("dsp_rowwise", 0)
        if transport == "p2p":
            # fused kernels against the standalone ones (which the block below checks against the oracle's index math)
            from videosys_b200 import kernels as K

            for (B, T, S, C) in ((1, 5, 9, 64), (2, 4, 8, 64), (2, 20, 30, 288)):
                full = synth.normalish(f"fz{T}{S}", (B, T, S, C)).to(BF)
                shard = dsp_oracle.split_sequence(full, world, dim=2)[rank].to(dev).contiguous()
                Sl = shard.shape[2]
                Tl = -(-(B * T) // world)
                p2p = comm.DspP2P(pm.sp_group, max(B * T * Sl, Tl * Sl * world) * C, dev)
                t6 = synth.normalish(f"fz.t{C}", (B, 6 * C), std=0.5).to(BF).to(dev)
                t06 = synth.normalish(f"fz.t0{C}", (B, 6 * C), std=0.5).to(BF).to(dev)
                tab = synth.normalish(f"fz.tab{C}", (6, C), std=0.3).to(BF).to(dev)
                mod = K.modulation_table(tab, t6, t06)
                xm_mask = torch.ones(B, T, dtype=torch.uint8, device=dev)
                xm_mask[0, 0] = 0
                x2 = shard.reshape(B, T * Sl, C)
                def fence():  # the test reuses the windows back to back, outside the block's own ordering
                    torch.cuda.synchronize()
                    dist.barrier()

                for rep in range(3):
                    ref_a = p2p.switch(K.ln_modulate(x2, mod, xm_mask, 0, 1, B, T, Sl).view(1, B * T, Sl, C), B * T, S, False).clone()
                    fence()
                    got_a = p2p.ln_modulate_push(x2, mod, xm_mask, 0, 1, B, T, Sl, S).clone()
                    fence()
                    back = p2p.switch(ref_a, B * T, S, True).reshape(B, T * Sl, C).clone()
                    fence()
                    ref_b = K.gate_residual(x2, back, mod, xm_mask, 2, B, T, Sl)
                    cache_ref = torch.empty_like(x2)
                    K.gate_residual(x2, back, mod, xm_mask, 2, B, T, Sl, cache_out=cache_ref)
                    p2p.branch_window(B, T, S, C).copy_(ref_a.reshape(-1, C))
                    cache = torch.empty_like(x2)
                    got_b = p2p.gate_residual_pull(x2, mod, xm_mask, 2, B, T, Sl, S, cache_out=cache)
                    fence()
                res[("fused", B, T, S)] = (torch.equal(ref_a, got_a), torch.equal(ref_b, got_b), torch.equal(cache_ref, cache))
                p2p.close()
        if transport.startswith("p2p"):
            for (T, S) in ((5, 9), (4, 8), (20, 30)):
                B, C = 2, 64
                full = synth.normalish(f"p2p{T}{S}", (B, T, S, C)).to(BF)
                shards = dsp_oracle.split_sequence(full, world, dim=2)
                Sl = shards[0].shape[2]
                p2p = comm.DspP2P(pm.sp_group, B * max(T * Sl, -(-T // world) * Sl * world) * C, dev)
                x = shards[rank].to(dev).contiguous()
                for rep in range(3):
                    a = p2p.switch(x, T, S, to_spatial_shard=False).clone()
                    b = p2p.switch(a, T, S, to_spatial_shard=True).clone()
                torch.cuda.synchronize()
                res[(T, S)] = (a.cpu(), b.cpu())
                dist.barrier()
                p2p.close()
        cfg = cases.small_model_cfg(depth=2)
        sd = synth.fill_state_dict(stdit3_state_dict_template(cfg, BF), "golden.")
        net = STDiT3(STDiT3Config(**cfg)).to(BF)
        net.load_state_dict(sd)
        net = net.to(dev).eval()
        net.enable_parallel(parallel_mgr=pm)
        inp = cases.forward_inputs(BF)
        gi = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
        out = net(**gi)
        torch.cuda.synchronize()
        res["forward"] = out.cpu()
        res["steps"] = _four_steps(net, gi, dev, graph=True).cpu()
        q.put((rank, res, None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        q.put((rank, None, traceback.format_exc()))


def _four_steps(net, gi, dev, graph):
    """6 denoising steps from a fixed latent (with graph=True: two eager, one capture + replay, three replays)."""
    from videosys_b200.core.graph_step import StepGraph

    st = StepGraph(net, 7.0, enabled=graph)
    fwd = {k: v for k, v in gi.items() if k not in ("x", "timestep")}
    z = gi["x"][:1].to(BF).contiguous()
    # step 1 eager (bf16 latent), step 2 eager (first sight of the fp32-latent key), step 3 captures, 4..6 replay
    for t in (900.0, 800.0, 700.0, 500.0, 300.0, 100.0):
        z = st.step(z, torch.tensor([t], device=dev), torch.tensor([0.1], device=dev), fwd)
    torch.cuda.synchronize()
    if graph:
        assert st.replays >= 3, st.replays
    del st  # the captured graphs hold NCCL work: release them before the process group goes away
    return z


def _single_rank_forward():
    from videosys_b200.models.transformers.open_sora_transformer_3d import STDiT3, STDiT3Config

    dev = torch.device("cuda:0")
    cfg = cases.small_model_cfg(depth=2)
    sd = synth.fill_state_dict(stdit3_state_dict_template(cfg, BF), "golden.")
    net = STDiT3(STDiT3Config(**cfg)).to(BF)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    inp = cases.forward_inputs(BF)
    gi = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    return net(**gi).cpu(), _four_steps(net, gi, dev, graph=False).cpu()


def _world():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return 8 if n >= 8 else (4 if n >= 4 else 2)


@pytest.mark.parametrize("transport", ["p2p", "p2p-scatter", "p2p-scatter-v1", "nccl"])
def test_dsp_two_gpus(transport):
    """Runs on 2, 4 or 8 ranks (whatever the box has): reshard vs the oracle, and sharded forward == unsharded."""
    _need(2)
    world, port = _world(), 29800 + (os.getpid() % 100) + ["p2p", "p2p-scatter", "p2p-scatter-v1", "nccl"].index(transport)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, transport), daemon=True) for r in range(world)]
    [p.start() for p in procs]
    got = {}
    try:
        for _ in range(world):
            r, res, err = q.get(timeout=240)
            assert err is None, err
            got[r] = res
        [p.join(timeout=60) for p in procs]
    finally:  # a rank that failed leaves its peers blocked in a collective: never leave them on the GPUs
        for p in procs:
            if p.is_alive():
                p.kill()
        [p.join(timeout=10) for p in procs]
    if transport == "p2p":
        for r in range(world):
            for k, v in got[r].items():
                if isinstance(k, tuple) and k[0] == "fused":
                    assert all(v), f"fused DSP kernels differ from the standalone ones: case {k[1:]} rank {r}: " \
                                   f"push {v[0]}, pull {v[1]}, pull cache {v[2]}"
    if transport.startswith("p2p"):
        for (T, S) in ((5, 9), (4, 8), (20, 30)):
            B, C = 2, 64
            full = synth.normalish(f"p2p{T}{S}", (B, T, S, C)).to(BF)
            shards = dsp_oracle.split_sequence(full, world, dim=2)
            sw, new_s, new_t = dsp_oracle.dynamic_switch([p.reshape(B, -1, C) for p in shards], T, S, False)
            for r in range(world):
                a, b = got[r][(T, S)]
                assert torch.equal(a.reshape(B, -1, C), sw[r]), f"switch to T-shard T={T} S={S} rank {r}"
                # back in the resident layout: real columns identical, pad columns zero (reference pads zeros)
                assert torch.equal(b, shards[r]), f"switch back T={T} S={S} rank {r}"
    ref, ref_steps = _single_rank_forward()
    for r in range(world):
        assert torch.equal(got[r]["forward"], ref), f"sp={world} ({transport}) forward differs from sp=1 on rank {r}"
        assert torch.equal(got[r]["steps"], ref_steps), f"sp={world} ({transport}) 6 graph-replayed steps differ from sp=1 eager on rank {r}"
