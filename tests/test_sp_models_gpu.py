"""CogVideoX, Latte and Open-Sora-Plan (v1.1.0, v1.2.0) under the reference's multi-GPU modes, on 2 real GPUs (run with
`gpurun --gpus 2 -- python -m pytest tests/test_sp_models_gpu.py -m gpu`):

  * CogVideoX head-scatter sequence parallelism (reference cogvideox_transformer_3d.py:44-165, :531-564) with and without
    row padding, and CFG parallelism (:488-503, :584-585);
  * Latte frame-sharded DSP (reference latte_transformer_3d.py:734-745, :826-843, :1300-1308, :1428-1429) with and without
    frame padding, CFG parallelism (:1198-1216, :1459-1461), and three PAB steps (attention broadcast + MLP skip) sharded.

The check is the strongest one available: every rank's output == the single-GPU forward, bit for bit (the exchanges are
permutations, every kernel is row- and head-independent).
"""
import os
import traceback

import pytest
import torch
import torch.multiprocessing as mp

from oracle import synth

pytestmark = pytest.mark.gpu

COGX = dict(num_attention_heads=4, attention_head_dim=64, in_channels=4, out_channels=4, time_embed_dim=64, text_embed_dim=48,
            num_layers=2, sample_width=16, sample_height=12, sample_frames=9, max_text_seq_length=16)
LATTE = dict(num_attention_heads=4, attention_head_dim=72, in_channels=4, out_channels=8, num_layers=2, sample_size=16,
             caption_channels=64, video_length=6)
# (name, latent shape): CogVideoX [B, F, C, H, W] -> F*(H/2)*(W/2) video rows; Latte [B, C, F, H, W]
COGX_CASES = {"even": (2, 3, 4, 12, 16), "padded": (2, 3, 4, 6, 10)}  # 144 rows / 45 rows (pad 1 at sp = 2)
LATTE_CASES = {"even": (2, 4, 6, 16, 16), "padded": (2, 4, 5, 16, 16)}  # 6 frames / 5 frames (pad 1 at sp = 2)
# Open-Sora-Plan v1.1.0 (RoPE tables must follow the temporal blocks' switch to a patch shard): 5 frames -> pad 1 at sp = 2
OSP = dict(num_attention_heads=4, attention_head_dim=72, in_channels=4, out_channels=8, num_layers=2, cross_attention_dim=288,
           sample_size=(16, 16), caption_channels=64, video_length=5, use_rope=True)
# Open-Sora-Plan v1.2.0: tokens split across the ranks, self-attention exchanges tokens for heads (head_dim 96: attn_mma)
OSP12 = dict(num_attention_heads=4, attention_head_dim=96, in_channels=4, out_channels=8, num_layers=2, cross_attention_dim=384,
             sample_size=(16, 16), sample_size_t=5, caption_channels=64, interpolation_scale_h=1.0, interpolation_scale_w=2.0,
             interpolation_scale_t=1.5, use_rope=True)
PAB_STEPS = [900, 880, 860]


def _need(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def _cogx_net(dev, dt):
    from videosys_b200.models.transformers.cogvideox_transformer_3d import CogVideoXTransformer3DModel

    net = CogVideoXTransformer3DModel(**COGX).to(dt)
    sd = synth.fill_state_dict(net.state_dict(), "spcogx.")
    for k in sd:
        if "norm" in k and k.endswith(".weight") and sd[k].dim() == 1:
            sd[k] = (1.0 + 0.2 * synth.uniform("spcogx." + k, tuple(sd[k].shape))).to(dt)
    net.load_state_dict(sd)
    return net.to(dev).eval()


def _latte_net(dev, dt):
    from videosys_b200.models.transformers.latte_transformer_3d import LatteT2V

    net = LatteT2V(**LATTE).to(dt)
    net.load_state_dict(synth.fill_state_dict(net.state_dict(), "splatte."))
    return net.to(dev).eval()


def _run_all(dev, mode=None):
    """Every case on this process's GPU; mode None = single GPU, "sp" / "cp" = enable_parallel over the 2 ranks."""
    from videosys_b200.core.pab import pab_mgr as P

    dt = torch.float16  # the reference's dtype for both models
    res = {}
    cogx, latte = _cogx_net(dev, dt), _latte_net(dev, dt)
    from videosys_b200.models.transformers.open_sora_plan_v110_transformer_3d import LatteT2V as OspT2V

    osp = OspT2V(**OSP).to(dt)
    osp.load_state_dict(synth.fill_state_dict(osp.state_dict(), "sposp."))
    osp = osp.to(dev).eval()
    from videosys_b200.models.transformers.open_sora_plan_v120_transformer_3d import OpenSoraT2V

    osp12 = OpenSoraT2V(**OSP12).to(dt)
    osp12.load_state_dict(synth.fill_state_dict(osp12.state_dict(), "sposp12."))
    osp12 = osp12.to(dev).eval()
    if mode is not None:
        osp12.enable_parallel(1, 2, enable_cp=(mode == "cp"))
        for net in (cogx, latte, osp):
            net.enable_parallel(1, 2, enable_cp=(mode == "cp"))
            pm = net.parallel_manager
            assert (pm.sp_size, pm.cp_size) == ((1, 2) if mode == "cp" else (2, 1))
    for name, shape in COGX_CASES.items():
        lat = synth.normalish("spcogx.lat" + name, shape).to(dt).to(dev)
        txt = synth.normalish("spcogx.txt", (shape[0], 16, 48)).to(dt).to(dev)
        ts = torch.tensor([499] * shape[0], dtype=torch.int64, device=dev)
        res["cogx." + name] = cogx(lat, txt, ts, return_dict=False)[0].cpu()
    for name, shape in LATTE_CASES.items():
        lat = synth.normalish("splatte.lat" + name, shape).to(dt).to(dev)
        txt = synth.normalish("splatte.txt", (shape[0], 20, 64)).to(dt).to(dev)
        ts = torch.tensor([999] * shape[0], dtype=torch.int64, device=dev)
        res["latte." + name] = latte(lat, timestep=ts, encoder_hidden_states=txt, return_dict=False)[0].cpu()
    lat = synth.normalish("sposp.lat", (2, 4, 5, 16, 16)).to(dt).to(dev)
    txt = synth.normalish("sposp.txt", (2, 1, 20, 64)).to(dt).to(dev)
    msk = torch.ones(2, 1, 20)
    msk[1, 0, 13:] = 0
    ts = torch.tensor([500, 500], dtype=torch.int64, device=dev)
    res["osp_v110.padded"] = osp(lat, timestep=ts, all_timesteps=[900, 500], encoder_hidden_states=txt,
                                 encoder_attention_mask=msk, return_dict=False)[0].cpu()
    res["osp_v120"] = osp12(lat, timestep=ts.float(), encoder_hidden_states=txt, encoder_attention_mask=msk, return_dict=False)[0].cpu()
    # PAB: caches live in the sharded layout (CogVideoX: the attention output of the local rows; Latte: gated outputs of
    # the local frames, MLP outputs through the manager)
    from videosys_b200.pipelines.latte.pipeline_latte import LattePABConfig

    mlp = {900: {"block": [0, 1], "skip_count": 2}}
    try:
        P.set_pab_manager(LattePABConfig(spatial_threshold=(0, 1001), temporal_threshold=(0, 1001), cross_threshold=(0, 1001),
                                         mlp_spatial_broadcast_config=mlp, mlp_temporal_broadcast_config=mlp))
        P.update_steps(len(PAB_STEPS))
        latte.reset_pab_state()
        shape = LATTE_CASES["padded"]
        outs = []
        for i, t in enumerate(PAB_STEPS):
            lat = synth.normalish(f"splatte.pab{i}", shape).to(dt).to(dev)
            txt = synth.normalish("splatte.txt", (shape[0], 20, 64)).to(dt).to(dev)
            ts = torch.tensor([t] * shape[0], dtype=torch.int64, device=dev)
            outs.append(latte(lat, timestep=ts, all_timesteps=PAB_STEPS, encoder_hidden_states=txt, return_dict=False,
                              ts_int=t)[0].cpu())
        res["latte.pab"] = torch.stack(outs)
        P.set_pab_manager(P.PABConfig(spatial_broadcast=True, spatial_threshold=[0, 1001], spatial_range=2))
        P.update_steps(len(PAB_STEPS))
        cogx.reset_pab_state()
        shape = COGX_CASES["padded"]
        outs = []
        for i, t in enumerate(PAB_STEPS):
            lat = synth.normalish(f"spcogx.pab{i}", shape).to(dt).to(dev)
            txt = synth.normalish("spcogx.txt", (shape[0], 16, 48)).to(dt).to(dev)
            ts = torch.tensor([t] * shape[0], dtype=torch.int64, device=dev)
            outs.append(cogx(lat, txt, ts, return_dict=False, ts_int=t)[0].cpu())
        res["cogx.pab"] = torch.stack(outs)
    finally:
        P.set_pab_manager(None)
    torch.cuda.synchronize()
    return res


def _worker(rank, world, port, q, mode):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as dist

        from videosys_b200.core.distributed.parallel_mgr import initialize

        initialize(rank, world)
        res = _run_all(torch.device("cuda", rank), mode)
        q.put((rank, res, None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        q.put((rank, None, traceback.format_exc()))


@pytest.mark.parametrize("mode", ["sp", "cp"])
def test_cogvideox_and_latte_two_gpus(mode):
    _need(2)
    world, port = 2, 29900 + (os.getpid() % 80) + (0 if mode == "sp" else 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode), daemon=True) for r in range(world)]
    [p.start() for p in procs]
    got = {}
    try:
        for _ in range(world):
            r, res, err = q.get(timeout=300)
            assert err is None, err
            got[r] = res
        [p.join(timeout=60) for p in procs]
    finally:  # a rank that failed leaves its peer blocked in a collective: never leave it on the GPU
        for p in procs:
            if p.is_alive():
                p.kill()
        [p.join(timeout=10) for p in procs]
    ref = _run_all(torch.device("cuda:0"))
    for r in range(world):
        for k, v in ref.items():
            assert got[r][k].shape == v.shape, (k, got[r][k].shape, v.shape)
            assert torch.isfinite(v.float()).all(), k
            assert torch.equal(got[r][k], v), f"{mode}=2 differs from one GPU: {k} on rank {r} " \
                                              f"(max abs diff {(got[r][k].float() - v.float()).abs().max().item():.3e})"
    print(f"[parity] CogVideoX + Latte + Open-Sora-Plan v1.1.0 / v1.2.0 under {mode}=2: {len(ref)} cases bit-identical to one GPU on both ranks")
