"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference modules for oracle pinning.

Imports VideoSys' OpenSora hot-path modules straight from ``/root/reference`` (read-only) by
registering a bare ``videosys`` namespace package (skipping ``videosys/__init__.py:1-12`` which
pulls every pipeline -> diffusers) and stubbing the third-party names that are not installed in
this image (timm ``Mlp``/``DropPath``, colossalai ``ProcessGroupMesh``, diffusers ``Attention``,
``rotary_embedding_torch.RotaryEmbedding``, imageio, omegaconf).  Recipe: SURVEY.md Appendix C.

Only ``tests/`` and ``oracle/gen_golden.py`` may import this file, and only inside the authoring
container: ``/root/reference`` does not exist on the GPU box, ``available()`` says so.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("VSB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "videosys"))


def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


class _Mlp(nn.Module):
    """timm.models.vision_transformer.Mlp semantics: fc2(act(fc1(x))), biases on, dropout 0."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, **_):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _RotaryEmbedding(nn.Module):
    """rotary_embedding_torch.RotaryEmbedding(dim) ('lang' freqs, theta 1e4), rotate_queries_or_keys only."""

    def __init__(self, dim, theta=10000):
        super().__init__()
        self.freqs = nn.Parameter(
            1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim)), requires_grad=False
        )

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        n = t.shape[seq_dim]
        pos = torch.arange(n, device=t.device, dtype=torch.float32)
        f = torch.einsum("i,j->ij", pos, self.freqs.float()).repeat_interleave(2, dim=-1)
        x1, x2 = t.reshape(*t.shape[:-1], -1, 2).unbind(-1)
        rh = torch.stack((-x2, x1), -1).flatten(-2)
        return (t * f.cos() + rh * f.sin()).type(t.dtype)


_LOADED = {}


def load():
    """Returns a namespace with the reference modules (cached)."""
    if _LOADED:
        return types.SimpleNamespace(**_LOADED)
    if not available():
        raise RuntimeError(f"reference tree not present at {REF}")
    import transformers  # noqa: F401  (must be imported before timm is stubbed)

    pkg = _mod("videosys")
    pkg.__path__ = [REF + "/videosys"]
    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers", DropPath=nn.Identity)
    _mod("timm.models.vision_transformer", Mlp=_Mlp)
    _mod("colossalai")
    _mod("colossalai.cluster")
    _mod(
        "colossalai.cluster.process_group_mesh",
        ProcessGroupMesh=type("ProcessGroupMesh", (), {"__init__": lambda s, *a: None}),
    )
    _mod("diffusers")
    _mod("diffusers.models")
    _mod("diffusers.models.attention", Attention=object)
    _mod("diffusers.models.attention_processor", AttnProcessor=object)
    _mod("imageio")
    _mod("omegaconf", DictConfig=dict, ListConfig=list, OmegaConf=object)
    _mod("rotary_embedding_torch", RotaryEmbedding=_RotaryEmbedding)
    _LOADED["stdit3"] = importlib.import_module("videosys.models.transformers.open_sora_transformer_3d")
    _LOADED["attentions"] = importlib.import_module("videosys.models.modules.attentions")
    _LOADED["normalization"] = importlib.import_module("videosys.models.modules.normalization")
    _LOADED["pab_mgr"] = importlib.import_module("videosys.core.pab.pab_mgr")
    _LOADED["comm"] = importlib.import_module("videosys.core.distributed.comm")
    _LOADED["rflow"] = importlib.import_module("videosys.schedulers.scheduling_rflow_open_sora")
    return types.SimpleNamespace(**_LOADED)


class SingleRankPM:
    """Stand-in for ParallelManager on one rank (parallel_mgr.py:14-39)."""

    sp_size = 1
    cp_size = 1
    sp_group = None
    cp_group = None


def build_stdit3(depth=1, hidden_size=1152, num_heads=16, dtype=torch.float32, **kw):
    ref = load()
    M = ref.stdit3
    net = M.STDiT3(M.STDiT3Config(depth=depth, hidden_size=hidden_size, num_heads=num_heads, **kw)).eval().to(dtype)
    net.parallel_manager = SingleRankPM()
    for b in [*net.spatial_blocks, *net.temporal_blocks]:
        b.parallel_manager = SingleRankPM()
    return net


def load_cogvideox_scheduler():
    """The reference's CogVideoXDDIMScheduler class (schedulers/scheduling_ddim_cogvideox.py), imported unmodified with
    the three diffusers base names it needs stubbed (ConfigMixin / register_to_config: keep the constructor kwargs as
    ``self.config``; SchedulerMixin; BaseOutput).  All scheduler arithmetic is the reference's own."""
    import functools
    import inspect

    if not available():
        raise RuntimeError(f"reference tree not present at {REF}")
    if "videosys" not in sys.modules:
        pkg = _mod("videosys")
        pkg.__path__ = [REF + "/videosys"]

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *a, **kw):
            sig = inspect.signature(init)
            bound = sig.bind(self, *a, **kw)
            bound.apply_defaults()
            self.config = types.SimpleNamespace(**{k: v for k, v in bound.arguments.items() if k != "self"})
            init(self, *a, **kw)

        return wrapper

    class BaseOutput:
        def __init_subclass__(cls, **kw):
            super().__init_subclass__(**kw)

    for name in ("diffusers", "diffusers.schedulers"):
        if name not in sys.modules:
            _mod(name)
    _mod("diffusers.configuration_utils", ConfigMixin=type("ConfigMixin", (), {}), register_to_config=register_to_config)
    _mod("diffusers.schedulers.scheduling_utils", KarrasDiffusionSchedulers=[], SchedulerMixin=type("SchedulerMixin", (), {}))
    _mod("diffusers.utils", BaseOutput=BaseOutput)
    return importlib.import_module("videosys.schedulers.scheduling_ddim_cogvideox").CogVideoXDDIMScheduler


# ---- Open-Sora-Plan v1.1.0: the reference module imported unmodified, diffusers LEAF classes restated --------------------------
class _RefGELU(nn.Module):
    """diffusers.models.activations.GELU(dim_in, dim_out, approximate, bias): Linear then F.gelu."""

    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, x, *a, **k):
        return torch.nn.functional.gelu(self.proj(x), approximate=self.approximate)


class _RefTimesteps(nn.Module):
    """diffusers Timesteps(num_channels, flip_sin_to_cos, downscale_freq_shift): get_timestep_embedding, scale 1."""

    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.n, self.flip, self.shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        import math

        half = self.n // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - self.shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        return torch.cat([emb[:, half:], emb[:, :half]], dim=-1) if self.flip else emb


class _RefTimestepEmbedding(nn.Module):
    """diffusers TimestepEmbedding(in_channels, time_embed_dim): linear_1 -> SiLU -> linear_2."""

    def __init__(self, in_channels, time_embed_dim, **_):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


def _register_to_config(init):
    import functools
    import inspect

    @functools.wraps(init)
    def wrapper(self, *a, **kw):
        bound = inspect.signature(init).bind(self, *a, **kw)
        bound.apply_defaults()
        self.__dict__["config"] = types.SimpleNamespace(**{k: v for k, v in bound.arguments.items() if k != "self"})
        init(self, *a, **kw)

    return wrapper


class _ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype


def load_osp_v110():
    """The reference's models/transformers/open_sora_plan_v110_transformer_3d.py, imported unmodified.  Everything the file
    defines itself (RoPE1D/2D, PatchEmbed, Attention + AttnProcessor2_0, FeedForward, both block classes, AdaLayerNormSingle,
    CaptionProjection, LatteT2V) runs as written; of the diffusers names it imports only three leaf classes carry arithmetic
    on the inference path and are restated here (GELU, Timesteps, TimestepEmbedding); the rest are placeholders."""
    if "osp_v110" in _LOADED:
        return _LOADED["osp_v110"]
    load()  # videosys namespace + the shared stubs
    ph = lambda name: type(name, (nn.Module,), {})  # noqa: E731  (placeholder class: never instantiated on this path)
    _mod("diffusers.configuration_utils", ConfigMixin=type("ConfigMixin", (), {}), register_to_config=_register_to_config)
    _mod("diffusers.models.activations", GEGLU=ph("GEGLU"), GELU=_RefGELU, ApproximateGELU=ph("ApproximateGELU"))
    names = ["AttnAddedKVProcessor", "AttnAddedKVProcessor2_0", "AttnProcessor", "CustomDiffusionAttnProcessor",
             "CustomDiffusionAttnProcessor2_0", "CustomDiffusionXFormersAttnProcessor", "LoRAAttnAddedKVProcessor",
             "LoRAAttnProcessor", "LoRAAttnProcessor2_0", "LoRAXFormersAttnProcessor", "SlicedAttnAddedKVProcessor",
             "SlicedAttnProcessor", "SpatialNorm", "XFormersAttnAddedKVProcessor", "XFormersAttnProcessor"]
    _mod("diffusers.models.attention_processor", **{n: ph(n) for n in names})
    _mod("diffusers.models.embeddings", SinusoidalPositionalEmbedding=ph("SinusoidalPositionalEmbedding"),
         TimestepEmbedding=_RefTimestepEmbedding, Timesteps=_RefTimesteps)
    _mod("diffusers.models.lora", LoRACompatibleConv=nn.Conv2d, LoRACompatibleLinear=nn.Linear)
    _mod("diffusers.models.modeling_utils", ModelMixin=_ModelMixin)
    _mod("diffusers.models.normalization", AdaLayerNorm=ph("AdaLayerNorm"), AdaLayerNormZero=ph("AdaLayerNormZero"))
    _mod("diffusers.utils", USE_PEFT_BACKEND=True, BaseOutput=type("BaseOutput", (), {}), deprecate=lambda *a, **k: None,
         is_xformers_available=lambda: False)
    _mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda cls: cls)
    m = importlib.import_module("videosys.models.transformers.open_sora_plan_v110_transformer_3d")
    _LOADED["osp_v110"] = m
    return m


def build_osp_v110(dtype=torch.float32, **cfg):
    M = load_osp_v110()
    net = M.LatteT2V(**cfg).eval().to(dtype)
    net.parallel_manager = SingleRankPM()
    for mod in net.modules():
        if hasattr(mod, "parallel_manager"):
            mod.parallel_manager = SingleRankPM()
    return net
