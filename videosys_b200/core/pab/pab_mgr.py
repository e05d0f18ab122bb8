"""Pyramid Attention Broadcast manager -- host-side mirror of the reference module of the same name.

Same public names and semantics as videosys/core/pab/pab_mgr.py (PABConfig :6-40, PABManager :43-174, module
functions :183-232) so pipelines and blocks written against the reference keep working.  The three attention gates
are one table-driven call into the C-ABI integer gate ``vsb_pab_gate`` (bit-exact: strict ``lo < t < hi``,
``count % range != 0``, the counter advances on every call and wraps modulo ``steps``).  The MLP-skip logic is the
Latte/OSP feature (unreachable for OpenSora in the reference: SURVEY.md fact 7) and is restated from :93-174.
"""
import logging
from typing import Optional

from ... import kernels

PAB_MANAGER: Optional["PABManager"] = None
_KINDS = ("cross", "spatial", "temporal")


class PABConfig:
    def __init__(
        self,
        cross_broadcast: bool = False,
        cross_threshold: list = None,
        cross_range: int = None,
        spatial_broadcast: bool = False,
        spatial_threshold: list = None,
        spatial_range: int = None,
        temporal_broadcast: bool = False,
        temporal_threshold: list = None,
        temporal_range: int = None,
        mlp_broadcast: bool = False,
        mlp_spatial_broadcast_config: dict = None,
        mlp_temporal_broadcast_config: dict = None,
    ):
        self.steps = None
        given = locals()
        for kind in _KINDS:
            for field in ("broadcast", "threshold", "range"):
                setattr(self, f"{kind}_{field}", given[f"{kind}_{field}"])
        self.mlp_broadcast = mlp_broadcast
        self.mlp_spatial_broadcast_config = mlp_spatial_broadcast_config
        self.mlp_temporal_broadcast_config = mlp_temporal_broadcast_config
        self.mlp_temporal_outputs = {}
        self.mlp_spatial_outputs = {}


class PABManager:
    def __init__(self, config: PABConfig):
        self.config = config
        parts = [
            f"{k} broadcast: {getattr(config, k + '_broadcast')}, range: {getattr(config, k + '_range')}, "
            f"threshold: {getattr(config, k + '_threshold')}"
            for k in ("spatial", "temporal", "cross")
        ]
        logging.info("Init Pyramid Attention Broadcast. " + "; ".join(parts) + f"; mlp broadcast: {config.mlp_broadcast}.")

    # -- attention gates ------------------------------------------------------------------------------------
    def _gate(self, kind: str, timestep, count: int):
        c = self.config
        on = bool(getattr(c, kind + "_broadcast"))
        thr = getattr(c, kind + "_threshold") or (0, 0)
        return kernels.pab_gate(on, timestep, count, getattr(c, kind + "_range") or 1, thr[0], thr[1], c.steps)

    def if_broadcast_cross(self, timestep: int, count: int):
        return self._gate("cross", timestep, count)

    def if_broadcast_temporal(self, timestep: int, count: int):
        return self._gate("temporal", timestep, count)

    def if_broadcast_spatial(self, timestep: int, count: int):
        return self._gate("spatial", timestep, count)

    # -- MLP skip (Latte / OSP) -----------------------------------------------------------------------------
    @staticmethod
    def _is_t_in_skip_config(all_timesteps, timestep, config):
        """First key (dict order) present in all_timesteps whose window [key .. key+skip_count] holds timestep."""
        window = None
        for key, spec in config.items():
            if key not in all_timesteps:
                continue
            i = all_timesteps.index(key)
            k = int(spec["skip_count"])
            window = all_timesteps[i : i + 1 + k]
            if timestep in window:
                return True, [all_timesteps[i], all_timesteps[i + k]]
        # on a miss the reference hands back the last examined window (or None); callers only read it on a hit
        return False, window

    def if_skip_mlp(self, timestep: int, count: int, block_idx: int, all_timesteps, is_temporal=False):
        c = self.config
        if not c.mlp_broadcast:
            return False, None, False, None
        cur = c.mlp_temporal_broadcast_config if is_temporal else c.mlp_spatial_broadcast_config
        hit, skip_range = self._is_t_in_skip_config(all_timesteps, timestep, cur)
        flag = next_flag = False
        if timestep is not None and timestep in cur and block_idx in cur[timestep]["block"]:
            next_flag = True  # compute now, keep the output for the following skip_count steps
            count = count + 1
        elif timestep is not None and hit and block_idx in cur[skip_range[0]]["block"]:
            flag = True
            count = 0
        return flag, count, next_flag, skip_range

    def _store(self, is_temporal):
        return self.config.mlp_temporal_outputs if is_temporal else self.config.mlp_spatial_outputs

    def save_skip_output(self, timestep, block_idx, ff_output, is_temporal=False):
        self._store(is_temporal)[(timestep, block_idx)] = ff_output

    def get_mlp_output(self, skip_range, timestep, block_idx, is_temporal=False):
        store = self._store(is_temporal)
        key = (skip_range[0], block_idx)
        out = store.get(key) if store is not None else None
        if out is None:
            raise ValueError(
                f"No stored MLP output found | t {timestep} |[{skip_range[0]}, {skip_range[-1]}] | block {block_idx}"
            )
        if timestep == skip_range[-1]:
            del store[key]
        return out

    def get_spatial_mlp_outputs(self):
        return self.config.mlp_spatial_outputs

    def get_temporal_mlp_outputs(self):
        return self.config.mlp_temporal_outputs


def set_pab_manager(config: PABConfig):
    global PAB_MANAGER
    PAB_MANAGER = PABManager(config) if config is not None else None


def enable_pab() -> bool:
    """True when any attention broadcast is on (mlp_broadcast alone does not count: reference :188-195)."""
    if PAB_MANAGER is None:
        return False
    c = PAB_MANAGER.config
    return bool(c.cross_broadcast or c.spatial_broadcast or c.temporal_broadcast)


def update_steps(steps: int):
    if PAB_MANAGER is not None:
        PAB_MANAGER.config.steps = steps


def if_broadcast_cross(timestep: int, count: int):
    return PAB_MANAGER.if_broadcast_cross(timestep, count) if enable_pab() else (False, count)


def if_broadcast_temporal(timestep: int, count: int):
    return PAB_MANAGER.if_broadcast_temporal(timestep, count) if enable_pab() else (False, count)


def if_broadcast_spatial(timestep: int, count: int):
    return PAB_MANAGER.if_broadcast_spatial(timestep, count) if enable_pab() else (False, count)


def if_broadcast_mlp(timestep: int, count: int, block_idx: int, all_timesteps, is_temporal=False):
    if not enable_pab():
        return False, count
    return PAB_MANAGER.if_skip_mlp(timestep, count, block_idx, all_timesteps, is_temporal)


def save_mlp_output(timestep: int, block_idx: int, ff_output, is_temporal=False):
    return PAB_MANAGER.save_skip_output(timestep, block_idx, ff_output, is_temporal)


def get_mlp_output(skip_range, timestep, block_idx: int, is_temporal=False):
    return PAB_MANAGER.get_mlp_output(skip_range, timestep, block_idx, is_temporal)
