"""TEST INFRASTRUCTURE ONLY -- restatement of the Pyramid Attention Broadcast integer gate.

Follows core/pab/pab_mgr.py:54-91 (gate), :188-195 (enable_pab ignores mlp_broadcast), :198-200
(update_steps), :203-218 (disabled wrappers return (False, count) untouched).
"""
from typing import Optional, Sequence, Tuple


class PABGate:
    KINDS = ("spatial", "temporal", "cross")

    def __init__(
        self,
        spatial: Optional[Tuple[bool, Sequence[int], int]] = None,
        temporal: Optional[Tuple[bool, Sequence[int], int]] = None,
        cross: Optional[Tuple[bool, Sequence[int], int]] = None,
        steps: Optional[int] = None,
    ):
        """Each kind is (broadcast_on, (lo, hi), range)."""
        off = (False, (0, 0), 1)
        self.cfg = {"spatial": spatial or off, "temporal": temporal or off, "cross": cross or off}
        self.steps = steps

    def enabled(self) -> bool:
        return any(self.cfg[k][0] for k in self.KINDS)

    def gate(self, kind: str, timestep: Optional[int], count: int):
        """-> (reuse_cached, next_count).  Strict lo < t < hi; count % range != 0; the counter advances
        on every call and wraps modulo steps (pab_mgr.py:54-65)."""
        if not self.enabled():
            return False, count
        on, (lo, hi), rng = self.cfg[kind]
        flag = bool(on and (timestep is not None) and (count % rng != 0) and (lo < timestep < hi))
        return flag, (count + 1) % self.steps

    def schedule(self, kind: str, timesteps: Sequence[int]) -> str:
        """Skip bitmap over a whole sampling run for one block (counter starts at 0)."""
        c, bits = 0, []
        for t in timesteps:
            f, c = self.gate(kind, t, c)
            bits.append("1" if f else "0")
        return "".join(bits)


def opensora_default(steps: int) -> PABGate:
    """OpenSoraPABConfig defaults: pipelines/open_sora/pipeline_open_sora.py:32-69 (mlp_broadcast unreachable)."""
    return PABGate((True, (450, 930), 2), (True, (450, 930), 4), (True, (450, 930), 6), steps)
