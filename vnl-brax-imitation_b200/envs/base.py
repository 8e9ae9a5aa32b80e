"""brax `State` mirror (`brax/envs/base.py`): the container `Env.reset/step` return."""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Any, Dict


class PipelineState(dict):
    """Leaves of the mjx pipeline state the reference's callers read, with attribute access
    (`state.pipeline_state.qpos`, reference train.py:293)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    @property
    def q(self):  # brax alias used by envs/rodent.py:314
        return self["qpos"]

    @property
    def qd(self):
        return self["qvel"]


@dataclass
class State:
    pipeline_state: PipelineState
    obs: Any
    reward: Any
    done: Any
    metrics: Dict[str, Any] = field(default_factory=dict)
    info: Dict[str, Any] = field(default_factory=dict)

    def replace(self, **kw) -> "State":
        return replace(self, **kw)
