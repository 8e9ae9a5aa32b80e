"""Host-side mirrors of the reference task envs (`envs/rodent.py`, ...) over the CUDA engine."""
from .ant import ANT_ENV_ARGS, AntTracking, ant_task_tables  # noqa: F401
from .base import State  # noqa: F401
from .humanoid import HUMANOID_ENV_ARGS, HumanoidTracking, humanoid_task_tables  # noqa: F401
from .rodent import RODENT_ENV_ARGS, RodentMultiClipTracking, RodentTracking, process_clips_gpu, rodent_task_tables, stack_clips  # noqa: F401

_REGISTRY = {"rodent": RodentTracking, "rodent_multiclip": RodentMultiClipTracking, "humanoidtracking": HumanoidTracking, "ant": AntTracking}


def register_environment(name, cls):  # `brax.envs.register_environment` (reference train.py:65-68)
    _REGISTRY[name] = cls


def get_environment(name, **kwargs):  # `brax.envs.get_environment` (reference train.py:86-90)
    return _REGISTRY[name](**kwargs)
