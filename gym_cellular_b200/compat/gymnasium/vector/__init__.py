"""`gymnasium.vector` look-alike: VectorEnv base class, AutoresetMode, batch_space."""
from enum import Enum

import numpy as np

from .. import spaces as _sp


class AutoresetMode(Enum):
    NEXT_STEP = "NextStep"
    SAME_STEP = "SameStep"
    DISABLED = "Disabled"


def batch_space(space, n=1):
    if isinstance(space, _sp.Discrete):
        return _sp.MultiDiscrete(np.full((n,), space.n, dtype=np.int64),
                                 start=np.full((n,), space.start, dtype=np.int64))
    if isinstance(space, _sp.MultiDiscrete):
        return _sp.MultiDiscrete(np.tile(space.nvec, (n,) + (1,) * space.nvec.ndim))
    if isinstance(space, _sp.MultiBinary):
        return _sp.MultiBinary((n,) + space.shape)
    if isinstance(space, _sp.Box):
        return _sp.Box(np.broadcast_to(space.low, (n,) + space.shape),
                       np.broadcast_to(space.high, (n,) + space.shape), dtype=space.dtype)
    if isinstance(space, _sp.Tuple):
        return _sp.Tuple([batch_space(s, n) for s in space.spaces])
    if isinstance(space, _sp.Dict):
        return _sp.Dict({k: batch_space(s, n) for k, s in space.spaces.items()})
    raise TypeError(f"cannot batch {space!r}")


class VectorEnv:
    metadata = {}
    spec = None
    render_mode = None
    closed = False
    num_envs = None
    observation_space = None
    action_space = None
    single_observation_space = None
    single_action_space = None

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def step(self, actions):
        raise NotImplementedError

    def render(self):
        return None

    def close(self, **kwargs):
        if self.closed:
            return
        self.close_extras(**kwargs)
        self.closed = True

    def close_extras(self, **kwargs):
        return None

    @property
    def unwrapped(self):
        return self

    def __del__(self):
        try:
            if not getattr(self, "closed", True):
                self.close()
        except Exception:
            pass


class utils:  # namespace shim: gymnasium.vector.utils.batch_space
    batch_space = staticmethod(batch_space)


from .async_vector_env import AsyncVectorEnv  # noqa: E402  (needs VectorEnv / batch_space above)

__all__ = ["VectorEnv", "AsyncVectorEnv", "AutoresetMode", "batch_space", "utils"]
