"""`gymnasium.envs.registration` look-alike: register / make / make_vec / registry / spec."""
import importlib
from dataclasses import dataclass, field


@dataclass
class EnvSpec:
    id: str
    entry_point: object = None
    max_episode_steps: object = None
    kwargs: dict = field(default_factory=dict)
    vector_entry_point: object = None
    reward_threshold: object = None
    nondeterministic: bool = False
    order_enforce: bool = True
    disable_env_checker: bool = False


registry = {}


def register(id, entry_point=None, max_episode_steps=None, kwargs=None, vector_entry_point=None, **extra):
    registry[id] = EnvSpec(id=id, entry_point=entry_point, max_episode_steps=max_episode_steps,
                           kwargs=dict(kwargs or {}), vector_entry_point=vector_entry_point)


def spec(id):
    if id not in registry:
        raise KeyError(f"No registered env with id: {id}")
    return registry[id]


def _load(entry_point):
    if callable(entry_point):
        return entry_point
    mod, _, attr = entry_point.partition(":")
    return getattr(importlib.import_module(mod), attr)


def make(id, max_episode_steps=None, **kwargs):
    s = spec(id) if isinstance(id, str) else id
    kw = dict(s.kwargs)
    kw.update(kwargs)
    env = _load(s.entry_point)(**kw)
    try:
        env.spec = s
    except AttributeError:
        pass
    return env


def make_vec(id, num_envs=1, vectorization_mode=None, vector_kwargs=None, wrappers=None, **kwargs):
    s = spec(id) if isinstance(id, str) else id
    if s.vector_entry_point is None:
        raise ValueError(f"{s.id} has no vector_entry_point; the compat stand-in cannot build "
                         "Sync/AsyncVectorEnv")
    kw = dict(s.kwargs)
    kw.update(kwargs)
    kw.update(vector_kwargs or {})
    env = _load(s.vector_entry_point)(num_envs=num_envs, **kw)
    try:
        env.spec = s
    except AttributeError:
        pass
    return env
