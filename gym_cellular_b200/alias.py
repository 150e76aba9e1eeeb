"""`install_alias()`: let code written against the reference keep its imports.

After `gym_cellular_b200.install_alias()`, `import gym_cellular`, `from gym_cellular.envs import
Cells3States3Actions3Env`, `from gym_cellular.envs.cells3states3actions3 import right_polarizing`,
`from gym_cellular.envs.utils import generalized_cellular2tabular` ... resolve to this package's
classes and functions (module layout of the reference: gym_cellular/__init__.py,
gym_cellular/envs/__init__.py:1-6, gym_cellular/envs/utils/__init__.py:1-2).  Nothing is aliased
when a real `gym_cellular` is importable, unless force=True.
"""
import importlib.util
import sys
import types


def install_alias(force=False):
    if "gym_cellular" in sys.modules and getattr(sys.modules["gym_cellular"], "__b200_alias__", False):
        return sys.modules["gym_cellular"]
    if not force and importlib.util.find_spec("gym_cellular") is not None:
        raise ImportError("a real `gym_cellular` package is importable; pass force=True to shadow it")
    from . import codec, envs, tables, registration  # noqa: F401  (registration registers the ids)

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__b200_alias__ = True
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    codec_fns = dict(generalized_cellular2tabular=codec.generalized_cellular2tabular,
                     generalized_tabular2cellular=codec.generalized_tabular2cellular,
                     cellular2tabular=codec.cellular2tabular, tabular2cellular=codec.tabular2cellular)
    utils = module("gym_cellular.envs.utils", **codec_fns)
    utils.generalized_space_transformations = module(
        "gym_cellular.envs.utils.generalized_space_transformations",
        generalized_cellular2tabular=codec.generalized_cellular2tabular,
        generalized_tabular2cellular=codec.generalized_tabular2cellular)
    utils.space_transformations = module("gym_cellular.envs.utils.space_transformations",
                                         cellular2tabular=codec.cellular2tabular, tabular2cellular=codec.tabular2cellular)
    rewards = dict(right_polarizing=tables.right_polarizing, multiple_optima=tables.multiple_optima)
    subs = {
        "cells3states3actions3": dict(Cells3States3Actions3Env=envs.Cells3States3Actions3Env, nonlinear=tables.nonlinear,
                                      PriorKnowledge=envs.PriorKnowledge, n_cells=3, **rewards),
        "cells2rest3": dict(Cells2Rest3Env=envs.Cells2Rest3Env, nonlinear=tables.nonlinear,
                            PriorKnowledge=envs.PriorKnowledge, n_cells=2, **rewards),
        "cells3resetVdeadlock": dict(Cells3ResetVDeadlockEnv=envs.Cells3ResetVDeadlockEnv, PriorKnowledge=envs.PriorKnowledge,
                                     right_polarizing=tables.right_polarizing, nonlinear=tables.nonlinear_right_polarizing,
                                     n_cells=3),
        "grid_world": dict(GridWorldEnv=envs.GridWorldEnv, PriorKnowledge=envs.GridWorldPriorKnowledge,
                           reward_func=envs.grid_reward_func, n_jurisdictions=2, grid_shape=(2, 2)),
    }
    env_classes = dict(Cells3States3Actions3Env=envs.Cells3States3Actions3Env, Cells2Rest3Env=envs.Cells2Rest3Env,
                       Cells3ResetVDeadlockEnv=envs.Cells3ResetVDeadlockEnv, GridWorldEnv=envs.GridWorldEnv,
                       DebugEnv=envs.DebugEnv, DeepPlanningDebugEnv=envs.DeepPlanningDebugEnv,
                       DeepExplorationDebugEnv=envs.DeepExplorationDebugEnv)
    envs_mod = module("gym_cellular.envs", utils=utils, **env_classes, **codec_fns)
    for name, attrs in subs.items():
        setattr(envs_mod, name, module(f"gym_cellular.envs.{name}", **attrs))
    envs_mod.debug = module("gym_cellular.envs.debug", DebugEnv=envs.DebugEnv, DeepPlanningDebugEnv=envs.DeepPlanningDebugEnv,
                            DeepExplorationDebugEnv=envs.DeepExplorationDebugEnv)
    root = module("gym_cellular", envs=envs_mod)
    root.__path__ = []                     # a package, so that `import gym_cellular.envs` resolves through sys.modules
    envs_mod.__path__ = []
    utils.__path__ = []
    return root
