"""Registers the reference's env ids (gym_cellular/__init__.py:4-45) with this package's classes and
adds a `vector_entry_point` per id, so both `gymnasium.make(id)` and `gymnasium.make_vec(id, N)` work."""
from ._gym import gym
from .vector_env import CellularVectorEnv

ENV_IDS = {
    "gym_cellular/Cells3States3Actions3-v0": ("gym_cellular_b200.envs:Cells3States3Actions3Env", "Cells3States3Actions3Vec"),
    "gym_cellular/Cells2Rest3-v0": ("gym_cellular_b200.envs:Cells2Rest3Env", "Cells2Rest3Vec"),
    "gym_cellular/Cells3ResetVDeadlock-v0": ("gym_cellular_b200.envs:Cells3ResetVDeadlockEnv", "Cells3ResetVDeadlockVec"),
    "gym_cellular/GridWorld-v0": ("gym_cellular_b200.envs:GridWorldEnv", "GridWorldVec"),
    "gym_cellular/Debug-v0": ("gym_cellular_b200.envs:DebugEnv", "DebugVec"),
    "gym_cellular/DeepPlanningDebug-v0": ("gym_cellular_b200.envs:DeepPlanningDebugEnv", "DeepPlanningDebugVec"),
    "gym_cellular/DeepExplorationDebug-v0": ("gym_cellular_b200.envs:DeepExplorationDebugEnv", "DeepExplorationDebugVec"),
}


def Cells3States3Actions3Vec(num_envs=1, **kwargs):
    return CellularVectorEnv(kind="cellular", num_envs=num_envs, n_cells=3, n_states=3, **kwargs)


def Cells2Rest3Vec(num_envs=1, **kwargs):
    return CellularVectorEnv(kind="cellular", num_envs=num_envs, n_cells=2, n_states=3, **kwargs)


def Cells3ResetVDeadlockVec(num_envs=1, **kwargs):
    kwargs.setdefault("stochastic", True)
    return CellularVectorEnv(kind="cellular", num_envs=num_envs, n_cells=3, n_states=3, **kwargs)


def GridWorldVec(num_envs=1, **kwargs):
    return CellularVectorEnv(kind="gridworld", num_envs=num_envs, **kwargs)


def DebugVec(num_envs=1, **kwargs):
    from . import tables
    return CellularVectorEnv(kind="cellular", num_envs=num_envs, cell_tables=tables.debug_tables(), **kwargs)


def DeepPlanningDebugVec(num_envs=1, **kwargs):
    from . import tables
    return CellularVectorEnv(kind="cellular", num_envs=num_envs, cell_tables=tables.deep_planning_tables(), **kwargs)


def DeepExplorationDebugVec(num_envs=1, **kwargs):
    from . import tables
    kwargs.setdefault("rng_episodic", False)
    return CellularVectorEnv(kind="cellular", num_envs=num_envs, cell_tables=tables.deep_exploration_tables(), **kwargs)


def register_all():
    for env_id, (entry, vec) in ENV_IDS.items():
        gym.register(id=env_id, entry_point=entry, max_episode_steps=None,
                     vector_entry_point=f"gym_cellular_b200.registration:{vec}")


register_all()
