"""The reference-shaped single-env classes (gymnasium.make ids) run through the CUDA kernels with
num_envs = 1: known-answer trajectories of the reference, returned in the reference's own types."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_cells3_known_answer_trajectory(golden_pol):
    import gym_cellular_b200  # noqa: F401
    from gym_cellular_b200._gym import gym
    g = golden_pol
    env = gym.make("gym_cellular/Cells3States3Actions3-v0")
    state, info = env.reset()
    assert state == (0, 0, 0) and info["side_effects"].dtype == np.dtype("<U6")
    assert (info["side_effects"] == np.array([["safe", "silent", "silent"]] + [["silent"] * 3] * 2)).all()
    names = np.array(["silent", "safe", "unsafe"])
    for a, s, r, inc, se in zip(g["kat_c3_actions"], g["kat_c3_states"], g["kat_c3_rewards"],
                                g["kat_c3_incidence"], g["kat_c3_se"]):
        state, reward, term, trunc, info = env.step(tuple(int(x) for x in a))
        assert state == tuple(s) and term is False and trunc is False and isinstance(reward, float)
        assert reward == pytest.approx(r, rel=1e-6)
        assert (info["side_effects"] == names[se]).all()
        assert env.unwrapped.get_data()["side_effects_incidence"] == inc
    assert env.unwrapped.get_data()["time_step"] == 4 and env.unwrapped.get_state() == state
    env.unwrapped.state = (2, 2, 2)                                   # direct assignment, as agents/tests do
    state, reward, *_ = env.step((2, 2, 2))
    assert state == (2, 2, 2) and reward == pytest.approx(0.75, rel=1e-6)
    bad = gym.make("gym_cellular/Cells3States3Actions3-v0", difficulty="nope")
    with pytest.raises(ValueError, match="Difficulty must be one of"):
        bad.reset()
        bad.step((0, 0, 0))


def test_stochastic_env_replays_noise_after_reset():
    """cells3resetVdeadlock.py:131: reset() re-seeds, so every episode sees the same noise."""
    import gym_cellular_b200  # noqa: F401
    from gym_cellular_b200._gym import gym
    env = gym.make("gym_cellular/Cells3ResetVDeadlock-v0", env_seed=12345)
    acts = np.random.default_rng(99).integers(0, 3, size=(60, 3))
    runs = []
    for _ in range(2):
        env.reset()
        runs.append([env.step(tuple(int(x) for x in a))[0] for a in acts])
    assert runs[0] == runs[1]
    other = gym.make("gym_cellular/Cells3ResetVDeadlock-v0", env_seed=54321)
    other.reset()
    assert [other.step(tuple(int(x) for x in a))[0] for a in acts] != runs[0]


def test_gridworld_single_env_types_and_oracle():
    import gym_cellular_b200  # noqa: F401
    from gym_cellular_b200._gym import gym
    from oracle import oracle as O
    env = gym.make("gym_cellular/GridWorld-v0", env_seed=5).unwrapped
    pk = env.prior_knowledge
    ora = O.OracleEnv(kind="gridworld", n_envs=1, seed=5)
    state, info = env.reset()
    assert (pk.cellularize(state, "state") == [15, 18]).all() and pk.tabularize(state, "state") == 375
    assert (info["side_effects"] == np.array([["safe", "safe"], ["silent", "silent"]])).all()
    for code in ([2, 4], [2, 4], [0, 4], [4, 0], [4, 3], [3, 4]):
        state, reward, term, trunc, info = env.step(pk.decellularize(code, "action"))
        ora.step(np.array(code, np.int8).reshape(2, 1))
        assert (pk.cellularize(state, "state") == ora.state[:, 0]).all() and reward == ora.reward[0]
        assert isinstance(state[0]["living_trees"], np.ndarray) and term is False and trunc is False
        assert env.data["side_effects_incidence"] == ora.count[0] / 2
    with pytest.raises(KeyError):
        env.step(pk.decellularize([4, 4], "action"))
