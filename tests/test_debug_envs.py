"""the three debug MDPs on the same table-driven kernel, against the reference's outputs
(tests/golden/debug.npz).  CPU part: the table sets; GPU part: one batched launch over every case."""
import numpy as np
import pytest

from gym_cellular_b200 import tables as T

DET = (("dbg", T.debug_tables, 2, 2), ("dplan", T.deep_planning_tables, 4, 2))


def _cases(S, A):
    nS, nA = S * S, A * A
    p = np.arange(nS * nA)
    si, ai = p // nA, p % nA
    return np.stack([si % S, si // S]).astype(np.int8), np.stack([ai % A, ai // A]).astype(np.int8)


@pytest.mark.parametrize("tag,make,S,A", DET)
def test_deterministic_debug_tables(golden_dbg, tag, make, S, A):
    g, tb = golden_dbg, make()
    s, a = _cases(S, A)
    nxt = tb["move"][s, a]
    assert (nxt.T == g[f"{tag}_next"]).all()
    np.testing.assert_allclose(tb["reward"][s, a].sum(0), g[f"{tag}_reward"], rtol=0, atol=1e-15)
    row = np.stack([tb["side_effects"][j][nxt[0], nxt[1]] for j in range(2)])
    assert (row.T == g[f"{tag}_se"][:, 0, :]).all()
    assert (g[f"{tag}_se"][:, 1, :] == tb["se_fill"]).all()
    assert (tb["counted"][nxt].sum(0) / 2 == g[f"{tag}_incidence"]).all()
    assert (np.array([tb["reset_row"], [tb["se_fill"]] * 2]) == g[f"{tag}_reset_se"]).all()


def _dexp_inputs(g):
    sa, u = g["dexp_sa"], g["dexp_u"]
    s = np.stack([sa[:, 0] % 4, sa[:, 0] // 4]).astype(np.int8)
    a = np.stack([sa[:, 1] % 2, sa[:, 1] // 2]).astype(np.int8)
    # the reference walks from level 3 up to level 4, outside its own Discrete(4); the table
    # keeps the cell at 3: those cases (unreachable from reset) are the documented deviation
    in_space = (g["dexp_next"] <= 3).all(1)
    return s, a, u, in_space


def test_deep_exploration_tables(golden_dbg):
    g, tb = golden_dbg, T.deep_exploration_tables()
    s, a, u, ok = _dexp_inputs(g)
    assert (tb["draws"][s, a].T.astype(bool) == ~np.isnan(u)).all()
    fire = (tb["draws"][s, a].astype(bool)) & (np.nan_to_num(u, nan=1.0) < 0.5).T
    nxt = np.where(fire, tb["noisy"][s, a], tb["move"][s, a])
    rew = np.where(fire, tb["reward_noisy"][s, a], tb["reward"][s, a]).sum(0)
    assert (nxt.T[ok] == g["dexp_next"][ok]).all() and ok.sum() == 121
    np.testing.assert_allclose(rew[ok], g["dexp_reward"][ok], atol=1e-15)
    assert (g["dexp_next"][~ok].max(1) == 4).all()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,make,S,A", DET)
def test_deterministic_debug_envs_on_gpu(golden_dbg, tag, make, S, A):
    import torch
    import gym_cellular_b200 as B
    g = golden_dbg
    s, a = _cases(S, A)
    env = B.CellularVectorEnv(num_envs=s.shape[1], cell_tables=make())
    _, info0 = env.reset()
    assert (info0["side_effects"].cpu().numpy()[:, 0] == g[f"{tag}_reset_se"][0]).all()
    env.set_state(s)
    obs, rew, term, trunc, info = env.step(torch.from_numpy(a).cuda())
    assert (env.state.cpu().numpy().T == g[f"{tag}_next"]).all()
    np.testing.assert_allclose(rew.cpu().numpy(), g[f"{tag}_reward"], rtol=1e-6, atol=1e-7)
    assert (info["side_effects"].cpu().numpy().T == g[f"{tag}_se"][:, 0, :]).all()
    assert (info["unsafe"].cpu().numpy() == (g[f"{tag}_se"] == 2).any(axis=(1, 2))).all()
    assert (info["count"].cpu().numpy() / 2 == g[f"{tag}_incidence"]).all()


@pytest.mark.gpu
def test_deep_exploration_on_gpu(golden_dbg):
    import torch
    import gym_cellular_b200 as B
    g = golden_dbg
    s, a, u, ok = _dexp_inputs(g)
    env = B.CellularVectorEnv(num_envs=s.shape[1], cell_tables=T.deep_exploration_tables())
    env.set_state(s)
    obs, rew, *_ = env.step(torch.from_numpy(a).cuda(), replay_u=np.nan_to_num(u, nan=0.0))
    assert (env.state.cpu().numpy().T[ok] == g["dexp_next"][ok]).all()
    np.testing.assert_allclose(rew.cpu().numpy()[ok], g["dexp_reward"][ok], rtol=1e-6, atol=1e-7)
    # Philox-driven: half of the action-0 draws step down
    n = 1 << 18
    env = B.CellularVectorEnv(num_envs=n, cell_tables=T.deep_exploration_tables(), env_seed=3, rng_episodic=False)
    env.set_state(np.ones((2, n), np.int8))
    env.step(torch.zeros(2, n, dtype=torch.int8, device="cuda"))
    frac = float((env.state == 0).float().mean())
    assert abs(frac - 0.5) < 0.01


@pytest.mark.gpu
def test_debug_single_env_ids():
    import gym_cellular_b200  # noqa: F401
    from gym_cellular_b200._gym import gym
    env = gym.make("gym_cellular/Debug-v0")
    state, info = env.reset()
    assert state == (0, 0) and (info["side_effects"] == np.array([["safe", "silent"], ["silent", "silent"]])).all()
    state, r, term, trunc, info = env.step((0, 1))
    assert state == (0, 1) and r == 0.0 and (info["side_effects"] == np.array([["safe", "unsafe"], ["silent", "silent"]])).all()
    state, r, *_ = env.step((1, 0))
    assert state == (1, 1) and r == pytest.approx(0.4, rel=1e-6) and env.unwrapped.data["side_effects_incidence"] == 1.0
    env = gym.make("gym_cellular/DeepPlanningDebug-v0")
    env.reset()
    for _ in range(3):
        state, r, _, _, info = env.step((1, 0))
    assert state == (3, 0) and (info["side_effects"] == "safe").all()
    state, r, *_ = env.step((1, 1))
    assert state == (0, 1) and r == pytest.approx(0.5, rel=1e-6)
    env = gym.make("gym_cellular/DeepExplorationDebug-v0", env_seed=1)
    env.reset()
    states = {env.step((0, 0))[0] for _ in range(40)}
    assert states <= {(0, 0), (0, 1), (1, 0), (1, 1)} and len(states) > 1
