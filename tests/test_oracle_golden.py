"""Pins the CPU oracle (oracle/gc_oracle.c) against outputs of the UNMODIFIED reference.

Fixtures: tests/golden/*.npz, written by tests/golden/make_golden.py which executes the reference's
own env.step()/reset()/prior_knowledge code.  Integer outputs must be bit-exact, rewards equal to
the reference's float64 to the last bit (the oracle accumulates in float64 in the same order).
"""
import numpy as np
import pytest

from conftest import detab
from oracle import oracle as O


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kats = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
            ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
            ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
             [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kats:
        assert O.philox4x32_10(ctr, key).tolist() == want


@pytest.mark.parametrize("tag,C", [("c3", 3), ("c2", 2)])
@pytest.mark.parametrize("difficulty", ["easy", "hard", "impossible"])
@pytest.mark.parametrize("reward,rid", [("right_polarizing", "right_polarizing"),
                                        ("multiple_optima", "multiple_optima"),
                                        ("nonlinear", "nonlinear_mo")])
def test_polarisation_exhaustive(golden_pol, tag, C, difficulty, reward, rid):
    g = golden_pol
    n_s = 3 ** C
    pairs = np.arange(n_s * n_s)
    env = O.OracleEnv(n_envs=len(pairs), n_cells=C, reward=rid, difficulty=difficulty)
    env.state[:] = detab(pairs // n_s, C, 3)
    env.step(detab(pairs % n_s, C, 3))
    assert (env.state.T == g[f"{tag}_next"]).all()
    assert (env.index == g[f"{tag}_next_tab"]).all()
    assert (env.reward == g[f"{tag}_reward_{reward}"]).all()          # float64, bit-exact
    se = g[f"{tag}_se_{difficulty}"]
    assert (se[:, 1:, :] == 0).all()                                  # only row 0 is ever written
    assert (env.se_row.T == se[:, 0, :]).all()
    assert (env.unsafe == (se == 2).any(axis=(1, 2))).all()
    assert (env.count / C == g[f"{tag}_incidence"]).all()
    assert not env.terminated.any() and not env.truncated.any() and (env.t == 1).all()


def test_polarisation_reset_and_kat(golden_pol):
    g = golden_pol
    for tag, C in (("c3", 3), ("c2", 2)):
        env = O.OracleEnv(n_envs=4, n_cells=C)
        assert (env.state.T == g[f"{tag}_reset_state"]).all() and (env.index == 0).all()
        assert g[f"{tag}_meta"].tolist() == [C, 3 ** C, 3 ** C, 3, 3]
    env = O.OracleEnv(n_envs=1)
    for a, s, r, inc, se in zip(g["kat_c3_actions"], g["kat_c3_states"], g["kat_c3_rewards"],
                                g["kat_c3_incidence"], g["kat_c3_se"]):
        env.step(a.reshape(3, 1))
        assert (env.state[:, 0] == s).all() and env.reward[0] == r
        assert env.count[0] / 3 == inc and (env.se_row[:, 0] == se[0]).all()


@pytest.mark.parametrize("mode", ["rs", "dl"])
def test_polarisation_noise_exhaustive(golden_pol, mode):
    g = golden_pol
    sa, u = g[f"noise_{mode}_sa"], g[f"noise_{mode}_u"]
    n = len(sa)
    assert n == 2744
    for reward, rid in (("nonlinear", "nonlinear_rp"), ("right_polarizing", "right_polarizing")):
        env = O.OracleEnv(n_envs=n, noise=True, deadlock=(mode == "dl"), replay=True, reward=rid)
        env.state[:] = detab(sa[:, 0], 3, 3)
        # NaN slots are cells the reference did not draw for: poison them so a wrong draw shows up
        env.step(detab(sa[:, 1], 3, 3), replay_u=np.where(np.isnan(u), 0.0, u))
        assert (env.state.T == g[f"noise_{mode}_next"]).all()
        assert (env.reward == g[f"noise_{mode}_reward_{reward}"]).all()
        se = g[f"noise_{mode}_se_easy"]
        assert (env.se_row.T == se[:, 0, :]).all() and (se[:, 1:, :] == 0).all()
        assert (env.count / 3 == g[f"noise_{mode}_incidence"]).all()


@pytest.mark.parametrize("mode", ["rs", "dl"])
@pytest.mark.parametrize("seed", [12345, 7])
def test_polarisation_noise_trajectory_real_mt19937(golden_pol, mode, seed):
    g = golden_pol
    tag = f"traj_{mode}_{seed}"
    acts, states, rewards, u = (g[f"{tag}_{k}"] for k in ("actions", "states", "rewards", "u"))
    env = O.OracleEnv(n_envs=1, noise=True, deadlock=(mode == "dl"), replay=True, reward="nonlinear_rp")
    for t in range(len(acts)):
        env.step(acts[t].reshape(3, 1), replay_u=np.where(np.isnan(u[t]), 0.0, u[t]).reshape(1, 3))
        assert (env.state[:, 0] == states[t]).all(), t
        assert env.reward[0] == rewards[t]


def _gw_replay(u0, bits, k):
    n = len(u0)
    u = np.zeros((n, 6))
    u[:, 0] = u0
    u[:, 1:5] = bits * 0.5 + 0.25          # randint(2) replayed as floor(u * 2)
    u[:, 5] = k * 0.5 + 0.25
    return u


def test_gridworld_exhaustive(golden_gw):
    g = golden_gw
    s, a = g["gw_case_state"], g["gw_case_action"]
    n = len(s)
    assert n == 120 * 24 * 33 and g["gw_meta"].tolist() == [2, 400, 25, 20, 5]
    env = O.OracleEnv(kind="gridworld", n_envs=n, replay=True)
    assert (env.initial_state() == g["gw_reset_cell"]).all() and env.index[0] == g["gw_reset_tab"]
    env.state[:] = s.T
    env.step(a.T.copy(), replay_u=_gw_replay(g["gw_case_u0"], g["gw_case_bits"], g["gw_case_k"]))
    assert (env.state.T == g["gw_case_next"]).all()
    assert (env.index == g["gw_case_tab"]).all()
    assert (env.reward == g["gw_case_reward"]).all()
    assert (env.count / 2 == g["gw_case_incidence"]).all()
    se = g["gw_case_se"]
    assert (se[:, 1, :] == 0).all() and (env.se_row.T == se[:, 0, :]).all() and not env.unsafe.any()


def test_gridworld_barren_states_draw_nothing(golden_gw):
    g = golden_gw
    s, a = g["gw_barren_state"], g["gw_barren_action"]
    env = O.OracleEnv(kind="gridworld", n_envs=len(s), replay=True)
    env.state[:] = s.T
    # a trigger value in slot 0 must be ignored: the reference does not draw when all is barren
    env.step(a.T.copy(), replay_u=np.zeros((len(s), 6)))
    assert (env.state.T == g["gw_barren_next"]).all() and (env.index == g["gw_barren_tab"]).all()
    assert (env.reward == g["gw_barren_reward"]).all() and (env.count == 2).all()
    assert (env.se_row.T == g["gw_barren_se"][:, 0, :]).all()


def test_gridworld_no_position_action_raises(golden_gw):
    env = O.OracleEnv(kind="gridworld", n_envs=1, replay=True)
    assert "position" in str(golden_gw["gw_noaction_error"])
    with pytest.raises(KeyError):
        env.step(np.array([[4], [4]], np.int8), replay_u=np.full((1, 6), 0.5))


def test_gridworld_trajectory_real_mt19937(golden_gw):
    g = golden_gw
    ep_len = int(g["gwtraj_ep_len"])
    acts = g["gwtraj_actions"]
    u = _gw_replay(np.nan_to_num(g["gwtraj_u0"], nan=0.0), np.maximum(g["gwtraj_bits"], 0),
                   np.maximum(g["gwtraj_k"], 0))
    env = O.OracleEnv(kind="gridworld", n_envs=1, replay=True, max_episode_steps=ep_len)
    for t in range(len(acts)):
        env.step(acts[t].reshape(2, 1), replay_u=u[t:t + 1])
        last = (t + 1) % ep_len == 0
        assert env.truncated[0] == last
        if not last:                       # at the time limit the oracle returns the reset state
            assert (env.state[:, 0] == g["gwtraj_cells"][t]).all(), t
            assert env.index[0] == g["gwtraj_tab"][t]
        else:
            assert (env.state[:, 0] == g["gw_reset_cell"]).all() and env.t[0] == 0
        assert env.reward[0] == g["gwtraj_reward"][t]
        assert env.count[0] / 2 == g["gwtraj_incidence"][t]
        assert (env.se_row[:, 0] == g["gwtraj_se"][t][0]).all()


def test_codec(golden_pol, golden_gw):
    g = golden_pol
    lens, mins, cells = g["codec_ragged_lens"], g["codec_ragged_mins"], g["codec_ragged_cells"]
    for i in range(0, len(cells), 7):
        assert O.encode_mixed_radix(cells[i], mins, lens) == i
        assert (O.decode_mixed_radix(i, mins, lens) == cells[i]).all()
    c16, t16 = g["codec_c16_cells"], g["codec_c16_tab"]
    idx = O.encode(c16.T.copy(), 4)
    assert idx.dtype == np.uint32 and (idx.astype(np.uint64) == t16).all() and idx[0] == 2 ** 32 - 1
    assert (O.decode(idx, 16, 4).T == c16).all()
    assert (O.encode(g["codec_fixed_cells"].T.copy(), 3) == g["codec_fixed_tab"]).all()
    assert (O.decode(np.arange(27, dtype=np.uint32), 3, 3).T == g["c3_detab"]).all()
    st = golden_gw["gw_states"]
    assert (O.encode(st.T.copy(), 20) == st[:, 0].astype(int) + 20 * st[:, 1].astype(int)).all()
    assert (O.encode(golden_gw["gw_actions"].T.copy(), 5) == golden_gw["gw_action_tab"]).all()


def test_generalisation_reduces_to_reference_rules():
    """The S=4 / C=16 rules (not in the reference) collapse to the pinned S=3 rules: every cell j>=2
    of a wide env behaves like the reference's cell 2, and levels map by role."""
    rng = np.random.default_rng(0)
    n = 4096
    wide = O.OracleEnv(n_envs=n, n_cells=16, n_states=3, reward="right_polarizing")
    st = rng.integers(0, 3, size=(16, n)).astype(np.int8)
    ac = rng.integers(0, 3, size=(16, n)).astype(np.int8)
    wide.state[:] = st
    wide.step(ac)
    for j in range(2, 16):
        ref = O.OracleEnv(n_envs=n, n_cells=3)
        ref.state[:] = st[[0, 1, j]]
        ref.step(ac[[0, 1, j]])
        assert (wide.state[[0, 1, j]] == ref.state).all()
        assert (wide.se_row[[0, 1, j]] == ref.se_row).all()


def test_python_port_matches_reference(golden_pol):
    """oracle/pyport.py (the timing stand-in for the reference's Python step loop) against the golden vectors."""
    from oracle.pyport import PolarisationEnv
    g = golden_pol
    names = np.array(["silent", "safe", "unsafe"])
    for tag, C in (("c3", 3), ("c2", 2)):
        env = PolarisationEnv(C, 3)
        n_s = 3 ** C
        for p in range(n_s * n_s):
            s, a = detab([p // n_s], C, 3)[:, 0], detab([p % n_s], C, 3)[:, 0]
            env.reset()
            env.state = tuple(int(x) for x in s)
            ns, r, term, trunc, info = env.step(tuple(int(x) for x in a))
            assert ns == tuple(g[f"{tag}_next"][p]) and r == g[f"{tag}_reward_right_polarizing"][p]
            assert (info["side_effects"] == names[g[f"{tag}_se_easy"][p]]).all()
            assert env.data["side_effects_incidence"] == g[f"{tag}_incidence"][p]
