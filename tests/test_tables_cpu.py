"""Host logic: the tables lowered in gym_cellular_b200/tables.py against the reference's outputs
(golden vectors), independently of the oracle and of any GPU."""
import numpy as np
import pytest

from conftest import detab
from gym_cellular_b200 import tables as T


@pytest.mark.parametrize("tag,C", [("c3", 3), ("c2", 2)])
def test_move_and_rewards(golden_pol, tag, C):
    g = golden_pol
    n_s = 3 ** C
    pairs = np.arange(n_s * n_s)
    s, a = detab(pairs // n_s, C, 3), detab(pairs % n_s, C, 3)
    assert (T.move_table(3, 3)[s, a].T == g[f"{tag}_next"]).all()
    for name, f in (("right_polarizing", T.right_polarizing), ("multiple_optima", T.multiple_optima),
                    ("nonlinear", T.nonlinear)):
        got = np.array([f(tuple(s[:, i]), tuple(a[:, i]), None) for i in pairs])
        assert (got == g[f"{tag}_reward_{name}"]).all()          # float64 bit-exact, same order of adds
        tab, log2 = T.lower_reward(f, C, 3, 3)
        assert log2 == (name == "nonlinear") and tab.shape == (3, 3)


@pytest.mark.parametrize("tag,C", [("c3", 3), ("c2", 2)])
@pytest.mark.parametrize("difficulty", T.DIFFICULTIES)
def test_side_effect_tables(golden_pol, tag, C, difficulty):
    g = golden_pol
    se = T.side_effect_tables(C, 3, difficulty)
    nxt = g[f"{tag}_next"].astype(int)
    want = g[f"{tag}_se_{difficulty}"][:, 0, :]
    for j in range(C):
        p = 1 if j == 0 else j
        assert (se[j][nxt[:, 0], nxt[:, p]] == want[:, j]).all()
    assert T.SE_NAMES[want].dtype == np.dtype("<U6")


def test_noise_tables(golden_pol):
    g = golden_pol
    for mode, dl in (("rs", False), ("dl", True)):
        move, noisy, draws = T.noise_tables(3, 3, deadlock=dl)
        sa, u = g[f"noise_{mode}_sa"], g[f"noise_{mode}_u"]
        s, a = detab(sa[:, 0], 3, 3), detab(sa[:, 1], 3, 3)
        assert (draws[s, a].T.astype(bool) == ~np.isnan(u)).all()      # who draws
        fire = (np.nan_to_num(u, nan=1.0) < 0.1).T
        assert (np.where(fire, noisy[s, a], move[s, a]).T == g[f"noise_{mode}_next"]).all()


def test_bad_difficulty_message(golden_pol):
    with pytest.raises(ValueError) as e:
        T.side_effect_tables(3, 3, "nope")
    assert str(e.value) == str(golden_pol["c3_bad_difficulty_msg"])


def test_callable_lowering():
    tab, log2 = T.lower_reward(lambda s, a, n: float(np.log2(1 + sum(0.1 * x * y for x, y in zip(s, a)))), 3, 3, 3)
    assert log2 and np.allclose(tab, 0.1 * np.outer(np.arange(3), np.arange(3)))
    with pytest.raises(ValueError):
        T.lower_reward(lambda s, a, n: float(s[0] * s[1]), 3, 3, 3)


def test_reward_callable_is_bound_to_the_env_shape():
    """The host callable handed to agents must agree with the device table: with four levels, level 2 is an
    interior level (role 1), not the polarised one -- which cannot be guessed from one (state, action) pair."""
    f4 = T.right_polarizing.for_shape(4, 4)
    tab4 = T.right_polarizing.table(4, 4)
    assert f4((2,), (2,)) == tab4[2, 2] == 0.10            # role 1, action == level
    assert T.right_polarizing((2,), (2,)) == 0.25          # three levels (the reference's shape): level 2 is polarised
    assert f4((3,), (3,)) == tab4[3, 3] == 0.25
    assert T.right_polarizing.for_shape(3, 3) is T.right_polarizing
    for s in range(4):
        for a in range(4):
            assert f4((s, 0), (a, 0)) == tab4[s, a] + tab4[0, 0]


def test_callable_depending_on_next_state_is_rejected():
    with pytest.raises(ValueError, match="next_state"):
        T.lower_reward(lambda s, a, n: float(sum(n)), 3, 3, 3)
