"""The gymnasium surface (ids, spaces, prior_knowledge, codec) without a GPU: the single-env classes
create their device env lazily, so everything up to reset()/step() is host-side."""
import numpy as np
import pytest

import gym_cellular_b200 as B
from gym_cellular_b200._gym import gym


def test_registered_ids_and_vector_entry_points():
    for env_id in ("gym_cellular/Cells3States3Actions3-v0", "gym_cellular/Cells2Rest3-v0",
                   "gym_cellular/Cells3ResetVDeadlock-v0", "gym_cellular/GridWorld-v0"):
        spec = gym.spec(env_id)
        assert spec.max_episode_steps is None                     # gym_cellular/__init__.py:7
        assert spec.vector_entry_point is not None
        env = gym.make(env_id)
        assert hasattr(env.unwrapped, "prior_knowledge")


@pytest.mark.parametrize("env_id,tag,C", [("gym_cellular/Cells3States3Actions3-v0", "c3", 3),
                                          ("gym_cellular/Cells2Rest3-v0", "c2", 2)])
def test_polarisation_prior_knowledge(golden_pol, env_id, tag, C):
    env = gym.make(env_id).unwrapped
    pk = env.prior_knowledge
    assert [pk.n_cells, pk.n_states, pk.n_actions, pk.n_intracellular_states,
            pk.n_intracellular_actions] == golden_pol[f"{tag}_meta"].tolist()
    assert pk.initial_state == tuple(golden_pol[f"{tag}_reset_state"]) and pk.confidence_level == 0.95
    assert pk.identical_intracellular_transitions is True and pk.reward_range == (0, 1)
    for i, cells in enumerate(golden_pol[f"{tag}_detab"]):
        assert pk.detabularize(i, pk.state_space) == tuple(cells)
        assert pk.tabularize(tuple(cells), pk.state_space) == i
        assert pk.initial_policy(tuple(cells)) == tuple(golden_pol[f"{tag}_initial_policy"][i])
    assert len(env.action_space) == C and env.observation_space is env.state_space
    assert gym.make(env_id, confidence_level=0.9).unwrapped.prior_knowledge.confidence_level == 0.9


def test_gridworld_prior_knowledge_and_codec(golden_gw):
    g = golden_gw
    env = gym.make("gym_cellular/GridWorld-v0").unwrapped
    pk = env.prior_knowledge
    assert [pk.n_cells, pk.n_states, pk.n_actions, pk.n_intracellular_states, pk.n_intracellular_actions] == g["gw_meta"].tolist()
    assert (pk.cellularize(pk.initial_state, "state") == g["gw_reset_cell"]).all()
    assert pk.tabularize(pk.initial_state, "state") == g["gw_reset_tab"]
    for i, (code, dec) in enumerate(zip(g["gw_states"].astype(int), g["gw_states_decoded"])):
        st = pk.decellularize(code, "state")
        J = 0 if "position" in st[0]["agt"] else 1
        assert J == dec[0] and (st[J]["agt"]["position"] == dec[1:3]).all() and "position" not in st[1 - J]["agt"]
        assert (st[0]["living_trees"].reshape(-1) == dec[3:7]).all() and (st[1]["living_trees"].reshape(-1) == dec[7:11]).all()
        assert (pk.cellularize(st, "state") == code).all()
        tab = pk.tabularize(st, "state")
        assert tab == code[0] + 20 * code[1] and (pk.cellularize(pk.detabularize(tab, "state"), "state") == code).all()
        assert (pk.cellularize(pk.initial_policy(st), "action") == g["gw_initial_policy"][i]).all()
    for a, t in zip(g["gw_actions"].astype(int), g["gw_action_tab"]):
        act = pk.decellularize(a, "action")
        assert pk.tabularize(act, "action") == t and (pk.cellularize(pk.detabularize(int(t), "action"), "action") == a).all()
    for _ in range(20):                                            # samplers: exactly one jurisdiction is named
        assert sum("position" in j["go_to"] for j in env.action_space.sample()) == 1
        assert sum("position" in j["agt"] for j in env.state_space.sample()) == 1
    with pytest.raises(ValueError):
        pk.cellularize(pk.initial_state, "nope")


def test_host_codec(golden_pol):
    g = golden_pol
    lens, mins, cells = g["codec_ragged_lens"], g["codec_ragged_mins"], g["codec_ragged_cells"]
    spaces = [range(int(m), int(m) + int(l)) for m, l in zip(mins, lens)]
    for i in range(0, len(cells), 5):
        assert B.generalized_cellular2tabular(list(cells[i]), spaces) == i
        assert B.generalized_tabular2cellular(i, spaces) == list(cells[i])
    sp16 = [range(0, 4)] * 16
    for row, tab in zip(g["codec_c16_cells"][:200], g["codec_c16_tab"][:200]):
        assert B.generalized_cellular2tabular([int(x) for x in row], sp16) == int(tab)
    for row, tab in zip(g["codec_fixed_cells"], g["codec_fixed_tab"]):
        assert B.cellular2tabular([int(x) for x in row], 3, 5) == tab
        assert (B.tabular2cellular(int(tab), 3, 5) == row).all()


def test_reward_callables_match_reference(golden_pol):
    assert B.right_polarizing((1, 1, 1), (2, 2, 2), None) == 0.8999999999999999
    assert B.nonlinear((1, 1, 1), (2, 2, 2), None) == 0.8073549220576041


def test_install_alias_keeps_reference_imports_working():
    import subprocess
    import sys
    code = (
        "import gym_cellular_b200 as b; b.install_alias()\n"
        "import gym_cellular\n"
        "from gym_cellular.envs import Cells3States3Actions3Env, GridWorldEnv, generalized_cellular2tabular\n"
        "from gym_cellular.envs.cells3states3actions3 import right_polarizing, nonlinear\n"
        "from gym_cellular.envs.cells3resetVdeadlock import nonlinear as nl_rp\n"
        "from gym_cellular.envs.utils.generalized_space_transformations import generalized_tabular2cellular\n"
        "from gym_cellular.envs.grid_world import PriorKnowledge\n"
        "from gym_cellular.envs.debug import DeepPlanningDebugEnv\n"
        "import gymnasium\n"
        "e = gymnasium.make('gym_cellular/Cells2Rest3-v0')\n"
        "assert right_polarizing((1, 1, 1), (2, 2, 2), None) == 0.8999999999999999\n"
        "assert abs(nl_rp((1, 1, 1), (2, 2, 2), None) - 0.9259994185562231) < 1e-12\n"
        "assert generalized_tabular2cellular(5, [range(3)] * 3) == [2, 1, 0] and PriorKnowledge().n_states == 400\n"
        "print('alias ok')\n")
    env = dict(__import__("os").environ, PYTHONPATH=__import__("conftest").REPO)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0 and "alias ok" in out.stdout, out.stderr[-800:]


def test_async_vector_env_stand_in():
    """The AsyncVectorEnv look-alike (worker processes, one pipe round trip per step) used as the CPU
    baseline harness: batched results equal stepping the same envs in-process."""
    from gym_cellular_b200._gym import gym
    from oracle.pyport import PolarisationEnv
    n, k = 6, 2
    fns = [lambda: PolarisationEnv(3, 3) for _ in range(n)]
    kw = {"envs_per_worker": k} if gym.vector.AsyncVectorEnv.__module__.startswith("gymnasium.vector.async_vector_env") and \
        "compat" in gym.__file__ else {}
    venv = gym.vector.AsyncVectorEnv(fns, shared_memory=False, **kw)
    try:
        if kw:
            assert venv.num_workers == n // k
        obs, _ = venv.reset()
        assert len(obs) == 3 and all((o == 0).all() for o in obs)
        local = [PolarisationEnv(3, 3) for _ in range(n)]
        for e in local:
            e.reset()
        rng = np.random.default_rng(1)
        for _ in range(12):
            acts = tuple(rng.integers(0, 3, n) for _ in range(3))
            obs, rew, term, trunc, infos = venv.step(acts)
            want = [e.step(tuple(int(a[i]) for a in acts)) for i, e in enumerate(local)]
            assert (np.stack(obs).T == np.array([w[0] for w in want])).all()
            assert (rew == np.array([w[1] for w in want])).all() and not term.any() and not trunc.any()
            assert (infos["per_env"][3]["side_effects"] == want[3][4]["side_effects"]).all()
    finally:
        venv.close()
