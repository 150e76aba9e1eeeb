#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by EXECUTING the unmodified reference.

Run in the build container only (the reference lives at /root/reference there and
does not travel to the GPU box):

    python tests/golden/make_golden.py [--reference /root/reference]

The reference imports `gymnasium`, which is absent from the image; it uses it only
structurally (base class, space constructors, `register`), so it is imported under
the structural stand-in shipped in gym_cellular_b200/compat (appended to sys.path,
real gymnasium wins if present).  No reference source is copied: every number in the
.npz files is an OUTPUT of `env.reset()` / `env.step()` / `prior_knowledge.*` of the
reference classes.

Random draws: `numpy.random.rand` / `numpy.random.randint` are looked up by the
reference at call time through the module attribute, so they are monkey-patched
either with a replay queue (exhaustive enumerations) or with a pass-through
recorder around the real legacy MT19937 (trajectories).

Encoding conventions used in the fixtures (ours, fixed here):
  side-effect strings -> codes  silent=0, safe=1, unsafe=2
  polarisation (state, action) pair index p = tab(state) * A**C + tab(action), where
    tab() is the reference's own `prior_knowledge.tabularize`
  grid world states/actions are the reference's own `cellularize` codes.
"""
import argparse
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))

SE_CODE = {"silent": 0, "safe": 1, "unsafe": 2}


def se_codes(mat):
    return np.vectorize(lambda s: SE_CODE[str(s)])(np.asarray(mat)).astype(np.int8)


class Replay:
    """Queue-backed stand-in for np.random.rand / np.random.randint."""

    def __init__(self):
        self.u = []
        self.ints = []
        self.n_rand = 0
        self.n_randint_calls = 0

    def load(self, u=(), ints=()):
        self.u = list(u)
        self.ints = list(ints)
        self.n_rand = 0
        self.n_randint_calls = 0

    def rand(self, *shape):
        assert shape == ()
        self.n_rand += 1
        return self.u.pop(0)

    def randint(self, low, high=None, size=None, dtype=int):
        assert high is None
        self.n_randint_calls += 1
        if size is None:
            v = self.ints.pop(0)
            assert 0 <= v < low
            return v
        n = int(np.prod(size))
        vals = [self.ints.pop(0) for _ in range(n)]
        return np.array(vals, dtype=int).reshape(size)


class Recorder:
    """Pass-through recorder around the real legacy generator."""

    def __init__(self):
        self._rand = np.random.rand
        self._randint = np.random.randint
        self.log = []

    def rand(self, *shape):
        v = self._rand(*shape)
        self.log.append(("rand", float(v)))
        return v

    def randint(self, low, high=None, size=None, dtype=int):
        v = self._randint(low, high, size, dtype)
        self.log.append(("randint", np.array(v).reshape(-1).tolist()))
        return v

    def take(self):
        out, self.log = self.log, []
        return out


def patch(obj):
    np.random.rand = obj.rand
    np.random.randint = obj.randint


def unpatch(saved):
    np.random.rand, np.random.randint = saved


# --------------------------------------------------------------------------------------
def gen_polarisation_det(gym, out):
    import gym_cellular.envs.cells3states3actions3 as m3
    import gym_cellular.envs.cells2rest3 as m2
    for tag, mod, env_id, C in (("c3", m3, "gym_cellular/Cells3States3Actions3-v0", 3),
                                ("c2", m2, "gym_cellular/Cells2Rest3-v0", 2)):
        rfuncs = {"right_polarizing": mod.right_polarizing, "multiple_optima": mod.multiple_optima,
                  "nonlinear": mod.nonlinear}
        n_s = 3 ** C
        nxt_ref = None
        for diff in ("easy", "hard", "impossible"):
            for rname, rf in rfuncs.items():
                env = gym.make(env_id, difficulty=diff, reward_func=rf)
                pk = env.prior_knowledge
                state0, info0 = env.reset()
                out[f"{tag}_reset_state"] = np.array(state0, dtype=np.int8)
                out[f"{tag}_reset_se"] = se_codes(info0["side_effects"])
                nxt = np.zeros((n_s * n_s, C), np.int8)
                rew = np.zeros(n_s * n_s, np.float64)
                se = np.zeros((n_s * n_s, C, C), np.int8)
                inc = np.zeros(n_s * n_s, np.float64)
                tab_next = np.zeros(n_s * n_s, np.int64)
                for si in range(n_s):
                    s = pk.detabularize(si, pk.state_space)
                    assert pk.tabularize(s, pk.state_space) == si
                    for ai in range(n_s):
                        a = pk.detabularize(ai, pk.action_space)
                        env.reset()
                        env.state = s
                        ns, r, term, trunc, info = env.step(a)
                        assert term is False and trunc is False
                        p = si * n_s + ai
                        nxt[p] = ns
                        rew[p] = r
                        se[p] = se_codes(info["side_effects"])
                        inc[p] = env.get_data()["side_effects_incidence"]
                        assert env.get_data()["time_step"] == 1 and env.get_data()["reward"] == r
                        tab_next[p] = pk.tabularize(ns, pk.state_space)
                if nxt_ref is None:
                    nxt_ref = nxt
                    out[f"{tag}_next"] = nxt
                    out[f"{tag}_next_tab"] = tab_next
                    out[f"{tag}_incidence"] = inc
                assert (nxt == nxt_ref).all()
                k = f"{tag}_reward_{rname}"
                if k in out:
                    assert (out[k] == rew).all()
                out[k] = rew
                k = f"{tag}_se_{diff}"
                if k in out:
                    assert (out[k] == se).all()
                out[k] = se
        # codec tables + metadata
        env = gym.make(env_id)
        pk = env.prior_knowledge
        out[f"{tag}_detab"] = np.array([pk.detabularize(i, pk.state_space) for i in range(n_s)], np.int8)
        out[f"{tag}_meta"] = np.array([pk.n_cells, pk.n_states, pk.n_actions, pk.n_intracellular_states,
                                       pk.n_intracellular_actions], np.int64)
        out[f"{tag}_initial_policy"] = np.array(
            [pk.initial_policy(pk.detabularize(i, pk.state_space)) for i in range(n_s)], np.int8)
        try:
            gym.make(env_id, difficulty="nope").reset()
            e = gym.make(env_id, difficulty="nope")
            e.reset()
            e.step(tuple([0] * C))
            raise AssertionError("expected ValueError")
        except ValueError as err:
            out[f"{tag}_bad_difficulty_msg"] = np.array(str(err))


def kat_b1(gym, out):
    """Appendix-B style short trajectory of the deterministic 3-cell env."""
    env = gym.make("gym_cellular/Cells3States3Actions3-v0")
    env.reset()
    acts = [(1, 2, 0), (2, 2, 2), (2, 0, 1), (0, 0, 0)]
    st, rw, inc, se = [], [], [], []
    for a in acts:
        s, r, _, _, info = env.step(a)
        st.append(s)
        rw.append(r)
        inc.append(env.get_data()["side_effects_incidence"])
        se.append(se_codes(info["side_effects"]))
    out["kat_c3_actions"] = np.array(acts, np.int8)
    out["kat_c3_states"] = np.array(st, np.int8)
    out["kat_c3_rewards"] = np.array(rw, np.float64)
    out["kat_c3_incidence"] = np.array(inc, np.float64)
    out["kat_c3_se"] = np.array(se, np.int8)


def gen_polarisation_noise(gym, out):
    import gym_cellular.envs.cells3resetVdeadlock as mn
    saved = (np.random.rand, np.random.randint)
    rp = Replay()
    patch(rp)
    try:
        LO, HI = 0.05, 0.95
        pk = gym.make("gym_cellular/Cells3ResetVDeadlock-v0", env_seed=1).prior_knowledge
        n_s = 27
        # 1) discover which cells draw, empirically, in the non-deadlock mode
        env = gym.make("gym_cellular/Cells3ResetVDeadlock-v0", env_seed=1, reward_func=mn.right_polarizing)
        env.reset()
        drawers = {}
        for si in range(n_s):
            s = pk.detabularize(si, pk.state_space)
            for ai in range(n_s):
                a = pk.detabularize(ai, pk.action_space)
                rp.load(u=[HI] * 8)
                env.state = s
                base, *_ = env.step(a)
                k = rp.n_rand
                cells = []
                for i in range(k):
                    rp.load(u=[LO if j == i else HI for j in range(k)])
                    env.state = s
                    ns, *_ = env.step(a)
                    assert rp.n_rand == k
                    diff = [c for c in range(3) if ns[c] != base[c]]
                    assert len(diff) == 1 and ns[diff[0]] == base[diff[0]] - 1
                    cells.append(diff[0])
                assert cells == sorted(cells), "draws are in cell order"
                drawers[(si, ai)] = cells
        # 2) enumerate every low/high pattern for both modes and the default reward (nonlinear)
        for dl in (False, True):
            rows_sa, rows_u, rows_next, rows_rew_nl, rows_rew_rp, rows_se, rows_inc = [], [], [], [], [], [], []
            e_nl = gym.make("gym_cellular/Cells3ResetVDeadlock-v0", env_seed=1, deadlock=dl)
            e_rp = gym.make("gym_cellular/Cells3ResetVDeadlock-v0", env_seed=1, deadlock=dl,
                            reward_func=mn.right_polarizing)
            e_nl.reset()
            e_rp.reset()
            for (si, ai), cells in drawers.items():
                s = pk.detabularize(si, pk.state_space)
                a = pk.detabularize(ai, pk.action_space)
                for pat in itertools.product((LO, HI), repeat=len(cells)):
                    u = np.full(3, np.nan)
                    for c, v in zip(cells, pat):
                        u[c] = v
                    rp.load(u=list(pat))
                    e_nl.state = s
                    ns, r, _, _, info = e_nl.step(a)
                    assert rp.n_rand == len(cells) and not rp.u
                    rp.load(u=list(pat))
                    e_rp.state = s
                    ns2, r2, *_ = e_rp.step(a)
                    assert ns2 == ns
                    rows_sa.append((si, ai))
                    rows_u.append(u)
                    rows_next.append(ns)
                    rows_rew_nl.append(r)
                    rows_rew_rp.append(r2)
                    rows_se.append(se_codes(info["side_effects"]))
                    rows_inc.append(e_nl.get_data()["side_effects_incidence"])
            t = "dl" if dl else "rs"
            out[f"noise_{t}_sa"] = np.array(rows_sa, np.int16)
            out[f"noise_{t}_u"] = np.array(rows_u, np.float64)
            out[f"noise_{t}_next"] = np.array(rows_next, np.int8)
            out[f"noise_{t}_reward_nonlinear"] = np.array(rows_rew_nl, np.float64)
            out[f"noise_{t}_reward_right_polarizing"] = np.array(rows_rew_rp, np.float64)
            out[f"noise_{t}_se_easy"] = np.array(rows_se, np.int8)
            out[f"noise_{t}_incidence"] = np.array(rows_inc, np.float64)
    finally:
        unpatch(saved)


def gen_polarisation_noise_traj(gym, out):
    """Real MT19937 trajectories (draws recorded into per-cell slots)."""
    saved = (np.random.rand, np.random.randint)
    try:
        for dl in (False, True):
            for seed, T, aseed in ((12345, 40, 99), (7, 3000, 5)):
                rec = Recorder()
                env = gym.make("gym_cellular/Cells3ResetVDeadlock-v0", env_seed=seed, deadlock=dl)
                patch(rec)
                acts = np.random.default_rng(aseed).integers(0, 3, size=(T, 3))
                states = np.zeros((2, T, 3), np.int8)
                rewards = np.zeros((2, T), np.float64)
                us = np.full((2, T, 3), np.nan)
                for ep in range(2):           # reset() re-seeds: both episodes must be identical
                    s, _ = env.reset()
                    rec.take()
                    for t in range(T):
                        prev = s
                        s, r, *_ = env.step(tuple(int(x) for x in acts[t]))
                        draws = [v for k, v in rec.take() if k == "rand"]
                        cells = [c for c in range(3) if (prev[c] == 1 and acts[t][c] != 0) or prev[c] == 2]
                        assert len(cells) == len(draws)
                        for c, v in zip(cells, draws):
                            us[ep, t, c] = v
                        states[ep, t] = s
                        rewards[ep, t] = r
                assert (states[0] == states[1]).all() and np.array_equal(us[0], us[1], equal_nan=True)
                tag = f"traj_{'dl' if dl else 'rs'}_{seed}"
                out[tag + "_actions"] = acts.astype(np.int8)
                out[tag + "_states"] = states[0]
                out[tag + "_rewards"] = rewards[0]
                out[tag + "_u"] = us[0]
                unpatch(saved)
    finally:
        unpatch(saved)


# --------------------------------------------------------------------------------------
def gen_gridworld(gym, out):
    saved = (np.random.rand, np.random.randint)
    rp = Replay()
    try:
        env = gym.make("gym_cellular/GridWorld-v0")
        pk = env.prior_knowledge
        s0, info0 = env.reset()
        out["gw_reset_cell"] = np.array(pk.cellularize(s0, "state"), np.int8)
        out["gw_reset_tab"] = np.array(pk.tabularize(s0, "state"), np.int64)
        out["gw_reset_se"] = se_codes(info0["side_effects"])
        out["gw_meta"] = np.array([pk.n_cells, pk.n_states, pk.n_actions, pk.n_intracellular_states,
                                   pk.n_intracellular_actions], np.int64)
        # valid states: exactly one jurisdiction holds the agent
        states = []
        for J in range(2):
            for p in range(4):
                for T0 in range(4):
                    for T1 in range(4):
                        code = [T0 + 4 * (p if J == 0 else 4), T1 + 4 * (p if J == 1 else 4)]
                        states.append(code)
        actions = [(a0, a1) for a0 in range(5) for a1 in range(5) if not (a0 == 4 and a1 == 4)]
        # codec round trips through the reference's own codec
        for code in states:
            st = pk.decellularize(np.array(code), "state")
            assert list(pk.cellularize(st, "state")) == code
            tab = pk.tabularize(st, "state")
            assert tab == code[0] + 20 * code[1]
            assert list(pk.cellularize(pk.detabularize(tab, "state"), "state")) == code
        out["gw_states"] = np.array(states, np.int8)
        out["gw_actions"] = np.array(actions, np.int8)
        out["gw_action_tab"] = np.array(
            [pk.tabularize(pk.decellularize(np.array(a), "action"), "action") for a in actions], np.int64)
        # decode of each valid state: agent (jurisdiction, row, col) and the two 2x2 tree arrays
        dec = np.zeros((len(states), 3 + 8), np.int8)
        for i, code in enumerate(states):
            st = pk.decellularize(np.array(code), "state")
            J = 0 if "position" in st[0]["agt"] else 1
            dec[i, 0] = J
            dec[i, 1:3] = st[J]["agt"]["position"]
            dec[i, 3:7] = st[0]["living_trees"].reshape(-1)
            dec[i, 7:11] = st[1]["living_trees"].reshape(-1)
        out["gw_states_decoded"] = dec
        # initial policy over all valid states
        out["gw_initial_policy"] = np.array(
            [pk.cellularize(pk.initial_policy(pk.decellularize(np.array(c), "state")), "action") for c in states],
            np.int8)
        patch(rp)
        outcomes = [(0.5, None, None)]
        for k in range(2):
            for bits in itertools.product((0, 1), repeat=4):
                outcomes.append((0.005, bits, k))
        rows = {n: [] for n in ("s", "a", "u0", "bits", "k", "ndraw", "next", "tab", "rew", "inc", "se")}
        env.reset()
        for code in states:
            if code[0] % 4 == 0 and code[1] % 4 == 0:
                continue          # both barren: the reference draws nothing (gen_gridworld_barren)
            for a in actions:
                act = pk.decellularize(np.array(a), "action")
                for (u0, bits, k) in outcomes:
                    rp.load(u=[u0], ints=(list(bits) + [k]) if bits is not None else [])
                    env.state = pk.decellularize(np.array(code), "state")
                    ns, r, term, trunc, info = env.step(act)
                    assert term is False and trunc is False
                    consumed = rp.n_rand + (5 if rp.n_randint_calls else 0)
                    assert rp.n_rand == 1
                    assert (rp.n_randint_calls == 2) == (bits is not None)
                    rows["s"].append(code)
                    rows["a"].append(a)
                    rows["u0"].append(u0)
                    rows["bits"].append(bits if bits is not None else (0, 0, 0, 0))
                    rows["k"].append(k if k is not None else 0)
                    rows["ndraw"].append(consumed)
                    rows["next"].append(pk.cellularize(ns, "state"))
                    rows["tab"].append(pk.tabularize(ns, "state"))
                    rows["rew"].append(r)
                    rows["inc"].append(env.data["side_effects_incidence"])
                    rows["se"].append(se_codes(info["side_effects"]))
        out["gw_case_state"] = np.array(rows["s"], np.int8)
        out["gw_case_action"] = np.array(rows["a"], np.int8)
        out["gw_case_u0"] = np.array(rows["u0"], np.float64)
        out["gw_case_bits"] = np.array(rows["bits"], np.int8)
        out["gw_case_k"] = np.array(rows["k"], np.int8)
        out["gw_case_ndraw"] = np.array(rows["ndraw"], np.int8)
        out["gw_case_next"] = np.array(rows["next"], np.int8)
        out["gw_case_tab"] = np.array(rows["tab"], np.int64)
        out["gw_case_reward"] = np.array(rows["rew"], np.float64)
        out["gw_case_incidence"] = np.array(rows["inc"], np.float64)
        out["gw_case_se"] = np.array(rows["se"], np.int8)
        # (4,4): KeyError
        try:
            rp.load(u=[0.5])
            env.reset()
            env.step(pk.decellularize(np.array([4, 4]), "action"))
            raise AssertionError("expected KeyError")
        except KeyError as err:
            out["gw_noaction_error"] = np.array(repr(err))
    finally:
        unpatch(saved)


def gen_gridworld_barren(gym, out):
    """Both-barren states consume no draw: enumerate them separately (8 states x 24 actions)."""
    saved = (np.random.rand, np.random.randint)
    rp = Replay()
    try:
        env = gym.make("gym_cellular/GridWorld-v0")
        pk = env.prior_knowledge
        env.reset()
        patch(rp)
        S, A, NX, TAB, RW, INC, SE = [], [], [], [], [], [], []
        for J in range(2):
            for p in range(4):
                code = [4 * (p if J == 0 else 4), 4 * (p if J == 1 else 4)]
                for a0 in range(5):
                    for a1 in range(5):
                        if a0 == 4 and a1 == 4:
                            continue
                        rp.load()
                        env.state = pk.decellularize(np.array(code), "state")
                        ns, r, _, _, info = env.step(pk.decellularize(np.array([a0, a1]), "action"))
                        assert rp.n_rand == 0 and rp.n_randint_calls == 0
                        S.append(code)
                        A.append((a0, a1))
                        NX.append(pk.cellularize(ns, "state"))
                        TAB.append(pk.tabularize(ns, "state"))
                        RW.append(r)
                        INC.append(env.data["side_effects_incidence"])
                        SE.append(se_codes(info["side_effects"]))
        out["gw_barren_state"] = np.array(S, np.int8)
        out["gw_barren_action"] = np.array(A, np.int8)
        out["gw_barren_next"] = np.array(NX, np.int8)
        out["gw_barren_tab"] = np.array(TAB, np.int64)
        out["gw_barren_reward"] = np.array(RW, np.float64)
        out["gw_barren_incidence"] = np.array(INC, np.float64)
        out["gw_barren_se"] = np.array(SE, np.int8)
    finally:
        unpatch(saved)


def gen_gridworld_traj(gym, out):
    """Real MT19937 trajectory; draws recorded per step as (u0, b0..b3, k), NaN/-1 when not drawn.

    Random play reaches the absorbing all-barren state after ~50 steps, so the trajectory is cut
    into episodes of EP_LEN steps with `env.reset()` in between (reset does not touch the RNG)."""
    saved = (np.random.rand, np.random.randint)
    try:
        EP_LEN, T = 64, 64 * 150
        env = gym.make("gym_cellular/GridWorld-v0")
        pk = env.prior_knowledge
        arng = np.random.default_rng(2024)
        jur = arng.integers(0, 2, size=T)
        pos = arng.integers(0, 4, size=T)
        acts = np.full((T, 2), 4, np.int8)
        acts[np.arange(T), jur] = pos
        np.random.seed(4242)
        rec = Recorder()
        patch(rec)
        s, _ = env.reset()
        rec.take()
        cells = np.zeros((T, 2), np.int8)
        tabs = np.zeros(T, np.int64)
        rews = np.zeros(T, np.float64)
        incs = np.zeros(T, np.float64)
        ses = np.zeros((T, 2, 2), np.int8)
        u0 = np.full(T, np.nan)
        bits = np.full((T, 4), -1, np.int8)
        kk = np.full(T, -1, np.int8)
        for t in range(T):
            if t and t % EP_LEN == 0:
                env.reset()
                assert not rec.take()
            s, r, _, _, info = env.step(pk.decellularize(acts[t].astype(int), "action"))
            log = rec.take()
            if log:
                assert log[0][0] == "rand"
                u0[t] = log[0][1]
                if len(log) > 1:
                    assert len(log) == 3 and len(log[1][1]) == 4 and len(log[2][1]) == 1
                    bits[t] = log[1][1]
                    kk[t] = log[2][1][0]
            cells[t] = pk.cellularize(s, "state")
            tabs[t] = pk.tabularize(s, "state")
            rews[t] = r
            incs[t] = env.data["side_effects_incidence"]
            ses[t] = se_codes(info["side_effects"])
        assert (kk >= 0).sum() >= 10, "trajectory should contain dispersal events"
        out["gwtraj_ep_len"] = np.array(EP_LEN, np.int64)
        out["gwtraj_actions"] = acts
        out["gwtraj_cells"] = cells
        out["gwtraj_tab"] = tabs
        out["gwtraj_reward"] = rews
        out["gwtraj_incidence"] = incs
        out["gwtraj_se"] = ses
        out["gwtraj_u0"] = u0
        out["gwtraj_bits"] = bits
        out["gwtraj_k"] = kk
    finally:
        unpatch(saved)


# --------------------------------------------------------------------------------------
def gen_debug(gym, out):
    """The three debug MDPs: exhaustive (state, action) tables."""
    saved = (np.random.rand, np.random.randint)
    rp = Replay()
    try:
        for tag, env_id, S, A in (("dbg", "gym_cellular/Debug-v0", 2, 2),
                                  ("dplan", "gym_cellular/DeepPlanningDebug-v0", 4, 2)):
            env = gym.make(env_id)
            pk = env.prior_knowledge
            s0, i0 = env.reset()
            out[f"{tag}_reset_se"] = se_codes(i0["side_effects"])
            nS, nA = S * S, A * A
            nxt = np.zeros((nS * nA, 2), np.int8)
            rew = np.zeros(nS * nA)
            se = np.zeros((nS * nA, 2, 2), np.int8)
            inc = np.zeros(nS * nA)
            for si in range(nS):
                for ai in range(nA):
                    env.reset()
                    env.state = pk.detabularize(si, pk.state_space)
                    ns, r, _, _, info = env.step(pk.detabularize(ai, pk.action_space))
                    p = si * nA + ai
                    nxt[p], rew[p], se[p] = ns, r, se_codes(info["side_effects"])
                    inc[p] = env.get_data()["side_effects_incidence"]
            out[f"{tag}_next"], out[f"{tag}_reward"], out[f"{tag}_se"], out[f"{tag}_incidence"] = nxt, rew, se, inc
        # DeepExploration: draws iff a==0, one per such cell, in cell order
        env = gym.make("gym_cellular/DeepExplorationDebug-v0")
        pk = env.prior_knowledge
        s0, i0 = env.reset()
        out["dexp_reset_se"] = se_codes(i0["side_effects"])
        patch(rp)
        SA, U, NX, RW = [], [], [], []
        for si in range(16):
            for ai in range(4):
                s = pk.detabularize(si, pk.state_space)
                a = pk.detabularize(ai, pk.action_space)
                cells = [c for c in range(2) if a[c] == 0]
                for pat in itertools.product((0.25, 0.75), repeat=len(cells)):
                    rp.load(u=list(pat))
                    env.state = s
                    ns, r, *_ = env.step(a)
                    assert rp.n_rand == len(cells)
                    u = np.full(2, np.nan)
                    for c, v in zip(cells, pat):
                        u[c] = v
                    SA.append((si, ai))
                    U.append(u)
                    NX.append(ns)
                    RW.append(r)
        out["dexp_sa"] = np.array(SA, np.int16)
        out["dexp_u"] = np.array(U)
        out["dexp_next"] = np.array(NX, np.int8)
        out["dexp_reward"] = np.array(RW)
    finally:
        unpatch(saved)


def gen_codec(out):
    from gym_cellular.envs.utils import (generalized_cellular2tabular, generalized_tabular2cellular,
                                         cellular2tabular, tabular2cellular)
    rng = np.random.default_rng(0)
    # ragged mixed radix with non-zero minima
    spaces = [range(2, 5), range(0, 7), range(-3, 0), range(10, 12), range(0, 1)]
    n = int(np.prod([len(s) for s in spaces]))
    cells = np.zeros((n, len(spaces)), np.int64)
    for i in range(n):
        c = generalized_tabular2cellular(i, spaces)
        assert generalized_cellular2tabular(c, spaces) == i
        cells[i] = c
    out["codec_ragged_lens"] = np.array([len(s) for s in spaces], np.int64)
    out["codec_ragged_mins"] = np.array([min(s) for s in spaces], np.int64)
    out["codec_ragged_cells"] = cells
    # 16 cells x 4 states: the 32-bit case (unsigned)
    sp16 = [range(0, 4)] * 16
    samples = rng.integers(0, 4, size=(4096, 16))
    samples[0] = 3
    samples[1] = 0
    tabs = np.array([generalized_cellular2tabular(list(map(int, r)), sp16) for r in samples], np.uint64)
    assert tabs[0] == 2 ** 32 - 1
    for r, t in zip(samples[:64], tabs[:64]):
        assert generalized_tabular2cellular(int(t), sp16) == list(r)
    out["codec_c16_cells"] = samples.astype(np.int8)
    out["codec_c16_tab"] = tabs
    # fixed-radix variant
    s3 = rng.integers(0, 3, size=(256, 5))
    out["codec_fixed_cells"] = s3.astype(np.int8)
    out["codec_fixed_tab"] = np.array([cellular2tabular(list(map(int, r)), 3, 5) for r in s3], np.int64)
    for r, t in zip(s3, out["codec_fixed_tab"]):
        assert (tabular2cellular(int(t), 3, 5) == r).all()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    try:
        import gymnasium as gym
    except ImportError:
        sys.path.append(os.path.join(REPO, "gym_cellular_b200", "compat"))
        import gymnasium as gym
    import gym_cellular  # noqa: F401  (the reference: registers its ids)
    assert os.path.realpath(gym_cellular.__file__).startswith(os.path.realpath(args.reference))

    pol, gw, dbg = {}, {}, {}
    gen_polarisation_det(gym, pol)
    kat_b1(gym, pol)
    gen_polarisation_noise(gym, pol)
    gen_polarisation_noise_traj(gym, pol)
    gen_codec(pol)
    gen_gridworld(gym, gw)
    gen_gridworld_barren(gym, gw)
    gen_gridworld_traj(gym, gw)
    gen_debug(gym, dbg)
    for name, d in (("polarisation", pol), ("gridworld", gw), ("debug", dbg)):
        path = os.path.join(args.out, f"{name}.npz")
        np.savez_compressed(path, **d)
        print(f"{path}: {len(d)} arrays, {os.path.getsize(path)} bytes")
    import platform
    with open(os.path.join(args.out, "PROVENANCE.txt"), "w") as f:
        f.write("generated by tests/golden/make_golden.py from the unmodified reference at "
                f"{args.reference}\npython {platform.python_version()} numpy {np.__version__} "
                f"gymnasium {'stand-in' if getattr(gym, 'IS_COMPAT_STANDIN', False) else gym.__version__}\n")
        for name, d in (("polarisation", pol), ("gridworld", gw), ("debug", dbg)):
            for k, v in d.items():
                f.write(f"{name}.npz:{k} shape={getattr(v, 'shape', ())} dtype={getattr(v, 'dtype', type(v))}\n")


if __name__ == "__main__":
    main()
