"""GPU parity tests of the packed layout (gc_step_packed / gc_step_host_packed through ctypes):
against the CPU oracle on seeded rollouts, against the golden vectors of the reference, and bit for bit
(reward bits included) against the int8 layout of the same library."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REWARD_RTOL = 1e-6      # north_star: "rewards within 1e-6 relative"
REWARD_ATOL = 1e-7


@pytest.fixture(scope="module")
def B():
    import gym_cellular_b200 as pkg
    assert torch.cuda.is_available()
    return pkg


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def pack(cells):
    from gym_cellular_b200.packed_env import pack_host
    return pack_host(cells)


def u32(t):
    return host(t).view(np.uint32)


def assert_matches_oracle(env, ora, final=None):
    n = env.num_envs
    assert (u32(env.packed_state) == pack(ora.state)).all()
    assert (host(env.state) == ora.state).all()
    assert (host(env.tabular_state()) == ora.index).all()
    assert (host(env.time_step) == ora.t).all()
    f = host(env._flags[:n])
    assert ((f & 1) == ora.unsafe).all()
    assert (((f >> 1) & 1) == ora.truncated).all()
    assert ((f >> 2) == ora.count).all()
    np.testing.assert_allclose(host(env._reward[:n]), ora.reward, rtol=REWARD_RTOL, atol=REWARD_ATOL)
    if final is not None:
        assert (u32(env._final[:n]) == pack(final)).all()


def _rollout(env, ora, T, rng, check_every=1, words=True):
    n, C = env.num_envs, env.n_cells
    for t in range(T):
        a = rng.integers(0, env.n_actions, size=(C, n)).astype(np.int8)
        if words and t % 2 == 0:
            w = torch.zeros(env.ld, dtype=torch.int32, device="cuda")
            w[:n] = dev(pack(a).view(np.int32))
            env.step_device(w)
        else:
            env.step_device(dev(a))                       # per-cell levels: packed by the device kernel
        ora.step(a)
        if (t + 1) % check_every == 0 or t == T - 1:
            assert_matches_oracle(env, ora)


@pytest.mark.parametrize("C", list(range(1, 17)))
def test_every_cell_count_vs_oracle(B, O, C):
    """Every instantiation of the packed kernel (1..16 cells), deterministic and with Philox noise, 2-4
    levels (separate index for S < 4), with the fused auto-reset, ragged batch size."""
    n = 3001
    for stochastic in (False, True):
        S = (4, 3, 2)[C % 3]
        difficulty = "hard" if (C >= 3 and C % 3) else "easy"
        env = B.PackedCellularVectorEnv(num_envs=n, n_cells=C, n_states=S, stochastic=stochastic, env_seed=C,
                                        max_episode_steps=5, difficulty=difficulty)
        ora = O.OracleEnv(n_envs=n, n_cells=C, n_states=S, noise=stochastic, rng_episodic=True, seed=C, max_episode_steps=5,
                          difficulty=difficulty, reward="nonlinear_rp" if stochastic else "right_polarizing")
        _rollout(env, ora, 8, np.random.default_rng(100 + C), check_every=4)
        s = env.stats()
        assert s["env_steps"] == 8 * n == ora.stats[0] and s["unsafe_steps"] == ora.stats[1]
        assert s["count_sum"] == ora.stats[2] and s["episodes_truncated"] == ora.stats[3]


@pytest.mark.parametrize("tag,C", [("c3", 3), ("c2", 2)])
@pytest.mark.parametrize("difficulty", ["easy", "hard", "impossible"])
@pytest.mark.parametrize("reward", ["right_polarizing", "multiple_optima", "nonlinear"])
def test_exhaustive_vs_reference(B, golden_pol, tag, C, difficulty, reward):
    """All 729 / 81 (state, action) pairs of the reference's 3- and 2-cell envs, against the outputs of the
    unmodified reference (tests/golden/polarisation.npz)."""
    from conftest import detab
    g = golden_pol
    n_s = 3 ** C
    pairs = np.arange(n_s * n_s)
    env = B.PackedCellularVectorEnv(num_envs=len(pairs), n_cells=C, difficulty=difficulty,
                                    reward_func=getattr(B.tables, reward), emit_side_effects=True)
    env.set_state(dev(detab(pairs // n_s, C, 3)))
    obs, rew, term, trunc, info = env.step(dev(detab(pairs % n_s, C, 3)))
    assert (host(env.state).T == g[f"{tag}_next"]).all()
    assert (u32(obs) == pack(g[f"{tag}_next"].T)).all()
    assert (host(env.tabular_state()) == g[f"{tag}_next_tab"]).all()
    np.testing.assert_allclose(host(rew), g[f"{tag}_reward_{reward}"], rtol=REWARD_RTOL, atol=REWARD_ATOL)
    se = g[f"{tag}_se_{difficulty}"]
    assert (host(env.side_effects_row()).T == se[:, 0, :]).all()
    assert (host(info["unsafe"]) == (se == 2).any(axis=(1, 2))).all()
    np.testing.assert_allclose(host(env.side_effects_incidence()), g[f"{tag}_incidence"], rtol=1e-6)
    assert not host(term).any() and not host(trunc).any() and (host(info["time_step"]) == 1).all()


@pytest.mark.parametrize("stochastic", [False, True])
@pytest.mark.parametrize("C,S", [(16, 4), (3, 3), (2, 3), (7, 2), (13, 4)])
def test_packed_equals_int8_bit_for_bit(B, C, S, stochastic):
    """The two layouts of the same library: identical states, indices, flags and reward BITS."""
    n = 50021
    kw = dict(num_envs=n, n_cells=C, n_states=S, stochastic=stochastic, env_seed=3, max_episode_steps=6,
              env_id_offset=4 * 10 ** 9)
    pk = B.PackedCellularVectorEnv(emit_side_effects=True, emit_final_obs=True, **kw)
    i8 = B.CellularVectorEnv(**kw)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in range(9):
        a = torch.randint(0, S, (C, n), dtype=torch.int8, device="cuda", generator=gen)
        pk.step_device(a)
        i8.step_device(a)
        assert torch.equal(pk.state, i8.state)
        assert torch.equal(pk.tabular_state(), i8.tabular_state())
        assert torch.equal(pk._reward[:n].view(torch.int32), i8._reward[:n].view(torch.int32))
        f = pk._flags[:n]
        assert torch.equal(f & 1, i8._unsafe[:n]) and torch.equal((f >> 1) & 1, i8._truncated[:n])
        assert torch.equal(f >> 2, i8._count[:n])
        assert torch.equal(pk.side_effects_row(), i8._se_row[:, :n])
        assert torch.equal(pk.time_step, i8.time_step)
    assert pk.stats() == i8.stats()


def test_config4_full_size_packed_equals_int8(B):
    """BASELINE config 4 at its full size (2^24 envs, 16 cells x 4 levels): the packed layout (replicated
    table variant) against the int8 layout, bit for bit, plus the size-independent properties."""
    n = 1 << 24
    pk = B.PackedCellularVectorEnv(num_envs=n, n_cells=16, n_states=4, env_seed=1)
    i8 = B.CellularVectorEnv(num_envs=n, n_cells=16, n_states=4, env_seed=1, emit_side_effects=False)
    gen = torch.Generator(device="cuda").manual_seed(6)
    for t in range(3):
        a = torch.randint(0, 4, (16, n), dtype=torch.int8, device="cuda", generator=gen)
        words = pk.pack(a)
        assert torch.equal(pk.unpack(words), a)                     # pack / unpack round trip
        pk.step_device(words)
        i8.step_device(a)
        # four levels: the state word is the tabular index
        assert torch.equal(pk.packed_state, i8._index[:n])
        assert torch.equal(pk._reward[:n].view(torch.int32), i8._reward[:n].view(torch.int32))
        f = pk._flags[:n]
        assert torch.equal(f & 1, i8._unsafe[:n]) and torch.equal(f >> 2, i8._count[:n])
    assert pk.stats() == i8.stats()
    assert pk.stats()["env_steps"] == 3 * n


@pytest.mark.parametrize("stochastic", [False, True])
def test_big_launch_variant_vs_oracle(B, O, stochastic):
    """Launches of >= 2^21 envs use the replicated-table variant: against the oracle (all host threads)."""
    import os
    n = (1 << 21) + 48
    threads = os.cpu_count() or 1
    env = B.PackedCellularVectorEnv(num_envs=n, n_cells=16, n_states=4, stochastic=stochastic, env_seed=8,
                                    max_episode_steps=3)
    ora = O.OracleEnv(n_envs=n, n_cells=16, n_states=4, noise=stochastic, rng_episodic=True, seed=8, max_episode_steps=3,
                      reward="nonlinear_rp" if stochastic else "right_polarizing")
    rng = np.random.default_rng(12)
    for t in range(4):
        a = rng.integers(0, 4, size=(16, n)).astype(np.int8)
        env.step_device(dev(a))
        ora.step_parallel(a, threads)
        assert (u32(env.packed_state) == ora.index).all() and (host(env.time_step) == ora.t).all()
        f = host(env._flags[:n])
        assert ((f & 1) == ora.unsafe).all() and ((f >> 2) == ora.count).all() and (((f >> 1) & 1) == ora.truncated).all()
        np.testing.assert_allclose(host(env._reward[:n]), ora.reward, rtol=REWARD_RTOL, atol=REWARD_ATOL)


def test_final_observation_and_side_effects(B, O):
    """`final_obs` = the oracle's next state BEFORE the time-limit auto-reset, for every truncated env;
    the packed row-0 side-effect codes against the oracle's."""
    n = 4099
    env = B.PackedCellularVectorEnv(num_envs=n, n_cells=5, n_states=3, stochastic=True, env_seed=2, max_episode_steps=4,
                                    emit_final_obs=True, emit_side_effects=True, difficulty="hard")
    ora = O.OracleEnv(n_envs=n, n_cells=5, n_states=3, noise=True, rng_episodic=True, seed=2, difficulty="hard",
                      reward="nonlinear_rp")                          # no time limit: its state is the pre-reset one
    lim = O.OracleEnv(n_envs=n, n_cells=5, n_states=3, noise=True, rng_episodic=True, seed=2, difficulty="hard",
                      reward="nonlinear_rp", max_episode_steps=4)
    rng = np.random.default_rng(3)
    for t in range(9):
        a = rng.integers(0, 3, size=(5, n)).astype(np.int8)
        obs, rew, term, trunc, infos = env.step(dev(a))
        # the unlimited oracle is re-synchronised to the limited one's state before every step
        ora.state[:] = lim.state
        ora.t[:] = lim.t
        ora.step(a)
        lim.step(a)
        assert_matches_oracle(env, lim, final=ora.state)
        assert (host(env.side_effects_row()) == lim.se_row).all()
        assert (host(trunc) == lim.truncated.astype(bool)).all() and not host(term).any()
        assert (host(infos["_final_obs"]) == lim.truncated.astype(bool)).all()
        assert (host(infos["unsafe"]) == lim.unsafe.astype(bool)).all() and (host(infos["count"]) == lim.count).all()


def test_host_path_packed(B, O):
    """gc_step_host_packed (numpy words in, numpy words / reward / flags out; chunked pipeline)."""
    n = 100003
    env = B.PackedCellularVectorEnv(num_envs=n, stochastic=True, env_seed=4, max_episode_steps=9, host_chunk_envs=16384)
    ora = O.OracleEnv(n_envs=n, noise=True, rng_episodic=True, seed=4, max_episode_steps=9, reward="nonlinear_rp")
    rng = np.random.default_rng(4)
    for t in range(12):
        a = rng.integers(0, 3, size=(3, n)).astype(np.int8)
        if t % 2:
            obs, rew, term, trunc, info = env.step(pack(a))
        else:
            obs, rew, term, trunc, info = env.step(a)
        ora.step(a)
        assert (obs == pack(ora.state)).all() and (info["tabular_state"] == ora.index).all()
        assert (info["observation"] == ora.state).all()
        assert (trunc == ora.truncated.astype(bool)).all() and not term.any()
        assert (info["unsafe"] == ora.unsafe.astype(bool)).all() and (info["count"] == ora.count).all()
        np.testing.assert_allclose(rew, ora.reward, rtol=REWARD_RTOL, atol=REWARD_ATOL)
    assert_matches_oracle(env, ora)
    # 16 x 4: the state word is the index; non-episodic noise exercises the device-resident step counter
    # that every chunk reads
    env = B.PackedCellularVectorEnv(num_envs=n, n_cells=16, n_states=4, stochastic=True, rng_episodic=False, env_seed=6,
                                    host_chunk_envs=50000)
    ora = O.OracleEnv(n_envs=n, n_cells=16, n_states=4, noise=True, rng_episodic=False, seed=6, reward="nonlinear_rp")
    assert env.host_bytes_per_env_step == (4, 9)
    for t in range(5):
        a = rng.integers(0, 4, size=(16, n)).astype(np.int8)
        obs, rew, term, trunc, info = env.step(pack(a))
        ora.step(a)
        assert (obs == ora.index).all() and (info["tabular_state"] == ora.index).all()
        assert (info["unsafe"] == ora.unsafe.astype(bool)).all()
        np.testing.assert_allclose(rew, ora.reward, rtol=REWARD_RTOL, atol=REWARD_ATOL)
    # a device-path step after host-path steps continues the same noise sequence
    a = rng.integers(0, 4, size=(16, n)).astype(np.int8)
    env.step_device(dev(a))
    ora.step(a)
    assert_matches_oracle(env, ora)


def test_host_path_with_final_obs_and_side_effect_rows(B, O):
    """A host-path step of an env that emits final observations / side-effect rows (the un-pipelined branch)."""
    n = 5003
    env = B.PackedCellularVectorEnv(num_envs=n, n_cells=5, n_states=3, stochastic=True, env_seed=8, max_episode_steps=3,
                                    emit_final_obs=True, emit_side_effects=True)
    lim = O.OracleEnv(n_envs=n, n_cells=5, n_states=3, noise=True, rng_episodic=True, seed=8, max_episode_steps=3,
                      reward="nonlinear_rp")
    ora = O.OracleEnv(n_envs=n, n_cells=5, n_states=3, noise=True, rng_episodic=True, seed=8, reward="nonlinear_rp")
    rng = np.random.default_rng(8)
    for t in range(7):
        a = rng.integers(0, 3, size=(5, n)).astype(np.int8)
        obs, rew, term, trunc, info = env.step(a if t % 2 else pack(a))
        ora.state[:] = lim.state
        ora.t[:] = lim.t
        ora.step(a)
        lim.step(a)
        assert isinstance(obs, np.ndarray) and (obs == pack(lim.state)).all()
        assert (info["final_obs"] == pack(ora.state)).all() and (info["_final_obs"] == lim.truncated.astype(bool)).all()
        assert (info["side_effects_packed"] == pack(lim.se_row)).all() and (host(env.side_effects_row()) == lim.se_row).all()
        assert (info["unsafe"] == lim.unsafe.astype(bool)).all() and (info["count"] == lim.count).all()
        assert (trunc == lim.truncated.astype(bool)).all() and not term.any()
        np.testing.assert_allclose(rew, lim.reward, rtol=REWARD_RTOL, atol=REWARD_ATOL)


@pytest.mark.parametrize("C,S,stochastic", [(3, 3, True), (16, 4, True), (7, 2, False)])
def test_rollout_on_packed_state(B, O, C, S, stochastic):
    """`rollout` of the packed env (unpack -> fused K-step kernel -> pack) against the oracle's, then ordinary steps."""
    n = 3001
    env = B.PackedCellularVectorEnv(num_envs=n, n_cells=C, n_states=S, stochastic=stochastic, env_seed=5, max_episode_steps=7)
    ora = O.OracleEnv(n_envs=n, n_cells=C, n_states=S, noise=stochastic, seed=5, rng_episodic=True, max_episode_steps=7,
                      reward="nonlinear_rp" if stochastic else "right_polarizing")
    for _ in range(2):
        ret, uns = env.rollout(9)
        oret, ouns = ora.rollout(9, None)
        assert (u32(env.packed_state) == pack(ora.state)).all()
        assert (host(env.time_step) == ora.t).all() and (host(env.tabular_state()) == ora.index).all()
        assert (host(uns) == ouns).all()
        np.testing.assert_allclose(host(ret), oret, rtol=1e-5, atol=1e-6)
    a = np.random.default_rng(1).integers(0, S, size=(C, n)).astype(np.int8)
    env.step_device(dev(a))
    ora.step(a)
    assert_matches_oracle(env, ora)
    assert env.stats()["env_steps"] == ora.stats[0] == 19 * n


def test_step_many_bound_and_graph(B, O):
    """gc_step_many (n pre-bound steps in one foreign call), bound steps and a captured CUDA graph, in the
    packed layout with non-episodic noise: 3 x 8 distinct steps, all reading the device-resident counter."""
    import ctypes as C
    n = 10007
    # (a final-observation buffer keeps the handle on separate launches: the one-launch path of small shards is
    # covered by tests/test_gpu_many.py)
    env = B.PackedCellularVectorEnv(num_envs=n, n_cells=4, n_states=4, stochastic=True, rng_episodic=False, env_seed=9,
                                    emit_final_obs=True)
    ora = O.OracleEnv(n_envs=n, n_cells=4, n_states=4, noise=True, rng_episodic=False, seed=9, reward="nonlinear_rp")
    rng = np.random.default_rng(5)
    acts = [rng.integers(0, 4, size=(4, n)).astype(np.int8) for _ in range(8)]
    ring = []
    for a in acts:
        w = torch.zeros(env.ld, dtype=torch.int32, device="cuda")
        w[:n] = dev(pack(a).view(np.int32))
        ring.append(w)
    slots = [env._bind(w) for w in ring]
    arr = (C.c_int32 * 8)(*slots)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    from gym_cellular_b200 import _lib
    _lib.check(env._lib.gc_step_many(env._h, arr, 8, 8, stream))
    for a in acts:
        ora.step(a)
    assert_matches_oracle(env, ora)
    graph = env.capture_steps(ring)            # capture launches the 8 steps once more (warm-up inside capture is not run)
    graph.replay()
    for a in acts:
        ora.step(a)
    assert_matches_oracle(env, ora)
    _lib.check(env._lib.gc_step_many(env._h, arr, 8, 11, stream))      # wraps around the slot list
    for i in range(11):
        ora.step(acts[i % 8])
    assert_matches_oracle(env, ora)
    assert env.sync_step_counter() == 27
    env.step_many(slots, 70)                               # 4 replays of the cached 16-step graph + 6 plain launches
    for i in range(70):
        ora.step(acts[i % 8])
    assert_matches_oracle(env, ora)
    assert env.sync_step_counter() == 97 and env.launch_count >= 97


def test_make_vector_env_layouts_and_debug_ids(B):
    for name in ("Cells3States3Actions3-v0", "Cells2Rest3-v0", "Cells3ResetVDeadlock-v0", "Debug-v0",
                 "DeepPlanningDebug-v0", "DeepExplorationDebug-v0"):
        a = B.make_vector_env("gym_cellular/" + name, 64, layout="packed")
        b = B.make_vector_env("gym_cellular/" + name, 64)
        gen = torch.Generator(device="cuda").manual_seed(1)
        for _ in range(6):
            act = torch.randint(0, a.n_actions, (a.n_cells, 64), dtype=torch.int8, device="cuda", generator=gen)
            a.step_device(act)
            b.step_device(act)
            assert torch.equal(a.state, b.state) and torch.equal(a._reward[:64], b._reward[:64])
            assert torch.equal(a._flags[:64] & 1, b._unsafe[:64])
    with pytest.raises(ValueError):
        B.make_vector_env("gym_cellular/GridWorld-v0", 64, layout="packed")
    with pytest.raises(ValueError):
        B.PackedCellularVectorEnv(num_envs=16, n_cells=4, n_states=5)


def test_example_packed_tabular_agent(B):
    """examples/packed_tabular_agent.py: a tabular agent consuming index / reward / flag of the packed env learns
    to report fewer unsafe steps; the env's own statistics agree with what the agent saw."""
    import importlib.util
    import sys
    from conftest import REPO
    spec = importlib.util.spec_from_file_location("packed_tabular_agent", REPO + "/examples/packed_tabular_agent.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    argv, sys.argv = sys.argv, ["x", "--envs", "8192", "--steps", "240"]
    try:
        res = mod.main()
    finally:
        sys.argv = argv
    assert res["env_steps"] == 8192 * 240 and res["kernel_launches"] >= 240
    assert abs(res["stats_reward_sum"] - res["total_reward"]) <= 1e-3 * abs(res["total_reward"])
    assert res["unsafe_rate_last_quarter"] < res["unsafe_rate"]          # the greedy phase reports fewer unsafe steps
