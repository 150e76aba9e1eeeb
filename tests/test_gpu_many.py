"""gc_step_many in ONE launch (gc_cell_fast.cu: cell_pair_many_kernel, gc_grid.cu: grid_many_kernel): small shards
whose bound slots share every buffer but the actions run all their steps in a single kernel.  The results must be
bit-identical to the same steps launched one by one (a twin handle that is kept off the fused path) and must
match the oracle stepped with the same actions."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

# GC_B200_STEP_MANY_FUSED=0 in the environment keeps every call on separate launches (the suite is also run that way)
FUSED = os.environ.get("GC_B200_STEP_MANY_FUSED", "1")[:1] != "0"

REWARD_RTOL = 1e-6
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


def ring_of(env, acts):
    ring = []
    for a in acts:
        t = torch.zeros(a.shape[0], env.ld, dtype=torch.int8, device="cuda")
        t[:, :a.shape[1]] = dev(a)
        ring.append(t)
    return ring


def same_outputs(a, b, n, se):
    assert torch.equal(a.state, b.state) and torch.equal(a.time_step, b.time_step)
    assert torch.equal(a._reward[:n], b._reward[:n])                       # bit for bit: the same arithmetic
    assert torch.equal(a._index[:n], b._index[:n]) and torch.equal(a._truncated[:n], b._truncated[:n])
    assert torch.equal(a._terminated[:n], b._terminated[:n])
    assert torch.equal(a._unsafe[:n], b._unsafe[:n]) and torch.equal(a._count[:n], b._count[:n])
    if se:
        assert torch.equal(a._se_row[:, :n], b._se_row[:, :n])
    assert a.stats() == b.stats()


def matches_oracle(env, ora, n, se):
    assert (host(env.state) == ora.state).all() and (host(env.tabular_state()) == ora.index).all()
    assert (host(env.time_step) == ora.t).all() and (host(env._truncated[:n]) == ora.truncated).all()
    assert (host(env._unsafe[:n]) == ora.unsafe).all() and (host(env._count[:n]) == ora.count).all()
    if se:
        assert (host(env._se_row[:, :n]) == ora.se_row).all()
    np.testing.assert_allclose(host(env._reward[:n]), ora.reward, rtol=REWARD_RTOL, atol=REWARD_ATOL)


CASES = [  # C, S, stochastic, deadlock, episodic, side-effect rows, time limit, n
    (3, 3, False, False, True, False, 0, 65536),      # BASELINE config 2
    (3, 3, True, True, True, True, 5, 20011),
    (3, 3, True, False, False, False, 7, 4099),
    (2, 3, False, False, True, True, 4, 1000),
    (1, 4, True, False, False, True, 3, 517),
    (4, 4, True, False, True, False, 6, 9001),
    (5, 4, True, True, False, True, 5, 3003),          # wide: 16-bit-half draws
    (8, 4, True, False, True, True, 9, 2049),
    (7, 2, False, False, True, False, 0, 5),
]


@pytest.mark.parametrize("C,S,stochastic,deadlock,episodic,se,limit,n", CASES)
def test_cellular_many_steps_in_one_launch(B, O, C, S, stochastic, deadlock, episodic, se, limit, n):
    off = 4 * 123456789 if C > 2 else 0
    kw = dict(num_envs=n, n_cells=C, n_states=S, stochastic=stochastic, deadlock=deadlock, rng_episodic=episodic, env_seed=C + S,
              max_episode_steps=limit or None, emit_side_effects=se, env_id_offset=off, difficulty="hard" if C % 2 else "easy")
    fused = B.CellularVectorEnv(**kw)
    plain = B.CellularVectorEnv(emit_final_obs=True, **kw)                 # a final-observation buffer keeps a handle off the fused path
    ora = O.OracleEnv(n_envs=n, n_cells=C, n_states=S, noise=stochastic, deadlock=deadlock, rng_episodic=episodic, seed=C + S,
                      max_episode_steps=limit, env_id_offset=off, difficulty="hard" if C % 2 else "easy",
                      reward="nonlinear_rp" if stochastic else "right_polarizing")
    rng = np.random.default_rng(n + C)
    acts = [rng.integers(0, S, size=(C, n)).astype(np.int8) for _ in range(3)]
    slots_f = [fused._bind(t) for t in ring_of(fused, acts)]
    slots_p = [plain._bind(t) for t in ring_of(plain, acts)]
    done = 0
    for steps in (7, 1, 5):                                                # 1: a single step takes the plain launch
        l_f, l_p = fused.launch_count, plain.launch_count
        order_f = [slots_f[(done + i) % 3] for i in range(3)]
        order_p = [slots_p[(done + i) % 3] for i in range(3)]
        fused.step_many(order_f, steps)
        plain.step_many(order_p, steps)
        assert fused.launch_count - l_f == (1 if FUSED else steps) and plain.launch_count - l_p == steps
        for i in range(steps):
            ora.step(acts[(done + i) % 3])
        done += steps
        same_outputs(fused, plain, n, se)
        matches_oracle(fused, ora, n, se)
        assert fused.sync_step_counter() == done == plain.sync_step_counter()
    # the handle goes on with ordinary steps (the device step counter and the registers' state were written back)
    a = rng.integers(0, S, size=(C, n)).astype(np.int8)
    fused.step_device(dev(a)); plain.step_device(dev(a)); ora.step(a)
    same_outputs(fused, plain, n, se)
    matches_oracle(fused, ora, n, se)
    s = fused.stats()
    assert s["env_steps"] == (done + 1) * n == ora.stats[0] and s["unsafe_steps"] == ora.stats[1] and s["count_sum"] == ora.stats[2]
    assert s["episodes_truncated"] == ora.stats[3]


@pytest.mark.parametrize("episodic,se,limit,n", [(False, False, 7, 20011), (True, True, 5, 1 << 20), (False, True, 0, 333)])
def test_gridworld_many_steps_in_one_launch(B, O, episodic, se, limit, n):
    kw = dict(kind="gridworld", num_envs=n, env_seed=5, max_episode_steps=limit or None, dispersal_prob=0.1, rng_episodic=episodic,
              emit_side_effects=se, env_id_offset=4 * 987654321)
    fused = B.CellularVectorEnv(**kw)
    plain = B.CellularVectorEnv(emit_final_obs=True, **kw)
    ora = O.OracleEnv(kind="gridworld", n_envs=n, seed=5, max_episode_steps=limit, dispersal_prob=0.1, rng_episodic=episodic,
                      env_id_offset=4 * 987654321)
    rng = np.random.default_rng(n)
    acts = []
    for _ in range(8):
        a = np.full((2, n), 4, np.int8)
        a[rng.integers(0, 2, n), np.arange(n)] = rng.integers(0, 4, n)
        acts.append(a)
    slots_f = [fused._bind(t) for t in ring_of(fused, acts)]
    slots_p = [plain._bind(t) for t in ring_of(plain, acts)]
    done = 0
    for steps in (19, 3):
        l_f = fused.launch_count
        fused.step_many([slots_f[(done + i) % 8] for i in range(8)], steps)
        plain.step_many([slots_p[(done + i) % 8] for i in range(8)], steps)
        assert fused.launch_count - l_f == (1 if FUSED else steps)
        for i in range(steps):
            ora.step(acts[(done + i) % 8])
        done += steps
        same_outputs(fused, plain, n, se)
        matches_oracle(fused, ora, n, se)
        assert fused.sync_step_counter() == done
    assert fused.stats()["env_steps"] == done * n == ora.stats[0] and fused.stats()["episodes_truncated"] == ora.stats[3]


def test_slots_with_their_own_outputs_stay_on_separate_launches(B, O):
    """Slots that do not share their output buffers cannot be fused (each step's outputs must land in its own slot):
    the call falls back to one launch per step."""
    n = 4096
    env = B.CellularVectorEnv(num_envs=n, emit_side_effects=False)
    ora = O.OracleEnv(n_envs=n)
    rng = np.random.default_rng(0)
    acts = [rng.integers(0, 3, size=(3, n)).astype(np.int8) for _ in range(2)]
    ring = ring_of(env, acts)
    own_reward = torch.zeros(env.ld, dtype=torch.float32, device="cuda")
    from gym_cellular_b200 import _lib
    from gym_cellular_b200.vector_env import _ptr
    s0 = env._bind(ring[0])
    s1 = 9
    _lib.check(env._lib.gc_bind_step(env._h, s1, _ptr(ring[1]), _ptr(env._state), _ptr(env._t), _ptr(own_reward), _ptr(env._index),
                                     _ptr(env._terminated), _ptr(env._truncated), _ptr(env._unsafe), _ptr(env._count), None,
                                     _ptr(env._stats)))
    before = env.launch_count
    env.step_many([s0, s1], 6)
    assert env.launch_count - before == 6
    for i in range(6):
        ora.step(acts[i % 2])
    assert (host(env.state) == ora.state).all()
    np.testing.assert_allclose(host(own_reward[:n]), ora.reward, rtol=REWARD_RTOL, atol=REWARD_ATOL)    # step 6 used slot s1


def pack_host(cells):
    w = np.zeros(cells.shape[1], np.uint32)
    for c in range(cells.shape[0]):
        w |= cells[c].astype(np.uint32) << np.uint32(2 * c)
    return w


@pytest.mark.parametrize("C,S,stochastic,episodic,se,limit,n", [
    (3, 3, False, True, False, 0, 65536), (3, 3, True, False, True, 5, 20011), (4, 4, True, True, False, 6, 9001),
    (8, 4, True, False, True, 7, 2049), (1, 2, False, True, False, 3, 33), (7, 4, False, True, False, 0, 5003)])
def test_packed_many_steps_in_one_launch(B, O, C, S, stochastic, episodic, se, limit, n):
    """The same for packed bindings (cell_packed_many_kernel): bit-identical to separate launches of a twin handle and
    to the int8 layout stepped through the oracle."""
    kw = dict(num_envs=n, n_cells=C, n_states=S, stochastic=stochastic, rng_episodic=episodic, env_seed=C + S,
              max_episode_steps=limit or None, emit_side_effects=se, env_id_offset=4 * 1000003)
    fused = B.PackedCellularVectorEnv(**kw)
    plain = B.PackedCellularVectorEnv(emit_final_obs=True, **kw)           # a final-observation buffer keeps it off the fused path
    ora = O.OracleEnv(n_envs=n, n_cells=C, n_states=S, noise=stochastic, rng_episodic=episodic, seed=C + S, max_episode_steps=limit,
                      env_id_offset=4 * 1000003, reward="nonlinear_rp" if stochastic else "right_polarizing")
    rng = np.random.default_rng(n)
    acts = [rng.integers(0, S, size=(C, n)).astype(np.int8) for _ in range(4)]

    def ring(env):
        out = []
        for a in acts:
            w = torch.zeros(env.ld, dtype=torch.int32, device="cuda")
            w[:n] = dev(pack_host(a).view(np.int32))
            out.append(w)
        return out
    slots_f = [fused._bind(w) for w in ring(fused)]
    slots_p = [plain._bind(w) for w in ring(plain)]
    done = 0
    for steps in (9, 6):
        l_f, l_p = fused.launch_count, plain.launch_count
        fused.step_many([slots_f[(done + i) % 4] for i in range(4)], steps)
        plain.step_many([slots_p[(done + i) % 4] for i in range(4)], steps)
        assert fused.launch_count - l_f == (1 if FUSED else steps) and plain.launch_count - l_p == steps
        for i in range(steps):
            ora.step(acts[(done + i) % 4])
        done += steps
        assert torch.equal(fused._state, plain._state) and torch.equal(fused._t, plain._t)
        assert torch.equal(fused._reward[:n], plain._reward[:n]) and torch.equal(fused._flags[:n], plain._flags[:n])
        if fused._index_ptr() is not None:
            assert torch.equal(fused.tabular_state(), plain.tabular_state())
        if se:
            assert torch.equal(fused._se_row[:n], plain._se_row[:n])
        assert fused.stats() == plain.stats()
        assert (host(fused._state[:n]).view(np.uint32) == pack_host(ora.state)).all() and (host(fused._t[:n]) == ora.t).all()
        assert (host(fused.tabular_state()) == ora.index).all()
        np.testing.assert_allclose(host(fused._reward[:n]), ora.reward, rtol=REWARD_RTOL, atol=REWARD_ATOL)
        assert fused.sync_step_counter() == done
    s = fused.stats()
    assert s["env_steps"] == done * n == ora.stats[0] and s["unsafe_steps"] == ora.stats[1] and s["count_sum"] == ora.stats[2]
    assert s["episodes_truncated"] == ora.stats[3]
