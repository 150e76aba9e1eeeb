"""Exact tabular model export: P[s, a, s'] and R[s, a] of an environment kind.

The model is not re-derived on the host: it is ENUMERATED BY THE STEP KERNEL.  Every (state, action)
pair of the tabular space is stepped once per random outcome in replay mode (the outcome is fed as
the uniform stream), and the outcomes are weighted by their probabilities:
  cellular family  each drawing cell fires independently with p = noise_prob  -> 2^C patterns
  grid world       no dispersal (1 - p), or dispersal (p) x 2 x 2 x 2 equally likely (b00, b10, k)
This is what tabular planners of the PilotExperimentation kind estimate from samples.
"""
import itertools

import numpy as np
import torch

from .vector_env import CellularVectorEnv


def exact_model(kind="cellular", **env_kwargs):
    """-> (P float64 [nS, nA, nS], R float64 [nS, nA], states_valid bool [nS]) as numpy arrays.

    Feasible for the reference-sized spaces (27 x 27, 9 x 9, 400 x 25, 16 x 4); raises for spaces
    beyond 2^20 state-action pairs."""
    probe = CellularVectorEnv(kind=kind, num_envs=16, **env_kwargs)
    C, S, A = probe.n_cells, probe.n_states, probe.n_actions
    nS, nA = S ** C, A ** C
    stochastic = probe.stochastic
    p_noise = probe._cfg.noise_prob
    p_disp = probe._cfg.dispersal_prob
    probe.close()
    if nS * nA > 1 << 20:
        raise ValueError("state-action space too large for a dense model")
    n = nS * nA
    pairs = torch.arange(n, device="cuda")
    s_idx, a_idx = pairs // nA, pairs % nA
    env = CellularVectorEnv(kind=kind, num_envs=n, **env_kwargs)
    states, actions = env.detabularize(s_idx, "state"), env.detabularize(a_idx, "action")
    valid = torch.ones(nS, dtype=torch.bool, device="cuda")
    if kind == "gridworld":
        # exactly one jurisdiction holds the agent (the others are not states of the reference)
        cells = env.detabularize(torch.arange(nS, device="cuda"), "state")
        valid = ((cells[0] >> 2) < 4) ^ ((cells[1] >> 2) < 4)
        actions = torch.where((actions[0] == 4) & (actions[1] == 4), torch.zeros_like(actions), actions)
        outcomes = [(1.0 - p_disp, [0.999999, 0.25, 0.25, 0.25, 0.25, 0.25])]
        for b00, b10, k in itertools.product((0, 1), repeat=3):
            outcomes.append((p_disp / 8, [0.0, 0.25 + 0.5 * b00, 0.25, 0.25 + 0.5 * b10, 0.25, 0.25 + 0.5 * k]))
    elif stochastic:
        outcomes = []
        for fire in itertools.product((0, 1), repeat=C):
            w = float(np.prod([p_noise if f else 1.0 - p_noise for f in fire]))
            outcomes.append((w, [0.0 if f else 0.999999 for f in fire]))
    else:
        outcomes = [(1.0, None)]
    P = torch.zeros(nS, nA, nS, dtype=torch.float64, device="cuda")
    R = torch.zeros(nS, nA, dtype=torch.float64, device="cuda")
    for w, u in outcomes:
        env.set_state(states, t=torch.zeros(n, dtype=torch.int32, device="cuda"))
        ru = None if u is None else torch.tensor(u, dtype=torch.float64, device="cuda").repeat(n, 1)
        if ru is None and (stochastic or kind == "gridworld"):
            ru = torch.full((n, C if kind == "cellular" else 6), 0.999999, dtype=torch.float64, device="cuda")
        env.step_device(actions, replay_u=ru)
        P.view(n, nS).index_put_((pairs, env.tabular_state()), torch.full((n,), w, dtype=torch.float64, device="cuda"),
                                 accumulate=True)
        R.view(n).add_(w * env._reward[:n].to(torch.float64))
    env.close()
    return P.cpu().numpy(), R.cpu().numpy(), valid.cpu().numpy()
