"""`gymnasium.vector.AsyncVectorEnv` look-alike on `multiprocessing`: worker processes that each own a
slice of the env instances and talk to the parent over one pipe, one round trip per `step()`.

gymnasium is not installed in the build image; this stand-in has the same constructor (`env_fns`,
`shared_memory`, `context`, `daemon`), the same `reset` / `step_async` / `step_wait` / `step` / `close`
protocol and the same one-message-per-worker-per-step cost model, so it can serve as the CPU baseline
harness BASELINE.json names ("gymnasium AsyncVectorEnv across all host cores").  One difference, stated:
gymnasium spawns one process per env; here `envs_per_worker` consecutive env_fns share a process
(`envs_per_worker=1` reproduces gymnasium's layout), because the baseline is specified as one worker per
core with `num_envs = workers x k`.  `shared_memory` is accepted and ignored (always False: the grid-world
observation is not a member of its declared space, SURVEY hard parts), there is no auto-reset (no env of
the family ever terminates, cells3states3actions3.py:121-122).
"""
import multiprocessing as mp

import numpy as np

from . import VectorEnv, batch_space


def _worker(pipe, parent_pipe, env_fns):
    parent_pipe.close()
    envs = [fn() for fn in env_fns]
    try:
        while True:
            cmd, data = pipe.recv()
            if cmd == "reset":
                pipe.send([env.reset(**data) if data else env.reset() for env in envs])
            elif cmd == "step":
                pipe.send([env.step(a) for env, a in zip(envs, data)])
            elif cmd == "spaces":
                pipe.send((getattr(envs[0], "observation_space", None), getattr(envs[0], "action_space", None)))
            elif cmd == "call":
                name, args, kwargs = data
                pipe.send([getattr(env, name)(*args, **kwargs) if callable(getattr(env, name)) else getattr(env, name)
                           for env in envs])
            elif cmd == "close":
                for env in envs:
                    if hasattr(env, "close"):
                        env.close()
                pipe.send(None)
                break
            else:
                raise RuntimeError(f"unknown command {cmd!r}")
    except (KeyboardInterrupt, EOFError):
        pass
    finally:
        pipe.close()


def _concat(items):
    """Per-env results -> batched result: tuples of scalars become a tuple of arrays (a batched
    Tuple(Discrete) observation), scalars an array; anything else stays a tuple of per-env objects."""
    first = items[0]
    if isinstance(first, (tuple, list)) and first and all(np.isscalar(x) for x in first):
        return tuple(np.asarray(col) for col in zip(*items))
    if np.isscalar(first):
        return np.asarray(items)
    return tuple(items)


class AsyncVectorEnv(VectorEnv):
    def __init__(self, env_fns, observation_space=None, action_space=None, shared_memory=False, copy=True,
                 context=None, daemon=True, worker=None, envs_per_worker=1):
        self.env_fns = list(env_fns)
        self.num_envs = len(self.env_fns)
        self.envs_per_worker = int(envs_per_worker)
        ctx = mp.get_context(context or "fork")
        self._pipes, self._procs, self._slices = [], [], []
        for lo in range(0, self.num_envs, self.envs_per_worker):
            fns = self.env_fns[lo:lo + self.envs_per_worker]
            parent, child = ctx.Pipe()
            proc = ctx.Process(target=worker or _worker, args=(child, parent, fns), daemon=daemon)
            proc.start()
            child.close()
            self._pipes.append(parent)
            self._procs.append(proc)
            self._slices.append((lo, lo + len(fns)))
        self._pipes[0].send(("spaces", None))
        obs_space, act_space = self._pipes[0].recv()
        self.single_observation_space = observation_space or obs_space
        self.single_action_space = action_space or act_space
        try:
            self.observation_space = batch_space(self.single_observation_space, self.num_envs)
            self.action_space = batch_space(self.single_action_space, self.num_envs)
        except TypeError:
            self.observation_space = self.action_space = None
        self._waiting = False
        self.closed = False

    @property
    def num_workers(self):
        return len(self._procs)

    def reset(self, *, seed=None, options=None):
        for pipe in self._pipes:
            pipe.send(("reset", None))
        results = [r for pipe in self._pipes for r in pipe.recv()]
        obs, infos = zip(*results)
        return _concat(list(obs)), {"per_env": infos}

    def step_async(self, actions):
        if self._waiting:
            raise RuntimeError("step_async called while a step is pending")
        # batched actions: a tuple of per-cell arrays (batched Tuple space) or a sequence of per-env actions
        if isinstance(actions, tuple) and len(actions) and isinstance(actions[0], np.ndarray) and \
                actions[0].shape == (self.num_envs,):
            per_env = list(zip(*(a.tolist() for a in actions)))
        else:
            per_env = list(actions)
        for pipe, (lo, hi) in zip(self._pipes, self._slices):
            pipe.send(("step", per_env[lo:hi]))
        self._waiting = True

    def step_wait(self, timeout=None):
        if not self._waiting:
            raise RuntimeError("step_wait called without step_async")
        results = [r for pipe in self._pipes for r in pipe.recv()]
        self._waiting = False
        obs, rew, term, trunc, infos = zip(*results)
        return (_concat(list(obs)), np.asarray(rew, np.float64), np.asarray(term, np.bool_),
                np.asarray(trunc, np.bool_), {"per_env": infos})

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def call(self, name, *args, **kwargs):
        for pipe in self._pipes:
            pipe.send(("call", (name, args, kwargs)))
        return tuple(r for pipe in self._pipes for r in pipe.recv())

    def close_extras(self, timeout=None, terminate=False):
        for pipe, proc in zip(self._pipes, self._procs):
            try:
                if proc.is_alive() and not terminate:
                    pipe.send(("close", None))
                    pipe.recv()
            except (BrokenPipeError, EOFError):
                pass
            pipe.close()
        for proc in self._procs:
            if terminate and proc.is_alive():
                proc.terminate()
            proc.join(timeout)
