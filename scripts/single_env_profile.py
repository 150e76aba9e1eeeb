"""Where the time of one single-env step() goes (the reference's gymnasium.make API on the CUDA path, num_envs = 1)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import gym_cellular_b200  # noqa: F401  (registers the ids)
from gym_cellular_b200._gym import gym

name = sys.argv[1] if len(sys.argv) > 1 else "gym_cellular/Cells3States3Actions3-v0"
e = gym.make(name)
e.reset()
act = (1, 2, 0) if "Grid" not in name else e.action_space.sample()
for _ in range(200):
    e.step(act)
t0 = time.perf_counter()
for _ in range(2000):
    e.step(act)
dt = time.perf_counter() - t0
print(f"{name}: {dt / 2000 * 1e6:.1f} us per step, {2000 / dt:.0f} steps/s")
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    e.step(act)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
