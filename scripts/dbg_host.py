import subprocess, sys
CASE = r'''
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import gym_cellular_b200 as B
kind, n, chunk, se, sto, dev_first = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] == "1", sys.argv[5] == "1", sys.argv[6] == "1"
kw = dict(stochastic=sto) if kind == "cellular" else {}
env = B.CellularVectorEnv(kind=kind, num_envs=n, host_chunk_envs=chunk, emit_side_effects=se, **kw)
a = np.zeros((env.n_cells, n), np.int8)
if dev_first:
    env.step_device(torch.from_numpy(a).cuda()); torch.cuda.synchronize()
for t in range(3):
    env.step(a)
torch.cuda.synchronize()
'''
for args in (("cellular", 100003, 16384, 1, 1, 0), ("cellular", 100003, 16384, 0, 1, 0), ("cellular", 100003, 16384, 1, 0, 0),
             ("cellular", 4096, 1 << 20, 1, 1, 0), ("cellular", 65536, 16384, 1, 1, 0), ("cellular", 100003, 1 << 20, 1, 1, 0),
             ("gridworld", 100003, 16384, 1, 1, 0), ("cellular", 4096, 1 << 20, 1, 1, 1)):
    r = subprocess.run([sys.executable, "-c", CASE] + [str(x) for x in args], capture_output=True, text=True,
                       env={**__import__("os").environ, "CUDA_LAUNCH_BLOCKING": "1"})
    print(args, "ok" if r.returncode == 0 else "FAIL " + r.stderr.strip().splitlines()[-1][:160], flush=True)
