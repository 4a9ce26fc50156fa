"""Synchronous F110HostVecEnv.step with different chunkings of 4096 envs (ms/step, best of 3 x 100 steps)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from f110_gymnasium_ros2_jazzy_b200 import F110HostVecEnv
from tests import helpers as H
m = H.golden_map('Shanghai_map'); cl = H.load('maps')['Shanghai_map__centerline_poses']
N = 4096
poses = cl[np.linspace(0, len(cl) - 1, N).round().astype(int)][:, None, :]
acts = torch.rand((120, N, 1, 2)).mul(torch.tensor([0.8, 20.])).sub(torch.tensor([0.4, 0.])).pin_memory().numpy()
SETS = {'r01': (1, 2, 3, 4, (1, 3), (1, 2), (1, 2, 3), (1, 2, 2), (1, 3, 4), (2, 3, 3), (1, 1, 2, 4), (1, 2, 3, 4), (3, 2)),
        'quick': ((1, 3), (1, 4), (1, 2), 2, (1, 1, 2), (1, 2, 5), (1, 3)),
        'two': ((1, 3), (1, 4), (1, 5), (1, 7), (1, 9), (1, 11), (1, 15), (1, 31), (1, 3), (1, 5, 10))}
for chunks in SETS[sys.argv[1] if len(sys.argv) > 1 else 'r01']:
    env = F110HostVecEnv(N, chunks=chunks, map_arrays=m, num_agents=1)
    env.reset(poses)
    for k in range(10): env.step(acts[k])
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(10, 110): env.step(acts[k])
        best = min(best, (time.perf_counter() - t0) / 100)
    print('chunks', chunks, 'bounds', env.bounds, 'ms/step %.4f' % (best * 1e3), flush=True)
    env.close()
