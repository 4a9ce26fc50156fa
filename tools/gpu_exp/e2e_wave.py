"""Synchronous F110HostVecEnv.step: the first chunk's lidar kernel on a share of the resident wave, so that the chunks'
kernels run side by side instead of one after the other (ms/step, best of 3 x 100 steps)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from f110_gymnasium_ros2_jazzy_b200 import F110HostVecEnv, workloads
m = workloads.shanghai_map()
N = 4096
poses = workloads.start_poses(N)
acts = torch.rand((120, N, 1, 2)).mul(torch.tensor([0.8, 20.])).sub(torch.tensor([0.4, 0.])).pin_memory().numpy()
CASES = [((1, 3), None), ((1, 3), (50, None)), ((1, 3), (33, None)), ((1, 3), (25, None)), ((1, 3), (17, None)),
         ((1, 2), (50, None)), ((1, 2), (33, None)), ((1, 1), (50, None)), ((1, 1), (50, 50)), ((2, 3), (50, None)),
         ((1, 3), (33, 67)), ((1, 1, 2), (25, 33, None)), ((1, 1, 2), (33, 50, None)), ((1, 2, 3), (25, 50, None)),
         ((1, 1, 1, 1), (25, 33, 50, None)), ((1, 3), None)]
for chunks, wave in CASES:
    env = F110HostVecEnv(N, chunks=chunks, map_arrays=m, num_agents=1, wave=wave)
    env.reset(poses)
    for k in range(10): env.step(acts[k])
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(10, 110): env.step(acts[k])
        best = min(best, (time.perf_counter() - t0) / 100)
    print('chunks', chunks, 'wave', wave, 'ms/step %.4f' % (best * 1e3), flush=True)
    env.close()
