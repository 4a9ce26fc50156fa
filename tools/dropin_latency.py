"""Latency of the N = 1 drop-in env (BASELINE configs 1/2 shape): F110Env.step through the C ABI with numpy in/out,
numpy-drawn lidar noise uploaded every step (the reference's stream), all outputs downloaded."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import f110_gymnasium_ros2_jazzy_b200 as f
from tests import helpers as H
d = tempfile.mkdtemp()
map_dir, name = H.write_map_files('Shanghai_map', d)
for A, poses, act in ((2, [[0, 0, 0], [3.0, 0.5, 0]], [[0.05, 3.0], [0.0, 2.0]]), (1, [[0, 0, 0]], [[0.0, 2.0]])):
    for noise in ('numpy', 'device'):
        env = f.make('f110_gym:f110-v0', map_dir=map_dir, map=name, map_ext='.png', num_agents=A, noise=noise)
        poses_a = np.array(poses, np.float64); a = np.array(act, np.float32)
        env.reset(options=poses_a)
        for _ in range(50): env.step(a)
        t0 = time.perf_counter(); n = 0
        for _ in range(2000):
            obs, r, term, trunc, info = env.step(a); n += 1
            if term: env.reset(options=poses_a)
        el = time.perf_counter() - t0
        print("F110Env num_agents=%d noise=%s: %.0f env-steps/s (%.1f us/step)" % (A, noise, n / el, 1e6 * el / n))
        env.close()
