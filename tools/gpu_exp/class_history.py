"""Record the longest ray (lookups) of every lidar work unit over consecutive steps of the C3 workload -> gpurun_out/class_history.npz
(input for choosing the launch-order predictor; development aid)."""
import argparse, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from f110_gymnasium_ros2_jazzy_b200 import F110VecEnv, _lib

ap = argparse.ArgumentParser(); ap.add_argument('--envs', type=int, default=4096); ap.add_argument('--steps', type=int, default=60)
a = ap.parse_args()
args = argparse.Namespace(map='Shanghai_map', map_upsample=1, agents=1, beams=1080, envs=a.envs)
map_arrays, poses = bench.load_workload(args, a.envs, 0, a.envs)
env = F110VecEnv(a.envs, num_agents=1, num_beams=1080, seed=42, device=0, auto_reset=True, outputs=('obs', 'reward', 'terminated'),
                 noise_std=0.01, count_lookups=True, map_arrays=map_arrays)
env.reset(poses)
acts = bench.action_stream(torch, a.steps + 100, a.envs, 1, torch.device('cuda', 0))
L = env.backend.lib
n = int(L.f110_debug_unit_timeline(env.backend.h, None, 0))
look = np.zeros((a.steps, n), np.uint16); term = np.zeros((a.steps, a.envs), np.uint8)
buf = np.zeros((n, 4), np.uint32)
for k in range(100):
    env.step(acts[k])
for k in range(a.steps):
    out = env.step(acts[100 + k])
    L.f110_debug_unit_timeline(env.backend.h, buf.ctypes.data_as(C.c_void_p), n)
    look[k] = np.minimum(buf[:, 2], 65535)
    term[k] = out[2].cpu().numpy() if isinstance(out, tuple) else env.backend.out['terminated'].cpu().numpy()
os.makedirs('gpurun_out', exist_ok=True)
np.savez_compressed('gpurun_out/class_history.npz', look=look, terminated=term)
print('saved', look.shape, 'mean of unit max', look.mean(), 'max', look.max())
