"""BASELINE config 5: DDPG-style rollout loop on device -- ego = the reference's actor architecture
(random init, seed 42), opponent = gap-follow kernel (or constant), E two-agent envs per GPU, observations consumed
in place by torch.  Prints rollout env-steps/s.  Not a bench.py line; results are recorded in profiles/."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f110_gymnasium_ros2_jazzy_b200 import Actor, DeviceRollout, F110VecEnv, ShapedReward  # noqa: E402
from tests import helpers as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=8192)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--opponent", default="gap_follow")
ap.add_argument("--reward", action="store_true", help="also evaluate the shaped reward on device")
ap.add_argument("--graph", action="store_true", help="capture one rollout step in a CUDA graph")
args = ap.parse_args()

# under torchrun: one process per GPU, `--envs` per GPU (BASELINE config 5 = 8 x 8192), env-index sharding, no collective on
# the step path; the max over ranks of the device time is reported
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = args.envs
m = H.golden_map('Shanghai_map')
cl = H.load('maps')['Shanghai_map__centerline_poses']
idx = np.linspace(0, len(cl) - 1, N * world).round().astype(int)[rank * N:(rank + 1) * N]
poses = np.stack([cl[idx], cl[(idx + 40) % len(cl)]], axis=1)
env = F110VecEnv(N, num_agents=2, map_arrays=m, outputs=('obs', 'reward', 'terminated', 'scans_f32'))
torch.manual_seed(42)
actor = Actor(1088, 2, [-0.4189, 0.0], [0.4189, 20.0]).cuda()
opp = 'gap_follow' if args.opponent == 'gap_follow' else (0.0, 1.5)
rfn = None
if args.reward:
    rfn = ShapedReward(N, H.load('reward')['centerline'], w_prog=5.0, alive_bonus=0.5, grace_steps_wall=25, grace_steps_opp=175,
                       w_lat=0.25, lat_cap=3.0, near_wall_dist=0.30 / 30, w_wall=0.30, wall_quantile=0.10, opp_safe_dist=0.60,
                       w_opp=0.30)
ro = DeviceRollout(env, actor, opponent=opp, reward_fn=rfn)
ro.reset(poses)
for _ in range(10):
    ro.step()
step = ro.step
if args.graph:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ro.step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ro.step()
    step = g.replay
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms, wall], dtype=torch.float64, device='cuda')
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, wall = float(t[0]), float(t[1])
    N = N * world
if rank == 0:
  print(json.dumps({"config": "C5 rollout: %d two-agent envs, actor 1088-128-128-2 + %s opponent%s%s" % (N, args.opponent, " + shaped reward" if args.reward else "", ", CUDA graph" if args.graph else ""),
                  "n_gpus": world, "env_steps_per_s": N * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                  "wall_ms_per_step": 1e3 * wall / args.steps, "rays_per_s": N * 2 * 1080 * args.steps / (ms * 1e-3)}))
if world > 1:
    dist.destroy_process_group()
