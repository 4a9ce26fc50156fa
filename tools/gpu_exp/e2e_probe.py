import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from f110_gymnasium_ros2_jazzy_b200 import F110HostVecEnv
from tests import helpers as H
m = H.golden_map('Shanghai_map'); cl = H.load('maps')['Shanghai_map__centerline_poses']
N = 4096
poses = cl[np.linspace(0, len(cl) - 1, N).round().astype(int)][:, None, :]
acts = torch.rand((120, N, 1, 2)).mul(torch.tensor([0.8, 20.])).sub(torch.tensor([0.4, 0.])).pin_memory().numpy()
for chunks in (1, 2, 4):
    for outs in (('obs', 'reward', 'terminated'), ('reward', 'terminated')):
        env = F110HostVecEnv(N, chunks=chunks, map_arrays=m, num_agents=1, outputs=outs)
        env.reset(poses)
        if 'obs' not in outs:
            for io in env._ios: io.obs = None      # no observation download: compute-only host path
        for k in range(10): env.step(acts[k]) if 'obs' in outs else (np.copyto(env._term, env.out['terminated'].numpy()), env._run(acts[k]))
        torch.cuda.synchronize()
        t0 = time.perf_counter(); tc = 0.0
        for k in range(10, 110):
            if 'obs' in outs:
                env.step(acts[k])
            else:
                np.copyto(env._term, env.out['terminated'].numpy()); env._run(acts[k])
        el = (time.perf_counter() - t0) / 100
        print('chunks', chunks, 'outputs', '+'.join(outs), 'ms/step %.3f' % (el * 1e3))
        env.close()
