"""Soak run: (1) a long auto-resetting batch checked for invariants, (2) a 64-env slice replayed on the oracle for
thousands of steps with injected noise (flags exact, state 1e-9, lidar 1e-6 on >= 99.9 % of beams)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f110_gymnasium_ros2_jazzy_b200 import ALL_OUTPUTS, BatchSim, F110VecEnv
from oracle.f110_oracle import Oracle
from tests import helpers as H

m = H.golden_map('Shanghai_map'); cl = H.load('maps')['Shanghai_map__centerline_poses']
# ---- (1) invariants over a long run
N, T = 4096, 20000
poses = cl[np.linspace(0, len(cl) - 1, N).round().astype(int)][:, None, :]
env = F110VecEnv(N, num_agents=1, map_arrays=m, outputs=('obs', 'reward', 'terminated', 'state'))
env.reset(poses)
g = torch.Generator(device='cuda'); g.manual_seed(7)
lo = torch.tensor([-0.4189, 0.0], device='cuda'); hi = torch.tensor([0.4189, 20.0], device='cuda')
t0 = time.perf_counter(); bad = 0
for t in range(T):
    act = lo + torch.rand((N, 1, 2), generator=g, device='cuda') * (hi - lo)
    obs, r, term, trunc, info = env.step(act)
    if t % 500 == 0:
        ok = bool(torch.isfinite(info['state']).all()) and float(obs[:, :1080].min()) >= 0.0 and float(obs[:, :1080].max()) <= 1.0
        bad += 0 if ok else 1
torch.cuda.synchronize()
st = env.backend.stats().cpu().numpy()
print("soak 1: %d envs x %d steps in %.1f s, invariant violations %d, episodes %.0f, mean episode steps %.1f"
      % (N, T, time.perf_counter() - t0, bad, st[0], st[1] / max(st[0], 1)))
# ---- (2) long differential run against the oracle
N, T = 64, 3000
poses = cl[np.linspace(0, len(cl) - 1, N).round().astype(int)][:, None, :]
sim = BatchSim(N, 1, outputs=ALL_OUTPUTS, noise_std=0.0); sim.set_map_arrays(*m)
orc = Oracle(N, 1, threads=os.cpu_count()); orc.set_map_arrays(*m)
rng = np.random.default_rng(99)
outl = beams = 0; worst = 0.0; term = np.ones(N, np.uint8)
for t in range(T):
    noise = rng.normal(0, 0.01, size=(N, 1, 1080))
    act = rng.uniform([-0.4189, 0], [0.4189, 14], size=(N, 1, 2)).astype(np.float32)
    a = sim.step(act, noise, term.copy(), poses); b = orc.step(act, noise, term.copy(), poses)
    torch.cuda.synchronize()
    for k in ('collisions', 'terminated', 'toggles'):
        assert np.array_equal(a[k].cpu().numpy(), b[k]), (k, t)
    worst = max(worst, np.abs(a['state'].cpu().numpy() - b['state']).max())
    d = np.abs(a['scans_f64'].cpu().numpy() - b['scans']); outl += int((d > 1e-6).sum()); beams += d.size
    term = b['terminated'].copy()
assert worst <= 1e-9 and outl <= 1e-3 * beams
print("soak 2: %d envs x %d steps vs oracle: flags exact, worst state diff %.2e, lidar outliers %d / %d" % (N, T, worst, outl, beams))
