"""Analysis helper (not product): distribution of distance-transform lookups per ray on the bench workload,
and the SIMT inflation (mean over warps of the max over 32 consecutive beams)."""
import sys
import numpy as np
from numba import njit
sys.path.insert(0, '.')
from tests import helpers as H

@njit(cache=False)
def counts(poses, dt, res, ox, oy, sines, cosines, B, inc, fov):
    H_, W_ = dt.shape
    out = np.zeros((poses.shape[0], B), np.int32)
    for p in range(poses.shape[0]):
        ti = 2000 * (poses[p, 2] - fov / 2.) / (2. * np.pi)
        ti = np.fmod(ti, 2000.)
        while ti < 0: ti += 2000
        for i in range(B):
            s = sines[int(ti)]; c = cosines[int(ti)]
            x = poses[p, 0]; y = poses[p, 1]
            n = 0; total = 0.0; d = 1.0
            while True:
                xr = x - ox; yr = y - oy
                if xr < 0 or xr >= W_ * res or yr < 0 or yr >= H_ * res: d = dt[H_-1, W_-1]
                else: d = dt[int(yr / res), int(xr / res)]
                n += 1; total += d
                if not (d > 1e-4 and total <= 30.0): break
                x += d * c; y += d * s
            out[p, i] = n
            ti += inc
            while ti >= 2000: ti -= 2000
    return out

dt, res, orig = H.golden_map('Shanghai_map')
cl = H.load('maps')['Shanghai_map__centerline_poses']
idx = np.linspace(0, len(cl) - 1, 1024).round().astype(int)
s, c, *_ = H.tables()
n = counts(cl[idx], dt, res, orig[0], orig[1], s, c, 1080, 2000 * (4.7 / 1079) / (2 * np.pi), 4.7)
print('mean', n.mean(), 'median', np.median(n), 'p90', np.percentile(n, 90), 'p99', np.percentile(n, 99), 'max', n.max())
flat = n.reshape(-1)
w = flat[: flat.size // 32 * 32].reshape(-1, 32)
print('warp-max mean', w.max(1).mean(), ' => SIMT efficiency', n.mean() / w.max(1).mean())
for cap in (8, 12, 16, 24, 32):
    print('cap', cap, 'rays unfinished %.2f%%' % (100 * (flat > cap).mean()), 'capped warp-max mean', np.minimum(w, cap).max(1).mean(),
          'residual iters share %.1f%%' % (100 * np.maximum(flat - cap, 0).sum() / flat.sum()))
hist = np.bincount(flat)
print('hist', hist[:40])
