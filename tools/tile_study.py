"""Access pattern of the lidar ray-march on the bench workload, replayed on the CPU (analysis, not product): what a
shared-memory tile around the car could serve, and what a block-linear map layout would save in memory traffic.

north_star asks for "TMA-staged shared-memory tiles where it fits".  Sphere tracing does not walk the map cell by cell, so
whether a tile fits is a property of the workload's lookups, which this script measures exactly (the reference's march,
laser_models.py:106-146, on the C3 start poses):

  * the share of lookups that land within T cells (Chebyshev) of the car's own cell, for the tiles that fit in shared memory
    beside a useful number of resident CTAs (T = 32: 65^2 fp64 cells = 33 KB ... T = 84: 169^2 = 223 KB, one CTA per SM);
    lookup 1 of every ray is the car's cell itself (one broadcast load per warp) and is listed separately;
  * the distinct 32-byte sectors and 128-byte lines one launch touches, with the map stored row-major (sector = 4 x 1
    cells, line = 16 x 1) and block-linear (sector = 2 x 2, line = 4 x 4): the memory traffic of a launch when the map does
    not stay in L2, and the most a re-tiled layout could save.

  python tools/tile_study.py [--envs 4096] [--upsample 1,4] [--out profiles/r02_tile_study.json]
"""
import argparse
import json
import os
import sys

import numpy as np
from numba import njit

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from f110_gymnasium_ros2_jazzy_b200 import workloads   # noqa: E402


@njit(cache=False)
def march(poses, dt, res, ox, oy, sines, cosines, B, inc, fov, rows, cols, first, car_r, car_c):
    H_, W_ = dt.shape
    k = 0
    for p in range(poses.shape[0]):
        ti = 2000 * (poses[p, 2] - fov / 2.) / (2. * np.pi)
        ti = np.fmod(ti, 2000.)
        while ti < 0:
            ti += 2000
        cr = int((poses[p, 1] - oy) / res)
        cc = int((poses[p, 0] - ox) / res)
        for i in range(B):
            s = sines[int(ti)]
            c = cosines[int(ti)]
            x = poses[p, 0]
            y = poses[p, 1]
            total = 0.0
            n = 0
            while True:
                xr = x - ox
                yr = y - oy
                if xr < 0 or xr >= W_ * res or yr < 0 or yr >= H_ * res:
                    r_, c_ = H_ - 1, W_ - 1
                else:
                    r_, c_ = int(yr / res), int(xr / res)
                d = dt[r_, c_]
                if k < rows.shape[0]:
                    rows[k] = r_
                    cols[k] = c_
                    first[k] = n == 0
                    car_r[k] = cr
                    car_c[k] = cc
                k += 1
                n += 1
                total += d
                if not (d > 1e-4 and total <= 30.0):
                    break
                x += d * c
                y += d * s
            ti += inc
            while ti >= 2000:
                ti -= 2000
    return k


def study(envs, upsample):
    dt, res, origin = workloads.shanghai_map(upsample)
    poses = workloads.start_poses(envs)[:, 0]
    th = np.linspace(0, 2 * np.pi, 2000, endpoint=False)    # laser_models.py:379-381 uses theta_dis = 2000
    sines, cosines = np.sin(th), np.cos(th)
    B, fov = 1080, 4.7
    inc = 2000 * (fov / (B - 1)) / (2 * np.pi)
    cap = int(envs * B * 12)
    rows = np.zeros(cap, np.int32); cols = np.zeros(cap, np.int32); first = np.zeros(cap, np.bool_)
    car_r = np.zeros(cap, np.int32); car_c = np.zeros(cap, np.int32)
    n = march(poses, dt, res, origin[0], origin[1], sines, cosines, B, inc, fov, rows, cols, first, car_r, car_c)
    assert n <= cap, (n, cap)
    rows, cols, first, car_r, car_c = rows[:n], cols[:n], first[:n], car_r[:n], car_c[:n]
    W = dt.shape[1]
    out = {'map_cells': '%d x %d' % dt.shape, 'map_mb': dt.nbytes / 1e6, 'envs': envs, 'lookups': int(n),
           'lookups_per_ray': n / (envs * B), 'first_lookup_share': float(first.mean())}
    cheb = np.maximum(np.abs(rows - car_r), np.abs(cols - car_c))
    later = ~first
    out['tile'] = {}
    for T in (16, 32, 48, 64, 84):
        side = 2 * T + 1
        out['tile'][str(T)] = {'tile_kb': side * side * 8 / 1024, 'metres_half_width': T * res,
                               'share_of_all_lookups_incl_first': float((cheb <= T).mean()),
                               'share_of_lookups_2_onwards': float((cheb[later] <= T).mean())}
    def distinct(key):
        return int(np.unique(key).size)
    r64, c64 = rows.astype(np.int64), cols.astype(np.int64)
    lay = {
        'sector_row_major_4x1': r64 * W + (c64 >> 2),
        'sector_block_2x2': (r64 >> 1) * W + (c64 >> 1),
        'line_row_major_16x1': r64 * W + (c64 >> 4),
        'line_block_4x4': (r64 >> 2) * W + (c64 >> 2),
    }
    out['distinct'] = {}
    for name, key in lay.items():
        d = distinct(key)
        unit = 32 if name.startswith('sector') else 128
        out['distinct'][name] = {'count': d, 'mb_per_launch': d * unit / 1e6, 'lookups_per_unit': n / d}
    out['distinct_cells'] = distinct(r64 * W + c64)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=4096)
    ap.add_argument('--upsample', default='1,4')
    ap.add_argument('--out', default='profiles/r02_tile_study.json')
    a = ap.parse_args()
    res = {'what': 'tools/tile_study.py: exact CPU replay of the lidar march on the C3 start poses (Shanghai, 1080 beams)',
           'runs': []}
    for k in [int(v) for v in a.upsample.split(',')]:
        r = study(a.envs, k)
        print(json.dumps(r, indent=1), flush=True)
        res['runs'].append(r)
    json.dump(res, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
