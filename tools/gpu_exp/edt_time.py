"""Device EDT against scipy: wall time of BatchSim.set_map_image (upload of the 1-byte mask, exact EDT on
the device, padding and install -- the whole call a user makes) and of scipy.ndimage.distance_transform_edt on this box's
host, on the Shanghai occupancy at 2000^2 and replicated to 4000^2 and 8000^2, with the result compared bit for bit.

  python tools/gpu_exp/edt_time.py [--sizes 1,2,4] [--out gpurun_out/edt_time.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='1,2,4')
    ap.add_argument('--out', default='gpurun_out/edt_time.json')
    ap.add_argument('--reps', type=int, default=5)
    a = ap.parse_args()
    import torch
    from scipy.ndimage import distance_transform_edt
    from f110_gymnasium_ros2_jazzy_b200 import BatchSim, workloads
    free0, res0, origin = workloads.shanghai_free_mask()
    sim = BatchSim(1, 1)
    rows = []
    for k in [int(v) for v in a.sizes.split(',')]:
        free = np.kron(free0, np.ones((k, k), bool)) if k > 1 else free0
        res = res0 / k
        mask = np.ascontiguousarray(free, np.uint8)
        t0 = time.perf_counter()
        ref = res * distance_transform_edt(np.where(free, 255., 0.))
        t_scipy = time.perf_counter() - t0
        row = {'cells': '%d x %d' % mask.shape, 'scipy_s': t_scipy}
        # 'serial': the general form (one thread per column, then one per row -- the round-1 kernels), forced
        for label, wide in (('parallel', '0'), ('serial', '1')):
            os.environ['F110_EDT_WIDE'] = wide
            sim.set_map_image(mask, res, origin)          # warm-up (allocations)
            torch.cuda.synchronize()
            ts, ks = [], []
            for _ in range(a.reps):
                t0 = time.perf_counter()
                sim.set_map_image(mask, res, origin)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
                ks.append(sim.edt_kernel_ms())
            row[label] = {'set_map_image_ms': 1e3 * min(ts), 'edt_kernels_ms': min(ks),
                          'bit_identical': bool(np.array_equal(sim.get_map(), ref))}
        os.environ['F110_EDT_WIDE'] = '0'
        # bytes the EDT must move at least: 1 read + 8 written per cell
        row['hbm_floor_ms'] = 9 * mask.size / 6.5e9
        rows.append(row)
        print(row, flush=True)
        del ref
    os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
    json.dump({'what': __doc__.split('\n\n')[0], 'rows': rows}, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
