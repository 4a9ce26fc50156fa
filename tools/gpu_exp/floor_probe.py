import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from f110_gymnasium_ros2_jazzy_b200 import BatchSim, _lib
from tests import helpers as H
from oracle.f110_oracle import Oracle
m = H.golden_map('Shanghai_map'); cl = H.load('maps')['Shanghai_map__centerline_poses']
orc = Oracle(1, 1); orc.set_map_arrays(*m)
L = _lib.load()
flush = torch.empty(256 * 2**20, dtype=torch.uint8, device='cuda')
for idx in (0, 700, 1500, 3000, 4200, 5600):
    pose = cl[idx].copy()
    # per-ray lookups from the oracle: scan one beam at a time is not exposed, so use total and a max estimate via sub-scans
    scan, tot = orc.scan(pose)
    sim = BatchSim(1, 1, outputs=('obs',), noise_std=0.0); sim.set_map_arrays(*m)
    sim.reset(pose[None, None])
    zero = torch.zeros((1, 1, 2), device='cuda')
    for warm in (True, False):
        _lib.check(L.f110_set_kernel_timing(sim.h, 1))
        for _ in range(20):
            if not warm: flush.fill_(1)
            sim.sim_reset(pose[None, None]); sim.step(zero)
        torch.cuda.synchronize()
        ms = (C.c_double * 3)(); cnt = C.c_int64(0)
        _lib.check(L.f110_get_kernel_timing(sim.h, ms, C.byref(cnt)))
        print("pose %d: lookups/ray mean %.1f | lidar kernel N=1 %s L2: %.1f us" % (idx, tot / 1080.0, "warm" if warm else "cold", 1e3 * ms[1] / cnt.value))
    sim.close()
