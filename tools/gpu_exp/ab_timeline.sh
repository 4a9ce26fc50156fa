#!/bin/bash
# Unit timelines of library variants (tools/gpu_exp/libs/*.so): where the lidar launch's time goes per start-time window.
cd "$(dirname "$0")/../.."
for lib in tools/gpu_exp/libs/${1:-*}.so; do
  name=$(basename $lib .so)
  F110_B200_LIB=$PWD/$lib python tools/unit_timeline.py --envs ${2:-4096} --flush > gpurun_out/timeline_$name.txt 2>&1
  cp gpurun_out/unit_timeline_${2:-4096}.npz gpurun_out/unit_timeline_$name.npz
  echo "== $name: $(head -2 gpurun_out/timeline_$name.txt | tr '\n' ' ')"
  python - $name <<'PY'
import sys, numpy as np
d = np.load('gpurun_out/unit_timeline_%s.npz' % sys.argv[1])
s, e, look = d['start_us'], d['end_us'], d['lookups']
dur = e - s
T = e.max()
for a, b in ((20, 60), (60, 80), (80, 90), (90, 95), (95, 100), (100, 105), (105, 110), (110, 120), (120, 140)):
    q = (s >= a) & (s < b) & (look < 24)
    if q.any():
        A = np.vstack([np.ones(q.sum()), look[q]]).T
        c = np.linalg.lstsq(A, dur[q], rcond=None)[0]
        act = ((s <= a) & (e > a)).sum()
        print('  start [%3d,%3d): n=%6d mean %.2f us = %.2f + %.3f/lookup; warps busy at %d us: %d' % (a, b, q.sum(), dur[q].mean(), c[0], c[1], a, act))
PY
done
