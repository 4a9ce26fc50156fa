"""Which non-finite command breaks which configuration (debug aid for test_non_finite_actions)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from tests.test_gpu_parity import GpuBackend
A = int(sys.argv[1]); bad = eval(sys.argv[2], {'nan': np.nan, 'inf': np.inf})
be = GpuBackend(1, A, 'Shanghai_map')
P = np.array([[[0., 0., 0.], [3.0, 0.5, 0.]]])[:, :A]
rng = np.random.default_rng(9)
for t in range(45):
    noise = rng.normal(0, 0.01, size=(1, A, 1080))
    if t == 0:
        g = be.reset(P, noise)
    else:
        act = np.tile(np.array([[0.05, 3.0], [0.0, 2.0]])[:A], (1, 1, 1))
        if t == 20: act[0, 0] = bad
        try:
            g = be.step(act, noise)
        except Exception as e:
            print('FAIL at step', t, type(e).__name__, str(e)[:200]); sys.exit(0)
    if t >= 20: print(t, g['state'][0, 0], flush=True)
print('ok', g['state'][0, 0])
''' % ROOT
for A in (1,):
    for bad in ('[0.05,nan]', '[inf,3.0]'):
        env = dict(os.environ, CUDA_LAUNCH_BLOCKING='1', F110_DEBUG_SYNC='1')
        r = subprocess.run([sys.executable, '-c', CHILD, str(A), bad], capture_output=True, text=True, env=env, timeout=300)
        lines = (r.stdout + r.stderr).strip().splitlines()
        print('A=%d bad=%s rc=%d:' % (A, bad, r.returncode), ' | '.join(lines[-3:])[:600], flush=True)
