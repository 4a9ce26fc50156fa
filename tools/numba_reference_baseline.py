"""The CPU baseline BASELINE.json's north_star names, timed with the UNMODIFIED reference: numba F110Env code under a
one-process-per-core lock-step vector runner (build container only: needs /root/reference and numba).

    python tools/numba_reference_baseline.py [--workers 0] [--steps 2000] -> profiles/numba_reference.json

gymnasium is not installed here, so gymnasium.vector.AsyncVectorEnv itself cannot run; this is the
equivalent the survey prescribes: one worker process per host core, each holding ONE reference env (RaceCar.scan_simulator is
a class-level static, base_classes.py:63-67), a pipe per worker, step() = send every worker its action, then collect every
observation (AsyncVectorEnv.step_async / step_wait), auto-reset to the start pose on done.

Two shapes are timed on the C3 workload (Shanghai map, 1080 beams, poses spread over the centerline, iid uniform actions):
  * sim1: Simulator.step with ONE agent -- the shape of BASELINE config C3 (F110Env itself cannot pack a 1-agent observation,
    f110_env.py:554,566) -- returning scan + pose, reset on collision;
  * env2: F110Env.step with the reference's default two agents, flat observation returned, reset on done.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
S_MIN, S_MAX, V_MIN, V_MAX = -0.4189, 0.4189, 0.0, 20.0


def worker(conn, shape, pose, seed):
    import warnings
    warnings.filterwarnings('ignore')
    from ref_loader import REF_MAPS, fresh_statics, load_reference
    ns = load_reference()
    fresh_statics(ns)
    if shape == 'sim1':
        params = ns.F110Env(map_dir=REF_MAPS, map='Shanghai_map', map_ext='.png', num_agents=2).params
        fresh_statics(ns)
        sim = ns.Simulator(params, 1, seed, time_step=0.01, integrator=ns.Integrator.RK4)
        sim.set_map(REF_MAPS + 'Shanghai_map.yaml', '.png')
        sim.reset(pose[None, :])

        def step(a):
            o = sim.step(a[None, :])
            done = bool(o['collisions'][0])
            if done:
                sim.reset(pose[None, :])
            return np.float32(o['scans'][0]), done
    else:
        env = ns.F110Env(map_dir=REF_MAPS, map='Shanghai_map', map_ext='.png', num_agents=2)
        poses = np.stack([pose, pose + np.array([3.0 * np.cos(pose[2]), 3.0 * np.sin(pose[2]), 0.0])])
        env.reset(options=poses)

        def step(a):
            obs, r, term, trunc, info = env.step(np.stack([a, np.array([0.0, 1.5], np.float32)]))
            if term:
                obs, info = env.reset(options=poses)
            return obs, bool(term)
    step(np.array([0.0, 1.0], np.float32))       # numba JIT warm-up
    conn.send('ready')
    while True:
        a = conn.recv()
        if a is None:
            break
        conn.send(step(a))
    conn.close()


def run(shape, workers, steps):
    m = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'maps.npz')))
    cl = m['Shanghai_map__centerline_poses']
    poses = cl[np.linspace(0, len(cl) - 1, workers).round().astype(int)]
    ctx = mp.get_context('fork')
    pipes, procs = [], []
    for k in range(workers):
        a, b = ctx.Pipe()
        p = ctx.Process(target=worker, args=(b, shape, poses[k].copy(), 42 + k), daemon=True)
        p.start()
        pipes.append(a); procs.append(p)
    for c in pipes:
        assert c.recv() == 'ready'
    rng = np.random.default_rng(1234)
    acts = rng.uniform([S_MIN, V_MIN], [S_MAX, V_MAX], size=(steps + 50, workers, 2)).astype(np.float32)
    dones = 0
    for t in range(50):                          # warm-up steps
        for k, c in enumerate(pipes):
            c.send(acts[t, k])
        for c in pipes:
            c.recv()
    t0 = time.perf_counter()
    for t in range(50, 50 + steps):
        for k, c in enumerate(pipes):            # step_async
            c.send(acts[t, k])
        for c in pipes:                          # step_wait
            dones += int(c.recv()[1])
    el = time.perf_counter() - t0
    for c in pipes:
        c.send(None)
    for p in procs:
        p.join(timeout=5)
    agents = 1 if shape == 'sim1' else 2
    return {'shape': shape, 'workers': workers, 'steps': steps, 'seconds': el, 'env_steps_per_s': workers * steps / el,
            'env_steps_per_s_per_core': steps / el, 'rays_per_s': workers * steps * agents * 1080 / el, 'episodes': dones}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--workers', type=int, default=0)
    ap.add_argument('--steps', type=int, default=2000)
    a = ap.parse_args()
    cores = len(os.sched_getaffinity(0))
    w = a.workers or cores
    import platform
    out = {'where': 'build container (NOT the GPU box): unmodified reference, numba %s' % __import__('numba').__version__,
           'cpu': platform.processor() or platform.machine(), 'cores': cores,
           'runner': 'multiprocessing pipes, one reference env per process, lock-step (AsyncVectorEnv semantics; gymnasium is not installed)',
           'workload': 'C3 sample: Shanghai_map, 1080 beams, centerline start poses, uniform random actions, reset on done',
           'results': [run('sim1', w, a.steps), run('env2', w, max(200, a.steps // 2)), run('sim1', 1, max(200, a.steps // 2))]}
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'numba_reference.json'), 'w'), indent=1)
    print(json.dumps(out, indent=1))
