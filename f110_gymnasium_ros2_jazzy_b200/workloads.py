"""The synthetic workloads of BASELINE.json as the package ships them: the Shanghai map of the reference's
rl_training/maps (binarised occupancy, bit-packed, with its yaml metadata) and the centerline start poses of
rl_training/maps/cenerlines/Shanghai_map.csv.  bench.py, smoke() and the tools build their inputs from here; the arrays are
the ones tests/golden/make_golden.py recorded from the reference tree (a CPU test keeps the two copies identical).
Nothing here reads /root/reference.
"""
import functools
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


@functools.lru_cache(maxsize=None)
def _shanghai():
    return dict(np.load(os.path.join(_DATA, 'shanghai_map.npz')))


def shanghai_free_mask():
    """-> (free bool [2000, 2000] with row 0 the bottom image row, resolution, origin[3]), after the reference's flip and
    threshold (laser_models.py:398-404)."""
    m = _shanghai()
    shape = tuple(int(v) for v in m['Shanghai_map__shape'])
    free = np.unpackbits(m['Shanghai_map__bits'])[:shape[0] * shape[1]].reshape(shape).astype(bool)
    return free, float(m['Shanghai_map__resolution']), [float(v) for v in m['Shanghai_map__origin']]


@functools.lru_cache(maxsize=4)
def shanghai_map(upsample=1):
    """-> (dt fp64 [H, W], resolution, origin): resolution * scipy EDT of the 255/0 image, as laser_models.py:40-53,398-425.
    upsample k > 1: BASELINE config C4's 'large maps' -- every pixel replicated k x k, resolution divided by k."""
    from scipy.ndimage import distance_transform_edt
    free, res, origin = shanghai_free_mask()
    if upsample > 1:
        free = np.kron(free, np.ones((upsample, upsample), bool))
        res = res / upsample
    return res * distance_transform_edt(np.where(free, 255., 0.)), res, origin


def centerline_poses():
    """[6687, 3] (x, y, yaw of the segment tangent) along the Shanghai centerline."""
    return _shanghai()['Shanghai_map__centerline_poses']


def start_poses(num_envs, num_agents=1, env_offset=0, total_envs=None, agent_gap=25):
    """env e of `total_envs` starts at centerline row round(linspace(0, 6686, total)[e]); further agents of the
    env start `agent_gap` rows (about 2 m) ahead.  -> [num_envs, num_agents, 3] for envs env_offset .. env_offset + num_envs."""
    cl = centerline_poses()
    total = total_envs or num_envs
    idx = np.linspace(0, len(cl) - 1, total).round().astype(int)[env_offset:env_offset + num_envs]
    poses = np.zeros((num_envs, num_agents, 3))
    for a in range(num_agents):
        poses[:, a] = cl[(idx + agent_gap * a) % len(cl)]
    return poses
