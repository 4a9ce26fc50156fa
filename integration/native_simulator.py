"""The reference-side binding: a drop-in `Simulator` for f110_gym/envs/base_classes.py backed by libf110_b200.so.

This is the file a maintainer of the reference would add as f110_gym/envs/native_simulator.py and import from
f110_env.py in place of `base_classes.Simulator` (f110_env.py:197-199).  It uses ctypes and numpy only -- no torch, and
nothing from the f110_gymnasium_ros2_jazzy_b200 Python package -- so the C ABI of include/f110_b200.h is the whole
interface.  Same constructor, methods, observation dict and exceptions as base_classes.Simulator (:464-643).
The library path comes from $F110_B200_LIB (default: libf110_b200.so on the loader path).
"""
import ctypes as C
import os

import numpy as np
import yaml
from PIL import Image
from scipy.ndimage import distance_transform_edt as edt          # laser_models.py:32

L = C.CDLL(os.environ.get("F110_B200_LIB", "libf110_b200.so"))
L.f110_last_error.restype = C.c_char_p
KEYS = ['mu', 'C_Sf', 'C_Sr', 'lf', 'lr', 'h', 'm', 'I', 's_min', 's_max', 'sv_min', 'sv_max', 'v_switch', 'a_max', 'v_min',
        'v_max', 'width', 'length']
NUM_BEAMS, FOV, THETA_DIS = 1080, 4.7, 2000                      # base_classes.py:121-123, laser_models.py:356-366


class Cfg(C.Structure):      # struct F110Config
    _fields_ = [(n, C.c_int32) for n in ("abi_version", "device", "num_envs", "num_agents", "num_beams", "theta_dis",
                                         "integrator", "ego_idx")] + \
               [("flags", C.c_uint32), ("host_stream_rank", C.c_uint32)] + \
               [(n, C.c_double) for n in ("fov", "eps", "max_range", "timestep", "lidar_dist", "ttc_thresh", "lidar_max",
                                          "noise_std")] + [("seed", C.c_uint64)]


class IO(C.Structure):       # struct F110StepIO
    _fields_ = [("actions", C.c_void_p), ("actions_f64", C.c_int32), ("host_flags", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("noise", "reset_mask", "reset_poses", "active_mask", "obs", "reward", "terminated",
                                          "scans_f64", "scans_f32", "state", "collisions", "toggles", "lap_times",
                                          "lap_counts", "time", "agent_poses")]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc == 0:
        return
    msg = L.f110_last_error().decode()
    raise {-2: ValueError, -4: IndexError, -5: ValueError, -6: SyntaxError}.get(rc, RuntimeError)(msg)


def _beam_tables(params):
    """RaceCar.__init__ class statics (base_classes.py:122-158): beam angles, their cosines, lidar-to-outline distances."""
    dist_sides, dist_fr = params['width'] / 2., (params['lf'] + params['lr']) / 2.
    inc = FOV / (NUM_BEAMS - 1)
    ang, cos, side = np.zeros(NUM_BEAMS), np.zeros(NUM_BEAMS), np.zeros(NUM_BEAMS)
    for i in range(NUM_BEAMS):
        a = -FOV / 2. + i * inc
        ang[i], cos[i] = a, np.cos(a)
        if a > 0:
            if a < np.pi / 2:
                side[i] = min(dist_sides / np.sin(a), dist_fr / np.cos(a))
            else:
                side[i] = min(dist_sides / np.cos(a - np.pi / 2.), dist_fr / np.sin(a - np.pi / 2.))
        else:
            if a > -np.pi / 2:
                side[i] = min(dist_sides / np.sin(-a), dist_fr / np.cos(-a))
            else:
                side[i] = min(dist_sides / np.cos(-a - np.pi / 2), dist_fr / np.sin(-a - np.pi / 2))
    return ang, cos, side


class Simulator(object):
    def __init__(self, params, num_agents, seed, time_step=0.01, ego_idx=0, integrator=None, lidar_dist=0.0):
        self.num_agents, self.seed, self.params, self.ego_idx = num_agents, seed, params, ego_idx
        self.time_step = time_step
        self.agent_poses = np.empty((num_agents, 3))
        self.collisions = np.zeros((num_agents,))
        cfg = Cfg(L.f110_abi_version(), 0, 1, num_agents, NUM_BEAMS, THETA_DIS, getattr(integrator, "value", integrator or 1), ego_idx, 0, 0,
                  FOV, 1e-4, 30.0, time_step, lidar_dist, 0.005, params.get("lidar_max", 30.0), 0.0, seed)
        self.h = C.c_void_p()
        _check(L.f110_create(C.byref(cfg), _p(np.array([params[k] for k in KEYS], np.float64)), C.byref(self.h)))
        th = np.linspace(0.0, 2 * np.pi, THETA_DIS)              # laser_models.py:379-381
        _check(L.f110_set_tables(self.h, _p(np.sin(th)), _p(np.cos(th))))
        _check(L.f110_set_beam_tables(self.h, *[_p(t) for t in _beam_tables(params)]))
        self.rngs = None

    def __del__(self):
        if getattr(self, 'h', None):
            L.f110_destroy(self.h)
            self.h = None

    def set_map(self, map_path, map_ext):                        # laser_models.py:383-427
        img = np.array(Image.open(os.path.splitext(map_path)[0] + map_ext).transpose(Image.FLIP_TOP_BOTTOM)).astype(np.float64)
        img[img <= 128.] = 0.
        img[img > 128.] = 255.
        meta = yaml.safe_load(open(map_path))
        o, res = meta['origin'], meta['resolution']
        dt = np.ascontiguousarray(res * edt(img))
        _check(L.f110_set_map(self.h, _p(dt), dt.shape[0], dt.shape[1], C.c_double(res), C.c_double(o[0]), C.c_double(o[1]),
                              C.c_double(np.cos(o[2])), C.c_double(np.sin(o[2]))))

    def update_params(self, params, agent_idx=-1):               # base_classes.py:527-547
        _check(L.f110_set_params(self.h, _p(np.array([params[k] for k in KEYS], np.float64)), agent_idx))

    def reset(self, poses):                                      # base_classes.py:627-643
        poses = np.ascontiguousarray(poses, np.float64)
        _check(L.f110_sim_reset_host(self.h, _p(poses), poses.shape[0], None))
        self.rngs = [np.random.default_rng(seed=self.seed) for _ in range(self.num_agents)]   # :204

    def step(self, control_inputs):                              # base_classes.py:566-625
        A = self.num_agents
        act = np.ascontiguousarray(control_inputs, np.float64)
        noise = np.stack([r.normal(0., 0.01, size=NUM_BEAMS) for r in self.rngs])   # the reference's stream, injected
        st, sc, col, ap = np.empty((A, 7)), np.empty((A, NUM_BEAMS)), np.empty(A, np.uint8), np.empty((A, 3))
        io = IO(actions=_p(act), actions_f64=1, noise=_p(noise), state=_p(st), scans_f64=_p(sc), collisions=_p(col),
                agent_poses=_p(ap))
        _check(L.f110_step_host(self.h, C.byref(io)))
        self.agent_poses = ap                                    # recorded before check_ttc (:587 vs :248-250)
        self.collisions = col.astype(np.float64)
        return {'ego_idx': self.ego_idx, 'scans': list(sc), 'poses_x': list(st[:, 0]), 'poses_y': list(st[:, 1]),
                'poses_theta': list(st[:, 4]), 'linear_vels_x': list(st[:, 3]), 'linear_vels_y': [0.] * A,
                'ang_vels_z': list(st[:, 5]), 'collisions': self.collisions}
