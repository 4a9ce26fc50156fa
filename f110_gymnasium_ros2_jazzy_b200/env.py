"""F110Env ('f110-v0') and its batched form F110VecEnv on the B200 backend.

F110Env keeps the reference's Gymnasium surface (f110_env.py:55-602): the same kwargs with the same
defaults, ``reset(seed=None, options=poses) -> (obs f32[1088], info)``, ``step(action (A, 2)) -> (obs,
reward, terminated, truncated, info)``, the same ``info`` keys and dtypes, ``update_map``,
``update_params``, ``add_render_callback``, ``render``, and the attributes consumers touch (``timestep``,
``sim``, ``lap_times``, ``lap_counts``, ``action_space``, ``observation_space``, ``unwrapped``).

Documented extensions:
  * ``num_agents=1`` works: the flat observation keeps its 1088 layout with the opponent slots = 0
    (the reference raises IndexError in _pack_flat_obs, f110_env.py:554,566).
  * ``obs_mode='scans'`` returns the (A, B) f32 scans instead of the flat vector, the shape
    jazzy_bridge/gym_bridge.py:113-114,265-267 indexes.
  * ``map`` may be given without ``map_dir`` as a path without extension (gym_bridge.py:77-80 style).
  * ``noise`` selects the lidar-noise source, see simulator.Simulator.
  * ``edt='device'`` builds the distance transform with the exact EDT kernel instead of scipy (same bits; update_map
    in milliseconds).
"""
import os

import numpy as np
import torch

from .backend import BatchSim, FAST_OUTPUTS
from .gym_compat import gym, spaces
from .maps import map_bounds
from .params import default_params
from .simulator import Integrator, Simulator

# rendering constants kept for API parity (f110_env.py:48-52)
VIDEO_W, VIDEO_H, WINDOW_W, WINDOW_H = 600, 400, 1000, 800


class F110Env(gym.Env):
    metadata = {'render_modes': ['human', 'human_fast'], 'render_fps': 30}

    renderer = None
    current_obs = None
    render_callbacks = []

    def __init__(self, **kwargs):
        # kwargs extraction, defaults as f110_env.py:104-185
        self.conf = kwargs.get('conf', None)
        self.seed = kwargs.get('seed', 42)
        if 'map_dir' in kwargs and 'map' in kwargs:
            self.map_dir = kwargs['map_dir']
            self.map_name = kwargs['map']
            self.map_path = self.map_dir + self.map_name + '.yaml'
        elif 'map' in kwargs:
            # extension: bare path without extension (the bridge's calling convention)
            base = os.path.splitext(kwargs['map'])[0]
            self.map_dir = os.path.dirname(base) + '/'
            self.map_name = os.path.basename(base)
            self.map_path = base + '.yaml'
        else:
            # the reference falls back to envs/maps/vegas.yaml, a file that does not exist in its tree
            raise FileNotFoundError("F110Env needs map_dir= and map= (the reference's default maps/vegas.yaml is absent)")
        self.map_ext = kwargs.get('map_ext', '.png')
        self.params = kwargs.get('params', None) or default_params()
        self.num_agents = kwargs.get('num_agents', 2)
        self.timestep = kwargs.get('timestep', 0.01)
        self.ego_idx = kwargs.get('ego_idx', 0)
        self.integrator = kwargs.get('integrator', Integrator.RK4)
        self.lidar_dist = kwargs.get('lidar_dist', 0.0)
        self.obs_mode = kwargs.get('obs_mode', 'flat')
        self.render_mode = kwargs.get('render_mode', None)
        noise = kwargs.get('noise', 'numpy')

        self.start_thresh = 0.1
        self.poses_x, self.poses_y, self.poses_theta = [], [], []
        self.collisions = np.zeros((self.num_agents,))
        self.lidar_max = self.params.get("lidar_max", 30.0)
        self.near_start = True
        self.num_toggles = 0
        self.lap_times = np.zeros((self.num_agents,))
        self.lap_counts = np.zeros((self.num_agents,))
        self.current_time = 0.0
        self.near_starts = np.array([True] * self.num_agents)
        self.toggle_list = np.zeros((self.num_agents,))
        self.start_xs = np.zeros((self.num_agents,))
        self.start_ys = np.zeros((self.num_agents,))
        self.start_thetas = np.zeros((self.num_agents,))
        self.start_rot = np.eye(2)

        self.sim = Simulator(self.params, self.num_agents, self.seed, time_step=self.timestep, ego_idx=self.ego_idx,
                             integrator=self.integrator, lidar_dist=self.lidar_dist, noise=noise,
                             device=kwargs.get('device', None), edt=kwargs.get('edt', 'host'))
        self.sim.set_map(self.map_path, self.map_ext)

        self.x_min, self.x_max, self.y_min, self.y_max = map_bounds(self.map_path, self.map_dir)
        self.render_obs = None

        low = np.array([self.params['s_min'], self.params['v_min']], dtype=np.float32)
        high = np.array([self.params['s_max'], self.params['v_max']], dtype=np.float32)
        self.action_space = spaces.Box(low=np.tile(low, (self.num_agents, 1)), high=np.tile(high, (self.num_agents, 1)),
                                       dtype=np.float32)
        nb = self.sim.num_beams
        low = np.array([0.0] * nb + [self.x_min, self.y_min, -np.pi, 0.0, self.x_min, self.y_min, -np.pi, 0.0], dtype=np.float32)
        high = np.array([1.0] * nb + [self.x_max, self.y_max, np.pi, 1.0, self.x_max, self.y_max, np.pi, 1.0], dtype=np.float32)
        self.observation_space = spaces.Box(low=low, high=high, dtype=np.float32)

    # ------------------------------------------------------------------ gym API
    def _finish(self, o):
        """Everything F110Env.step does after Simulator.step (f110_env.py:389-421), read back from the device."""
        A = self.num_agents
        st = o['state'][0]
        obs_dict = self.sim._observations(o)
        # the obs dict carries the lap arrays as they were BEFORE this step's _check_done (f110_env.py:389-390)
        obs_dict['lap_times'] = self.lap_times.astype(np.float32)
        obs_dict['lap_counts'] = self.lap_counts.astype(np.float32)
        F110Env.current_obs = obs_dict
        self.render_obs = {k: obs_dict[k] for k in ('ego_idx', 'poses_x', 'poses_y', 'poses_theta', 'lap_times',
                                                    'lap_counts', 'scans')}
        reward = self.timestep
        self.current_time = float(o['time'][0])
        self.poses_x, self.poses_y, self.poses_theta = obs_dict['poses_x'], obs_dict['poses_y'], obs_dict['poses_theta']
        self.collisions = obs_dict['collisions']
        self.toggle_list = o['toggles'][0].astype(np.float64)
        self.lap_times = o['lap_times'][0].copy()
        self.lap_counts = o['lap_counts'][0].copy()
        terminated = bool(o['terminated'][0])
        toggle_done = self.toggle_list >= 4
        if self.obs_mode == 'scans':
            obs = o['scans_f32'][0].copy()
        else:
            obs = o['obs'][0].copy()
        info = {
            "ego_idx": int(self.ego_idx),
            "poses_x": st[:, 0].astype(np.float32),
            "poses_y": st[:, 1].astype(np.float32),
            "poses_theta": st[:, 4].astype(np.float32),
            "linear_vels_x": st[:, 3].astype(np.float32),
            "linear_vels_y": np.zeros(A, np.float32),
            "ang_vels_z": st[:, 5].astype(np.float32),
            "collisions": o['collisions'][0].astype(np.int8),
            "lap_times": self.lap_times.astype(np.float32),
            "lap_counts": self.lap_counts.astype(np.float32),
            "scans": [o['scans_f32'][0, i].copy() for i in range(A)],
            "checkpoint_done": toggle_done,
            "time": float(self.current_time),
        }
        return obs, reward, terminated, False, info

    def step(self, action):
        action = np.asarray(action)
        if action.dtype != np.float32:
            action = action.astype(np.float64)
        o = self.sim._step_raw(action.reshape(1, self.num_agents, 2))
        return self._finish(o)

    def reset(self, seed=None, options=None):
        poses = options
        if poses is None:
            # the reference dereferences None here (f110_env.py:438,448)
            raise TypeError("'NoneType' object is not subscriptable: F110Env.reset needs options=poses (num_agents, 3)")
        poses = np.asarray(poses, dtype=np.float64)
        if poses.shape[0] != self.num_agents:
            raise ValueError('Number of poses for reset does not match number of agents.')
        self.num_toggles = 0
        self.near_start = True
        self.near_starts = np.array([True] * self.num_agents)
        self.start_xs, self.start_ys, self.start_thetas = poses[:, 0], poses[:, 1], poses[:, 2]
        th = -self.start_thetas[self.ego_idx]
        self.start_rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        # RaceCar.reset re-seeds every car's generator (base_classes.py:204)
        self.sim._rngs = [np.random.default_rng(seed=self.sim.seed) for _ in range(self.num_agents)]
        o = self.sim._step_raw(None, reset_mask=np.ones(1, np.uint8), reset_poses=poses[None])
        obs, reward, terminated, truncated, info = self._finish(o)
        return obs, info

    def update_map(self, map_path, map_ext):
        self.sim.set_map(map_path, map_ext)

    def update_params(self, params, index=-1):
        self.sim.update_params(params, agent_idx=index)

    def add_render_callback(self, callback_func):
        F110Env.render_callbacks.append(callback_func)

    def render(self, mode='human'):
        """The pyglet viewer (rendering.py) is visualisation only and out of scope; callbacks still run."""
        assert mode in ['human', 'human_fast']
        for render_callback in F110Env.render_callbacks:
            render_callback(F110Env.renderer)

    def close(self):
        if getattr(self, 'sim', None) is not None:
            self.sim.backend.close()


class F110VecEnv(object):
    """N F110Envs stepped as one batch on one GPU; torch CUDA tensors in and out (zero-copy).

    ``step(actions [N, A, 2])`` -> ``(obs [N, B+8] f32, reward [N] f32, terminated [N] u8, truncated [N] u8, info)``.
    The returned tensors are persistent buffers rewritten by the next step.  With ``auto_reset=True`` an env
    that terminated is reset to its start poses by the NEXT step call, which is then that env's zero-action
    reset step (gymnasium's next-step autoreset; F110Env.reset semantics f110_env.py:425-472).
    """

    def __init__(self, num_envs, map_dir=None, map=None, map_ext='.png', num_agents=2, params=None, seed=42,
                 timestep=0.01, ego_idx=0, integrator=Integrator.RK4, lidar_dist=0.0, device=None, auto_reset=True,
                 outputs=FAST_OUTPUTS, noise_std=0.01, num_beams=1080, fov=4.7, count_lookups=False, map_arrays=None,
                 cuda_graph=False):
        self.num_envs, self.num_agents = num_envs, num_agents
        self.timestep = timestep
        self.auto_reset = auto_reset
        outs = tuple(dict.fromkeys(tuple(outputs) + ('obs', 'reward', 'terminated')))
        self.backend = BatchSim(num_envs, num_agents, params=params, seed=seed, timestep=timestep, integrator=integrator,
                                ego_idx=ego_idx, lidar_dist=lidar_dist, noise_std=noise_std, device=device, outputs=outs,
                                num_beams=num_beams, fov=fov, count_lookups=count_lookups)
        if map_arrays is not None:
            self.backend.set_map_arrays(*map_arrays)
        else:
            self.backend.set_map(map_dir + map + '.yaml', map_ext)
        self.device = self.backend.device
        self.start_poses = None
        self.truncated = torch.zeros(num_envs, dtype=torch.uint8, device=self.device)
        self.single_observation_shape = (num_beams + 8,)
        self.single_action_shape = (num_agents, 2)
        # cuda_graph=True: the step (3-4 kernel launches) is captured once and replayed, one launch per step on the host side
        self.cuda_graph = cuda_graph
        self._graph = None
        self._act = torch.zeros((num_envs, num_agents, 2), dtype=torch.float32, device=self.device)
        self._steps_eager = 0

    def reset(self, poses, noise=None):
        """poses [N, A, 3] (or [A, 3], broadcast to every env)."""
        p = torch.as_tensor(np.asarray(poses, np.float64) if not torch.is_tensor(poses) else poses, dtype=torch.float64)
        if p.ndim == 2:
            p = p[None].expand(self.num_envs, -1, -1)
        self.start_poses = p.to(self.device).contiguous()
        self._graph = None          # the captured step holds the old start-pose tensor
        self._steps_eager = 0
        o = self.backend.reset(self.start_poses, noise)
        return o['obs'], o

    def step(self, actions, noise=None):
        o = self.backend.out
        if self.cuda_graph and noise is None and self.auto_reset and self.start_poses is not None:
            self._act.copy_(torch.as_tensor(actions, device=self.device).reshape(self._act.shape))
            if self._graph is None and self._steps_eager >= 1:
                # capture on a side stream (the first step ran eagerly and warmed everything up); capture does not execute
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        self.backend.step(self._act, None, reset_mask=o['terminated'], reset_poses=self.start_poses)
                torch.cuda.current_stream(self.device).wait_stream(side)
                self._graph = g
            if self._graph is not None:
                self._graph.replay()
                return o['obs'], o['reward'], o['terminated'], self.truncated, o
            self._steps_eager += 1
            o = self.backend.step(self._act, None, reset_mask=o['terminated'], reset_poses=self.start_poses)
            return o['obs'], o['reward'], o['terminated'], self.truncated, o
        if self.auto_reset:
            # `terminated` of the previous step doubles as this step's reset mask (read by K1 before K3 rewrites it)
            o = self.backend.step(actions, noise, reset_mask=o['terminated'], reset_poses=self.start_poses)
        else:
            o = self.backend.step(actions, noise)
        return o['obs'], o['reward'], o['terminated'], self.truncated, o

    def close(self):
        self.backend.close()


class F110HostVecEnv(object):
    """N F110Envs for a HOST-side consumer: numpy actions in, pinned-host observations out, every step.

    This is the end-to-end shape of the reference's own use (train_ddpg.py:160-202 reads the observation on the
    host every step).  The envs are split over ``chunks`` independent library handles, each with its own stream;
    a step enqueues, per chunk, the action upload, the three kernels and the observation download
    (f110_step_host_async), then waits for all of them (f110_host_sync) -- so one chunk's PCIe traffic overlaps
    the other chunks' kernels.  Auto-reset as in F110VecEnv.
    """

    def __init__(self, num_envs, chunks=2, map_arrays=None, map_dir=None, map=None, map_ext='.png', num_agents=1,
                 seed=42, device=None, outputs=FAST_OUTPUTS, num_beams=1080, **kw):
        chunks = max(1, min(chunks, num_envs))
        self.num_envs, self.num_agents, self.num_beams = num_envs, num_agents, num_beams
        self.bounds = [(num_envs * k) // chunks for k in range(chunks + 1)]
        self.parts = []
        for k in range(chunks):
            n = self.bounds[k + 1] - self.bounds[k]
            b = BatchSim(n, num_agents, seed=seed + 7919 * k, device=device, outputs=('obs',), num_beams=num_beams,
                         host_stream_rank=k + 1, **kw)
            if map_arrays is not None:
                b.set_map_arrays(*map_arrays)
            else:
                b.set_map(map_dir + map + '.yaml', map_ext)
            self.parts.append(b)
        outs = tuple(dict.fromkeys(tuple(outputs) + ('obs', 'reward', 'terminated')))
        self.out = {}
        from .backend import _OUT_SPECS
        for key in outs:
            shape, dtype = _OUT_SPECS[key]
            self.out[key] = torch.zeros(shape(num_envs, num_agents, num_beams), dtype=dtype, pin_memory=True)
        self._views = [{key: t[self.bounds[k]:self.bounds[k + 1]] for key, t in self.out.items()} for k in range(chunks)]
        self._np = {key: t.numpy() for key, t in self.out.items()}     # numpy views of the pinned outputs, made once
        self.start_poses = None
        self._term_t = torch.ones(num_envs, dtype=torch.uint8, pin_memory=True)
        self._term = self._term_t.numpy()

    def _build_ios(self):
        """The F110StepIO of every chunk, built once: all buffers are persistent (pinned outputs, reset mask, start
        poses); only the action pointer changes from step to step."""
        import ctypes as C
        from . import _lib
        K = len(self.parts)
        self._ios = (_lib.F110StepIO * K)()
        self._handles = (C.c_void_p * K)(*[b.h for b in self.parts])
        for k in range(K):
            lo = self.bounds[k]
            io = self._ios[k]
            io.reset_mask = self._term_t.data_ptr() + lo
            io.reset_poses = self._poses_t.data_ptr() + lo * self.num_agents * 3 * 8
            for key, t in self._views[k].items():
                setattr(io, key, t.data_ptr())
        self._lib = _lib

    def _run(self, actions):
        K = len(self.parts)
        if actions is None:
            for k in range(K):
                self._ios[k].actions = None
        else:
            a = np.ascontiguousarray(actions)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            assert a.size == self.num_envs * self.num_agents * 2
            item = a.dtype.itemsize * self.num_agents * 2
            base = a.ctypes.data
            f64 = int(a.dtype == np.float64)
            for k in range(K):
                self._ios[k].actions = base + self.bounds[k] * item
                self._ios[k].actions_f64 = f64
            self._keep = a
        self._lib.check(self._lib.load().f110_step_host_multi(self._handles, self._ios, K))
        return self.out

    def reset(self, poses):
        p = np.asarray(poses, np.float64)
        if p.ndim == 2:
            p = np.broadcast_to(p[None], (self.num_envs,) + p.shape)
        self._poses_t = torch.from_numpy(np.ascontiguousarray(p)).pin_memory()
        self.start_poses = self._poses_t.numpy()
        self._term[:] = 1
        self._build_ios()
        o = self._run(None)
        return self._np['obs'], o

    def step(self, actions):
        """actions: numpy (or pinned tensor viewed as numpy) [N, A, 2] f32/f64."""
        # the previous step's `terminated` (still in the pinned output buffer) is this step's reset mask
        np.copyto(self._term, self._np['terminated'])
        o = self._run(actions)
        return self._np['obs'], self._np['reward'], self._np['terminated'], None, o

    # ---- pipelined use (EnvPool-style): the chunks step independently, so the host can prepare chunk k's next
    # actions while the other chunks' kernels and downloads are in flight ACROSS step boundaries.  Per chunk the
    # order is recv(k) -> read chunk_out(k) -> send(k, actions); a chunk's buffers must not be touched between
    # send and recv.
    def chunk_slice(self, k):
        return slice(self.bounds[k], self.bounds[k + 1])

    def send(self, k, actions):
        """Enqueue one step of chunk k (upload, kernels, download) and return at once.  actions: [n_k, A, 2]."""
        lo, hi = self.bounds[k], self.bounds[k + 1]
        np.copyto(self._term[lo:hi], self._np['terminated'][lo:hi])
        a = np.ascontiguousarray(actions)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        assert a.size == (hi - lo) * self.num_agents * 2
        io = self._ios[k]
        io.actions = a.ctypes.data
        io.actions_f64 = int(a.dtype == np.float64)
        self._keep_k = getattr(self, '_keep_k', {})
        self._keep_k[k] = a
        self._lib.check(self._lib.load().f110_step_host_async(self.parts[k].h, self._ios[k]))

    def recv(self, k):
        """Wait for chunk k's step; returns (obs, reward, terminated) views of the pinned buffers for that chunk."""
        self._lib.check(self._lib.load().f110_host_sync(self.parts[k].h))
        sl = self.chunk_slice(k)
        return self._np['obs'][sl], self._np['reward'][sl], self._np['terminated'][sl]

    def close(self):
        for b in self.parts:
            b.close()
