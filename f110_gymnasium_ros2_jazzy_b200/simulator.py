"""Simulator: the reference's lower boundary (base_classes.py:464-643) on top of the CUDA backend.

Same constructor, same methods, same return dictionary and the same exceptions; the physics, lidar,
GJK and iTTC run in libf110_b200.so.  One instance drives N = 1 env (the reference's shape); the
batched form is backend.BatchSim / env.F110VecEnv.
"""
from enum import Enum

import numpy as np
import torch

from .backend import ALL_OUTPUTS, BatchSim


class Integrator(Enum):   # base_classes.py:40-42
    RK4 = 1
    Euler = 2


class RaceCar(object):
    """Read-only view of one vehicle of a Simulator, with the attributes of the reference's RaceCar that its consumers
    and tests read (base_classes.py:45-158): ``state`` fp64[7] = [x, y, steer, v, yaw, yaw_rate, slip], ``params``,
    ``in_collision`` (iTTC), ``is_ego``, ``time_step``, ``num_beams``, ``fov``, ``ttc_thresh``, ``steer_buffer_size``.
    The vehicle itself lives on the device; ``state`` is the host copy of the last step's output."""
    ttc_thresh = 0.005          # base_classes.py:115
    steer_buffer_size = 2       # base_classes.py:109

    def __init__(self, sim, index):
        self._sim, self._index = sim, index

    @property
    def state(self):
        return self._sim._state[self._index].copy()

    @property
    def params(self):
        return self._sim._agent_params[self._index]

    @property
    def in_collision(self):
        return bool(self._sim._ttc_collisions[self._index])

    @property
    def is_ego(self):
        return self._index == self._sim.ego_idx

    @property
    def time_step(self):
        return self._sim.time_step

    @property
    def num_beams(self):
        return self._sim.num_beams

    @property
    def fov(self):
        return self._sim.fov


class Simulator(object):
    """Drop-in for f110_gym.envs.base_classes.Simulator.

    Lidar noise: ``noise='numpy'`` (default) draws ``Generator.normal(0, 0.01, num_beams)`` per agent per
    step on the host from ``default_rng(seed)`` re-seeded on reset -- bit-identical to the reference's
    stream (base_classes.py:119,204; laser_models.py:450-452) -- and uploads it; ``noise='device'`` uses the
    library's Philox stream (same N(0, 0.01^2) law, different bits); ``noise=None`` disables it.
    """

    def __init__(self, params, num_agents, seed, time_step=0.01, ego_idx=0, integrator=Integrator.RK4, lidar_dist=0.0,
                 noise='numpy', device=None, num_beams=1080, fov=4.7, edt='host'):
        if not isinstance(integrator, Integrator) and integrator not in (1, 2):
            raise SyntaxError("Invalid Integrator Specified. Provided %s. Please choose RK4 or Euler" % (integrator,))
        self.num_agents = num_agents
        self.seed = seed
        self.time_step = time_step
        self.ego_idx = ego_idx
        self.params = params
        self.num_beams = num_beams
        self.fov = fov
        self.noise_mode = noise
        self.edt = edt    # 'host': scipy as the reference; 'device': exact EDT kernel (bit-identical map, ~25x faster)
        self.agent_poses = np.empty((self.num_agents, 3))
        self.collisions = np.zeros((self.num_agents,))
        self.collision_idx = -1 * np.ones((self.num_agents,))
        self.backend = BatchSim(1, num_agents, params=params, seed=seed, timestep=time_step, integrator=integrator,
                                ego_idx=ego_idx, lidar_dist=lidar_dist, num_beams=num_beams, fov=fov,
                                noise_std=0.01 if noise == 'device' else 0.0, device=device, outputs=ALL_OUTPUTS)
        # every RaceCar owns default_rng(seed) (all agents share the seed => identical streams)
        self._rngs = None
        self._pending_reset = None
        self.last = None
        self._blocks = {}        # action dtype -> pinned host blocks
        self._state = np.zeros((self.num_agents, 7))
        self._ttc_collisions = np.zeros(self.num_agents, bool)
        self._agent_params = [dict(params) for _ in range(self.num_agents)]
        self.agents = [RaceCar(self, i) for i in range(self.num_agents)]   # base_classes.py:504-510

    def set_map(self, map_path, map_ext):
        self.backend.set_map(map_path, map_ext, edt=self.edt)

    def update_params(self, params, agent_idx=-1):
        if agent_idx >= self.num_agents:
            raise IndexError('Index given is out of bounds for list of agents.')
        self.backend.update_params(params, agent_idx)
        for i in range(self.num_agents):
            if agent_idx < 0 or i == agent_idx:
                self._agent_params[i] = dict(params)

    def reset(self, poses):
        poses = np.asarray(poses, dtype=np.float64)
        if poses.shape[0] != self.num_agents:
            raise ValueError('Number of poses for reset does not match number of agents.')
        self.backend.sim_reset(poses[None])
        torch.cuda.current_stream(self.backend.device).synchronize()   # the host-buffer step runs on the library's own stream
        self._rngs = [np.random.default_rng(seed=self.seed) for _ in range(self.num_agents)]
        self._state = np.zeros((self.num_agents, 7))
        self._state[:, 0:2] = poses[:, 0:2]
        self._state[:, 4] = poses[:, 2]
        self._ttc_collisions = np.zeros(self.num_agents, bool)

    def _fill_noise(self, dst):
        """dst [1, A, B]: one normal(0, 0.01) draw per car from its own generator (laser_models.py:450-452)."""
        if self._rngs is None:
            raise AttributeError("scan_rng is only created by reset() for cars other than the first (base_classes.py:119,204)")
        for i, r in enumerate(self._rngs):
            dst[0, i] = r.normal(0., 0.01, size=self.num_beams)

    def _step_raw(self, control_inputs, reset_mask=None, reset_poses=None):
        """One f110_step_host call: actions / noise up, kernels, every output down, one sync.  Inputs and outputs live
        in one pinned block per direction (BatchSim.host_blocks), so each direction is a single PCIe copy.
        The returned arrays are views of the output block (rewritten by the next step); callers copy what they keep."""
        dt = np.dtype(np.float32) if control_inputs is None else control_inputs.dtype
        hb = self._blocks.get(dt)
        if hb is None:
            hb = self._blocks[dt] = self.backend.host_blocks(ALL_OUTPUTS, actions_dtype=dt, noise=self.noise_mode == 'numpy')
        if control_inputs is None:
            hb.actions[...] = 0
        else:
            hb.actions[...] = control_inputs
        if hb.noise is not None:
            self._fill_noise(hb.noise)
        if reset_mask is None:
            hb.reset_mask[...] = 0
        else:
            hb.reset_mask[...] = reset_mask
            hb.reset_poses[...] = reset_poses
        self.last = self.backend.step_host_blocks(hb)
        return self.last

    def step(self, control_inputs):
        control_inputs = np.asarray(control_inputs)
        if control_inputs.dtype != np.float32:
            control_inputs = control_inputs.astype(np.float64)
        o = self._step_raw(control_inputs.reshape(1, self.num_agents, 2))
        return self._observations(o)

    def _observations(self, o):
        st = o['state'][0]
        # recorded before check_ttc zeroes a colliding car's yaw (base_classes.py:587 vs :248-250)
        self.agent_poses = o['agent_poses'][0].copy()
        self._state = st
        # RaceCar.in_collision is the iTTC flag alone; an iTTC hit zeroes state[3:] (base_classes.py:246-252), GJK does not
        self._ttc_collisions = (o['collisions'][0] != 0) & np.all(st[:, 3:] == 0, axis=1)
        self.collisions = o['collisions'][0].astype(np.float64)
        observations = {'ego_idx': self.ego_idx,
                        'scans': [o['scans_f64'][0, i].copy() for i in range(self.num_agents)],
                        'poses_x': [st[i, 0] for i in range(self.num_agents)],
                        'poses_y': [st[i, 1] for i in range(self.num_agents)],
                        'poses_theta': [st[i, 4] for i in range(self.num_agents)],
                        'linear_vels_x': [st[i, 3] for i in range(self.num_agents)],
                        'linear_vels_y': [0. for _ in range(self.num_agents)],
                        'ang_vels_z': [st[i, 5] for i in range(self.num_agents)],
                        'collisions': self.collisions}
        return observations
