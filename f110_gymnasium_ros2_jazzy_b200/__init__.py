"""f110_gymnasium_ros2_jazzy_b200 -- the F110Env step hot path on B200 (sm_100a).

Public surface
  F110Env, Simulator, Integrator   drop-ins for f110_gym.envs.{F110Env, Simulator, Integrator}
  F110VecEnv, BatchSim             the batched, device-resident forms
  make('f110_gym:f110-v0', **kw)   works with or without gymnasium installed
  shard_range / EpisodeStats       multi-GPU sharding by env index + NCCL-reduced episode statistics
"""
from .gym_compat import HAVE_GYMNASIUM, make, register
from .simulator import Integrator, Simulator
from .env import F110Env, F110HostVecEnv, F110VecEnv
from .backend import ALL_OUTPUTS, FAST_OUTPUTS, BatchSim
from .params import default_params
from .dist import EpisodeStats, shard_range
from .rollout import Actor, DeviceReplayBuffer, DeviceRollout, ShapedReward, gap_follow_actions

# same id the reference registers (f110_gym/__init__.py:2-5)
try:
    register(id='f110-v0', entry_point='f110_gymnasium_ros2_jazzy_b200.env:F110Env')
except Exception:  # already registered
    pass

__all__ = ['F110Env', 'F110VecEnv', 'F110HostVecEnv', 'Simulator', 'Integrator', 'BatchSim', 'make', 'default_params', 'shard_range',
           'EpisodeStats', 'Actor', 'DeviceReplayBuffer', 'DeviceRollout', 'ShapedReward', 'gap_follow_actions', 'ALL_OUTPUTS', 'FAST_OUTPUTS', 'HAVE_GYMNASIUM']
