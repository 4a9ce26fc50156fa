"""Drop-in module id for the reference's ``gym.make('f110_gym:f110-v0', ...)`` (f110_gymnasium/gym/f110_gym/__init__.py:1-5).

gymnasium resolves ``'f110_gym:f110-v0'`` by importing a module called ``f110_gym`` and looking the id up afterwards; the
reference's consumers (rl_training/train_ddpg.py:58-65, jazzy_bridge gym_bridge.py:77-80) use exactly that string.  With
this alias on the path the id resolves to the B200 ``F110Env``; nothing else lives here.
"""
from f110_gymnasium_ros2_jazzy_b200.gym_compat import register

try:
    register(id='f110-v0', entry_point='f110_gym.envs:F110Env')
except Exception:   # gymnasium raises when the id is registered twice
    pass
