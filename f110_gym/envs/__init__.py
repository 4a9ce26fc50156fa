"""``f110_gym.envs`` as the reference exposes it (f110_gym/envs/__init__.py): the env and the classes below it."""
from f110_gymnasium_ros2_jazzy_b200.env import F110Env
from f110_gymnasium_ros2_jazzy_b200.simulator import Integrator, RaceCar, Simulator

__all__ = ['F110Env', 'Simulator', 'RaceCar', 'Integrator']
