"""gymnasium if it is installed, otherwise the few names F110Env needs (Env, spaces.Box, register).

The reference env subclasses gymnasium.Env and registers 'f110-v0' (f110_gym/__init__.py:1-5).  This
image has no gymnasium; the shim keeps the same attribute surface so consumers written against the
reference (train_ddpg.py:58-68, gym_bridge.py:77-80) run unchanged either way.
"""
import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as gym
    from gymnasium import spaces
    from gymnasium.envs.registration import register
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent
    HAVE_GYMNASIUM = False

    class _Box(object):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def __repr__(self):
            return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

    class _Env(object):
        metadata = {}
        render_mode = None
        spec = None

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class _Spaces(object):
        Box = _Box

    class _Gym(object):
        Env = _Env
        spaces = _Spaces

    gym = _Gym
    spaces = _Spaces
    _REGISTRY = {}

    def register(id, entry_point=None, **kwargs):
        _REGISTRY[id] = (entry_point, kwargs)


def make(env_id, **kwargs):
    """gym.make as gymnasium resolves ``'module:id'``: import the module (which registers the id), then build the id's entry
    point.  With gymnasium installed this IS gymnasium.make; without it the shim registry above is used, so
    ``make('f110_gym:f110-v0', ...)`` -- the string every consumer of the reference uses -- behaves the same either way."""
    import importlib
    if HAVE_GYMNASIUM:
        return gym.make(env_id, **kwargs)
    module, _, name = env_id.rpartition(':')
    if module:
        importlib.import_module(module)
    if name not in _REGISTRY:
        raise ValueError("unknown env id %r" % env_id)
    entry_point, reg_kwargs = _REGISTRY[name]
    if isinstance(entry_point, str):
        mod, _, attr = entry_point.partition(':')
        entry_point = getattr(importlib.import_module(mod), attr)
    kw = dict(reg_kwargs.get('kwargs', {}))
    kw.update(kwargs)
    return entry_point(**kw)
