"""
Gymnasium interface (reference: /root/reference/loco_mujoco/environments/gymnasium.py:11-173).
`gymnasium` itself is optional: when it is importable the wrapper subclasses gymnasium.Env and is registered as
"LocoMujoco" (reference registration: environments/humanoids/__init__.py:20-25); otherwise the same class works
stand-alone through `loco_mujoco_b200.make_gym("LocoMujoco", env_name=...)`.
"""
import numpy as np

from .base import LocoEnv

try:                                    # pragma: no cover - depends on the host environment
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    _Base = _gym.Env
except Exception:                       # gymnasium is not installed in the build image
    _gym = None
    _Base = object

    class _BoxSpace:
        def __init__(self, low, high, shape, dtype):
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)
            self.shape, self.dtype = tuple(shape), dtype

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi).astype(self.dtype)

    class _spaces:
        Box = _BoxSpace


class GymnasiumWrapper(_Base):
    """step -> (obs, reward, terminated, truncated=False, info); reset(seed, options) -> (obs, {})."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 100}

    def __init__(self, env_name, render_mode=None, **kwargs):
        self.spec = None
        self._env = LocoEnv.make(env_name, **kwargs)
        self.render_mode = render_mode
        self.metadata = dict(self.metadata, render_fps=1.0 / self._env.dt)
        self.observation_space = self._convert_space(self._env.info.observation_space)
        self.action_space = self._convert_space(self._env.info.action_space)

    def step(self, action):
        obs, reward, absorbing, info = self._env.step(action)
        return obs, reward, absorbing, False, info

    def reset(self, *, seed=None, options=None):
        if seed is not None and _gym is not None:
            super().reset(seed=seed)
        return self._env.reset(), {}

    def render(self):
        return self._env.render()

    def close(self):
        self._env.stop()

    def create_dataset(self, **kwargs):
        return self._env.create_dataset(**kwargs)

    def play_trajectory(self, **kwargs):
        return self._env.play_trajectory(**kwargs)

    @property
    def unwrapped(self):
        return self._env

    @staticmethod
    def _convert_space(space):
        low, high = np.min(space.low), np.max(space.high)
        return _spaces.Box(low, high, shape=space.shape, dtype=np.float64) if _gym is None else \
            _spaces.Box(low, high, space.shape, np.float64)


def make_gym(env_id, **kwargs):
    if env_id != "LocoMujoco":
        raise KeyError(env_id)
    return GymnasiumWrapper(**kwargs)


if _gym is not None:                    # pragma: no cover
    try:
        _gym.register("LocoMujoco", entry_point="loco_mujoco_b200.environments.gymnasium:GymnasiumWrapper")
    except Exception:
        pass
