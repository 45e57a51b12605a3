"""gym shim -- TEST ORACLE infrastructure: the three names the reference's
masurvival_env.py uses from gym 0.21 (`gym.Env`, `gym.spaces`,
`gym.spaces.Space`), so the reference package imports without gym installed."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}
    action_space = None
    observation_space = None

    def reset(self, **kwargs):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError

    def close(self):
        pass
