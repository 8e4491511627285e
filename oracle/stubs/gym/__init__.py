"""Minimal stand-in for the `gym` package (TEST INFRASTRUCTURE ONLY).

The reference (opherlieber/rltime) imports gym at module import time
(rltime/models/torch/torch_model.py:2, rltime/env_wrappers/common.py:1-11) but
the replay/learner hot path never steps a gym env.  This stub provides just the
names those imports need so the unmodified reference can be executed in the
build container as the parity oracle (oracle/gen_golden.py).  It never ships
with the product package.
"""
from . import spaces  # noqa: F401


class Env:
    observation_space = None
    action_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = env.observation_space
        self.action_space = env.action_space

    def __getattr__(self, name):
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        obs, rew, done, info = self.env.step(action)
        return self.observation(obs), rew, done, info

    def observation(self, observation):
        raise NotImplementedError


def make(*args, **kwargs):
    raise RuntimeError("gym stub: no real environments are available")
