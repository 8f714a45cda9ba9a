from typing import TypeVar

ObsType = TypeVar("ObsType")
ActType = TypeVar("ActType")


class Env:
    pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)
