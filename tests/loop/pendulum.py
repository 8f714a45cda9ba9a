"""Pendulum stand-in for loop-parity tests (SURVEY.md §8c "Env-side oracle"): neither gym nor its
Pendulum-v1 exist in this image, so the reference loop (golden generation) and the B200 loop (GPU test)
both run THIS environment.  It exposes exactly what the reference touches: reset(), step(a) ->
(obs, r, done, info), seed(s), action_space / observation_space with low/high/shape/dtype, and a 200-step
time limit that reports `info["TimeLimit.truncated"]` (reference utils/wrapping_env.py:101-107,
training/off_sampler.py:69-73).  Dynamics: the textbook torque-limited pendulum swing-up."""
import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)


class PendulumStandIn:
    max_speed, max_torque, dt, g, m, l = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0

    def __init__(self):
        self.action_space = Box([-self.max_torque], [self.max_torque])
        self.observation_space = Box([-1.0, -1.0, -self.max_speed], [1.0, 1.0, self.max_speed])
        self.rng = np.random.default_rng(0)
        self.state = np.zeros(2)

    def seed(self, seed=None):
        self.rng = np.random.default_rng(seed)
        return [seed]

    def _obs(self):
        th, thdot = self.state
        return np.array([np.cos(th), np.sin(th), thdot], dtype=np.float32)

    def reset(self, **kwargs):
        self.state = self.rng.uniform(low=[-np.pi, -1.0], high=[np.pi, 1.0])
        return self._obs()

    def step(self, u):
        th, thdot = self.state
        u = float(np.clip(np.asarray(u, dtype=np.float64).reshape(-1)[0], -self.max_torque, self.max_torque))
        ang = ((th + np.pi) % (2 * np.pi)) - np.pi
        cost = ang ** 2 + 0.1 * thdot ** 2 + 0.001 * u ** 2
        thdot = thdot + (3 * self.g / (2 * self.l) * np.sin(th) + 3.0 / (self.m * self.l ** 2) * u) * self.dt
        thdot = float(np.clip(thdot, -self.max_speed, self.max_speed))
        th = th + thdot * self.dt
        self.state = np.array([th, thdot])
        return self._obs(), -float(cost), False, {}


class StepLimit:
    """gym.wrappers.TimeLimit semantics (gym 0.23): done after `max_steps`, truncated = not already done."""

    def __init__(self, env, max_steps=200):
        self.env, self.max_steps, self.elapsed = env, max_steps, 0
        self.action_space, self.observation_space = env.action_space, env.observation_space

    def seed(self, seed=None):
        return self.env.seed(seed)

    def reset(self, **kwargs):
        self.elapsed = 0
        return self.env.reset(**kwargs)

    def step(self, action):
        obs, rew, done, info = self.env.step(action)
        self.elapsed += 1
        if self.elapsed >= self.max_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return obs, rew, done, info

    @property
    def state(self):
        return self.env.state


def env_creator(**kwargs):
    return StepLimit(PendulumStandIn(), 200)
