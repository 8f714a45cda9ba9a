"""CPU sampler and evaluator for standalone runs of the drop-in trainer (the GPU box has no reference
checkout).  They restate what the trainer needs from the reference's `OffSampler.sample`
(training/off_sampler.py:41-93) and `Evaluator.run_evaluation` (training/evaluator.py:34-83): one env,
a CPU `ApproxContainer` whose policy acts, the same torch-RNG consumption (one [1,A] normal per env step),
the 8-tuple transition format, timeouts stored as done=False."""
import numpy as np
import torch

from .pendulum import env_creator


class Sampler:
    def __init__(self, container_cls, **kwargs):
        self.env = env_creator(**kwargs)
        self.env.seed(kwargs["seed"])
        self.obs, self.info = self.env.reset().astype(np.float32), {}
        self.networks = container_cls(**kwargs)
        self.sample_batch_size = kwargs["batch_size_per_sampler"]
        self.reward_scale = kwargs["reward_scale"]
        self.total_sample_number = 0

    def sample(self):
        self.total_sample_number += self.sample_batch_size
        out = []
        lo, hi = self.env.action_space.low, self.env.action_space.high
        for _ in range(self.sample_batch_size):
            batch_obs = torch.from_numpy(np.expand_dims(self.obs, axis=0).astype("float32"))
            logits = self.networks.policy(batch_obs)
            action, logp = self.networks.create_action_distributions(logits).sample()
            action = np.array(action.detach()[0].numpy())
            logp = logp.detach()[0].numpy()
            next_obs, reward, done, next_info = self.env.step(action.clip(lo, hi).astype(np.float32))
            next_obs = next_obs.astype(np.float32)
            next_info.setdefault("TimeLimit.truncated", False)
            if next_info["TimeLimit.truncated"]:
                done = False
            reward = (reward + 0.0) * self.reward_scale        # env-side shaping wrapper (utils/initialization.py:35-42)
            out.append((self.obs.copy(), self.info, action, self.reward_scale * reward, next_obs.copy(), done, logp, next_info))
            self.obs, self.info = next_obs, next_info
            if done or next_info["TimeLimit.truncated"]:
                self.obs, self.info = self.env.reset().astype(np.float32), {}
        return out, {"Time/Sampler time [ms]-RL iter": 0.0}

    def get_total_sample_number(self):
        return self.total_sample_number


class Evaluator:
    def __init__(self, container_cls, **kwargs):
        self.env = env_creator(**kwargs)
        self.env.seed(kwargs["seed"])
        self.networks = container_cls(**kwargs)
        self.num_eval_episode = kwargs["num_eval_episode"]
        self.returns = []

    def run_an_episode(self):
        obs, done, truncated, total = self.env.reset().astype(np.float32), False, False, 0.0
        rewards = []
        while not (done or truncated):
            logits = self.networks.policy(torch.from_numpy(obs[None].astype("float32")))
            action = self.networks.create_action_distributions(logits).mode().detach().numpy()[0]
            obs, reward, done, info = self.env.step(action.astype(np.float32))
            obs = obs.astype(np.float32)
            truncated = info.get("TimeLimit.truncated", False)
            rewards.append(reward)
        return sum(rewards)

    def run_evaluation(self, iteration):
        r = float(np.mean([self.run_an_episode() for _ in range(self.num_eval_episode)]))
        self.returns.append((iteration, r))
        return r


def loop_kwargs(cfg_kwargs: dict, seed: int, save_folder: str, **over) -> dict:
    """What example_train/main.py + utils/init_args.py hand to the five factories (pendulum defaults)."""
    kw = dict(cfg_kwargs)
    kw.update(env_id="gym_pendulumstandin", trainer="off_serial_trainer", seed=seed, reward_scale=1, reward_shift=None,
              max_episode_steps=None, is_render=False, sampler_name="off_sampler", sample_batch_size=20,
              batch_size_per_sampler=20, noise_params=None, buffer_name="replay_buffer", buffer_warm_size=1000,
              buffer_max_size=100000, replay_batch_size=256, sample_interval=1, evaluator_name="evaluator",
              num_eval_episode=2, eval_interval=50, eval_save=False, max_iteration=120, ini_network_dir=None,
              log_save_interval=1000, apprfunc_save_interval=100000, save_folder=save_folder, additional_info={},
              use_gpu=False, enable_cuda=False, policy_func_name="StochaPolicy", action_type="continu")
    kw.update(over)
    return kw
