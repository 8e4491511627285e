"""Minimal vectorised actor over fake uint8-frame environments (test infrastructure).

Implements the reference's ActingInterface surface used by trainers
(rltime/acting/acting_interface.py:2-90) and produces samples exactly the way
rltime/acting/actor.py:97-149 does: policy.actor_predict on the last state, env step,
policy.make_input_state(obs, dones), one dict per env."""
import numpy as np


class _Space:
    def __init__(self, shape=None, n=None, spaces=None):
        self.shape, self.n = shape, n
        if spaces is not None:
            self.spaces = spaces      # gym.spaces.Tuple surface


class FakeVecActor:
    def __init__(self, num_envs=4, frame_shape=(4, 84, 84), num_actions=4, seed=0, done_p=0.05, eps=0.3,
                 extra_dim=0, vector_obs=False):
        """extra_dim > 0: tuple observations (frame, extra float32 vector), the layout of
        ExtraFeaturesEnvWrapper (rltime/env_wrappers/common.py:221-233); vector_obs: 1-D float32
        observations (CartPole-like, configs/cartpole_*.json)."""
        self.num_envs, self.frame_shape, self.num_actions = num_envs, tuple(frame_shape), num_actions
        self.extra_dim, self.vector_obs = extra_dim, vector_obs
        self.rs = np.random.RandomState(seed)
        if vector_obs:
            self.pool = self.rs.randn(32, *self.frame_shape).astype(np.float32)
        else:
            self.pool = self.rs.randint(0, 255, (32,) + self.frame_shape).astype(np.uint8)
        self.done_p, self.eps = done_p, eps
        self.policy = None
        self.last_state = None
        self.progress = 0.0
        self.updates = 0

    def get_spaces(self):
        obs = _Space(shape=self.frame_shape)
        if self.extra_dim:
            obs = _Space(spaces=(obs, _Space(shape=(self.extra_dim,))))
        return obs, _Space(n=self.num_actions)

    def _obs(self):
        o = self.pool[self.rs.randint(0, 32, self.num_envs)]
        if self.extra_dim:
            return (o, self.rs.randn(self.num_envs, self.extra_dim).astype(np.float32))
        return o

    def get_env_count(self):
        return self.num_envs

    def set_actor_policy(self, policy):
        self.policy = policy
        self.last_state = policy.make_input_state(self._obs(), np.ones(self.num_envs, dtype=bool))

    def update_state(self, progress, policy_state=None):
        self.progress = progress
        self.updates += 1

    def close(self):
        pass

    def get_samples(self, min_samples):
        iters = (max(1, min_samples) + self.num_envs - 1) // self.num_envs
        samples = []
        for _ in range(iters):
            pred = self.policy.actor_predict(self.last_state, timesteps=1)
            explore = self.rs.rand(self.num_envs) < self.eps
            pred["actions"] = np.where(explore, self.rs.randint(0, self.num_actions, self.num_envs),
                                       pred["actions"]).astype(np.int64)
            obs = self._obs()
            rewards = self.rs.randn(self.num_envs) + (pred["actions"] == 1)   # action 1 pays
            dones = self.rs.rand(self.num_envs) < self.done_p
            states = self.policy.make_input_state(obs, np.array(dones))
            for i in range(self.num_envs):
                def take(tree):
                    if isinstance(tree, dict):
                        return {k: take(v) for k, v in tree.items()}
                    if isinstance(tree, (tuple, list)):
                        return tuple(take(v) for v in tree)
                    return tree[i]
                samples.append({"policy_output": take(pred), "next_state": take(states),
                                "reward": rewards[i], "done": dones[i], "info": {}, "env_id": i})
            self.last_state = states
        return samples
