"""Seeded synthetic transition streams.

One object produces the *same* transitions in two shapes:

* ``next_samples()``  -> list of per-env sample dicts in the reference's acting
  schema (rltime/acting/acting_interface.py:58-90, rltime/acting/actor.py:132-145):
  ``{policy_output{actions,qvalues}, next_state{x, layer0_state{}, layer1_state{hx,cx,initials},
  layer2_state{}}, reward, done, info, env_id}``
* the raw numpy arrays behind them (``last_arrays``) for batched ingest.

Used by the golden-vector generator, the parity tests and bench.py so that the
reference, the oracle and the CUDA path all see identical seeded transitions.
"""
import numpy as np


class SyntheticStream:
    def __init__(self, num_envs=32, frame_shape=(4, 84, 84), num_actions=6,
                 lstm_units=512, seed=1, done_mode="periodic", done_period=500,
                 done_p=0.01, pool=256, recurrent=True, env_id_base=0,
                 clip_rewards=False, pooled_state=False):
        self.num_envs = int(num_envs)
        self.frame_shape = tuple(frame_shape)
        self.num_actions = int(num_actions)
        self.lstm_units = int(lstm_units)
        self.recurrent = bool(recurrent)
        self.done_mode = done_mode
        self.done_period = int(done_period)
        self.done_p = float(done_p)
        self.env_id_base = env_id_base
        self.clip_rewards = clip_rewards
        # pooled_state: hx / cx rows are views into a 256-row pool (a 1M-transition host-side
        # reference then holds 1M references instead of 4 GB of LSTM states)
        self.pooled_state = bool(pooled_state)
        rs = np.random.RandomState(seed)
        # Pooled frames: frame(g) = pool[g & (pool-1)] keeps a 1M-transition CPU
        # reference within RAM while the device gather cost is unchanged.
        assert pool & (pool - 1) == 0
        self.pool = rs.randint(0, 255, (pool,) + self.frame_shape).astype(np.uint8)
        self._rs = rs
        self._state_pool = (np.random.RandomState(seed + 1000).randn(pool, self.lstm_units).astype(np.float32)
                            if self.pooled_state and self.recurrent else None)
        self._step = 0          # vector steps taken
        self._count = 0         # transitions generated
        self._prev_done = np.ones(self.num_envs, dtype=bool)
        self.last_arrays = None

    # -- raw arrays ---------------------------------------------------------
    def next_arrays(self, env_subset=None):
        """Generates one transition for each env in env_subset (default: all, in order)."""
        envs = np.arange(self.num_envs) if env_subset is None else np.asarray(env_subset)
        m = len(envs)
        rs = self._rs
        g = self._count + np.arange(m)
        frame_idx = g & (len(self.pool) - 1)
        reward = rs.randn(m)
        if self.clip_rewards:
            reward = np.sign(reward)
        if self.done_mode == "periodic":
            done = ((self._step + 1 + 37 * envs) % self.done_period) == 0
        elif self.done_mode == "bernoulli":
            done = rs.rand(m) < self.done_p
        else:
            done = np.zeros(m, dtype=bool)
        action = rs.randint(0, self.num_actions, m).astype(np.int64)
        qvalues = rs.randn(m, self.num_actions).astype(np.float32)
        out = {
            "env": envs.astype(np.int64) + self.env_id_base, "frame_idx": frame_idx,
            "reward": reward, "done": done, "action": action, "qvalues": qvalues,
        }
        if self.recurrent and self.pooled_state:
            out["hx_idx"] = frame_idx
            out["cx_idx"] = (frame_idx + 7) & (len(self.pool) - 1)
        elif self.recurrent:
            out["hx"] = rs.randn(m, self.lstm_units).astype(np.float32)
            out["cx"] = rs.randn(m, self.lstm_units).astype(np.float32)
        if self.recurrent:
            out["initials"] = self._prev_done[envs].astype(np.float32)
        self._prev_done[envs] = done
        self._count += m
        self._step += 1
        self.last_arrays = out
        return out

    def frames(self, arrays):
        return self.pool[arrays["frame_idx"]]

    # -- reference-schema dicts --------------------------------------------
    def next_samples(self, env_subset=None):
        a = self.next_arrays(env_subset)
        return self.samples_from_arrays(a)

    def samples_from_arrays(self, a):
        samples = []
        for i in range(len(a["env"])):
            state = {"x": self.pool[a["frame_idx"][i]], "layer0_state": {}}
            if self.recurrent:
                if "hx_idx" in a:     # pooled: basic indexing = views of the pool rows
                    hx, cx = self._state_pool[a["hx_idx"][i]], self._state_pool[a["cx_idx"][i]]
                else:
                    hx, cx = a["hx"][i], a["cx"][i]
                state["layer1_state"] = {"hx": hx, "cx": cx, "initials": a["initials"][i]}
                state["layer2_state"] = {}
            else:
                state["layer1_state"] = {}
            samples.append({
                "policy_output": {"actions": a["action"][i], "qvalues": a["qvalues"][i]},
                "next_state": state,
                "reward": a["reward"][i],
                "done": a["done"][i],
                "info": {},
                "env_id": int(a["env"][i]),
            })
        return samples
