"""CPU oracle for the history-buffer half of the hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  The product
path (rltime_b200.history -> librltime_b200.so) never routes through it.

What it is: a struct-of-arrays restatement, in plain Python + numpy, of the reference's
multi-env / multi-step history buffers (opherlieber/rltime @ 504c8405):

    History                         rltime/history/history.py:8-335
    ReplayHistoryBuffer             rltime/history/replay_history.py:6-184
    PrioritizedReplayHistoryBuffer  rltime/history/prioritized_replay_history.py:10-356
    OnlineHistoryBuffer             rltime/history/online_history.py:4-120
    SumSegmentTree/MinSegmentTree   rltime/history/data_structures/segment_tree.py:10-156
    StateStore.stack (cpu)          rltime/general/backend.py:136-153

Parity is PINNED: tests/test_oracle_golden.py replays the seeded scenarios of
oracle/scenario.py and compares every output field against tests/golden/replay_*.npz,
which oracle/gen_golden.py produced by executing the unmodified reference in the build
container (sampled indices, loss indices, n-step counts, frames: bit-exact; fp64
returns / weights / tree sums: bit-exact).

Canonical-semantics decisions:
  * all priority arithmetic is fp64 (numpy-1.x behaviour the reference was written
    against; under numpy 2 the caller widens fp32 losses exactly, see scenario.py);
  * instead of one dict per transition, transitions live in per-env columns indexed by
    the env-local running offset (the reference's sample["env_buffer_offset"],
    prioritized_replay_history.py:143-150); evicted entries are dropped by advancing
    `first[env]` (== the reference's _env_sample_offsets).
"""
import random
from collections import deque

import numpy as np


# --------------------------------------------------------------------------- trees
class _Tree:
    """Array-embedded binary tree, root at 1, leaves at cap+i
    (segment_tree.py:10-101).  Parent = op(left, right), recomputed bottom-up."""

    def __init__(self, capacity, neutral, op):
        assert capacity > 0 and capacity & (capacity - 1) == 0
        self.cap = capacity
        self.v = np.full(2 * capacity, neutral, dtype=np.float64)
        self.op = op

    def set(self, idx, val):
        v = self.v
        i = idx + self.cap
        v[i] = val
        i >>= 1
        while i >= 1:
            v[i] = self.op(v[2 * i], v[2 * i + 1])
            i >>= 1

    def get(self, idx):
        assert 0 <= idx < self.cap
        return self.v[self.cap + idx]

    def root(self):
        # reduce(0, cap) hits the `start == node_start and end == node_end` case at the
        # root (segment_tree.py:44-45), i.e. returns _value[1] untouched.
        return self.v[1]


class SumTree(_Tree):
    def __init__(self, capacity):
        super().__init__(capacity, 0.0, lambda a, b: a + b)

    def find_prefixsum_idx(self, mass):
        """segment_tree.py:116-142: go left iff node[2i] > mass, else subtract, go right."""
        v = self.v
        i = 1
        while i < self.cap:
            left = v[2 * i]
            if left > mass:
                i = 2 * i
            else:
                mass -= left
                i = 2 * i + 1
        return i - self.cap


class MinTree(_Tree):
    def __init__(self, capacity):
        super().__init__(capacity, np.inf, min)


# --------------------------------------------------------------------------- stacking
def _stack_tree(items):
    """Recursive stack of nested dict/tuple leaves (general/utils.py:25-53)."""
    first = items[0]
    if isinstance(first, dict):
        return {k: _stack_tree([it[k] for it in items]) for k in first}
    if isinstance(first, (tuple, list)):
        return type(first)(_stack_tree([it[i] for it in items]) for i in range(len(first)))
    if first is None:
        return None
    return np.stack(items)


def _map_tree(x, f):
    if isinstance(x, dict):
        return {k: _map_tree(v, f) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return type(x)(_map_tree(v, f) for v in x)
    if x is None:
        return None
    return f(x)


class _EnvColumns:
    """Per-env transition columns indexed by absolute env offset."""
    __slots__ = ("next_state", "reward", "done", "policy_output", "ret", "nstep",
                 "mask", "loss", "prio", "first", "base_state")

    def __init__(self):
        self.next_state = []
        self.reward = []
        self.done = []
        self.policy_output = []
        self.ret = []      # lazily grown n-step return          (history.py:83-108)
        self.nstep = []    # number of steps folded in so far
        self.mask = []     # target_mask
        self.loss = []     # PER per-transition loss
        self.prio = []     # PER prioritization index (or None)
        self.first = 0     # offset of the oldest live transition
        self.base_state = None  # state of offset 0 (== its own next_state, history.py:159-163)

    def __len__(self):
        return len(self.reward) - self.first


# --------------------------------------------------------------------------- base
class HistoryOracle:
    """history.py:8-335."""

    def __init__(self, nstep_target, nstep_train, prefix_steps=0,
                 discount_function=None, state_store=None):
        assert nstep_target == 1 or discount_function is not None
        self.nstep_target = nstep_target
        self.nstep_train = nstep_train
        self.prefix_steps = prefix_steps
        self.discount_function = discount_function
        self.envs = {}      # env_id -> _EnvColumns, insertion-ordered like History.buffer
        self.last_sampled_idxes = None

    # hooks
    def _added(self, env_id, offset):
        pass

    def _removed(self, env_id, offset):
        pass

    def update(self, new_samples):
        """history.py:123-176."""
        for s in new_samples:
            env_id = s["env_id"]
            c = self.envs.get(env_id)
            if c is None:
                c = self.envs[env_id] = _EnvColumns()
                c.base_state = s["next_state"]
            c.next_state.append(s["next_state"])
            c.reward.append(s["reward"])
            c.done.append(s["done"])
            c.policy_output.append(s["policy_output"])
            c.ret.append(float(s["reward"]))
            c.nstep.append(1)
            c.mask.append(1 - s["done"])
            c.loss.append(None)
            c.prio.append(None)
            self._added(env_id, len(c.reward) - 1)
        return {}

    def _state_of(self, c, offset):
        return c.base_state if offset == 0 else c.next_state[offset - 1]

    def _drop_oldest(self, env_id, amount=1):
        """history.py:110-121: removal is always oldest-first within an env."""
        c = self.envs[env_id]
        assert len(c) >= amount
        for _ in range(amount):
            self._removed(env_id, c.first)
            k = c.first
            # release payload; next_state[k] stays reachable as the state of k+1
            if k >= 1:
                c.next_state[k - 1] = None
            c.policy_output[k] = None
            c.first += 1

    def _grow_nstep(self, c, k, nstep_target):
        """history.py:71-108 (lazy, cached, only ever grows)."""
        end = min(k + nstep_target, len(c.reward))
        for t in range(k + c.nstep[k], end):
            if c.mask[k]:
                c.ret[k] += self.discount_function(c.nstep[k], c.reward[t],
                                                   c.policy_output[t])
            c.nstep[k] += 1
            if c.done[t]:
                c.mask[k] = 0.

    def _range(self, env_id, index, steps, fixed_target=False):
        """history.py:178-201; `index` is relative to the env's oldest live transition."""
        c = self.envs[env_id]
        assert index >= 0 and index + steps <= len(c)
        out = []
        n = self.nstep_target
        start = c.first + index
        for k in range(start, start + steps):
            if fixed_target:
                n = min(n, start + steps - k)
            self._grow_nstep(c, k, n)
            out.append((c, k))
        return out

    def _batch(self, ranges, extra=None):
        """history.py:203-286 -> dict of (S, B, ...) arrays."""
        B = len(ranges)
        S = len(ranges[0])
        rows = [ranges[b][t] for t in range(S) for b in range(B)]  # time-major
        td = {
            "returns": np.stack([c.ret[k] for c, k in rows]),
            "nsteps": np.stack([c.nstep[k] for c, k in rows]),
            "target_masks": np.stack([c.mask[k] for c, k in rows]),
            "policy_outputs": _stack_tree([c.policy_output[k] for c, k in rows]),
        }
        # target state of k = next_state of k + nstep - 1 (history.py:100-102)
        td["states"] = _stack_tree([self._state_of(c, k) for c, k in rows])
        td["target_states"] = _stack_tree(
            [c.next_state[k + c.nstep[k] - 1] for c, k in rows])
        td = _map_tree(td, lambda x: x.reshape((S, B) + x.shape[1:]))
        if extra is not None:
            assert len(extra) == B
            td["extra_data"] = {k: np.stack([e[k] for e in extra], axis=1)
                                for k in extra[0]}
        else:
            td["extra_data"] = {}
        return td

    def update_losses(self, indices, losses):
        pass


# --------------------------------------------------------------------------- uniform
class ReplayOracle(HistoryOracle):
    """replay_history.py:6-184."""

    def __init__(self, size, train_frequency, avoid_episode_crossing=False, **kw):
        super().__init__(**kw)
        self.size = size
        self.train_frequency = train_frequency
        self.avoid_episode_crossing = avoid_episode_crossing
        self.fifo = deque()     # env ids in arrival order (linear_history)
        self.train_quota = 0

    def needed_feed_count(self, mbatch_size, num_envs):
        if not self.train_frequency:
            return 0
        if self.train_quota > 0:
            return None
        return max(int(-self.train_quota / self.train_frequency), num_envs)

    def _added(self, env_id, offset):
        if len(self.fifo) >= self.size:
            assert len(self.fifo) == self.size
            self._drop_oldest(self.fifo.popleft(), 1)
        self.fifo.append(env_id)
        if self.train_frequency:
            self.train_quota += self.train_frequency

    def _refine(self, env_id, start, amount):
        """replay_history.py:142-171."""
        if not self.avoid_episode_crossing:
            return start
        c = self.envs[env_id]
        for i in range(amount - 1):
            if c.done[c.first + start + i]:
                if i < amount / 2:
                    start = max(start - (amount - i - 1), 0)
                else:
                    start = min(start + i + 1, len(c) - amount)
                break
        return start

    def get_train_data(self, mbatch_size, train_progress=None):
        if self.train_frequency:
            self.train_quota -= mbatch_size * self.nstep_train
            assert self.train_quota < 100 * mbatch_size * self.nstep_train
            assert self.train_quota > -100 * mbatch_size * self.nstep_train
        return self._get_train_data(mbatch_size, train_progress)

    def _get_train_data(self, mbatch_size, train_progress):
        """replay_history.py:93-140."""
        S = self.nstep_train + self.prefix_steps
        avail = {}
        total = 0
        for env_id, c in self.envs.items():
            a = len(c) - (S + self.nstep_target - 1)
            if a > 0:
                avail[env_id] = a
                total += a
        if total < mbatch_size:
            assert len(self.fifo) < self.size
            return None
        choices = np.random.choice(total, mbatch_size)
        ranges = []
        for choice in choices:
            for env_id, a in avail.items():
                if choice < a:
                    ranges.append(self._range(env_id, self._refine(env_id, choice, S), S))
                    break
                choice -= a
            else:
                raise AssertionError("choice out of range")
        return self._batch(ranges)


# --------------------------------------------------------------------------- PER
def anneal_value(base_value, progress, anneal_mode, default_target=0.0):
    """general/utils.py:85-103."""
    assert progress >= 0
    progress = min(progress, 1.0)
    if anneal_mode is False or anneal_mode is None:
        return base_value
    target = default_target if anneal_mode is True else float(anneal_mode)
    return base_value + (target - base_value) * progress


class PrioritizedReplayOracle(ReplayOracle):
    """prioritized_replay_history.py:10-356."""

    def __init__(self, alpha=0.6, beta=0.4, beta_anneal=False, eps=1e-6, overlap=None,
                 max_weight_factor=0.9, global_importance_scaling=False, **kw):
        super().__init__(**kw)
        self.alpha, self.beta, self.beta_anneal, self.eps = alpha, beta, beta_anneal, eps
        self.max_weight_factor = max_weight_factor
        self.global_importance_scaling = global_importance_scaling
        T = self.nstep_train
        if overlap is None:
            overlap = int(T / 2)
        elif overlap < 0:
            overlap = T + overlap
            assert overlap >= 0
        assert overlap < T
        self.gap = T - overlap
        self.target_capacity = int(self.size / self.gap)
        cap = 1
        while cap < self.target_capacity:
            cap *= 2
        self.sum_tree = SumTree(cap)
        self.min_tree = MinTree(cap) if global_importance_scaling else None
        self.max_loss = 1.0
        self.free = deque(range(self.target_capacity))
        self.seq = [None] * self.target_capacity   # idx -> (env_id, base offset)

    def _added(self, env_id, offset):
        """prioritized_replay_history.py:136-172 (eviction happens first, in super)."""
        super()._added(env_id, offset)
        assert len(self.free) > 0
        c = self.envs[env_id]
        c.loss[offset] = self.max_loss
        base = offset - self.nstep_train + 1 - self.nstep_target + 1
        if base % self.gap == 0 and base >= c.first + self.prefix_steps:
            idx = self.free.popleft()
            c.prio[base] = idx
            self.seq[idx] = (env_id, base)
            self._recalc(idx)

    def _recalc(self, idx):
        """prioritized_replay_history.py:174-208."""
        env_id, base = self.seq[idx]
        c = self.envs[env_id]
        assert c.prio[base] == idx and base % self.gap == 0 and base >= c.first
        T = self.nstep_train
        if T == 1:
            w = c.loss[base]
        else:
            losses = c.loss[base:base + T]
            assert len(losses) == T
            w = self.max_weight_factor * np.max(losses) + \
                (1 - self.max_weight_factor) * np.mean(losses)
        prio = w ** self.alpha
        self.sum_tree.set(idx, prio)
        if self.min_tree is not None:
            self.min_tree.set(idx, prio)

    def _removed(self, env_id, offset):
        """prioritized_replay_history.py:210-230."""
        c = self.envs[env_id]
        assert offset == c.first
        k = c.first + self.prefix_steps
        assert k < len(c.reward), "reference indexes buffer[env][prefix_steps] here"
        if k % self.gap == 0 and c.prio[k] is not None:
            idx = c.prio[k]
            c.prio[k] = None
            self.sum_tree.set(idx, 0)
            if self.min_tree is not None:
                self.min_tree.set(idx, np.inf)
            self.free.append(idx)
            self.seq[idx] = None

    def _sample_proportional(self, batch_size):
        """prioritized_replay_history.py:232-241; RNG = module-global random (MT19937)."""
        p_total = self.sum_tree.root()
        every = p_total / batch_size
        return [self.sum_tree.find_prefixsum_idx(random.random() * every + i * every)
                for i in range(batch_size)]

    def update_losses(self, indices, losses):
        """prioritized_replay_history.py:243-279."""
        affected = {}
        T = self.nstep_train
        for (env_id, offset), loss in zip(indices, losses):
            env_id = int(env_id)
            offset = int(offset)
            c = self.envs[env_id]
            if offset < c.first:
                continue
            c.loss[offset] = abs(loss) + self.eps
            base = offset - (offset % self.gap)
            while base + T > offset and base >= c.first:
                idx = c.prio[base]
                if idx is not None:
                    affected[idx] = True
                base -= self.gap
        for idx in affected:
            self._recalc(idx)

    def _get_train_data(self, mbatch_size, train_progress):
        """prioritized_replay_history.py:281-356."""
        idxes = self._sample_proportional(mbatch_size)   # RNG advances even if None below
        self.last_sampled_idxes = idxes
        beta = anneal_value(self.beta, train_progress, self.beta_anneal, 1.0)
        total_items = len(self.seq) - len(self.free)
        S = self.prefix_steps + self.nstep_train
        if total_items < mbatch_size:
            assert len(self.fifo) < self.size
            return None
        ranges, extra = [], []
        p_sum = self.sum_tree.root()
        for idx in idxes:
            assert self.seq[idx] is not None
            env_id, base = self.seq[idx]
            c = self.envs[env_id]
            start = base - c.first - self.prefix_steps
            assert start >= 0
            start = self._refine(env_id, start, S)
            ranges.append(self._range(env_id, start, S))
            weight = ((self.sum_tree.get(idx) / p_sum) * total_items) ** (-beta)
            li = [(-1, -1)] * self.prefix_steps + \
                 [(env_id, o) for o in range(base, base + self.nstep_train)]
            extra.append({"importance_weights": np.array([weight] * S),
                          "loss_indices": np.array(li)})
        td = self._batch(ranges, extra)
        if self.global_importance_scaling:
            p_min = self.min_tree.root() / p_sum
            max_weight = (p_min * total_items) ** (-beta)
        else:
            max_weight = np.max(td["extra_data"]["importance_weights"])
        self.last_weight_max = float(max_weight)      # test aid: what this batch was normalised by
        td["extra_data"]["importance_weights"] /= max_weight
        return td


# --------------------------------------------------------------------------- online
class OnlineOracle(HistoryOracle):
    """online_history.py:4-120."""

    def __init__(self, max_delayed_steps=5000, fixed_target=True, **kw):
        super().__init__(**kw)
        self.last_env = None
        self.max_delayed_steps = max_delayed_steps
        self.fixed_target = fixed_target

    def _can_train(self, mbatch_size):
        return sum(int(len(c) / self.nstep_train) for c in self.envs.values()) >= mbatch_size

    def update(self, samples):
        ret = super().update(samples)
        discarded = 0
        for env_id, c in self.envs.items():
            if len(c) > self.max_delayed_steps:
                rm = len(c) - self.max_delayed_steps
                self._drop_oldest(env_id, rm)
                discarded += rm
        ret["discarded_steps"] = discarded
        return ret

    def needed_feed_count(self, mbatch_size, num_envs):
        return None if self._can_train(mbatch_size) else num_envs

    def get_train_data(self, mbatch_size, train_progress=None):
        assert self.prefix_steps == 0
        T = self.nstep_train
        if not self._can_train(mbatch_size):
            return None
        ids = sorted(self.envs.keys())
        # NB `not self.last_env` is also true for env id 0 (online_history.py:95-97)
        i = 0 if not self.last_env else (ids.index(self.last_env) + 1) % len(ids)
        ranges = []
        td_parts = []
        while len(ranges) < mbatch_size:
            env_id = ids[i]
            if len(self.envs[env_id]) >= T:
                rng = self._range(env_id, 0, T, self.fixed_target)
                ranges.append(rng)
                # the reference builds the batch after removal; payload of removed rows
                # stays alive there by reference, so snapshot before dropping.
                td_parts.append(self._snapshot(rng))
                self._drop_oldest(env_id, T)
                self.last_env = env_id
            i = (i + 1) % len(ids)
        return self._batch_from_snapshots(td_parts)

    def _snapshot(self, rng):
        return [dict(ret=c.ret[k], nstep=c.nstep[k], mask=c.mask[k], po=c.policy_output[k],
                     state=self._state_of(c, k), target=c.next_state[k + c.nstep[k] - 1])
                for c, k in rng]

    def _batch_from_snapshots(self, parts):
        B, S = len(parts), len(parts[0])
        rows = [parts[b][t] for t in range(S) for b in range(B)]
        td = {
            "returns": np.stack([r["ret"] for r in rows]),
            "nsteps": np.stack([r["nstep"] for r in rows]),
            "target_masks": np.stack([r["mask"] for r in rows]),
            "policy_outputs": _stack_tree([r["po"] for r in rows]),
            "states": _stack_tree([r["state"] for r in rows]),
            "target_states": _stack_tree([r["target"] for r in rows]),
        }
        td = _map_tree(td, lambda x: x.reshape((S, B) + x.shape[1:]))
        td["extra_data"] = {}
        return td


def get_types():
    return {"online": OnlineOracle, "replay": ReplayOracle,
            "prioritized_replay": PrioritizedReplayOracle}
