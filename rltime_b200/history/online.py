"""Online multi-step history buffer (host side; SURVEY.md section 8 row a11: no kernel).

Mirror of rltime/history/online_history.py:4-120 on top of the History base semantics
(rltime/history/history.py:71-201): per-env queues of freshly acted transitions, handed out
once as (nstep_train, mbatch) batches in round-robin env order and then discarded.  The
transitions of one training batch are a few KB to a few MB, so this stays a plain numpy
structure; the stacked state leaves are moved to the policy device (one H2D per leaf) the way
StateStore.stack does for the reference (rltime/general/backend.py:136-153).
"""
import numpy as np


def _stack(items):
    first = items[0]
    if isinstance(first, dict):
        return {k: _stack([it[k] for it in items]) for k in first}
    if isinstance(first, (tuple, list)):
        return type(first)(_stack([it[i] for it in items]) for i in range(len(first)))
    if first is None:
        return None
    return np.stack(items)


def _apply(tree, f):
    if isinstance(tree, dict):
        return {k: _apply(v, f) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_apply(v, f) for v in tree)
    return None if tree is None else f(tree)


class _Queue:
    """Transitions of one env awaiting training (oldest first)."""

    def __init__(self):
        self.items = []          # dicts: next_state, reward, done, policy_output, state
        self.prev_next_state = None


class OnlineHistoryBuffer:
    def __init__(self, max_delayed_steps=5000, fixed_target=True, *, nstep_target, nstep_train,
                 prefix_steps=0, discount_function=None, state_store=None, device=None):
        assert nstep_target == 1 or discount_function is not None, \
            "History buffer must get a 'discount_function' for nstep_target>1"
        self.nstep_target, self.nstep_train, self.prefix_steps = nstep_target, nstep_train, prefix_steps
        self.discount_function = discount_function
        self.max_delayed_steps = max_delayed_steps
        self.fixed_target = fixed_target
        self.state_store = state_store
        self.device = device
        self.last_env = None
        self._q = {}             # env_id -> _Queue, in first-appearance order

    def update(self, new_samples):
        if hasattr(new_samples, "unpack"):
            new_samples.unpack()
        for s in new_samples:
            ns = s["next_state"]
            if hasattr(ns, "get_object"):
                ns = ns.get_object()
            q = self._q.get(s["env_id"])
            if q is None:
                q = self._q[s["env_id"]] = _Queue()
                state = ns                       # first ever sample of an env (history.py:159-163)
            else:
                state = q.prev_next_state
            q.items.append({"next_state": ns, "reward": s["reward"], "done": s["done"],
                            "policy_output": s["policy_output"], "state": state})
            q.prev_next_state = ns
        discarded = 0
        for q in self._q.values():                # online_history.py:69-74
            extra = len(q.items) - self.max_delayed_steps
            if extra > 0:
                del q.items[:extra]
                discarded += extra
        return {"discarded_steps": discarded}

    def _ready(self, mbatch_size):
        return sum(len(q.items) // self.nstep_train for q in self._q.values()) >= mbatch_size

    def needed_feed_count(self, mbatch_size, num_envs):
        return None if self._ready(mbatch_size) else num_envs

    def _rows(self, items):
        """n-step returns / targets of the first nstep_train transitions (history.py:71-108,
        178-201), target capped at the sequence end when fixed_target."""
        T = self.nstep_train
        out = []
        for i in range(T):
            n = min(self.nstep_target, T - i) if self.fixed_target else self.nstep_target
            avail = min(n, len(items) - i)
            ret = float(items[i]["reward"])
            mask = 1 - items[i]["done"]
            for k in range(1, avail):
                if mask:
                    ret += self.discount_function(k, items[i + k]["reward"], items[i + k]["policy_output"])
                if items[i + k]["done"]:
                    mask = 0.
            out.append({"states": items[i]["state"], "target_states": items[i + avail - 1]["next_state"],
                        "returns": ret, "nsteps": avail, "target_masks": mask,
                        "policy_outputs": items[i]["policy_output"]})
        return out

    def get_train_data(self, mbatch_size, train_progress=None):
        assert self.prefix_steps == 0, "Online history does not support prefix/burnin steps"
        T = self.nstep_train
        if not self._ready(mbatch_size):
            return None
        ids = sorted(self._q.keys())
        # `not self.last_env` is also true for env id 0, as in the reference (online_history.py:95-97)
        i = 0 if not self.last_env else (ids.index(self.last_env) + 1) % len(ids)
        cols = []
        while len(cols) < mbatch_size:
            q = self._q[ids[i]]
            if len(q.items) >= T:
                cols.append(self._rows(q.items))
                del q.items[:T]
                self.last_env = ids[i]
            i = (i + 1) % len(ids)
        B = len(cols)
        rows = [cols[b][t] for t in range(T) for b in range(B)]      # time-major
        td = {k: _stack([r[k] for r in rows]) for k in
              ("returns", "nsteps", "target_masks", "policy_outputs", "states", "target_states")}
        td = _apply(td, lambda x: x.reshape((T, B) + x.shape[1:]))
        if self.device is not None:
            import torch
            for k in ("states", "target_states"):
                td[k] = _apply(td[k], lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(self.device))
        td["extra_data"] = {}
        return td

    def update_losses(self, indices, losses):
        pass
