"""Device-resident history buffers: host-side mirror of the rltime.history interface.

Drop-in classes for the reference's

    ReplayHistoryBuffer             rltime/history/replay_history.py:6
    PrioritizedReplayHistoryBuffer  rltime/history/prioritized_replay_history.py:10

with the same constructor kwargs and the History method surface the trainer calls
(rltime/history/history.py:123,288,296,332; callers training/multi_step_trainer.py:133-150,
251-268 and training/torch/dqn.py:73-81): update / needed_feed_count / get_train_data /
update_losses.  Storage, sum-tree sampling, n-step assembly, importance weights and the
batch gather run in librltime_b200.so (CUDA, sm_100a); this file only flattens the nested
sample dicts into leaves, keeps the train-quota arithmetic and rebuilds the nested,
time-major (S, B, ...) train-data dict out of borrowed device tensors.

Streams: every replay kernel / copy runs on the buffer's OWN CUDA stream, so ingest, the
priority write-back and the next draw + gather overlap whatever the caller's stream is doing
(the learner's backward pass).  get_train_data makes the caller's current stream wait for the
gather, so the returned tensors are safe to use there; a batch stays valid until the
second-next get_train_data call (3 rotating slots, like StateStore's history,
rltime/general/backend.py:88,149-152).  Host arrays handed to update() / update_arrays() must
stay unmodified until the next update_losses*() / sync() (the reference keeps them by
reference forever, history.py:159-171).

There is no CPU fallback: without a CUDA device / the built library construction raises.
"""
import ctypes as C
import random

import numpy as np

from .. import _lib


def _anneal_value(base_value, progress, anneal_mode, default_target=0.0):
    # rltime/general/utils.py:85-103
    assert progress >= 0
    progress = min(progress, 1.0)
    if anneal_mode is False or anneal_mode is None:
        return base_value
    target = default_target if anneal_mode is True else float(anneal_mode)
    return base_value + (target - base_value) * progress


def _flatten(tree, prefix=()):
    """Nested dict/tuple of arrays -> [(path, leaf)], depth-first in insertion order."""
    if isinstance(tree, dict):
        out = []
        for k, v in tree.items():
            out += _flatten(v, prefix + (k,))
        return out
    if isinstance(tree, (tuple, list)):
        out = []
        for i, v in enumerate(tree):
            out += _flatten(v, prefix + (i,))
        return out
    if tree is None:
        return []
    return [(prefix, tree)]


def _skeleton(tree):
    """Structure of a nested sample with leaves replaced by their flat index."""
    counter = [0]

    def rec(t):
        if isinstance(t, dict):
            return {k: rec(v) for k, v in t.items()}
        if isinstance(t, (tuple, list)):
            return type(t)(rec(v) for v in t)
        if t is None:
            return None
        i = counter[0]
        counter[0] += 1
        return ("leaf", i)
    return rec(tree)


def _rebuild(skel, leaves):
    if isinstance(skel, tuple) and len(skel) == 2 and skel[0] == "leaf":
        return leaves[skel[1]]
    if isinstance(skel, dict):
        return {k: _rebuild(v, leaves) for k, v in skel.items()}
    if isinstance(skel, (tuple, list)):
        return type(skel)(_rebuild(v, leaves) for v in skel)
    return None


class _Leaf:
    def __init__(self, arr):
        arr = np.asarray(arr)
        self.dtype = arr.dtype
        self.shape = arr.shape
        self.nbytes = int(arr.dtype.itemsize * int(np.prod(arr.shape, dtype=np.int64)))
        self.typestr = arr.dtype.str


def extract_gamma(discount_function):
    """The trainer always passes discount(nstep, reward, po) = (gamma ** nstep) * reward
    (rltime/training/multi_step_trainer.py:70-74); recover gamma and verify the form."""
    g = float(discount_function(1, 1.0, None))
    for k, r in ((0, 2.0), (2, 3.0), (5, -0.5)):
        if float(discount_function(k, r, None)) != (g ** k) * r:
            raise NotImplementedError(
                "device history buffers support discount_function(n, r, _) == (gamma**n)*r only")
    return g


class DeviceReplayHistoryBuffer:
    """Uniform multi-step / multi-env replay on the GPU (replay_history.py:6-184)."""

    _KIND = _lib.RT_KIND_UNIFORM

    def __init__(self, size, train_frequency, avoid_episode_crossing=False, *,
                 nstep_target, nstep_train, prefix_steps=0, discount_function=None,
                 state_store=None, gamma=None, max_envs=64, device=None, output="torch"):
        import torch
        if not torch.cuda.is_available():
            raise _lib.RtError("rltime_b200 history buffers need a CUDA device (no CPU fallback)")
        self.avoid_episode_crossing = bool(avoid_episode_crossing)
        assert nstep_target == 1 or discount_function is not None or gamma is not None, \
            "History buffer must get a 'discount_function' for nstep_target>1"
        self._lib = _lib.load()
        self.size = int(size)
        self.train_frequency = train_frequency
        self._pre_consumed = []
        self.nstep_target = int(nstep_target)
        self.nstep_train = int(nstep_train)
        self.prefix_steps = int(prefix_steps)
        self.discount_function = discount_function
        self.state_store = state_store   # accepted for signature parity; storage is on device
        if gamma is None:
            gamma = extract_gamma(discount_function) if discount_function is not None else 1.0
        self.gamma = float(gamma)
        self.max_envs = int(max_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        assert output in ("torch", "numpy")
        self.output = output
        self._h = None
        self._env_index = {}        # caller's env id -> dense index
        self._env_code = {}         # caller's env id -> integer code stored with the transitions
        self._code_index = {}       # that code -> dense index (update_losses gets the codes back)
        self._state_leaves = None
        self._po_leaves = None
        self._state_skel = None
        self._po_skel = None
        self.last_sampled_idxes = None
        self.last_batch = None
        self._own = torch.cuda.Stream(self.device)
        self._prev_mark = None      # caller-stream event recorded when the previous draw was requested
        self._views = {}            # (ptr, shape, typestr) -> borrowed tensor view (the batch slots rotate
                                    # over three fixed device buffers, so every view is built once)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(self._own.cuda_stream)

    def sync(self):
        """Blocks until everything this buffer enqueued (ingest copies included) has run."""
        self._own.synchronize()

    def _begin_draw(self):
        """Slot-reuse fence: the gather of draw k overwrites the slot draw k-3 used; whatever the
        caller had enqueued on its stream when draw k-1 was requested (the consumers of draws
        <= k-2) must be done first.  It does not wait for the consumer of draw k-1, which is what
        overlaps with this draw."""
        import torch
        cur = torch.cuda.current_stream(self.device)
        if self._prev_mark is not None:
            self._own.wait_event(self._prev_mark)
        self._prev_mark = torch.cuda.Event()
        self._prev_mark.record(cur)
        return cur

    def _end_draw(self, cur):
        cur.wait_stream(self._own)

    def _extra_config(self, cfg):
        cfg.overlap = 0
        cfg.avoid_episode_crossing = 1 if self.avoid_episode_crossing else 0

    def _create(self, sample):
        self._state_skel = _skeleton(sample["next_state"])
        self._po_skel = _skeleton(sample["policy_output"])
        self._state_leaves = [_Leaf(v) for _, v in _flatten(sample["next_state"])]
        self._po_leaves = [_Leaf(v) for _, v in _flatten(sample["policy_output"])]
        self._create_from_leaves()

    def _create_from_leaves(self):
        cfg = _lib.ReplayConfig()
        cfg.size = self.size
        cfg.kind = self._KIND
        cfg.nstep_train = self.nstep_train
        cfg.prefix_steps = self.prefix_steps
        cfg.nstep_target = self.nstep_target
        cfg.gamma = self.gamma
        cfg.alpha = cfg.eps = cfg.max_weight_factor = 0.0
        cfg.global_importance_scaling = 0
        cfg.max_envs = self.max_envs
        cfg.device = self.device.index or 0
        assert 1 <= len(self._state_leaves) <= _lib.RT_MAX_FIELDS
        assert len(self._po_leaves) <= _lib.RT_MAX_FIELDS
        cfg.num_state_fields = len(self._state_leaves)
        for i, l in enumerate(self._state_leaves):
            cfg.state_field_bytes[i] = l.nbytes
        cfg.num_po_fields = len(self._po_leaves)
        for i, l in enumerate(self._po_leaves):
            cfg.po_field_bytes[i] = l.nbytes
        self._extra_config(cfg)
        h = C.c_void_p()
        _lib.check(self._lib.rt_replay_create(C.byref(cfg), C.byref(h)))
        self._h = h
        # the train quota (replay_history.py:62-75,173-184) is kept by the library: rt_replay_append adds,
        # rt_replay_consume_quota takes, rt_replay_needed_feed answers
        _lib.check(self._lib.rt_replay_set_train_frequency(self._h, float(self.train_frequency or 0)))
        for mb in self._pre_consumed:     # get_train_data calls made before the first sample arrived
            _lib.check(self._lib.rt_replay_consume_quota(self._h, mb))
        self._pre_consumed = []

    def close(self):
        if self._h is not None:
            self._views = {}
            self._lib.rt_replay_destroy(self._h)
            self._h = None

    def _view(self, ptr, shape, typestr):
        key = (ptr, shape, typestr)
        t = self._views.get(key)
        if t is None:
            t = self._views[key] = _lib.as_tensor(ptr, shape, typestr, self.device)
        return t

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dense_env(self, env_id):
        e = self._env_index.get(env_id)
        if e is None:
            e = len(self._env_index)
            if e >= self.max_envs:
                raise _lib.RtError("more than max_envs=%d distinct env ids" % self.max_envs)
            self._env_index[env_id] = e
            # integer code echoed in loss_indices: the id itself when it is an integer, else a code
            # outside the usual id range (remote actors use e.g. (actor, env) tuples or strings)
            code = int(env_id) if isinstance(env_id, (int, np.integer)) else (1 << 40) + e
            self._env_code[env_id] = code
            self._code_index[code] = e
        return e

    # ------------------------------------------------------------------ History API
    def update(self, new_samples):
        """History.update (history.py:123-176) for a list of acting samples
        (acting/acting_interface.py:58-90)."""
        if hasattr(new_samples, "unpack"):      # SharedSampleList (general/backend.py:208)
            new_samples.unpack()
        samples = list(new_samples)
        if not samples:
            return {}
        for s in samples:
            if hasattr(s["next_state"], "get_object"):   # ObjectWrapper
                s["next_state"] = s["next_state"].get_object()
        if self._h is None:
            self._create(samples[0])
        m = len(samples)
        env = np.empty(m, dtype=np.int32)
        env_ids = np.empty(m, dtype=np.int64)
        reward = np.empty(m, dtype=np.float64)
        done = np.empty(m, dtype=np.uint8)
        state_cols = [np.empty((m,) + l.shape, dtype=l.dtype) for l in self._state_leaves]
        po_cols = [np.empty((m,) + l.shape, dtype=l.dtype) for l in self._po_leaves]
        for i, s in enumerate(samples):
            env[i] = self._dense_env(s["env_id"])
            env_ids[i] = self._env_code[s["env_id"]]
            reward[i] = float(s["reward"])
            done[i] = 1 if s["done"] else 0
            for col, (_, v) in zip(state_cols, _flatten(s["next_state"])):
                col[i] = v
            for col, (_, v) in zip(po_cols, _flatten(s["policy_output"])):
                col[i] = v
        self._append(env, env_ids, reward, done, state_cols, po_cols, on_device=False)
        return {}

    def update_arrays(self, env_ids, reward, done, state_leaves, po_leaves):
        """Batched ingest: one call per vector step.  Leaves are numpy
        arrays or CUDA torch tensors of shape (m, ...), in flattened-leaf order."""
        import torch
        on_device = isinstance(state_leaves[0], torch.Tensor) and state_leaves[0].is_cuda
        if self._h is None:
            def one(x):
                return x[0].cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x[0])
            self._state_leaves = [_Leaf(one(x)) for x in state_leaves]
            self._po_leaves = [_Leaf(one(x)) for x in po_leaves]
            self._state_skel = {"x": ("leaf", 0)} if len(state_leaves) == 1 else \
                tuple(("leaf", i) for i in range(len(state_leaves)))
            self._po_skel = tuple(("leaf", i) for i in range(len(po_leaves)))
            self._create_from_leaves()
        env_ids = np.ascontiguousarray(env_ids, dtype=np.int64)
        env = np.fromiter((self._dense_env(int(e)) for e in env_ids), dtype=np.int32,
                          count=len(env_ids))
        reward = np.ascontiguousarray(reward, dtype=np.float64)
        done = np.ascontiguousarray(done, dtype=np.uint8)
        self._append(env, env_ids, reward, done, list(state_leaves), list(po_leaves), on_device)

    def set_structure(self, state_skeleton, po_skeleton):
        """Optional: nested structure used to rebuild states / policy_outputs dicts when the
        buffer is fed through update_arrays (leaves referenced as ("leaf", i))."""
        self._state_skel, self._po_skel = state_skeleton, po_skeleton

    def _append(self, env, env_ids, reward, done, state_cols, po_cols, on_device):
        m = len(env)
        keep = []

        def ptrs(cols, leaves):
            arr = (C.c_void_p * max(len(cols), 1))()
            for i, (c, l) in enumerate(zip(cols, leaves)):
                if on_device:
                    c = c.contiguous()
                    assert c.numel() * c.element_size() == m * l.nbytes
                    arr[i] = c.data_ptr()
                else:
                    c = np.ascontiguousarray(c, dtype=l.dtype)
                    assert c.nbytes == m * l.nbytes
                    arr[i] = c.ctypes.data
                keep.append(c)
            return arr
        sp = ptrs(state_cols, self._state_leaves)
        pp = ptrs(po_cols, self._po_leaves)
        if on_device:
            import torch
            self._own.wait_stream(torch.cuda.current_stream(self.device))
        _lib.check(self._lib.rt_replay_append(
            self._h, m, env.ctypes.data, env_ids.ctypes.data, reward.ctypes.data,
            done.ctypes.data, C.cast(sp, C.c_void_p), C.cast(pp, C.c_void_p),
            1 if on_device else 0, self._stream()))
        if on_device:
            # the copies are stream-ordered; keep the sources alive until they ran
            self._own.synchronize()

    def profile_gather(self, enable=True):
        _lib.check(self._lib.rt_replay_profile(self._h, 1 if enable else 0))

    def gather_time(self):
        """(total device ms, launches) of the gather kernel since the last call."""
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self._lib.rt_replay_gather_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def needed_feed_count(self, mbatch_size, num_envs):
        # replay_history.py:62-75
        if not self.train_frequency:
            return 0
        if self._h is None:              # nothing fed yet: quota = -(what the early get_train_data calls took)
            taken = sum(self._pre_consumed) * self.nstep_train
            return max(int(taken / self.train_frequency), num_envs)
        n = self._lib.rt_replay_needed_feed(self._h, mbatch_size, num_envs)
        return None if n < 0 else int(n)

    @property
    def train_quota(self):
        return 0 if self._h is None else self._lib.rt_replay_train_quota(self._h)

    def get_train_data(self, mbatch_size, train_progress=None):
        # replay_history.py:173-184
        if self.train_frequency:
            if self._h is None:
                self._pre_consumed.append(int(mbatch_size))
            else:
                _lib.check(self._lib.rt_replay_consume_quota(self._h, mbatch_size))
        return self._get_train_data(mbatch_size, train_progress)

    def draw(self, mbatch_size, train_progress=None):
        """get_train_data for a consumer that reads the raw device batch (self.last_batch, the input of
        DeviceLearner.step) and does not need the nested dict of tensor views: same quota / RNG / draw,
        returns True or None."""
        self._raw_only = True
        try:
            return self.get_train_data(mbatch_size, train_progress)
        finally:
            self._raw_only = False

    def _get_train_data(self, mbatch_size, train_progress):
        # replay_history.py:93-140
        if self._h is None:
            return None
        total_available = self._lib.rt_replay_uniform_available(self._h)
        if total_available < mbatch_size:
            assert self._lib.rt_replay_len(self._h) < self.size
            return None
        choices = np.ascontiguousarray(np.random.choice(total_available, mbatch_size),
                                       dtype=np.int64)
        cur = self._begin_draw()
        _lib.check(self._lib.rt_replay_sample_uniform(
            self._h, mbatch_size, choices.ctypes.data, self._stream()))
        self._end_draw(cur)
        return self._wrap_batch()

    def _wrap_batch(self):
        b = _lib.Batch()
        _lib.check(self._lib.rt_replay_batch(self._h, C.byref(b)))
        self.last_batch = b      # raw device view of the draw (input of DeviceLearner.step)
        if getattr(self, "_raw_only", False):
            if b.importance_weights:
                self._last_idx_tensor = self._view(b.idxes, (b.B,), "<i4")
            return True
        B, S, n = b.B, b.S, b.n
        dev = self.device
        states, targets = [], []
        for f, l in enumerate(self._state_leaves):
            full = self._view(b.all_states[f], (S + n, B) + l.shape, l.typestr)
            states.append(full[:S])
            targets.append(full[n:])
        pos = [self._view(b.policy_outputs[f], (S, B) + l.shape, l.typestr)
               for f, l in enumerate(self._po_leaves)]
        td = {
            "returns": self._view(b.returns, (S, B), "<f8"),
            "nsteps": self._view(b.nsteps, (S, B), "<i8"),
            "target_masks": self._view(b.target_masks, (S, B), "<f8"),
            "policy_outputs": _rebuild(self._po_skel, pos),
            "states": _rebuild(self._state_skel, states),
            "target_states": _rebuild(self._state_skel, targets),
            "extra_data": {},
        }
        if b.importance_weights:
            td["extra_data"] = {
                "importance_weights": self._view(b.importance_weights, (S, B), "<f8"),
                "loss_indices": self._view(b.loss_indices, (S, B, 2), "<i8"),
            }
            self._last_idx_tensor = self._view(b.idxes, (B,), "<i4")
        if self.output == "numpy":
            # the reference hands numpy for everything except the state tensors
            # (history.py:235-241 vs general/backend.py:136-153)
            for k in ("returns", "nsteps", "target_masks"):
                td[k] = td[k].cpu().numpy()
            td["policy_outputs"] = _map(td["policy_outputs"], lambda t: t.cpu().numpy())
            td["extra_data"] = _map(td["extra_data"], lambda t: t.cpu().numpy())
        return td

    def update_losses(self, indices, losses):
        pass


def _map(tree, f):
    if isinstance(tree, dict):
        return {k: _map(v, f) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_map(v, f) for v in tree)
    if tree is None:
        return None
    return f(tree)


class DevicePrioritizedReplayHistoryBuffer(DeviceReplayHistoryBuffer):
    """Weighted, overlapped-sequence prioritized replay on the GPU
    (prioritized_replay_history.py:10-356)."""

    _KIND = _lib.RT_KIND_PRIORITIZED

    def __init__(self, alpha=0.6, beta=0.4, beta_anneal=False, eps=1e-6, overlap=None,
                 max_weight_factor=0.9, global_importance_scaling=False, **kwargs):
        super().__init__(**kwargs)
        if self.avoid_episode_crossing:
            # the reference shifts prioritized sequences too (prioritized_replay_history.py:317-318);
            # the device draw has no refine step yet, and silently ignoring the flag would change what
            # is trained on
            raise NotImplementedError(
                "avoid_episode_crossing=True is supported by the uniform device replay buffer only")
        self._alpha, self._beta, self._beta_anneal, self._eps = alpha, beta, beta_anneal, eps
        self._max_weight_factor = max_weight_factor
        self._global_importance_scaling = global_importance_scaling
        # prioritized_replay_history.py:97-105
        if overlap is None:
            overlap = int(self.nstep_train / 2)
        elif overlap < 0:
            overlap = self.nstep_train + overlap
            assert overlap >= 0
        assert overlap < self.nstep_train, "Overlap must be < nstep_train"
        self._overlap = overlap
        self._gap = self.nstep_train - overlap
        self._last_idx_tensor = None
        # sharded replay: callable(device f64[1] tensor) that replaces its value by the maximum
        # over all shards in place (e.g. torch.distributed.all_reduce(MAX) or rt_comm_allreduce_max_f64), on
        # the CURRENT stream.  The importance weights of a draw are then normalised by the global maximum
        # weight instead of this shard's (prioritized_replay_history.py:347-354 on the union of the shards).
        self.global_weight_max = None

    def _extra_config(self, cfg):
        cfg.overlap = self._overlap
        cfg.alpha = self._alpha
        cfg.eps = self._eps
        cfg.max_weight_factor = self._max_weight_factor
        cfg.global_importance_scaling = 1 if self._global_importance_scaling else 0

    @property
    def last_sampled_idxes(self):
        return None if self._last_idx_tensor is None else self._last_idx_tensor.cpu().tolist()

    @last_sampled_idxes.setter
    def last_sampled_idxes(self, v):
        pass

    def tree_sum(self):
        out = C.c_double()
        _lib.check(self._lib.rt_replay_tree_sum(self._h, C.byref(out), self._stream()))
        return out.value

    def _get_train_data(self, mbatch_size, train_progress):
        # prioritized_replay_history.py:281-356.  The B uniforms come from the module-global
        # MT19937 stream exactly like _sample_proportional (:238) and are consumed even
        # when None is returned (:284 precedes :295-299).
        uniforms = (C.c_double * mbatch_size)(*[random.random() for _ in range(mbatch_size)])
        beta = _anneal_value(self._beta, train_progress if train_progress is not None else 0,
                             self._beta_anneal, 1.0)
        if self._h is None:
            return None
        cur = self._begin_draw()
        rc = _lib.check(self._lib.rt_replay_sample_prioritized(
            self._h, mbatch_size, float(beta), C.cast(uniforms, C.c_void_p), self._stream()))
        if rc != _lib.RT_NEED_MORE_DATA and self.global_weight_max is not None:
            self._rescale_to_global_max()
        self._end_draw(cur)
        if rc == _lib.RT_NEED_MORE_DATA:
            self._last_idx_tensor = None
            return None
        return self._wrap_batch()

    def _rescale_to_global_max(self):
        import torch
        b = _lib.Batch()
        _lib.check(self._lib.rt_replay_batch(self._h, C.byref(b)))
        with torch.cuda.stream(self._own):
            local = _lib.as_tensor(b.weight_max, (1,), "<f8", self.device)
            g = local.clone()
            self.global_weight_max(g)
            w = _lib.as_tensor(b.importance_weights, (b.S * b.B,), "<f8", self.device)
            w.mul_(local / g)

    def update_losses(self, indices, losses):
        """prioritized_replay_history.py:243-279.  indices: (M, 2) rows of
        (env_id, env_offset); losses: (M,).  Losses are widened to fp64 exactly (canonical
        semantics)."""
        import torch
        if isinstance(indices, torch.Tensor):
            indices = indices.cpu().numpy()
        if isinstance(losses, torch.Tensor):
            losses = losses.detach().cpu().numpy()
        indices = np.asarray(indices).reshape(-1, 2)
        losses = np.ascontiguousarray(losses, dtype=np.float64).reshape(-1)
        assert len(indices) == len(losses)
        pairs = np.empty((len(indices), 2), dtype=np.int64)
        try:
            pairs[:, 0] = [self._code_index[int(e)] for e in indices[:, 0]]
        except KeyError as ex:
            raise KeyError("update_losses: unknown env id %s (prefix rows (-1,-1) must not be "
                           "sent back)" % ex)
        pairs[:, 1] = indices[:, 1]
        _lib.check(self._lib.rt_replay_update_losses(
            self._h, len(pairs), pairs.ctypes.data, losses.ctypes.data, self._stream()))

    def update_losses_device(self, td_abs, ready=None):
        """Priority write-back for the rows trained from the last draw, |td| still on the
        device (fp32, (T*B,) time-major).  `ready(stream_ptr)`, when given, makes the buffer's
        stream wait for just the point where td_abs is final (DeviceLearner.wait_loss: before the
        backward pass), so the write-back and the next draw overlap the rest of the update;
        otherwise the buffer's stream waits for everything enqueued on the caller's stream."""
        assert td_abs.is_cuda and td_abs.dtype.itemsize == 4 and td_abs.is_contiguous()
        if ready is not None:
            ready(self._stream())
        else:
            import torch
            self._own.wait_stream(torch.cuda.current_stream(self.device))
        _lib.check(self._lib.rt_replay_update_losses_last(
            self._h, C.c_void_p(td_abs.data_ptr()), self._stream()))
