"""Host-side handle of the CUDA learner (rt_learner_* in include/rltime_b200.h).

Mirrors the parts of the reference a trainer touches on the learner side:
  IQNPolicy / DQNPolicy state (state_dict names, get_state / load_state,
      rltime/policies/torch/torch_policy.py:97-101),
  TorchTrainer.train_batch / calc_target_values / set_lr (training/torch/torch_trainer.py),
  PolicyTrainer.sync_target (training/policy_trainer.py:68-70).
No CPU fallback: construction fails without the built library and a CUDA device.
"""
import ctypes as C
import io

import numpy as np

from . import _lib


class DeviceLearner:
    def __init__(self, in_shape, conv, lstm_units, fc_size, num_actions, num_quantiles=32,
                 embedding_dim=64, dueling=True, *, mbatch, nstep_train, burn_in=0,
                 nstep_target=1, gamma=0.99, double_q=False, rnn_bootstrap=False,
                 vf_scale_epsilon=None, huber_kappa=1.0, clip_grad=None, adam_epsilon=1e-8,
                 lr=1e-3, loss_aggregation="mean", seed=0, device=None, gemm="tf32", policy="iqn",
                 loss_mode="huber", loss_timestep_aggregation=None, clip_grad_dynamic_alpha=None,
                 pre_fc=(), extra_dim=0, rnn_steps_train=None):
        """policy="iqn": IQNPolicy + IQN trainer (policies/torch/iqn.py, training/torch/iqn.py);
        policy="dqn": DQNPolicy + DQN trainer (policies/torch/dqn.py, training/torch/dqn.py:
        Rainbow-style dueling / double-Q / n-step / PER without the quantile layer).
        in_shape: (C, H, W) uint8 frames, or (D,) float32 vectors with conv == [] (MLP models);
        pre_fc: FC modules (models/torch/modules/fc.py) between the CNN / observation and the LSTM /
        last FC module, one list of layer widths per module; extra_dim: width of the extra feature
        vector of a tuple observation, fed to the LSTM (models/torch/sequential.py:146-165)."""
        import torch
        if not torch.cuda.is_available():
            raise _lib.RtError("rltime_b200 learner needs a CUDA device (no CPU fallback)")
        self._lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        md = _lib.ModelDesc()
        in_shape = tuple(int(d) for d in in_shape)
        if len(conv) == 0:
            assert len(in_shape) >= 1, "MLP models take a float32 vector observation"
            in_shape = (int(np.prod(in_shape)), 1, 1)
        md.in_c, md.in_h, md.in_w = in_shape
        self.obs_dtype = np.uint8 if len(conv) else np.float32
        md.num_conv = len(conv)
        md.extra_dim = int(extra_dim)
        self.X = int(extra_dim)
        module = 1 if len(conv) else 0
        k = 0
        for sizes in pre_fc:
            for j, sz in enumerate(sizes):
                assert k < _lib.RT_MAX_PRE_FC, "too many FC layers in front of the LSTM / last FC module"
                md.pre_fc_size[k], md.pre_fc_module[k], md.pre_fc_sub[k] = int(sz), module, j
                k += 1
            module += 1
        md.num_pre_fc = k
        # indices of the LSTM / last FC module in SequentialModel.layers ('layer{i}_state' keys)
        self.lstm_module = module
        self.fc_module = module + (1 if lstm_units else 0)
        for i, (f, k, s) in enumerate(conv):
            md.conv_filters[i], md.conv_kernel[i], md.conv_stride[i] = f, k, s
        md.lstm_units = lstm_units
        md.fc_size = fc_size
        md.num_actions = num_actions
        assert policy in ("iqn", "dqn")
        self.policy = policy
        if policy == "dqn":
            num_quantiles = 0      # ABI: 0 quantiles = plain DQNPolicy
        md.num_quantiles = num_quantiles
        md.embedding_dim = embedding_dim
        md.dueling = 1 if dueling else 0
        td = _lib.TrainDesc()
        td.mbatch, td.nstep_train, td.burn_in, td.nstep_target = mbatch, nstep_train, burn_in, nstep_target
        td.double_q = 1 if double_q else 0
        td.rnn_bootstrap = 1 if rnn_bootstrap else 0
        assert loss_aggregation in ("mean", "sum")
        td.loss_sum = 1 if loss_aggregation == "sum" else 0
        assert loss_timestep_aggregation in (None, "mean", "sum")
        td.loss_timestep_agg = {None: 0, "mean": 1, "sum": 2}[loss_timestep_aggregation]
        assert loss_mode in ("huber", "mse")
        assert loss_mode == "huber" or policy == "dqn", "IQN uses the quantile-Huber loss"
        td.loss_mse = 1 if loss_mode == "mse" else 0
        td.clip_grad_dynamic_alpha = -1.0 if clip_grad_dynamic_alpha is None else float(clip_grad_dynamic_alpha)
        td.rnn_steps_train = int(rnn_steps_train or 0)
        assert not td.rnn_steps_train or nstep_train % td.rnn_steps_train == 0, \
            "nstep_train must be divisible by rnn_steps_train"
        td.gamma = gamma
        td.vf_scale_epsilon = vf_scale_epsilon or 0.0
        td.huber_kappa = huber_kappa
        td.clip_grad = clip_grad or 0.0
        td.adam_epsilon = adam_epsilon
        td.lr = lr
        td.seed = seed
        # "tf32": tcgen05 TF32 products on operands rounded to nearest at their producers (default);
        # "tf32_trunc": the same kernels on raw fp32 operands (the tensor core truncates them);
        # "fp32": CUDA-core fp32 GEMMs (in-library parity reference)
        modes = {"fp32": _lib.RT_GEMM_FP32_SIMT, "tf32_trunc": _lib.RT_GEMM_TF32_TCGEN05,
                 "tf32": _lib.RT_GEMM_TF32_RN}
        assert gemm in modes, gemm
        td.gemm_mode = modes[gemm]
        self.B, self.T, self.P, self.n = mbatch, nstep_train, burn_in, nstep_target
        self.Nq, self.A, self.U = max(num_quantiles, 1), num_actions, lstm_units
        h = C.c_void_p()
        _lib.check(self._lib.rt_learner_create(C.byref(md), C.byref(td), self.device.index or 0,
                                               C.byref(h)))
        self._h = h
        self.param_info = []
        for i in range(self._lib.rt_learner_num_params(self._h)):
            name = C.create_string_buffer(128)
            shape = (C.c_int64 * 4)()
            nd = C.c_int32()
            _lib.check(self._lib.rt_learner_param_info(self._h, i, name, 128, shape, C.byref(nd)))
            self.param_info.append((name.value.decode(), tuple(shape[:nd.value])))
        # leaf order of a flattened next_state: x (, extra), hx, cx, initials
        e = 1 if self.X else 0
        self.io = _lib.LearnerIO(0, 1 + e, 2 + e, 3 + e, 0, 1 if self.X else -1) if lstm_units else \
            _lib.LearnerIO(0, -1, -1, -1, 0, -1)

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.rt_learner_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- parameters ---------------------------------------------------------------
    def load_state_dict(self, sd, which=_lib.RT_BUF_ONLINE):
        arrs = []
        for name, shape in self.param_info:
            a = sd[name]
            a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.shape == shape, (name, a.shape, shape)
            arrs.append(a)
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        _lib.check(self._lib.rt_learner_load_params(self._h, which, C.cast(ptrs, C.c_void_p)))

    def state_dict(self, which=_lib.RT_BUF_ONLINE):
        import torch
        arrs = [np.empty(shape, dtype=np.float32) for _, shape in self.param_info]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        _lib.check(self._lib.rt_learner_get_params(self._h, which, C.cast(ptrs, C.c_void_p)))
        return {name: torch.from_numpy(a) for (name, _), a in zip(self.param_info, arrs)}

    def get_state(self):
        """Same bytes contract as TorchPolicy.get_state (torch_policy.py:97-98): a torch-saved
        state_dict with the reference's parameter names (+ the embedding_range buffer)."""
        import torch
        sd = self.state_dict()
        if self.policy == "iqn":
            sd["embedding_range"] = torch.arange(
                1, dict(self.param_info)["quantile_layer.weight"][1] + 1, dtype=torch.float32)
        f = io.BytesIO()
        torch.save(sd, f)
        return f.getvalue()

    def load_state(self, state):
        import torch
        sd = torch.load(io.BytesIO(state), map_location="cpu")
        self.load_state_dict(sd)

    # ---- checkpoint / resume ------------------------------------------
    def training_state(self):
        """Everything needed to resume training bit-exactly: online / target weights, Adam
        moments, Adam step counter, learning rate, the counter of the device RNG behind the IQN
        quantile fractions and the dynamic grad-clip moving average (the reference checkpoint,
        policy_trainer.py:170-185, holds the policy weights only and cannot resume)."""
        st = self.optimizer_state()
        st.update({"online": self.state_dict(_lib.RT_BUF_ONLINE), "target": self.state_dict(_lib.RT_BUF_TARGET)})
        return st

    def optimizer_state(self):
        """training_state() without the two weight sets (what a trainer checkpoint stores next to
        policy_state)."""
        steps, lr = C.c_int64(), C.c_double()
        _lib.check(self._lib.rt_learner_get_opt_state(self._h, C.byref(steps), C.byref(lr)))
        rng, ema, ema_init = C.c_uint64(), C.c_float(), C.c_int32()
        _lib.check(self._lib.rt_learner_get_aux_state(self._h, C.byref(rng), C.byref(ema), C.byref(ema_init)))
        return {"adam_m": self.state_dict(_lib.RT_BUF_ADAM_M), "adam_v": self.state_dict(_lib.RT_BUF_ADAM_V),
                "adam_steps": steps.value, "lr": lr.value, "rng_counter": rng.value,
                "clip_ema": ema.value, "clip_ema_init": ema_init.value}

    def load_training_state(self, st):
        if "online" in st:
            self.load_state_dict(st["online"], _lib.RT_BUF_ONLINE)
            self.load_state_dict(st["target"], _lib.RT_BUF_TARGET)
        self.load_state_dict(st["adam_m"], _lib.RT_BUF_ADAM_M)
        self.load_state_dict(st["adam_v"], _lib.RT_BUF_ADAM_V)
        _lib.check(self._lib.rt_learner_set_opt_state(self._h, int(st["adam_steps"]), float(st["lr"])))
        _lib.check(self._lib.rt_learner_set_aux_state(
            self._h, int(st.get("rng_counter", 0)), float(st.get("clip_ema", 0.0)),
            int(st.get("clip_ema_init", 0))))

    def params_changed(self):
        """After writing flat(RT_BUF_ONLINE / RT_BUF_TARGET) directly (e.g. a broadcast)."""
        _lib.check(self._lib.rt_learner_params_changed(self._h, self._stream()))

    def sync_target(self):
        _lib.check(self._lib.rt_learner_sync_target(self._h, self._stream()))

    def set_lr(self, lr):
        _lib.check(self._lib.rt_learner_set_lr(self._h, float(lr)))

    # ---- update -------------------------------------------------------------------
    def step(self, batch, taus=None, io=None):
        """batch: ctypes _lib.Batch (from a device history buffer or hand-built)."""
        tp, keep = self._tau_ptrs(taus)
        _lib.check(self._lib.rt_learner_step(self._h, C.byref(batch), C.byref(io or self.io), tp,
                                             self._stream()))

    def prefetch(self, batch, stream_ptr, io=None):
        """Frame conversion of `batch` on another stream (the replay buffer's), off the update's critical
        path; the next step() on this batch picks it up (rt_learner_prefetch)."""
        _lib.check(self._lib.rt_learner_prefetch(self._h, C.byref(batch), C.byref(io or self.io), stream_ptr))

    def _tau_ptrs(self, taus):
        if taus is None or self.policy == "dqn":
            return None, []
        keep = []
        arr = (C.c_void_p * 3)()
        for i, t in enumerate(taus):
            a = np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t,
                                     dtype=np.float32)
            assert a.size == self.T * self.B * self.Nq
            keep.append(a)
            arr[i] = a.ctypes.data
        keep.append(arr)
        return C.cast(arr, C.c_void_p), keep

    # ---- data parallelism inside the library (rt_comm_*) ---------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id (rank 0 creates it, every rank passes the same bytes to comm_init)."""
        buf = (C.c_uint8 * 128)()
        _lib.check(_lib.load().rt_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def comm_init(self, unique_id, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _lib.check(self._lib.rt_comm_init(self._h, C.cast(buf, C.c_void_p), int(rank), int(world)))
        self.world = int(world)

    def comm_broadcast_params(self, root=0):
        _lib.check(self._lib.rt_comm_broadcast_params(self._h, int(root), self._stream()))

    def step_dp(self, batch, taus=None, io=None):
        """One data-parallel update: local gradients, in-library NCCL sum (overlapping the conv backward),
        identical clip + Adam on the gradient mean."""
        tp, keep = self._tau_ptrs(taus)
        _lib.check(self._lib.rt_learner_step_dp(self._h, C.byref(batch), C.byref(io or self.io), tp,
                                                self._stream()))

    def compute_grads(self, batch, taus=None, io=None):
        """Data-parallel phase 1: everything of step() up to and including the backward pass."""
        tp, keep = self._tau_ptrs(taus)
        _lib.check(self._lib.rt_learner_compute_grads(self._h, C.byref(batch), C.byref(io or self.io),
                                                      tp, self._stream()))

    def apply_grads(self, grad_scale=1.0):
        """Data-parallel phase 2: grad-norm / clip / Adam on the (all-reduced) flat gradient."""
        _lib.check(self._lib.rt_learner_apply_grads(self._h, float(grad_scale), self._stream()))

    def flat(self, which=_lib.RT_BUF_GRAD):
        """Borrowed CUDA tensor over one flat fp32 buffer (gradient by default)."""
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self._lib.rt_learner_flat_buffer(self._h, which, C.byref(p), C.byref(n)))
        return _lib.as_tensor(p.value, (n.value,), "<f4", self.device)

    def profile_gemms(self, enable=True):
        _lib.check(self._lib.rt_learner_profile(self._h, 1 if enable else 0))

    def gemm_time(self):
        """(device ms, algorithmic flops, launches) of all GEMM-shaped launches since the last call."""
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        _lib.check(self._lib.rt_learner_gemm_time(self._h, C.byref(ms), C.byref(fl), C.byref(n)))
        return ms.value, fl.value, n.value

    def gemm_launches(self, cap=16384):
        """[(algorithmic flops, device ms)] of every timed GEMM-shaped launch since profiling was
        enabled; call before gemm_time(), which resets the record."""
        fl = np.zeros(cap, dtype=np.float64)
        ms = np.zeros(cap, dtype=np.float64)
        n = C.c_int64()
        _lib.check(self._lib.rt_learner_gemm_launches(self._h, cap, fl.ctypes.data, ms.ctypes.data,
                                                      C.byref(n)))
        return list(zip(fl[:n.value].tolist(), ms[:n.value].tolist()))

    def gemm_shapes(self, cap=16384):
        """[(kind, M, N, K, transA, transB)] of the timed launches, same order as gemm_launches()."""
        sh = np.zeros((cap, 6), dtype=np.int32)
        n = C.c_int64()
        _lib.check(self._lib.rt_learner_gemm_shapes(self._h, cap, sh.ctypes.data, C.byref(n)))
        return [tuple(int(v) for v in r) for r in sh[:n.value]]

    def wait_late_grads(self, stream_ptr=None):
        """(first, count) of the flat gradient range that is final before the conv backward of the
        last compute_grads(); with a stream pointer, also makes that stream wait for it."""
        first, count = C.c_int64(), C.c_int64()
        sp = C.c_void_p(-1) if stream_ptr is None else stream_ptr
        _lib.check(self._lib.rt_learner_wait_late_grads(self._h, sp, C.byref(first), C.byref(count)))
        return first.value, count.value

    def wait_loss(self, stream_ptr):
        """Makes the CUDA stream `stream_ptr` wait until the |td| / losses of the last enqueued
        step are final (they are before its backward pass)."""
        _lib.check(self._lib.rt_learner_wait_loss(self._h, stream_ptr))

    def loss(self):
        """{'qloss', 'td_mean'} of the last step, read back as soon as the forward pass is done
        (does not wait for the backward pass / Adam; stats() does)."""
        import torch
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.device)
        a, b = C.c_float(), C.c_float()
        _lib.check(self._lib.rt_learner_read_loss(self._h, C.byref(a), C.byref(b),
                                                  C.c_void_p(self._side.cuda_stream)))
        return {"qloss": a.value, "td_mean" if self.policy == "iqn" else "qvalue": b.value}

    def td_abs(self):
        p = C.c_void_p()
        _lib.check(self._lib.rt_learner_td_abs(self._h, C.byref(p)))
        return _lib.as_tensor(p.value, (self.T * self.B,), "<f4", self.device)

    def stats(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _lib.check(self._lib.rt_learner_read_stats(self._h, C.byref(a), C.byref(b), C.byref(c),
                                                   self._stream()))
        # the second slot is td_mean for IQN (iqn.py:128) and the mean chosen q-value for DQN
        # (dqn.py:159-160)
        return {"qloss": a.value, "td_mean" if self.policy == "iqn" else "qvalue": b.value,
                "grad_norm": c.value}

    def debug(self, name, shape=None):
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self._lib.rt_learner_debug_tensor(self._h, name.encode(), C.byref(p), C.byref(n)))
        t = _lib.as_tensor(p.value, (n.value,), "<f4", self.device)
        return t if shape is None else t[:int(np.prod(shape))].view(*shape)


def batch_from_tensors(all_x, all_hx, all_cx, all_initials, returns, nsteps, target_masks, actions,
                       importance_weights, n, all_extra=None, targets=None):
    """Builds an rt_batch over caller-owned CUDA tensors (time-major, (S+n, B, ...) / (S, B)).
    Leaf order = DeviceLearner.io: x (, extra), hx, cx, initials.  `targets`: optional dict of
    separately stacked target states {"x", "extra", "hx", "cx", "initials"} of (S, B, ...) each
    (batches whose n-step varies per row); the all_* tensors then only hold the S training rows.
    Returns (Batch, keepalive list)."""
    b = _lib.Batch()
    S, B = returns.shape
    b.B, b.S, b.n = B, S, n
    keep = []

    def ptr(t, dtype):
        assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), (t.dtype, dtype)
        keep.append(t)
        return t.data_ptr()
    import torch
    b.all_states[0] = ptr(all_x, all_x.dtype)
    assert all_x.dtype in (torch.uint8, torch.float32)
    nf = 1
    if all_extra is not None:
        b.all_states[nf] = ptr(all_extra, torch.float32)
        nf += 1
    if all_hx is not None:
        b.all_states[nf] = ptr(all_hx, torch.float32)
        b.all_states[nf + 1] = ptr(all_cx, torch.float32)
        b.all_states[nf + 2] = ptr(all_initials, torch.float32)
        nf += 3
    if targets is not None:
        order = ["x"] + (["extra"] if all_extra is not None else []) + \
            (["hx", "cx", "initials"] if all_hx is not None else [])
        for i, k in enumerate(order):
            b.target_states[i] = ptr(targets[k], all_x.dtype if k == "x" else torch.float32)
    b.num_state_fields = nf
    b.policy_outputs[0] = ptr(actions, torch.int64)
    b.num_po_fields = 1
    b.returns = ptr(returns, torch.float64)
    b.nsteps = ptr(nsteps, torch.int64)
    b.target_masks = ptr(target_masks, torch.float64)
    if importance_weights is not None:
        b.importance_weights = ptr(importance_weights, torch.float64)
    return b, keep
