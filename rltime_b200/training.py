"""Trainer-side mirror: IQN / DQN trainers with the reference's constructor / train() surface whose
learner, history buffer and acting-time inference all run in librltime_b200.so.

Mirrors (paths in the reference tree):
  PolicyTrainer.train / sample_actors / _process_new_samples   training/policy_trainer.py:256-325
  MultiStepTrainer._train (the act / feed / train loop)        training/multi_step_trainer.py:152-379
  TorchTrainer._train, DQN._train, IQN.create_policy           training/torch/{torch_trainer,dqn,iqn}.py
  DQNPolicy.actor_predict, SequentialModel.make_input_state,
  LSTM.get_state                                               policies/torch/dqn.py:132-148,
                                                               models/torch/sequential.py:128-144,
                                                               models/torch/modules/lstm.py:131-161
Environment stepping / exploration stay in the host-side `actors` object, which only needs the
reference's ActingInterface (acting/acting_interface.py:2-90): get_spaces, get_env_count,
get_samples, set_actor_policy, update_state.
"""
import ctypes as C
import logging
import time

import numpy as np

from . import _lib
from .history import get_types as _history_types
from .init import init_params
from .learner import DeviceLearner, batch_from_tensors


def _layer_kind(layer):
    t = layer.get("type")
    if isinstance(t, str):
        return t
    return {"CNN": "cnn", "FC": "fc", "LSTM": "lstm"}.get(getattr(t, "__name__", ""), str(t))


def parse_model_config(model_config):
    """{'type': 'sequential', 'args': {'layer_configs': [...]}} (configs/models/*.json,
    models/torch/sequential.py:15-66) -> dict(conv, pre_fc, lstm_units, fc_size).

    Supported module sequences: [cnn] -> fc* -> [lstm] -> fc, i.e. nature_cnn_fc512,
    nature_cnn_lstm512_fc512, nature_cnn_fc512_lstm512_fc512 and mlp_2x64 of the reference's
    configs/models.  Every FC module but the last may use fc_count > 1 (fc.py:18-24)."""
    t = model_config.get("type")
    assert t in ("sequential", None) or not isinstance(t, str), "only the sequential model family is supported"
    args = model_config.get("args", {})
    assert args.get("extra_input_layer") is None, \
        "extra_input_layer: only the default (the LSTM layer, sequential.py:67-79) is supported"
    layers = args["layer_configs"]
    kinds = [_layer_kind(l) for l in layers]
    assert kinds and kinds[-1] == "fc", \
        "the last module must be an FC module (IQN injected before an LSTM is not supported), got %s" % kinds
    conv, pre_fc, lstm_units = [], [], 0
    i = 0
    if kinds[0] == "cnn":
        conv = [(int(l["filters"]), int(l["kernel"]), int(l["stride"])) for l in layers[0]["args"]["layers"]]
        i = 1

    def fc_args(layer):
        a = layer.get("args", {})
        assert not a.get("batch_norm", False) and a.get("activation", "relu") == "relu", \
            "FC modules: only ReLU without batch-norm is supported"
        return int(a["fc_size"]), int(a.get("fc_count", 1))
    while i < len(layers) - 1 and kinds[i] == "fc":
        size, count = fc_args(layers[i])
        pre_fc.append([size] * count)
        i += 1
    if i < len(layers) - 1:
        assert kinds[i] == "lstm" and i == len(layers) - 2, \
            "supported topologies: [cnn] -> fc* -> [lstm] -> fc (got %s)" % kinds
        lstm_units = int(layers[i]["args"]["num_units"])
    size, count = fc_args(layers[-1])
    assert count == 1, "fc_count > 1 is supported in every FC module but the last one"
    return {"conv": conv, "pre_fc": pre_fc, "lstm_units": lstm_units, "fc_size": size}


def parse_observation_space(space):
    """Box -> (shape, 0); Tuple(main Box, 1-D extra Boxes...) -> (main shape, total extra width)
    (models/torch/torch_model.py:28-56)."""
    if hasattr(space, "spaces"):
        main = space.spaces[0]
        extra = 0
        for s in space.spaces[1:]:
            assert len(s.shape) == 1, "extra observation spaces must be 1-D"
            extra += int(s.shape[0])
        return tuple(main.shape), extra
    return tuple(space.shape), 0


class DevicePolicy:
    """Actor-facing policy backed by the learner's online network (rt_learner_act).

    The recurrent state stays on the device between calls: actor_predict reads the previous call's
    h / c there and the episode-start mask is applied by the LSTM kernel, so a vector step costs one
    H2D of the observations (pinned staging) and one D2H of [q-values | h | c] -- the copy of h, c is
    needed because the reference's acting schema stores them with every transition
    (acting/acting_interface.py:83-90, lstm.py:157-161)."""

    def __init__(self, learner, num_actions):
        self.learner = learner
        self.num_actions = num_actions
        self._dev_state = None    # output slot holding the (h, c) produced by the last actor_predict
        self._host_state = None   # their host copies (what make_input_state hands out)
        self._buf_key = None

    def is_recurrent(self):
        return self.learner.U > 0

    def get_state(self):
        return self.learner.get_state()

    def load_state(self, state):
        self.learner.load_state(state)

    def make_input_state(self, inp, initials):
        """SequentialModel.make_input_state + LSTM.get_state (sequential.py:128-144, lstm.py:131-161):
        {'x': obs, 'layer{i}_state': ...} with the last recurrent state masked by the episode starts."""
        initials = np.asarray(initials).astype("float32")
        L = self.learner
        state = {"x": inp}
        li = L.lstm_module if L.U else -1
        for i in range(L.fc_module + 1):
            state["layer%d_state" % i] = {}
        if self.is_recurrent():
            E, U = len(initials), L.U
            if self._host_state is None:
                assert np.all(initials), "first call must be all-initial states"
                hx = np.zeros((E, U), np.float32)
                cx = np.zeros((E, U), np.float32)
            else:
                hx, cx = self._host_state
            mask = (1 - initials)[:, None]
            self._handed_hx = hx * mask      # identity tag: actor_predict recognises its own state
            state["layer%d_state" % li] = {"hx": self._handed_hx, "cx": cx * mask, "initials": initials}
        return state

    def _buffers(self, E, obs_shape, obs_dtype):
        """Persistent staging for E envs: ONE pinned input block [obs | initials | extra | taus] with its device
        twin and two alternating device output blocks [q | h | c] (the previous call's block is this call's
        recurrent input), so that every pointer rt_learner_act sees repeats and the step replays from a graph."""
        import torch
        key = (E, tuple(obs_shape), np.dtype(obs_dtype).str)
        if self._buf_key == key:
            return
        L = self.learner
        up = lambda n: (int(n) + 255) // 256 * 256
        obs_bytes = int(np.prod(obs_shape)) * np.dtype(obs_dtype).itemsize
        self._off = off = {"ini": up(obs_bytes)}
        off["extra"] = off["ini"] + up(4 * E)
        off["tau"] = off["extra"] + up(4 * E * L.X)
        off["end"] = off["tau"] + up(4 * E * max(L.Nq, 1))
        self._in_host = torch.empty(off["end"], dtype=torch.uint8).pin_memory()
        self._in_dev = torch.empty(off["end"], dtype=torch.uint8, device=L.device)
        host = self._in_host.numpy()
        self._x_np = host[:obs_bytes].view(obs_dtype).reshape(obs_shape)
        self._ini_np = host[off["ini"]:off["ini"] + 4 * E].view(np.float32)
        self._extra_np = host[off["extra"]:off["extra"] + 4 * E * L.X].view(np.float32).reshape(E, L.X)
        self._tau_np = host[off["tau"]:off["tau"] + 4 * E * max(L.Nq, 1)].view(np.float32)
        A, U = self.num_actions, L.U
        self._out_dev = [torch.zeros(E * A + 2 * E * U, dtype=torch.float32, device=L.device) for _ in range(2)]
        self._out_host = torch.empty(E * A + 2 * E * U, dtype=torch.float32).pin_memory()
        self._slot = 0
        self._dev_state = None
        self._buf_key = key

    def actor_predict(self, state, timesteps=1, for_eval=False, taus=None):
        """DQNPolicy.actor_predict (policies/torch/dqn.py:132-148) for one time-step."""
        import torch
        assert timesteps == 1, "acting runs one time-step at a time"
        L = self.learner
        obs = state["x"]
        extra = None
        if isinstance(obs, (tuple, list)):
            extra = np.concatenate([np.asarray(e, dtype=np.float32).reshape(len(e), -1) for e in obs[1:]], axis=1)
            obs = obs[0]
        obs = np.asarray(obs)
        E, A, U = obs.shape[0], self.num_actions, L.U
        self._buffers(E, obs.shape, L.obs_dtype)
        off, base = self._off, self._in_dev.data_ptr()
        self._x_np[...] = obs
        null = C.c_void_p()
        ex_p = null
        if L.X:
            assert extra is not None and extra.shape[1] == L.X, "tuple observation with %d extra features expected" % L.X
            self._extra_np[...] = extra
            ex_p = C.c_void_p(base + off["extra"])
        out, prev = self._out_dev[self._slot], self._out_dev[1 - self._slot]
        if U:
            ls = state["layer%d_state" % L.lstm_module]
            self._ini_np[...] = np.asarray(ls["initials"], dtype=np.float32)
            if not (self._dev_state == 1 - self._slot and ls["hx"] is getattr(self, "_handed_hx", None)):
                # a state this policy did not produce itself: upload it where the previous call's would be.
                # (Its own state is still on the device; the LSTM kernel applies the episode-start mask --
                # the same values as the host-side masking of make_input_state.)
                hc = np.stack([np.asarray(ls["hx"], dtype=np.float32), np.asarray(ls["cx"], dtype=np.float32)])
                prev[E * A:].copy_(torch.from_numpy(hc.reshape(-1)))
            f4 = lambda t, o: C.c_void_p(t.data_ptr() + 4 * o)
            ptrs = [C.c_void_p(base), ex_p, f4(prev, E * A), f4(prev, E * A + E * U), C.c_void_p(base + off["ini"])]
            outs = [f4(out, 0), f4(out, E * A), f4(out, E * A + E * U)]
        else:
            ptrs = [C.c_void_p(base), ex_p, null, null, null]
            outs = [C.c_void_p(out.data_ptr()), null, null]
        tp = null
        if taus is not None:
            self._tau_np[...] = np.asarray(taus, dtype=np.float32).reshape(-1)
            tp = C.c_void_p(self._in_host.data_ptr() + off["tau"])     # pinned: the library's own copy stays async
        self._in_dev[:off["tau"]].copy_(self._in_host[:off["tau"]], non_blocking=True)
        _lib.check(L._lib.rt_learner_act(L._h, E, *ptrs, tp, *outs, L._stream()))
        self._out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(L.device).synchronize()
        res = self._out_host.numpy().copy()
        if U:
            self._dev_state = self._slot
            self._slot = 1 - self._slot
            self._host_state = (res[E * A:E * A + E * U].reshape(E, U), res[E * A + E * U:].reshape(E, U))
        qv = res[:E * A].reshape(E, A)
        return {"actions": np.argmax(qv, axis=1), "qvalues": qv}


class _ValueLog:
    """The slice of rltime.general.value_log.ValueLog this trainer needs: grouped keys aggregated by
    mean / sum / max since the last get()."""

    def __init__(self):
        self.data = {}

    def log(self, key, val, agg="mean", group=None, keep=False):
        self.data.setdefault((group, key), [agg, [], keep])[1].append(float(val))

    def get(self):
        out = {}
        for (group, key), (agg, vals, keep) in list(self.data.items()):
            if not vals:
                continue
            v = {"mean": np.mean, "sum": np.sum, "max": np.max}[agg](vals)
            dst = out
            if group:
                for g in group.split("->"):
                    dst = dst.setdefault(g, {})
            dst[key] = float(v)
            if not keep:
                self.data[(group, key)][1] = []
        return out


class IQNTrainer:
    """Same call surface as the reference's `IQN` trainer: IQNTrainer(logger, actors, model_config,
    policy_args).train(**training_args).  `DQNTrainer` below is the non-distributional sibling
    (training/torch/dqn.py: type "dqn", e.g. Rainbow-style dueling + double-Q + n-step + PER)."""
    POLICY = "iqn"

    def __init__(self, logger, actors, model_config, policy_args=None):
        self.logger = logger
        self.actors = actors
        self.model_config = model_config
        self.policy_args = dict(policy_args or {})
        self.steps = 0
        self.updates = 0
        self.log = {}
        self.policy = None
        self.value_log = _ValueLog()
        self._timer = None
        # test / reproducibility hooks: (online, target) state_dicts to start from instead of a fresh
        # initialisation, and a callable update_index -> [tau_target, tau_select, tau_train] replacing the
        # device RNG (the reference draws the fractions with torch.rand, policies/torch/iqn.py:88)
        self.initial_state_dicts = None
        self.tau_source = None

    # -- PolicyTrainer._start_timer / _end_timer (policy_trainer.py:228-244) ----------------------
    def _start_timer(self, name):
        self._timer = (name, time.time())

    def _end_timer(self):
        name, t0 = self._timer
        ms = (time.time() - t0) * 1000.0
        self.value_log.log(name, ms, agg="mean", group="timings_mean_ms")
        self.value_log.log(name, ms, agg="sum", group="timings_total_ms")

    # -- PolicyTrainer.init_policies -------------------------------------------------------
    def _build(self, t):
        obs_space, act_space = self.actors.get_spaces()
        m = parse_model_config(self.model_config)
        in_shape, extra_dim = parse_observation_space(obs_space)
        pa = self.policy_args
        assert self.POLICY == "dqn" or pa.get("injection_layer", -1) == -1, "only injection_layer=-1 is supported"
        assert not pa.get("dueling_value_layer_hidden_size"), "dueling_value_layer_hidden_size: only the default"
        mbatch = t["mbatch_size"] or self.actors.get_env_count()
        nstep_target = t["nstep_target"] or t["nstep_train"]
        rnn_steps = t["rnn_steps_train"] or t["nstep_train"]
        assert t["nstep_train"] % rnn_steps == 0, "nstep_train must be divisible by rnn_steps_train"
        self.learner = DeviceLearner(
            in_shape, m["conv"], m["lstm_units"], m["fc_size"], int(act_space.n),
            pa.get("num_sampling_quantiles", 32), pa.get("embedding_dim", 64), pa.get("dueling", False),
            mbatch=mbatch, nstep_train=t["nstep_train"], burn_in=t["burn_in_timesteps"],
            nstep_target=nstep_target, gamma=t["gamma"], double_q=t["double_q"],
            rnn_bootstrap=t["rnn_bootstrap"], vf_scale_epsilon=t["vf_scale_epsilon"],
            huber_kappa=t["huber_kappa"], clip_grad=t["clip_grad"], adam_epsilon=t["adam_epsilon"],
            # the reference's train_init ignores `lr` (Adam default 1e-3) until set_lr runs
            lr=1e-3, loss_aggregation=t["loss_aggregation"], seed=t.get("seed", 0), policy=self.POLICY,
            loss_mode=t["loss_mode"], loss_timestep_aggregation=t["loss_timestep_aggregation"],
            clip_grad_dynamic_alpha=t["clip_grad_dynamic_alpha"], pre_fc=m["pre_fc"], extra_dim=extra_dim,
            rnn_steps_train=rnn_steps, gemm=pa.get("gemm", "tf32"))
        lstm_units = m["lstm_units"]
        if self.initial_state_dicts is not None:
            p0, p1 = self.initial_state_dicts
        else:
            p0 = init_params(self.learner.param_info, lstm_units, seed=t.get("seed", 0))
            p1 = p0 if not t["target_update_freq"] else \
                init_params(self.learner.param_info, lstm_units, seed=t.get("seed", 0) + 1)
        self.learner.load_state_dict(p0, _lib.RT_BUF_ONLINE)
        self.learner.load_state_dict(p1, _lib.RT_BUF_TARGET)
        self.policy = DevicePolicy(self.learner, int(act_space.n))
        self.actors.set_actor_policy(self.policy)
        hist_cls = t["history_mode"].get("type", "replay")
        if isinstance(hist_cls, str):
            hist_cls = _history_types()[hist_cls]
        g = t["gamma"]
        self.history_buffer = hist_cls(
            **t["history_mode"].get("args", {}), nstep_target=nstep_target, nstep_train=t["nstep_train"],
            prefix_steps=t["burn_in_timesteps"], discount_function=lambda n, r, po: (g ** n) * r)
        self._assert_nsteps = t["rnn_bootstrap"]
        return mbatch, nstep_target

    # -- device batch of a host-side (online) history buffer ------------------------------------
    def _host_batch(self, td, nstep_target):
        """Builds the learner's batch view from a (T, B, ...) train-data dict of host arrays / tensors
        with separately stacked target states (history.py:268-270; OnlineHistoryBuffer)."""
        import torch
        dev = self.learner.device

        def dev_t(x, dtype=None):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            t = t.to(dev)
            return t.contiguous() if dtype is None else t.to(dtype).contiguous()
        L = self.learner

        def leaves(states):
            x = states["x"]
            out = {}
            if isinstance(x, (tuple, list)):
                out["extra"] = torch.cat([dev_t(e, torch.float32).reshape(e.shape[0], e.shape[1], -1) for e in x[1:]], -1)
                x = x[0]
            out["x"] = dev_t(x, torch.uint8 if L.obs_dtype == np.uint8 else torch.float32)
            if L.U:
                ls = states["layer%d_state" % L.lstm_module]
                out.update({k: dev_t(ls[k], torch.float32) for k in ("hx", "cx", "initials")})
            return out
        s, tg = leaves(td["states"]), leaves(td["target_states"])
        b, keep = batch_from_tensors(
            s["x"], s.get("hx"), s.get("cx"), s.get("initials"), dev_t(td["returns"], torch.float64),
            dev_t(td["nsteps"], torch.int64), dev_t(td["target_masks"], torch.float64),
            dev_t(td["policy_outputs"]["actions"], torch.int64), None, nstep_target,
            all_extra=s.get("extra"), targets=tg)
        self._host_keep = keep + list(tg.values())
        return b

    def train(self, total_steps, log_freq=10000, target_update_freq=0, clip_rewards=False,
              early_stop_steps=None, episode_history_windows=(10, 100), *, gamma, nstep_train, lr,
              history_mode={"type": "replay"}, mbatch_size=None, nstep_target=None, lr_anneal=False,
              epochs=1, minibatches=1, warmup_steps=0, actor_update_frequency_steps=1000,
              burn_in_timesteps=0, rnn_steps_train=None, rnn_bootstrap=False, async_history=False,
              clip_grad=None, clip_grad_dynamic_alpha=None, adam_epsilon=1e-8, vf_scale_epsilon=None,
              double_q=False, loss_mode="huber", huber_kappa=1.0, loss_aggregation="mean",
              loss_timestep_aggregation=None, seed=0):
        assert epochs == 1 and minibatches == 1, "epochs / minibatches > 1 are PPO options"
        assert loss_mode == "huber" or self.POLICY == "dqn", "IQN always uses the quantile-Huber loss"
        # async_history (multi_step_trainer.py:143-150, history/parallel_history.py:118-130) asks for the
        # batch of update k+1 to be produced while update k trains.  The device buffers do exactly that
        # inside the library (the replay stream's draw + gather overlap the backward pass) with the
        # reference's synchronous priority order, so the flag is accepted and needs no second process.
        self.async_history = bool(async_history)
        t = dict(locals())
        t.pop("self")
        mbatch, n_target = self._build(t)
        learner, hist = self.learner, self.history_buffer
        env_count = self.actors.get_env_count()
        self.actors.update_state(progress=0.0)
        base_lr = lr
        t_start = self._ts_start = time.time()
        self._ts_steps = self._ts_trained = 0
        actors_last_update_steps = 0
        rnn_steps = rnn_steps_train or nstep_train
        while True:
            progress = self.steps / total_steps
            if progress >= 1.0 or (early_stop_steps is not None and self.steps >= early_stop_steps):
                break
            warming_up = self.steps < warmup_steps
            n = hist.needed_feed_count(mbatch, env_count)
            if n is not None:
                if warming_up:
                    n = max(n, env_count)
                self._start_timer("sample_actors")
                samples = self.actors.get_samples(n)
                if samples:
                    for s in samples:
                        if s["done"]:
                            self.value_log.log("episodes", 1, agg="sum", group="this_interval")
                    if clip_rewards:
                        for s in samples:
                            s["reward"] = np.sign(s["reward"])
                    before = self.steps
                    self.steps += len(samples)
                    self._ts_steps += len(samples)
                    if target_update_freq > 0 and \
                            self.steps // target_update_freq != before // target_update_freq:
                        learner.sync_target()
                        self.log["target_syncs"] = self.log.get("target_syncs", 0) + 1
                    if self.steps // log_freq != before // log_freq:
                        self._log_checkpoint(t_start)
                    self._end_timer()
                    self._start_timer("history_update")
                    hist.update(samples)
                    self._end_timer()
            self._start_timer("get_train_data")
            # device buffers: draw() = get_train_data without building the nested dict of tensor views
            # (the learner reads the raw device batch)
            td = (hist.draw if hasattr(hist, "draw") else hist.get_train_data)(mbatch, train_progress=progress)
            if td is None:
                continue
            self._end_timer()
            if warming_up:
                continue
            self._start_timer("train")
            device_hist = hasattr(hist, "last_batch") and hasattr(hist, "_stream")
            if device_hist:
                # replay buffers guarantee the fixed n-step that rnn_bootstrap needs (multi_step_trainer.py:299-303)
                batch = hist.last_batch
                learner.prefetch(batch, hist._stream())      # frame conversion on the replay stream
            else:
                assert not rnn_bootstrap or np.all(np.asarray(td["nsteps"]) == n_target), \
                    "rnn_bootstrap is only supported with history buffers which guarantee the fixed target nstep"
                batch = self._host_batch(td, n_target)      # host-side buffer (online history)
            learner.step(batch, None if self.tau_source is None else self.tau_source(self.updates))
            if hasattr(hist, "update_losses_device"):
                hist.update_losses_device(learner.td_abs(), ready=learner.wait_loss)
            if not target_update_freq:
                # no separate target policy: the reference bootstraps from the online policy object itself
                # (policy_trainer.py:56-58), i.e. always from the current weights
                learner.sync_target()
            self.updates += 1
            self._ts_trained += mbatch * nstep_train
            self.value_log.log("batch_size", mbatch * nstep_train, group="train")
            self.value_log.log("steps_trained", mbatch * nstep_train, agg="sum", group="this_interval")
            if lr_anneal not in (False, None):
                anneal_to = 0.0 if lr_anneal is True else float(lr_anneal)
                lr = base_lr - progress * (base_lr - anneal_to)
                learner.set_lr(lr)
            self.value_log.log("lr", lr, group="train")
            # multi_step_trainer.py:363-373
            if not actor_update_frequency_steps or \
                    self.steps - actors_last_update_steps >= actor_update_frequency_steps:
                self.actors.update_state(progress=progress)
                actors_last_update_steps = self.steps
            self._end_timer()
        self._log_checkpoint(t_start)
        if hasattr(hist, "close") and async_history:
            hist.close()
        logging.getLogger().info("Training finished")

    def _log_checkpoint(self, t_start):
        """PolicyTrainer._log_checkpoint (policy_trainer.py:187-226): same group / key names."""
        now = time.time()
        st = self.learner.stats() if self.updates else {}
        for k, v in st.items():
            self.value_log.log(k, v, group="train")
        dt = now - self._ts_start + 1e-5
        vl = self.value_log
        vl.log("steps_acted_per_second", int(self._ts_steps / dt), group="this_interval")
        vl.log("steps_trained_per_second", int(self._ts_trained / dt), group="this_interval")
        vl.log("train_ratio", self._ts_trained / max(self._ts_steps, 1), group="this_interval")
        vl.log("seconds", now - self._ts_start, group="this_interval")
        vl.log("seconds", now - t_start, group="total")
        vl.log("steps_acted", self._ts_steps, group="this_interval")
        vl.log("steps_acted", self.steps, group="total")
        info = vl.get()
        self.log.update({"steps": self.steps, "updates": self.updates, "seconds": now - t_start,
                         **{"train." + k: v for k, v in st.items()}})
        info.update({k: v for k, v in self.log.items() if k not in info})
        self.last_log = info
        if self.logger is not None:
            if hasattr(self.logger, "log_result"):
                self.logger.log_result("train", info, self.steps)
            if hasattr(self.logger, "save_checkpoint"):
                # same payload as PolicyTrainer._save_checkpoint (policy_trainer.py:175-185); train_state
                # additionally carries what a resume needs (the reference stores {} and cannot resume)
                self.logger.save_checkpoint({"policy_state": self.learner.get_state(),
                                             "train_state": self.learner.optimizer_state()}, self.steps)
        self._ts_start = now
        self._ts_steps = self._ts_trained = 0


class DQNTrainer(IQNTrainer):
    """Reference `DQN` trainer surface (training/torch/dqn.py:8-50) on the same device engine:
    plain DQNPolicy heads (dueling optional), huber / mse TD loss, signed TD errors reported to
    the prioritized buffer."""
    POLICY = "dqn"
