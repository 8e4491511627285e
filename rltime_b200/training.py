"""Trainer-side mirror: an IQN trainer with the reference's constructor / train() surface whose
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
from .learner import DeviceLearner


def parse_model_config(model_config):
    """{'type': 'sequential', 'args': {'layer_configs': [cnn, (lstm), fc]}} (configs/models/*.json)
    -> (conv list, lstm_units, fc_size)."""
    assert model_config.get("type") in ("sequential", None) or not isinstance(model_config.get("type"), str), \
        "only the sequential model family is supported"
    layers = model_config["args"]["layer_configs"]
    kinds = [l["type"] for l in layers]
    assert kinds in (["cnn", "lstm", "fc"], ["cnn", "fc"]), \
        "supported topologies: cnn -> [lstm] -> fc (got %s)" % kinds
    conv = [(int(l["filters"]), int(l["kernel"]), int(l["stride"])) for l in layers[0]["args"]["layers"]]
    lstm_units = int(layers[1]["args"]["num_units"]) if "lstm" in kinds else 0
    fc_args = layers[-1].get("args", {})
    assert fc_args.get("fc_count", 1) == 1 and not fc_args.get("batch_norm", False) and \
        fc_args.get("activation", "relu") == "relu", "only a single ReLU FC layer is supported"
    return conv, lstm_units, int(fc_args["fc_size"])


class DevicePolicy:
    """Actor-facing policy backed by the learner's online network (rt_learner_act)."""

    def __init__(self, learner, num_actions):
        self.learner = learner
        self.num_actions = num_actions
        self._last = None        # (h, c) device tensors of the previous act call

    def is_recurrent(self):
        return self.learner.U > 0

    def get_state(self):
        return self.learner.get_state()

    def load_state(self, state):
        self.learner.load_state(state)

    def make_input_state(self, inp, initials):
        """SequentialModel.make_input_state + LSTM.get_state: last recurrent state masked by
        the episode-start flags."""
        initials = np.asarray(initials).astype("float32")
        state = {"x": inp, "layer0_state": {}}
        if self.is_recurrent():
            E, U = len(initials), self.learner.U
            if self._last is None:
                assert np.all(initials), "first call must be all-initial states"
                hx = np.zeros((E, U), np.float32)
                cx = np.zeros((E, U), np.float32)
            else:
                hx, cx = (t.cpu().numpy() for t in self._last)
            mask = (1 - initials)[:, None]
            state["layer1_state"] = {"hx": hx * mask, "cx": cx * mask, "initials": initials}
            state["layer2_state"] = {}
        else:
            state["layer1_state"] = {}
        return state

    def actor_predict(self, state, timesteps=1, for_eval=False):
        import torch
        assert timesteps == 1, "acting runs one time-step at a time"
        L = self.learner
        dev = L.device
        x = torch.as_tensor(np.ascontiguousarray(state["x"]), device=dev)
        assert x.dtype == torch.uint8, "observations must be uint8 frames"
        E = x.shape[0]
        q = torch.empty(E, self.num_actions, dtype=torch.float32, device=dev)
        null = C.c_void_p()
        if L.U:
            ls = state["layer1_state"]
            hx = torch.as_tensor(np.ascontiguousarray(ls["hx"], dtype=np.float32), device=dev)
            cx = torch.as_tensor(np.ascontiguousarray(ls["cx"], dtype=np.float32), device=dev)
            ini = torch.as_tensor(np.ascontiguousarray(ls["initials"], dtype=np.float32), device=dev)
            h_out, c_out = torch.empty_like(hx), torch.empty_like(cx)
            ptrs = [C.c_void_p(t.data_ptr()) for t in (x, hx, cx, ini)]
            outs = [C.c_void_p(t.data_ptr()) for t in (q, h_out, c_out)]
        else:
            ptrs = [C.c_void_p(x.data_ptr()), null, null, null]
            outs = [C.c_void_p(q.data_ptr()), null, null]
        _lib.check(L._lib.rt_learner_act(L._h, E, *ptrs, null, *outs, L._stream()))
        if L.U:
            self._last = (h_out, c_out)
        qv = q.cpu().numpy()
        return {"actions": np.argmax(qv, axis=1), "qvalues": qv}


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

    # -- PolicyTrainer.init_policies -------------------------------------------------------
    def _build(self, t):
        obs_space, act_space = self.actors.get_spaces()
        conv, lstm_units, fc_size = parse_model_config(self.model_config)
        pa = self.policy_args
        assert self.POLICY == "dqn" or pa.get("injection_layer", -1) == -1, "only injection_layer=-1 is supported"
        mbatch = t["mbatch_size"] or self.actors.get_env_count()
        nstep_target = t["nstep_target"] or t["nstep_train"]
        self.learner = DeviceLearner(
            tuple(obs_space.shape), conv, lstm_units, fc_size, int(act_space.n),
            pa.get("num_sampling_quantiles", 32), pa.get("embedding_dim", 64), pa.get("dueling", False),
            mbatch=mbatch, nstep_train=t["nstep_train"], burn_in=t["burn_in_timesteps"],
            nstep_target=nstep_target, gamma=t["gamma"], double_q=t["double_q"],
            rnn_bootstrap=t["rnn_bootstrap"], vf_scale_epsilon=t["vf_scale_epsilon"],
            huber_kappa=t["huber_kappa"], clip_grad=t["clip_grad"], adam_epsilon=t["adam_epsilon"],
            # the reference's train_init ignores `lr` (Adam default 1e-3) until set_lr runs
            lr=1e-3, loss_aggregation=t["loss_aggregation"], seed=t.get("seed", 0), policy=self.POLICY,
            loss_mode=t["loss_mode"], loss_timestep_aggregation=t["loss_timestep_aggregation"],
            clip_grad_dynamic_alpha=t["clip_grad_dynamic_alpha"])
        p0 = init_params(self.learner.param_info, lstm_units, seed=t.get("seed", 0))
        self.learner.load_state_dict(p0, _lib.RT_BUF_ONLINE)
        self.learner.load_state_dict(p0 if not t["target_update_freq"] else
                                     init_params(self.learner.param_info, lstm_units, seed=t.get("seed", 0) + 1),
                                     _lib.RT_BUF_TARGET)
        self.policy = DevicePolicy(self.learner, int(act_space.n))
        self.actors.set_actor_policy(self.policy)
        hist_cls = t["history_mode"].get("type", "replay")
        if isinstance(hist_cls, str):
            hist_cls = _history_types()[hist_cls]
        g = t["gamma"]
        self.history_buffer = hist_cls(
            **t["history_mode"].get("args", {}), nstep_target=nstep_target, nstep_train=t["nstep_train"],
            prefix_steps=t["burn_in_timesteps"], discount_function=lambda n, r, po: (g ** n) * r)
        return mbatch

    def train(self, total_steps, log_freq=10000, target_update_freq=0, clip_rewards=False,
              early_stop_steps=None, episode_history_windows=(10, 100), *, gamma, nstep_train, lr,
              history_mode={"type": "replay"}, mbatch_size=None, nstep_target=None, lr_anneal=False,
              epochs=1, minibatches=1, warmup_steps=0, actor_update_frequency_steps=1000,
              burn_in_timesteps=0, rnn_steps_train=None, rnn_bootstrap=False, async_history=False,
              clip_grad=None, clip_grad_dynamic_alpha=None, adam_epsilon=1e-8, vf_scale_epsilon=None,
              double_q=False, loss_mode="huber", huber_kappa=1.0, loss_aggregation="mean",
              loss_timestep_aggregation=None, seed=0):
        assert epochs == 1 and minibatches == 1, "epochs / minibatches > 1 are PPO options"
        assert rnn_steps_train in (None, nstep_train), "rnn_steps_train != nstep_train is not supported"
        assert loss_mode == "huber" or self.POLICY == "dqn", "IQN always uses the quantile-Huber loss"
        assert not async_history, "the device buffer needs no separate history process"
        t = dict(locals())
        t.pop("self")
        mbatch = self._build(t)
        learner, hist = self.learner, self.history_buffer
        env_count = self.actors.get_env_count()
        self.actors.update_state(progress=0.0)
        base_lr = lr
        t_start = time.time()
        while True:
            progress = self.steps / total_steps
            if progress >= 1.0 or (early_stop_steps is not None and self.steps >= early_stop_steps):
                break
            warming_up = self.steps < warmup_steps
            n = hist.needed_feed_count(mbatch, env_count)
            if n is not None:
                if warming_up:
                    n = max(n, env_count)
                samples = self.actors.get_samples(n)
                if samples:
                    if clip_rewards:
                        for s in samples:
                            s["reward"] = np.sign(s["reward"])
                    before = self.steps
                    self.steps += len(samples)
                    if target_update_freq > 0 and \
                            self.steps // target_update_freq != before // target_update_freq:
                        learner.sync_target()
                        self.log["target_syncs"] = self.log.get("target_syncs", 0) + 1
                    if self.steps // log_freq != before // log_freq:
                        self._log_checkpoint(t_start)
                    hist.update(samples)
            td = hist.get_train_data(mbatch, train_progress=progress)
            if td is None or warming_up:
                continue
            learner.step(hist.last_batch)
            if hasattr(hist, "update_losses_device"):
                hist.update_losses_device(learner.td_abs(), ready=learner.wait_loss)
            self.updates += 1
            if lr_anneal not in (False, None):
                anneal_to = 0.0 if lr_anneal is True else float(lr_anneal)
                learner.set_lr(base_lr - progress * (base_lr - anneal_to))
            if not actor_update_frequency_steps or self.updates % 16 == 0:
                self.actors.update_state(progress=progress)
        self._log_checkpoint(t_start)
        logging.getLogger().info("Training finished")

    def _log_checkpoint(self, t_start):
        st = self.learner.stats() if self.updates else {}
        self.log.update({"steps": self.steps, "updates": self.updates,
                         "seconds": time.time() - t_start, **{"train." + k: v for k, v in st.items()}})
        if self.logger is not None:
            if hasattr(self.logger, "log_result"):
                self.logger.log_result("train", dict(self.log), self.steps)
            if hasattr(self.logger, "save_checkpoint"):
                # same checkpoint payload as PolicyTrainer._save_checkpoint (policy_trainer.py:175-185)
                self.logger.save_checkpoint({"policy_state": self.learner.get_state(), "train_state": {}},
                                            self.steps)


class DQNTrainer(IQNTrainer):
    """Reference `DQN` trainer surface (training/torch/dqn.py:8-50) on the same device engine:
    plain DQNPolicy heads (dueling optional), huber / mse TD loss, signed TD errors reported to
    the prioritized buffer."""
    POLICY = "dqn"
