"""Drop-in evidence with the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh):

1. the reference's own IQN trainer (rltime/training/torch/iqn.py through PolicyTrainer.train /
   MultiStepTrainer._train) runs with `history_mode.type = rltime_b200.history.PrioritizedReplayHistoryBuffer`
   plugged in through its type registry (general/type_registry.py:29-39: a class is taken as-is);
2. SURVEY.md section 4 integration test: the reference trainer (its own PER buffer, torch CPU fp32) and
   rltime_b200.training.IQNTrainer (device PER buffer + CUDA learner) consume identical seeded
   transitions, identical initial weights and identical quantile fractions: the sampled prioritization
   indices are bit-identical and the per-row |td| signal agrees within 1e-4 on every one of >= 20 updates.
"""
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_arm  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_arm.available(), reason="baseline/_ref not installed")]

E, T, N_TARGET, B, A, U, NQ = 4, 6, 2, 4, 4, 32, 8
FRAME = (4, 84, 84)
MODEL = {"type": "sequential", "args": {"layer_configs": [
    {"type": "cnn", "args": {"layers": [{"filters": 16, "kernel": 8, "stride": 4},
                                         {"filters": 16, "kernel": 4, "stride": 2}]}},
    {"type": "lstm", "args": {"num_units": U}},
    {"type": "fc", "args": {"fc_size": 48}}]}}
POLICY_ARGS = dict(dueling=True, embedding_dim=16, num_sampling_quantiles=NQ, injection_layer=-1)
TRAIN = dict(log_freq=10 ** 9, target_update_freq=96, clip_rewards=True, gamma=0.99, mbatch_size=B,
             nstep_train=T, nstep_target=N_TARGET, lr=1e-3, adam_epsilon=1e-5, double_q=True,
             rnn_bootstrap=True, clip_grad=40.0, warmup_steps=160)
PER = dict(size=2000, train_frequency=4, alpha=0.9, beta=0.6)
FEED = (B * T) // PER["train_frequency"]


def _stream(seed=3):
    from rltime_b200.synthetic import SyntheticStream
    return SyntheticStream(num_envs=E, frame_shape=FRAME, num_actions=A, lstm_units=U, seed=seed,
                           done_mode="bernoulli", done_p=0.03, pool=64)


class _Actors(ref_arm.SyntheticActors):
    """Synthetic transitions + a snapshot of the trainer's initial online / target weights."""

    def __init__(self, stream, trainer_ref=None):
        super().__init__(stream, threads=0)
        self.trainer_ref = trainer_ref
        self.initial = None

    def set_actor_policy(self, policy):
        self.policy = policy
        tr = self.trainer_ref() if self.trainer_ref else None
        if tr is not None and hasattr(policy, "state_dict"):
            self.initial = ({k: v.detach().clone() for k, v in tr.policy.state_dict().items()},
                            {k: v.detach().clone() for k, v in tr.target_policy.state_dict().items()})


import torch  # noqa: E402

_TORCH_RAND = torch.rand     # the real one: _run_reference swaps torch.rand for a deterministic queue


def _tau(k, n):
    return _TORCH_RAND(n, generator=torch.Generator().manual_seed(1000 + k))


def _run_reference(updates):
    """The reference trainer with its own buffer on CPU; returns (idx log, |td| log, initial weights)."""
    ref_arm._import_reference()
    import torch
    from rltime.history.prioritized_replay_history import PrioritizedReplayHistoryBuffer
    from rltime.training.torch.iqn import IQN
    idx_log, td_log = [], []

    class Recording(PrioritizedReplayHistoryBuffer):
        def _sample_proportional(self, batch_size):
            res = super()._sample_proportional(batch_size)
            self._last = list(res)
            return res

        def update_losses(self, indices, losses):
            idx_log.append(list(self._last))
            td_log.append(np.array(losses, dtype=np.float32, copy=True))
            super().update_losses(indices, np.asarray(losses, dtype=np.float64))

    calls = [0]
    orig_rand = torch.rand

    def fake_rand(*size, device=None, **kw):
        assert len(size) == 1 and not kw
        t = _tau(calls[0], size[0])
        calls[0] += 1
        return t
    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    holder = {}
    actors = _Actors(_stream(), trainer_ref=lambda: holder.get("tr"))
    tr = holder["tr"] = IQN(logger=None, actors=actors, model_config=MODEL,
                           policy_args=dict(POLICY_ARGS, cuda=False))
    torch.rand = fake_rand
    try:
        tr.train(total_steps=TRAIN["warmup_steps"] + (updates + 1) * FEED,
                 history_mode={"type": Recording, "args": dict(PER)}, **TRAIN)
    finally:
        torch.rand = orig_rand
    return idx_log, td_log, actors.initial


def test_reference_trainer_vs_device_trainer_on_identical_transitions():
    import torch
    from rltime_b200.training import IQNTrainer
    updates = 24
    ref_idx, ref_td, initial = _run_reference(updates)
    assert len(ref_idx) >= updates
    random.seed(0)
    np.random.seed(0)
    actors = _Actors(_stream())
    tr = IQNTrainer(None, actors, MODEL, dict(POLICY_ARGS, gemm="fp32"))
    tr.initial_state_dicts = ({k: v for k, v in initial[0].items() if k != "embedding_range"},
                              {k: v for k, v in initial[1].items() if k != "embedding_range"})
    M = T * B
    tr.tau_source = lambda u: [_tau(3 * u + j, M * NQ) for j in range(3)]
    got_idx, got_td = [], []
    from rltime_b200.history import DevicePrioritizedReplayHistoryBuffer

    class Recording(DevicePrioritizedReplayHistoryBuffer):
        def update_losses_device(self, td_abs, ready=None):
            torch.cuda.synchronize()
            got_idx.append(self.last_sampled_idxes)
            got_td.append(td_abs.cpu().numpy().copy())
            super().update_losses_device(td_abs, ready=ready)
    tr.train(total_steps=TRAIN["warmup_steps"] + (updates + 1) * FEED,
             history_mode={"type": Recording, "args": dict(PER, max_envs=E)}, **TRAIN)
    assert len(got_idx) >= updates
    for u in range(updates):
        assert got_idx[u] == [int(i) for i in ref_idx[u]], "update %d: sampled indices differ" % u
        np.testing.assert_allclose(got_td[u], ref_td[u], rtol=0, atol=1e-4, err_msg="update %d |td|" % u)


def test_reference_trainer_runs_on_the_device_history_buffer():
    """history_mode.type = the device buffer class inside the reference's own trainer and policy (torch
    CUDA): the trainer's burn-in writes, reshapes and loss write-back all go through the borrowed device
    tensors of the plug-in."""
    ref_arm._import_reference()
    import torch
    from rltime.training.torch.iqn import IQN
    from rltime_b200.history import PrioritizedReplayHistoryBuffer as DevicePER
    random.seed(1)
    np.random.seed(1)
    torch.manual_seed(1)
    calls = []

    class Counting(DevicePER):
        def update_losses(self, indices, losses):
            calls.append(float(np.mean(np.abs(losses))))
            super().update_losses(indices, losses)
    actors = _Actors(_stream(seed=9))
    tr = IQN(logger=None, actors=actors, model_config=MODEL, policy_args=dict(POLICY_ARGS, cuda=True))
    updates = 12
    tr.train(total_steps=TRAIN["warmup_steps"] + (updates + 1) * FEED, burn_in_timesteps=2,
             history_mode={"type": Counting, "args": dict(PER, max_envs=E, output="numpy")}, **TRAIN)
    assert len(calls) >= updates and np.all(np.isfinite(calls))
    vals = tr.value_log.get()["train"]
    assert np.isfinite(vals["qloss"]) and vals["grad_norm"] > 0
