"""End-to-end: IQNTrainer (reference trainer call surface) driving fake envs through the device
history buffer, the CUDA learner and rt_learner_act acting inference."""
import io

import numpy as np
import pytest

MODEL = {"type": "sequential", "args": {"layer_configs": [
    {"type": "cnn", "args": {"layers": [{"filters": 16, "kernel": 8, "stride": 4},
                                         {"filters": 16, "kernel": 4, "stride": 2}]}},
    {"type": "lstm", "args": {"num_units": 64}},
    {"type": "fc", "args": {"fc_size": 64}}]}}


class _Logger:
    def __init__(self):
        self.results, self.checkpoints = [], []

    def log_result(self, name, info, step):
        self.results.append((name, dict(info), step))

    def save_checkpoint(self, data, step):
        self.checkpoints.append((data, step))


@pytest.mark.gpu
def test_iqn_trainer_end_to_end_with_fake_envs():
    import torch
    from rltime_b200.training import IQNTrainer
    from tests.fake_actor import FakeVecActor
    actors = FakeVecActor(num_envs=4, num_actions=4, seed=1)
    logger = _Logger()
    tr = IQNTrainer(logger, actors, MODEL, {"dueling": True, "num_sampling_quantiles": 8, "embedding_dim": 16})
    tr.train(total_steps=1200, log_freq=400, target_update_freq=300, clip_rewards=True, gamma=0.99,
             nstep_train=4, nstep_target=2, lr=3e-4, lr_anneal=True, mbatch_size=4, warmup_steps=200,
             burn_in_timesteps=2, rnn_bootstrap=True, double_q=True, clip_grad=40.0, adam_epsilon=1e-5,
             history_mode={"type": "prioritized_replay",
                           "args": {"size": 600, "train_frequency": 4, "alpha": 0.9, "beta": 0.6,
                                    "max_envs": 4}})
    assert tr.steps >= 1200 and tr.updates > 100
    assert tr.log["target_syncs"] >= 3
    st = tr.learner.stats()
    assert np.isfinite(st["qloss"]) and np.isfinite(st["grad_norm"]) and st["grad_norm"] > 0
    # train quota: ~train_frequency trained transitions per acted transition after warm-up
    trained = tr.updates * 4 * 4
    assert 0.5 < trained / (4.0 * (tr.steps - 200)) < 1.5
    # checkpoints carry a torch-loadable state_dict with the reference's parameter names
    assert logger.checkpoints and logger.results
    sd = torch.load(io.BytesIO(logger.checkpoints[-1][0]["policy_state"]), map_location="cpu")
    assert "model.layers.1.lstm_cell.weight_hh" in sd and "quantile_layer.weight" in sd and \
        "embedding_range" in sd
    # the weights moved away from their initialisation and stayed finite
    from rltime_b200.init import init_params
    p0 = init_params(tr.learner.param_info, 64, seed=0)
    moved = max(float((sd[k] - p0[k]).abs().max()) for k in p0)
    assert moved > 1e-4 and all(torch.isfinite(v).all() for v in sd.values())
    # acting inference is consistent with the learner's own forward: greedy action = argmax of q
    pred = tr.policy.actor_predict(actors.last_state, timesteps=1)
    assert pred["qvalues"].shape == (4, 4) and (pred["actions"] == pred["qvalues"].argmax(1)).all()


MODEL_FF = {"type": "sequential", "args": {"layer_configs": [
    {"type": "cnn", "args": {"layers": [{"filters": 16, "kernel": 8, "stride": 4},
                                         {"filters": 16, "kernel": 4, "stride": 2}]}},
    {"type": "fc", "args": {"fc_size": 64}}]}}


@pytest.mark.gpu
def test_dqn_trainer_rainbow_style_end_to_end():
    """BASELINE config 4 family: Rainbow-style DQN (dueling + double-Q + n-step 3 + prioritized
    replay of single transitions with the min tree for global importance scaling)."""
    import torch
    from rltime_b200.training import DQNTrainer
    from tests.fake_actor import FakeVecActor
    actors = FakeVecActor(num_envs=4, num_actions=4, seed=2)
    logger = _Logger()
    tr = DQNTrainer(logger, actors, MODEL_FF, {"dueling": True})
    tr.train(total_steps=1000, log_freq=500, target_update_freq=250, clip_rewards=True, gamma=0.99,
             nstep_train=1, nstep_target=3, lr=1e-4, mbatch_size=16, warmup_steps=200, double_q=True,
             clip_grad=10.0, adam_epsilon=1.5e-4, loss_mode="huber",
             history_mode={"type": "prioritized_replay",
                           "args": {"size": 500, "train_frequency": 4, "alpha": 0.5, "beta": 0.4,
                                    "beta_anneal": True, "global_importance_scaling": True,
                                    "max_envs": 4}})
    assert tr.steps >= 1000 and tr.updates > 100
    st = tr.learner.stats()
    assert np.isfinite(st["qloss"]) and np.isfinite(st["qvalue"]) and st["grad_norm"] > 0
    sd = torch.load(io.BytesIO(logger.checkpoints[-1][0]["policy_state"]), map_location="cpu")
    assert "value_hidden_layer.weight" in sd and "quantile_layer.weight" not in sd
    pred = tr.policy.actor_predict(actors.last_state, timesteps=1)
    assert pred["qvalues"].shape == (4, 4) and (pred["actions"] == pred["qvalues"].argmax(1)).all()


@pytest.mark.gpu
def test_pipelined_update_loop_is_bit_identical_to_the_serialised_one():
    """The priority write-back / next draw overlap the backward pass (own replay stream +
    rt_learner_wait_loss) and the update is replayed from CUDA graphs: sampled indices, |td| and
    the trained weights must not change by a bit against the same loop with one-by-one launches and
    a full device synchronisation between every call."""
    import random
    import torch
    from rltime_b200.history import DevicePrioritizedReplayHistoryBuffer
    from rltime_b200.init import init_params
    from rltime_b200.learner import DeviceLearner
    E, T, n, B, A, U = 8, 6, 2, 8, 4, 64
    dev = torch.device("cuda", 0)

    def run(pipelined):
        import os
        # the serialised run also issues every launch one by one; the pipelined run replays the
        # update from CUDA graphs (read at learner construction)
        os.environ["RT_GRAPHS"] = "1" if pipelined else "0"
        random.seed(5)
        rs = np.random.RandomState(3)
        hist = DevicePrioritizedReplayHistoryBuffer(
            size=4096, train_frequency=None, alpha=0.9, beta=0.6, nstep_target=n, nstep_train=T,
            prefix_steps=0, gamma=0.99, max_envs=E, device=dev)
        L = DeviceLearner((4, 84, 84), [(16, 8, 4), (16, 4, 2)], U, 64, A, 8, 16, True, mbatch=B,
                          nstep_train=T, nstep_target=n, double_q=True, rnn_bootstrap=True,
                          clip_grad=40.0, seed=11, device=dev)
        L.load_state_dict(init_params(L.param_info, U, seed=1), 0)
        L.load_state_dict(init_params(L.param_info, U, seed=2), 1)

        def feed(m_steps):
            m = m_steps * E
            env = np.arange(m) % E
            hist.update_arrays(env, np.sign(rs.randn(m)), (rs.rand(m) < 0.02).astype(np.uint8),
                               [rs.randint(0, 255, (m, 4, 84, 84)).astype(np.uint8),
                                rs.randn(m, U).astype(np.float32), rs.randn(m, U).astype(np.float32),
                                np.zeros(m, np.float32)],
                               [rs.randint(0, A, m).astype(np.int64), rs.randn(m, A).astype(np.float32)])
        feed(200)
        idx_log, td_log = [], []
        for it in range(40):
            if it % 3 == 0:
                feed(2)
            assert hist.get_train_data(B, 0.0) is not None
            if not pipelined:
                torch.cuda.synchronize()
            L.step(hist.last_batch)
            if pipelined:
                hist.update_losses_device(L.td_abs(), ready=L.wait_loss)
            else:
                torch.cuda.synchronize()
                hist.update_losses_device(L.td_abs())
                torch.cuda.synchronize()
            # async device-side copies on the caller's stream: no host sync inside the loop
            idx_log.append(hist._last_idx_tensor.clone())
            td_log.append(L.td_abs().clone())
        torch.cuda.synchronize()
        # the early loss read-back (after the forward pass) sees the same values as stats()
        early, full = L.loss(), L.stats()
        assert early["qloss"] == full["qloss"] and early["td_mean"] == full["td_mean"]
        first, count = L.wait_late_grads()
        assert 0 < first and first + count == L.flat().numel()
        return ([t.cpu().tolist() for t in idx_log], torch.stack(td_log).cpu().numpy(),
                L.flat(0).cpu().numpy().copy(), hist.tree_sum())

    try:
        a, b = run(False), run(True)
    finally:
        import os
        os.environ.pop("RT_GRAPHS", None)
    assert a[0] == b[0]
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[2], b[2])
    assert a[3] == b[3]
