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


MODEL_MLP = {"type": "sequential", "args": {"layer_configs": [
    {"type": "fc", "args": {"fc_size": 64}}, {"type": "fc", "args": {"fc_size": 64}}]}}


@pytest.mark.gpu
def test_dqn_trainer_mlp_with_online_history():
    """BASELINE config 1 family: DQN on a 1-D float observation with the reference's mlp_2x64 model
    (configs/models/mlp_2x64.json) and the ONLINE n-step history buffer (a host-side structure: the
    trainer uploads its separately stacked states / target states, per-row n-steps)."""
    from rltime_b200.training import DQNTrainer
    from tests.fake_actor import FakeVecActor
    actors = FakeVecActor(num_envs=2, frame_shape=(4,), num_actions=2, seed=4, vector_obs=True)
    logger = _Logger()
    tr = DQNTrainer(logger, actors, MODEL_MLP, {"dueling": False})
    tr.train(total_steps=600, log_freq=200, target_update_freq=100, gamma=0.99, nstep_train=5, lr=1e-3,
             mbatch_size=2, double_q=True, history_mode={"type": "online"})
    assert tr.steps >= 600 and tr.updates >= 50
    st = tr.learner.stats()
    assert np.isfinite(st["qloss"]) and st["grad_norm"] > 0
    names = [n for n, _ in tr.learner.param_info]
    assert names[:4] == ["model.layers.0.layers.0.0.weight", "model.layers.0.layers.0.0.bias",
                         "model.layers.1.layers.0.0.weight", "model.layers.1.layers.0.0.bias"]
    pred = tr.policy.actor_predict(actors.last_state, timesteps=1)
    assert pred["qvalues"].shape == (2, 2)
    # the reference's log groups (policy_trainer.py:187-244)
    info = logger.results[-1][1]
    for grp in ("timings_mean_ms", "timings_total_ms", "this_interval", "train"):
        assert grp in info, grp
    assert {"get_train_data", "train", "sample_actors", "history_update"} <= set(info["timings_mean_ms"])
    assert {"steps_trained", "steps_acted", "train_ratio", "seconds"} <= set(info["this_interval"])


MODEL_LSTM1 = {"type": "sequential", "args": {"layer_configs": [
    {"type": "cnn", "args": {"layers": [{"filters": 16, "kernel": 8, "stride": 4},
                                         {"filters": 16, "kernel": 4, "stride": 2}]}},
    {"type": "lstm", "args": {"num_units": 64}},
    {"type": "fc", "args": {"fc_size": 64}}]}}


@pytest.mark.gpu
def test_iqn_trainer_tuple_observation_rnn_steps_async_flag():
    """The shipped configs/atari_iqn_lstm.json wrapper stack: single-frame observations plus the
    extra feature vector of ExtraFeaturesEnvWrapper, fed to the LSTM; rnn_steps_train < nstep_train;
    async_history=True is accepted (the device buffer prefetches inside the library)."""
    import torch
    from rltime_b200.training import IQNTrainer
    from tests.fake_actor import FakeVecActor
    actors = FakeVecActor(num_envs=4, frame_shape=(1, 84, 84), num_actions=4, seed=5, extra_dim=6)
    logger = _Logger()
    tr = IQNTrainer(logger, actors, MODEL_LSTM1, {"dueling": True, "num_sampling_quantiles": 8, "embedding_dim": 16})
    tr.train(total_steps=900, log_freq=300, target_update_freq=300, clip_rewards=True, gamma=0.99,
             nstep_train=8, nstep_target=2, rnn_steps_train=4, lr=3e-4, mbatch_size=4, warmup_steps=200,
             rnn_bootstrap=True, double_q=True, clip_grad=40.0, adam_epsilon=1e-5, async_history=True,
             actor_update_frequency_steps=64,
             history_mode={"type": "prioritized_replay",
                           "args": {"size": 600, "train_frequency": 4, "alpha": 0.9, "beta": 0.6, "max_envs": 4}})
    assert tr.steps >= 900 and tr.updates > 30
    st = tr.learner.stats()
    assert np.isfinite(st["qloss"]) and st["grad_norm"] > 0
    sd = torch.load(io.BytesIO(logger.checkpoints[-1][0]["policy_state"]), map_location="cpu")
    assert tuple(sd["model.layers.1.lstm_cell.weight_ih"].shape) == (4 * 64, 16 * 9 * 9 + 6)
    assert "adam_steps" in logger.checkpoints[-1][0]["train_state"]
    # actor refresh cadence follows actor_update_frequency_steps (multi_step_trainer.py:363-373)
    assert 3 <= actors.updates <= 2 + (tr.steps - 200) // 64 + 1


@pytest.mark.gpu
def test_acting_inference_matches_oracle():
    """rt_learner_act through DevicePolicy.actor_predict against the oracle's IQNPolicy
    forward at timesteps = 1 with injected quantile fractions: q-values within 1e-4, identical greedy
    actions, the carried LSTM state within 1e-5 -- over eight consecutive vector steps with episode
    resets in between (the state stays on the device between the calls; from the fifth step on the library
    replays the step from its CUDA graph)."""
    import torch
    from oracle import learner_oracle as lo
    from rltime_b200.learner import DeviceLearner
    from rltime_b200.training import DevicePolicy
    E, U, A, Nq = 6, 32, 5, 8
    conv = [(16, 8, 4), (16, 4, 2)]
    spec = lo.ModelSpec((4, 84, 84), conv, U, 48, A, Nq, 16, True)
    p = spec.init_params(3)
    L = DeviceLearner((4, 84, 84), conv, U, 48, A, Nq, 16, True, mbatch=E, nstep_train=2, nstep_target=1,
                      double_q=True, rnn_bootstrap=True, gemm="fp32")
    try:
        L.load_state_dict(p, 0)
        L.load_state_dict(p, 1)
        pol = DevicePolicy(L, A)
        rs = np.random.RandomState(0)
        h = torch.zeros(E, U)
        c = torch.zeros(E, U)
        initials = np.ones(E, dtype=bool)
        for step in range(8):
            obs = rs.randint(0, 256, (E, 4, 84, 84)).astype(np.uint8)
            state = pol.make_input_state(obs, initials)
            taus = torch.rand(E * Nq, generator=torch.Generator().manual_seed(step))
            got = pol.actor_predict(state, taus=taus.numpy())
            ini = torch.from_numpy(initials.astype(np.float32))
            st = {"x": torch.from_numpy(obs), "layer1_state": {"hx": h, "cx": c, "initials": ini}}
            with torch.no_grad():
                q, (h, c) = lo.predict(spec, p, st, 1, taus)
            want = q.mean(1).numpy()
            np.testing.assert_allclose(got["qvalues"], want, rtol=0, atol=1e-4)
            assert (got["actions"] == want.argmax(1)).all()
            np.testing.assert_allclose(pol._host_state[0], h.numpy(), rtol=0, atol=1e-5)
            np.testing.assert_allclose(pol._host_state[1], c.numpy(), rtol=0, atol=1e-5)
            initials = rs.rand(E) < 0.3
    finally:
        L.close()


@pytest.mark.gpu
def test_iqn_trainer_uniform_device_replay():
    """history_mode "replay" (the shipped atari_iqn_lstm.json uses the uniform buffer): the trainer takes the
    raw device batch of the uniform device buffer, which has no priority write-back."""
    from rltime_b200.training import IQNTrainer
    from tests.fake_actor import FakeVecActor
    actors = FakeVecActor(num_envs=4, num_actions=4, seed=6)
    tr = IQNTrainer(_Logger(), actors, MODEL, {"dueling": True, "num_sampling_quantiles": 8, "embedding_dim": 16})
    tr.train(total_steps=700, log_freq=350, target_update_freq=200, clip_rewards=True, gamma=0.99,
             nstep_train=4, nstep_target=2, lr=3e-4, mbatch_size=4, warmup_steps=200, rnn_bootstrap=True,
             double_q=True, clip_grad=40.0, adam_epsilon=1e-5,
             history_mode={"type": "replay", "args": {"size": 600, "train_frequency": 4, "max_envs": 4}})
    assert tr.steps >= 700 and tr.updates > 50
    st = tr.learner.stats()
    assert np.isfinite(st["qloss"]) and st["grad_norm"] > 0


@pytest.mark.gpu
def test_frame_prefetch_with_burn_in_is_bit_identical():
    """rt_learner_prefetch converts the training window's frames into the batch slot's private buffer on the
    replay stream; the burn-in passes of the same update convert THEIR frames into the default buffer.  With and
    without the prefetch the trained weights and the |td| signal must not differ by a bit."""
    import random
    import torch
    from rltime_b200.history import DevicePrioritizedReplayHistoryBuffer
    from rltime_b200.init import init_params
    from rltime_b200.learner import DeviceLearner
    E, T, P, n, B, A, U = 8, 6, 3, 2, 8, 4, 64
    dev = torch.device("cuda", 0)

    def run(prefetch):
        random.seed(5)
        rs = np.random.RandomState(3)
        hist = DevicePrioritizedReplayHistoryBuffer(
            size=4096, train_frequency=None, alpha=0.9, beta=0.6, nstep_target=n, nstep_train=T,
            prefix_steps=P, gamma=0.99, max_envs=E, device=dev)
        L = DeviceLearner((4, 84, 84), [(32, 8, 4), (64, 4, 2)], U, 64, A, 8, 16, True, mbatch=B,
                          nstep_train=T, burn_in=P, nstep_target=n, double_q=True, rnn_bootstrap=True,
                          clip_grad=40.0, seed=11, device=dev)
        L.load_state_dict(init_params(L.param_info, U, seed=1), 0)
        L.load_state_dict(init_params(L.param_info, U, seed=2), 1)
        m = 200 * E
        env = np.arange(m) % E
        hist.update_arrays(env, np.sign(rs.randn(m)), (rs.rand(m) < 0.02).astype(np.uint8),
                           [rs.randint(0, 255, (m, 4, 84, 84)).astype(np.uint8),
                            rs.randn(m, U).astype(np.float32), rs.randn(m, U).astype(np.float32),
                            np.zeros(m, np.float32)],
                           [rs.randint(0, A, m).astype(np.int64), rs.randn(m, A).astype(np.float32)])
        tds = []
        for it in range(12):
            assert hist.draw(B, 0.0) is not None
            if prefetch:
                L.prefetch(hist.last_batch, hist._stream())
            L.step(hist.last_batch)
            hist.update_losses_device(L.td_abs(), ready=L.wait_loss)
            tds.append(L.td_abs().clone())
        torch.cuda.synchronize()
        w = {k: v.clone() for k, v in L.state_dict(0).items()}
        tds = [t.cpu().numpy() for t in tds]
        L.close()
        hist.close()
        return w, tds
    w0, t0 = run(False)
    w1, t1 = run(True)
    for a, b in zip(t0, t1):
        np.testing.assert_array_equal(a, b)
    for k in w0:
        np.testing.assert_array_equal(w0[k].numpy(), w1[k].numpy(), err_msg=k)


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
