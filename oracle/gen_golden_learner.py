"""Generates tests/golden/learner_<case>.npz by driving the UNMODIFIED reference trainer /
policy classes (opherlieber/rltime at /root/reference) on CPU, with the IQN quantile
fractions injected at the reference's torch.rand call (rltime/policies/torch/iqn.py:88).

TEST INFRASTRUCTURE.  Run in the build container:  python oracle/gen_golden_learner.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("RLTIME_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "stubs"), REF, ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle.learner_oracle import ModelSpec  # noqa: E402

CASES = {
    # nature-CNN -> LSTM -> FC family, shrunk; burn-in + double-Q + rnn_bootstrap + vf-rescale
    "iqn_lstm_small": dict(
        in_shape=(2, 16, 16), conv=[(8, 4, 2), (8, 3, 1)], lstm=16, fc=32, actions=3, nq=4,
        embed=8, dueling=True, B=3, T=4, P=2, n=2, gamma=0.99, double_q=True,
        rnn_bootstrap=True, vf_eps=1e-3, clip_grad=40.0, adam_eps=1e-5, updates=2),
    # no burn-in, no vf-rescale, clipping active, single-Q, Nq=8, non-square conv strides
    "iqn_lstm_clip": dict(
        in_shape=(4, 20, 20), conv=[(6, 8, 4), (8, 2, 1), (8, 2, 1)], lstm=24, fc=16,
        actions=5, nq=8, embed=16, dueling=True, B=4, T=5, P=0, n=3, gamma=0.997,
        double_q=False, rnn_bootstrap=True, vf_eps=None, clip_grad=0.05, adam_eps=1e-8,
        updates=2),
    # targets computed one time-step at a time from the stored LSTM states
    "iqn_lstm_nobootstrap": dict(
        in_shape=(1, 12, 12), conv=[(4, 4, 2)], lstm=8, fc=8, actions=2, nq=4, embed=4,
        dueling=False, B=2, T=3, P=0, n=1, gamma=0.9, double_q=True, rnn_bootstrap=False,
        vf_eps=None, clip_grad=None, adam_eps=1e-8, updates=1),
    # config-2 family: CNN -> FC, quantile layer injected after the CNN
    "iqn_cnn_fc": dict(
        in_shape=(4, 14, 14), conv=[(8, 4, 2), (4, 3, 1)], lstm=0, fc=24, actions=4, nq=8,
        embed=8, dueling=True, B=6, T=1, P=0, n=3, gamma=0.99, double_q=True,
        rnn_bootstrap=False, vf_eps=None, clip_grad=10.0, adam_eps=1.5e-4, updates=2),
    # config-4 family (Rainbow-style DQN): CNN -> FC, dueling + double-Q + n-step + PER weights,
    # plain DQNPolicy / DQN trainer (training/torch/dqn.py), huber loss
    "dqn_rainbow": dict(
        in_shape=(4, 14, 14), conv=[(8, 4, 2), (4, 3, 1)], lstm=0, fc=24, actions=4, nq=1,
        embed=0, dueling=True, B=6, T=1, P=0, n=3, gamma=0.99, double_q=True,
        rnn_bootstrap=False, vf_eps=None, clip_grad=10.0, adam_eps=1.5e-4, updates=2,
        policy="dqn"),
    # recurrent DQN: burn-in, single-Q, vf-rescale, mse loss, sum over the batch of per-sequence
    # means over time (loss_timestep_aggregation), dynamic (EMA) gradient clipping
    "dqn_lstm_mse": dict(
        in_shape=(2, 16, 16), conv=[(8, 4, 2), (8, 3, 1)], lstm=16, fc=32, actions=3, nq=1,
        embed=0, dueling=False, B=3, T=4, P=2, n=2, gamma=0.99, double_q=False,
        rnn_bootstrap=True, vf_eps=1e-3, clip_grad=0.8, adam_eps=1e-5, updates=3,
        policy="dqn", loss_mode="mse", loss_agg="sum", loss_ts_agg="mean", clip_dyn_alpha=0.9),
    # the shipped atari_iqn_lstm wrapper stack: tuple observation (frame, extra features = one-hot last action,
    # reward, timestep: env_wrappers/common.py:221-233) with the extra vector concatenated at the LSTM input
    "iqn_lstm_extra": dict(
        in_shape=(1, 16, 16), conv=[(8, 4, 2), (8, 3, 1)], lstm=16, fc=32, actions=3, nq=4,
        embed=8, dueling=True, B=3, T=4, P=2, n=2, gamma=0.99, double_q=True,
        rnn_bootstrap=True, vf_eps=None, clip_grad=40.0, adam_eps=1e-5, updates=2, extra=5),
    # config-1 family: MLP (configs/models/mlp_2x64.json shrunk) on a 1-D float observation, DQN
    "dqn_mlp": dict(
        in_shape=(4,), conv=[], pre_fc=[[16]], lstm=0, fc=16, actions=2, nq=1, embed=0,
        dueling=False, B=8, T=1, P=0, n=3, gamma=0.99, double_q=True, rnn_bootstrap=False,
        vf_eps=None, clip_grad=None, adam_eps=1e-8, updates=2, policy="dqn"),
    # CNN -> FC (fc_count 2) -> LSTM -> FC (configs/models/nature_cnn_fc512_lstm512_fc512.json family), IQN
    "iqn_cnn_fc_lstm_fc": dict(
        in_shape=(2, 16, 16), conv=[(8, 4, 2), (8, 3, 1)], pre_fc=[[12, 12], [20]], lstm=16, fc=24, actions=3,
        nq=4, embed=8, dueling=True, B=3, T=4, P=0, n=2, gamma=0.99, double_q=True,
        rnn_bootstrap=True, vf_eps=1e-3, clip_grad=40.0, adam_eps=1e-5, updates=2),
    # rnn_steps_train < nstep_train: the LSTM views the T*B rows as (rnn_steps, T*B/rnn_steps)
    # (multi_step_trainer.py:192-216,239,305; lstm.py:60-79)
    "iqn_lstm_rnnsteps": dict(
        in_shape=(2, 16, 16), conv=[(8, 4, 2), (8, 3, 1)], lstm=16, fc=32, actions=3, nq=4,
        embed=8, dueling=True, B=3, T=6, P=0, n=2, gamma=0.99, double_q=True,
        rnn_bootstrap=True, vf_eps=None, clip_grad=40.0, adam_eps=1e-5, updates=2, rnn_steps=2),
    # configs/ple_flappy_bird_iqn_lstm.json family: RGB frames without stacking, taller than wide (the shipped
    # warp is 120 x 80 x 3), plus the extra-features tuple of the atari_iqn_lstm wrapper stack
    "iqn_lstm_rgb_rect": dict(
        in_shape=(3, 30, 20), conv=[(8, 8, 4), (8, 3, 1), (8, 2, 1)], lstm=16, fc=32, actions=2, nq=4,
        embed=8, dueling=True, B=3, T=4, P=1, n=2, gamma=0.99, double_q=True,
        rnn_bootstrap=True, vf_eps=1e-3, clip_grad=40.0, adam_eps=1e-5, updates=2, extra=4),
    # IQN with mean over time then sum over the batch
    "iqn_lstm_tsagg": dict(
        in_shape=(1, 12, 12), conv=[(4, 4, 2)], lstm=8, fc=8, actions=2, nq=4, embed=4,
        dueling=True, B=2, T=3, P=0, n=1, gamma=0.9, double_q=True, rnn_bootstrap=True,
        vf_eps=None, clip_grad=None, adam_eps=1e-8, updates=1, loss_agg="sum", loss_ts_agg="mean"),
}


def make_spec(c):
    return ModelSpec(c["in_shape"], c["conv"], c["lstm"], c["fc"], c["actions"], c["nq"],
                     c["embed"], c["dueling"], policy=c.get("policy", "iqn"),
                     pre_fc=c.get("pre_fc", ()), extra_dim=c.get("extra", 0))


def make_batch(c, seed):
    """(S+n, B) overlapped state stack + per-row scalars, like History._make_train_batch."""
    rs = np.random.RandomState(seed)
    S, B, n = c["T"] + c["P"], c["B"], c["n"]
    U = max(c["lstm"], 1)
    b = {
        "all_x": rs.randint(0, 256, (S + n, B) + tuple(c["in_shape"])).astype(np.uint8) if c["conv"]
        else rs.randn(S + n, B, *c["in_shape"]).astype(np.float32),
        "returns": rs.randn(S, B),
        "nsteps": np.full((S, B), n, dtype=np.int64),
        "target_masks": (rs.rand(S, B) > 0.2).astype(np.float64),
        "actions": rs.randint(0, c["actions"], (S, B)).astype(np.int64),
        "importance_weights": rs.rand(S, B) * 0.9 + 0.1,
    }
    if c["lstm"]:
        b["all_hx"] = rs.randn(S + n, B, U).astype(np.float32)
        b["all_cx"] = rs.randn(S + n, B, U).astype(np.float32)
        b["all_initials"] = (rs.rand(S + n, B) < 0.15).astype(np.float32)
    if c.get("extra"):
        b["all_extra"] = rs.randn(S + n, B, c["extra"]).astype(np.float32)
    return b


def model_config(c):
    layers = []
    if c["conv"]:
        layers.append({"type": "cnn", "args": {"layers": [
            {"filters": f, "kernel": k, "stride": s} for f, k, s in c["conv"]]}})
    for sizes in c.get("pre_fc", ()):
        # an FC module has ONE fc_size for all its fc_count layers (fc.py:18-24)
        assert len(set(sizes)) == 1
        layers.append({"type": "fc", "args": {"fc_size": sizes[0], "fc_count": len(sizes)}})
    if c["lstm"]:
        layers.append({"type": "lstm", "args": {"num_units": c["lstm"]}})
    layers.append({"type": "fc", "args": {"fc_size": c["fc"]}})
    return {"type": "sequential", "args": {"layer_configs": layers}}


class TauQueue:
    """Replaces torch.rand while the reference runs; hands out pre-drawn tau vectors."""

    def __init__(self, gen):
        self.gen = gen
        self.log = []
        self._orig = torch.rand

    def __call__(self, *size, device=None, **kw):
        assert len(size) == 1 and not kw
        t = self._orig(size[0], generator=self.gen)
        self.log.append(t.clone())
        return t

    def __enter__(self):
        torch.rand = self
        return self

    def __exit__(self, *a):
        torch.rand = self._orig


def run_case(name, c):
    import gym
    from rltime.training.torch.iqn import IQN
    from rltime.training.torch.dqn import DQN
    from rltime.general.utils import deep_apply
    from rltime.general.value_log import ValueLog

    spec = make_spec(c)
    p_online = spec.init_params(seed=11)
    p_target = spec.init_params(seed=12)
    if c["conv"]:
        obs_space = gym.spaces.Box(0, 255, c["in_shape"], dtype=np.uint8)
    else:
        obs_space = gym.spaces.Box(-10, 10, c["in_shape"], dtype=np.float32)
    if c.get("extra"):
        obs_space = gym.spaces.Tuple((obs_space, gym.spaces.Box(-10, 10, (c["extra"],), dtype=np.float32)))
    act_space = gym.spaces.Discrete(c["actions"])
    dqn = c.get("policy", "iqn") == "dqn"
    if dqn:
        policy_args = dict(dueling=c["dueling"], cuda=False)
        tr = DQN(logger=None, actors=None, model_config=model_config(c), policy_args=policy_args)
    else:
        policy_args = dict(dueling=c["dueling"], num_sampling_quantiles=c["nq"],
                           embedding_dim=c["embed"], cuda=False)
        tr = IQN(logger=None, actors=None, model_config=model_config(c), policy_args=policy_args)

    def make_policy(params):
        pol = tr.create_policy(model_config=model_config(c), observation_space=obs_space,
                               action_space=act_space, **policy_args)
        sd = pol.state_dict()
        extra = set() if dqn else {"embedding_range"}
        assert set(sd.keys()) == set(params.keys()) | extra, (sorted(sd.keys()), sorted(params.keys()))
        for k, v in params.items():
            assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
        pol.load_state_dict({**{k: v.clone() for k, v in params.items()},
                             **{k: sd[k] for k in extra}})
        return pol
    tr.policy = make_policy(p_online)
    tr.target_policy = make_policy(p_target)
    # trainer attributes normally set by train()/_train() (policy_trainer.py:284-311,
    # dqn.py:40-47, torch_trainer.py:32-42, multi_step_trainer.py:217-219)
    tr.gamma = c["gamma"]
    tr.double_q = c["double_q"]
    tr.loss_mode, tr.huber_kappa = c.get("loss_mode", "huber"), 1.0
    tr.loss_aggregation = tr._get_aggregator(c.get("loss_agg", "mean"))
    tr.loss_timestep_aggregation = tr._get_aggregator(c["loss_ts_agg"]) if c.get("loss_ts_agg") else None
    tr.clip_grad = c["clip_grad"]
    tr.clip_grad_dynamic_alpha = c.get("clip_dyn_alpha")
    tr._grad_norm_moving_average = None
    tr.adam_epsilon = c["adam_eps"]
    tr.vf_scale_epsilon = c["vf_eps"]
    tr.clip_rewards = False
    tr.value_log = ValueLog()
    tr.ts_steps_trained = 0
    reported = []

    class Hist:
        def update_losses(self, idx, losses):
            reported.append(np.array(losses, copy=True))
    tr.history_buffer = Hist()
    tr.train_init(None)

    S, B, n, P, T = c["T"] + c["P"], c["B"], c["n"], c["P"], c["T"]
    out = {"online/" + k: v.numpy() for k, v in p_online.items()}
    out.update({"target/" + k: v.numpy() for k, v in p_target.items()})
    gen = torch.Generator().manual_seed(5)
    for u in range(c["updates"]):
        b = make_batch(c, seed=100 + u)
        for k, v in b.items():
            out["u%d/batch/%s" % (u, k)] = v
        # the replay buffer hands states/target_states as views of one stacked tensor
        # (history.py:254-265 + general/backend.py:143-147)
        x0 = torch.from_numpy(b["all_x"].copy())
        all_states = {"x": (x0, torch.from_numpy(b["all_extra"].copy())) if c.get("extra") else x0}
        for li in range(spec.fc_layer_index + 1):
            all_states["layer%d_state" % li] = {}
        if c["lstm"]:
            all_states["layer%d_state" % spec.lstm_index] = {
                "hx": torch.from_numpy(b["all_hx"].copy()),
                "cx": torch.from_numpy(b["all_cx"].copy()),
                "initials": torch.from_numpy(b["all_initials"].copy())}
        train_data = {
            "states": deep_apply(all_states, lambda x: x[:S]),
            "target_states": deep_apply(all_states, lambda x: x[n:]),
            "returns": b["returns"].copy(), "nsteps": b["nsteps"].copy(),
            "target_masks": b["target_masks"].copy(),
            "policy_outputs": {"actions": b["actions"].copy()},
            "extra_data": {"importance_weights": b["importance_weights"].copy(),
                           "loss_indices": np.zeros((S, B, 2), dtype=np.int64)},
        }
        with TauQueue(gen) as tq:
            # == MultiStepTrainer._train loop body (multi_step_trainer.py:278-340) ==
            if P:
                train_data = tr._burn_in(train_data, P, do_target_states=c["rnn_bootstrap"])
            train_data = deep_apply(
                train_data, lambda x: x.reshape((x.shape[0] * x.shape[1],) + x.shape[2:]))
            targets = tr.calc_target_values(
                train_data["returns"], train_data["target_states"], train_data["target_masks"],
                nsteps=train_data["nsteps"], timesteps=1 if not c["rnn_bootstrap"] else c.get("rnn_steps", T))
            params_before = {k: v.detach().clone() for k, v in tr.policy.state_dict().items()}
            tr.train_batch(train_data["states"], targets, train_data["policy_outputs"],
                           train_data["extra_data"], c.get("rnn_steps", T))
        taus = tq.log
        names = (["burn_online"] + (["burn_target"] if c["rnn_bootstrap"] else []) if P else []) + \
            ["target", "select", "train"]
        if dqn:
            names = []
        assert len(taus) == len(names), (len(taus), names)
        for nm, t in zip(names, taus):
            out["u%d/tau/%s" % (u, nm)] = t.numpy()
        out["u%d/targets" % u] = targets.numpy()
        out["u%d/report" % u] = reported[-1]
        vals = tr.value_log.get()["train"]
        out["u%d/qloss" % u] = np.float64(vals["qloss"])
        out["u%d/td_mean" % u] = np.float64(vals["qvalue" if dqn else "td_mean"])
        out["u%d/grad_norm" % u] = np.float64(vals["grad_norm"])
        for k, v in tr.policy.named_parameters():
            out["u%d/grad/%s" % (u, k)] = v.grad.detach().numpy().copy()   # post-clip grads
        for k, v in tr.policy.state_dict().items():
            if k != "embedding_range":
                out["u%d/after/%s" % (u, k)] = v.detach().numpy().copy()
        del params_before
    path = os.path.join(ROOT, "tests", "golden", "learner_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-22s fields=%d size=%.1f KB qloss=%s" % (
        name, len(out), os.path.getsize(path) / 1024,
        [float(out["u%d/qloss" % u]) for u in range(c["updates"])]))


if __name__ == "__main__":
    torch.manual_seed(0)
    only = sys.argv[1:]
    for name, c in CASES.items():
        if not only or name in only:
            run_case(name, c)
