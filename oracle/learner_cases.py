"""Shared helpers to turn tests/golden/learner_*.npz into oracle / CUDA learner inputs
(TEST INFRASTRUCTURE)."""
import numpy as np
import torch

from oracle.learner_oracle import ModelSpec

# mirrors oracle/gen_golden_learner.py:CASES (kept in sync by test_learner_oracle_golden)
from oracle.gen_golden_learner import CASES  # noqa: F401  (pure-data dict; no reference import at module level)


def spec_of(c):
    return ModelSpec(c["in_shape"], c["conv"], c["lstm"], c["fc"], c["actions"], c["nq"],
                     c["embed"], c["dueling"], policy=c.get("policy", "iqn"),
                     pre_fc=c.get("pre_fc", ()), extra_dim=c.get("extra", 0))


def update_kwargs(c):
    """Optional trainer arguments of a case -> learner_oracle.learner_update keywords."""
    return dict(aggregation=c.get("loss_agg", "mean"), timestep_aggregation=c.get("loss_ts_agg"),
                loss_mode=c.get("loss_mode", "huber"), rnn_steps_train=c.get("rnn_steps"))


def params_of(g, prefix):
    return {k[len(prefix) + 1:]: torch.from_numpy(v.copy()) for k, v in g.items()
            if k.startswith(prefix + "/")}


def batch_of(g, c, u):
    S, n = c["T"] + c["P"], c["n"]
    b = {k.split("/")[-1]: v for k, v in g.items() if k.startswith("u%d/batch/" % u)}

    # states / target_states are VIEWS of one overlapped stack, exactly what the replay
    # buffer hands out (history.py:254-265): the reference's burn-in writes through them
    # (multi_step_trainer.py:117-126), which matters when prefix_steps == nstep_target.
    allt = {k: torch.from_numpy(b[k].copy()) for k in b if k.startswith("all_")}

    spec = spec_of(c)

    def st(lo, hi):
        s = {"x": (allt["all_x"][lo:hi], allt["all_extra"][lo:hi]) if c.get("extra") else allt["all_x"][lo:hi]}
        if c["lstm"]:
            s["layer%d_state" % spec.lstm_index] = {
                "hx": allt["all_hx"][lo:hi], "cx": allt["all_cx"][lo:hi],
                "initials": allt["all_initials"][lo:hi]}
        return s
    return {
        "states": st(0, S), "target_states": st(n, S + n),
        "returns": torch.from_numpy(b["returns"].copy()),
        "nsteps": torch.from_numpy(b["nsteps"].copy()),
        "target_masks": torch.from_numpy(b["target_masks"].copy()),
        "actions": torch.from_numpy(b["actions"].copy()),
        "importance_weights": torch.from_numpy(b["importance_weights"].copy()),
    }, b


def taus_of(g, u):
    return {k.split("/")[-1]: torch.from_numpy(v.copy()) for k, v in g.items()
            if k.startswith("u%d/tau/" % u)}
