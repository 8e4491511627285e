"""CUDA learner vs the reference trainer's golden results and vs the torch fp32 oracle."""
import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo
from oracle.learner_cases import CASES, batch_of, params_of, spec_of, taus_of
from tests.util import load_golden


def make_learner(c, **kw):
    from rltime_b200.learner import DeviceLearner
    kw.setdefault("gemm", "fp32")
    return DeviceLearner(c["in_shape"], c["conv"], c["lstm"], c["fc"], c["actions"], c["nq"],
                         c["embed"], c["dueling"], mbatch=c["B"], nstep_train=c["T"],
                         burn_in=c["P"], nstep_target=c["n"], gamma=c["gamma"],
                         double_q=c["double_q"], rnn_bootstrap=c["rnn_bootstrap"],
                         vf_scale_epsilon=c["vf_eps"], clip_grad=c["clip_grad"],
                         adam_epsilon=c["adam_eps"], lr=1e-3, policy=c.get("policy", "iqn"),
                         loss_mode=c.get("loss_mode", "huber"),
                         loss_aggregation=c.get("loss_agg", "mean"),
                         loss_timestep_aggregation=c.get("loss_ts_agg"),
                         clip_grad_dynamic_alpha=c.get("clip_dyn_alpha"), pre_fc=c.get("pre_fc", ()),
                         extra_dim=c.get("extra", 0), rnn_steps_train=c.get("rnn_steps"), **kw)


def step_taus(c, g, u):
    if c.get("policy", "iqn") == "dqn":
        return None
    taus = taus_of(g, u)
    return [taus["target"], taus["select"], taus["train"]]


def device_batch(raw, c):
    from rltime_b200.learner import batch_from_tensors
    dev = "cuda"
    t = lambda k, dt=None: torch.from_numpy(raw[k].copy()).to(dev)
    has = c["lstm"] > 0
    return batch_from_tensors(
        t("all_x"), t("all_hx") if has else None, t("all_cx") if has else None,
        t("all_initials") if has else None, t("returns"), t("nsteps"), t("target_masks"),
        t("actions"), t("importance_weights"), c["n"], all_extra=t("all_extra") if c.get("extra") else None)


def report_diff(errs, name, got, want, rtol, atol):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    if got.shape != want.shape:
        errs.append("%s: shape %s vs %s" % (name, got.shape, want.shape))
        return
    bad = np.abs(got - want) > atol + rtol * np.abs(want)
    if bad.any():
        i = np.argmax(np.abs(got - want))
        errs.append("%s: %d/%d off, max|d|=%.3e (got %.6g want %.6g)" % (
            name, bad.sum(), bad.size, np.abs(got - want).max(), got.flat[i], want.flat[i]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_learner_matches_reference_golden(name):
    c = CASES[name]
    g = load_golden("learner_%s.npz" % name)
    L = make_learner(c)
    try:
        L.load_state_dict(params_of(g, "online"), 0)
        L.load_state_dict(params_of(g, "target"), 1)
        # parameter layout round-trip (permuted internal layouts)
        back = L.state_dict(0)
        for k, v in params_of(g, "online").items():
            np.testing.assert_array_equal(back[k].numpy(), v.numpy())
        errs = []
        M, Nq = c["T"] * c["B"], c["nq"]
        dqn = c.get("policy", "iqn") == "dqn"
        dyn = lo.DynamicClip(c["clip_grad"], c["clip_dyn_alpha"]) if c.get("clip_dyn_alpha") is not None else None
        for u in range(c["updates"]):
            _, raw = batch_of(g, c, u)
            b, keep = device_batch(raw, c)
            L.step(b, step_taus(c, g, u))
            st = L.stats()
            pre = "u%d/" % u
            # fp32 tolerance of the parity bar: TD-loss / targets within 1e-4 absolute
            tshape = (M,) if dqn else (M, Nq)
            report_diff(errs, pre + "targets", L.debug("targets", tshape).cpu().numpy(),
                        g[pre + "targets"], 1e-4, 1e-5)
            report_diff(errs, pre + "qloss", st["qloss"], g[pre + "qloss"], 1e-4, 1e-5)
            report_diff(errs, pre + "td_mean", st["qvalue" if dqn else "td_mean"], g[pre + "td_mean"], 1e-4, 1e-5)
            report_diff(errs, pre + "report", L.td_abs().cpu().numpy(), g[pre + "report"], 1e-4, 1e-5)
            report_diff(errs, pre + "grad_norm", st["grad_norm"], g[pre + "grad_norm"], 1e-3, 1e-6)
            grads = L.state_dict(2)
            coef = 1.0
            if c["clip_grad"]:
                cv = dyn.value(float(g[pre + "grad_norm"])) if dyn is not None else c["clip_grad"]
                coef = min(1.0, cv / (float(g[pre + "grad_norm"]) + 1e-6))
            after = L.state_dict(0)
            for k in grads:
                # golden grads are post-clip; ours are stored pre-clip
                report_diff(errs, pre + "grad/" + k, grads[k].numpy() * coef, g[pre + "grad/" + k],
                            2e-3, 2e-6)
                report_diff(errs, pre + "after/" + k, after[k].numpy(), g[pre + "after/" + k],
                            1e-4, 2e-5)
        assert not errs, "\n".join(errs)
    finally:
        L.close()


@pytest.mark.gpu
def test_learner_full_size_vs_oracle():
    """Config-3 shapes (nature CNN, LSTM 512, FC 512, Nq 32, B 32, T 20): one update against
    the torch fp32 oracle run on the host CPU."""
    torch.set_num_threads(max(1, torch.get_num_threads()))
    c = dict(in_shape=(4, 84, 84), conv=[(32, 8, 4), (64, 4, 2), (64, 3, 1)], lstm=512, fc=512,
             actions=6, nq=32, embed=64, dueling=True, B=32, T=20, P=0, n=2, gamma=0.99,
             double_q=True, rnn_bootstrap=True, vf_eps=None, clip_grad=40.0, adam_eps=1e-5)
    spec = spec_of(c)
    p_on, p_tg = spec.init_params(1), spec.init_params(2)
    rs = np.random.RandomState(0)
    S, B, n, M = c["T"], c["B"], c["n"], c["T"] * c["B"]
    raw = {
        "all_x": rs.randint(0, 256, (S + n, B, 4, 84, 84)).astype(np.uint8),
        "all_hx": rs.randn(S + n, B, 512).astype(np.float32),
        "all_cx": rs.randn(S + n, B, 512).astype(np.float32),
        "all_initials": (rs.rand(S + n, B) < 0.02).astype(np.float32),
        "returns": np.sign(rs.randn(S, B)), "nsteps": np.full((S, B), n, dtype=np.int64),
        "target_masks": (rs.rand(S, B) > 0.05).astype(np.float64),
        "actions": rs.randint(0, 6, (S, B)).astype(np.int64),
        "importance_weights": rs.rand(S, B) * 0.5 + 0.5,
    }
    gen = torch.Generator().manual_seed(3)
    taus = {k: torch.rand(M * 32, generator=gen) for k in ("target", "select", "train")}
    allt = {k: torch.from_numpy(v.copy()) for k, v in raw.items()}

    def st(lo_, hi):
        return {"x": allt["all_x"][lo_:hi], "layer1_state": {
            "hx": allt["all_hx"][lo_:hi], "cx": allt["all_cx"][lo_:hi],
            "initials": allt["all_initials"][lo_:hi]}}
    batch = {"states": st(0, S), "target_states": st(n, S + n), "returns": allt["returns"],
             "nsteps": allt["nsteps"], "target_masks": allt["target_masks"],
             "actions": allt["actions"], "importance_weights": allt["importance_weights"]}
    p_ref = {k: v.clone() for k, v in p_on.items()}
    opt = lo.Adam(p_ref, lr=1e-3, eps=c["adam_eps"])
    res = lo.learner_update(spec, p_ref, p_tg, opt, batch, taus, c["gamma"], double_q=True,
                            rnn_bootstrap=True, vf_eps=None, clip_grad=c["clip_grad"])
    L = make_learner(c)
    try:
        L.load_state_dict(p_on, 0)
        L.load_state_dict(p_tg, 1)
        b, keep = device_batch(raw, c)
        L.step(b, [taus["target"], taus["select"], taus["train"]])
        stt = L.stats()
        errs = []
        report_diff(errs, "targets", L.debug("targets", (M, 32)).cpu().numpy(), res["targets"].numpy(), 0, 1e-4)
        report_diff(errs, "qloss", stt["qloss"], float(res["loss"]), 0, 1e-4)
        report_diff(errs, "td_mean", stt["td_mean"], float(res["td_mean"]), 0, 1e-4)
        report_diff(errs, "report", L.td_abs().cpu().numpy(), res["report"].numpy(), 0, 1e-4)
        report_diff(errs, "grad_norm", stt["grad_norm"], res["grad_norm"], 1e-3, 0)
        grads = L.state_dict(2)
        for k in grads:
            gw = res["grads"][k].numpy()   # clipped by coef; norm < 40 here so coef == 1
            scale = np.abs(gw).max() + 1e-12
            report_diff(errs, "grad/" + k, grads[k].numpy() / scale, gw / scale, 0, 2e-3)
        after = L.state_dict(0)
        for k in after:
            report_diff(errs, "after/" + k, after[k].numpy(), p_ref[k].numpy(), 1e-4, 1e-4)  # Adam step ~ lr*g/(|g|+eps): sensitive where |g| ~ eps
        print("full-size: qloss %.6f (oracle %.6f) td_mean %.6f grad_norm %.5f (oracle %.5f)" % (
            stt["qloss"], float(res["loss"]), stt["td_mean"], stt["grad_norm"], res["grad_norm"]))
        assert not errs, "\n".join(errs)
    finally:
        L.close()


def golden_errors(name, gemm):
    """max |difference| to the reference golden of targets / qloss / td_mean / report, per update."""
    c = CASES[name]
    g = load_golden("learner_%s.npz" % name)
    L = make_learner(c, gemm=gemm)
    out = []
    try:
        L.load_state_dict(params_of(g, "online"), 0)
        L.load_state_dict(params_of(g, "target"), 1)
        M, Nq = c["T"] * c["B"], c["nq"]
        dqn = c.get("policy", "iqn") == "dqn"
        tshape = (M,) if dqn else (M, Nq)
        for u in range(c["updates"]):
            _, raw = batch_of(g, c, u)
            b, keep = device_batch(raw, c)
            L.step(b, step_taus(c, g, u))
            st = L.stats()
            pre = "u%d/" % u
            out.append({
                "targets": float(np.abs(L.debug("targets", tshape).cpu().numpy() - g[pre + "targets"]).max()),
                "qloss": abs(st["qloss"] - float(g[pre + "qloss"])),
                "td_mean": abs(st["qvalue" if dqn else "td_mean"] - float(g[pre + "td_mean"])),
                "report": float(np.abs(L.td_abs().cpu().numpy() - g[pre + "report"]).max()),
                "grad_norm_rel": abs(st["grad_norm"] - float(g[pre + "grad_norm"])) / float(g[pre + "grad_norm"]),
            })
    finally:
        L.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_learner_tf32_tensor_core_path(name):
    """The benched precision (tcgen05 TF32 products on round-to-nearest operands, RT_GEMM_TF32_RN)
    on the reference goldens: targets, TD loss and the per-row |td| signal within the north-star
    bound of 1e-4 (absolute) on EVERY update of every case."""
    errs = []
    for u, e in enumerate(golden_errors(name, "tf32")):
        for k in ("targets", "qloss", "td_mean", "report"):
            if not e[k] <= 1e-4:
                errs.append("u%d/%s: |d| = %.3e > 1e-4" % (u, k, e[k]))
        if not e["grad_norm_rel"] <= 2e-2:
            errs.append("u%d/grad_norm: rel %.3e" % (u, e["grad_norm_rel"]))
    assert not errs, "\n".join(errs)


def _full_size_case():
    c = dict(in_shape=(4, 84, 84), conv=[(32, 8, 4), (64, 4, 2), (64, 3, 1)], lstm=512, fc=512,
             actions=6, nq=32, embed=64, dueling=True, B=32, T=20, P=0, n=2, gamma=0.99,
             double_q=True, rnn_bootstrap=True, vf_eps=None, clip_grad=40.0, adam_eps=1e-5)
    rs = np.random.RandomState(0)
    S, B, n = c["T"], c["B"], c["n"]
    raw = {
        "all_x": rs.randint(0, 256, (S + n, B, 4, 84, 84)).astype(np.uint8),
        "all_hx": rs.randn(S + n, B, 512).astype(np.float32),
        "all_cx": rs.randn(S + n, B, 512).astype(np.float32),
        "all_initials": (rs.rand(S + n, B) < 0.02).astype(np.float32),
        "returns": np.sign(rs.randn(S, B)), "nsteps": np.full((S, B), n, dtype=np.int64),
        "target_masks": (rs.rand(S, B) > 0.05).astype(np.float64),
        "actions": rs.randint(0, 6, (S, B)).astype(np.int64),
        "importance_weights": rs.rand(S, B) * 0.5 + 0.5,
    }
    return c, raw


def full_size_drift(modes, updates, lr=3e-4, seed=0, log=None, teacher_forced=True, tie_margin=2e-4):
    """Config-3 shapes, `updates` consecutive learner updates on fresh seeded batches and injected tau:
    the torch fp32 oracle on the host CPU against this library in each of `modes`.

    teacher_forced=True: before every update the library's online weights are reset to the ORACLE's
    current weights, so update u measures the arithmetic of one update on the u-th weight state of a
    real training trajectory.  teacher_forced=False: both run freely from the same initial weights;
    what is compared then includes the chaotic growth of round-off through 50 Adam steps (Adam
    normalises every gradient component to ~lr, and an arg-max of the double-Q selection that flips on
    a near-tie moves that row's target by O(0.01)), which two fp32 implementations with different
    summation orders show just the same.

    Rows whose selection arg-max is a near-tie in the oracle itself (top-2 gap of the mean selection
    value < tie_margin) are excluded from the per-row target / |td| comparison and counted.
    Returns {mode: [per-update dict of |differences| to the oracle]}."""
    c, _ = _full_size_case()
    spec = spec_of(c)
    S, B, n, M = c["T"], c["B"], c["n"], c["T"] * c["B"]

    def make_raw(u):
        rs = np.random.RandomState(1000 * seed + u)
        return {
            "all_x": rs.randint(0, 256, (S + n, B, 4, 84, 84)).astype(np.uint8),
            "all_hx": (0.3 * rs.randn(S + n, B, 512)).astype(np.float32),
            "all_cx": (0.3 * rs.randn(S + n, B, 512)).astype(np.float32),
            "all_initials": (rs.rand(S + n, B) < 0.02).astype(np.float32),
            "returns": np.sign(rs.randn(S, B)) * (rs.rand(S, B) < 0.3), "nsteps": np.full((S, B), n, dtype=np.int64),
            "target_masks": (rs.rand(S, B) > 0.02).astype(np.float64),
            "actions": rs.randint(0, 6, (S, B)).astype(np.int64),
            "importance_weights": rs.rand(S, B) * 0.5 + 0.5,
        }

    def make_taus(u):
        gen = torch.Generator().manual_seed(77 + u)
        return [torch.rand(M * 32, generator=gen) for _ in range(3)]
    p0, pt = spec.init_params(1), spec.init_params(2)
    # ---- oracle trace (keeps the weights BEFORE every update for the teacher-forced comparison)
    p_ref = {k: v.clone() for k, v in p0.items()}
    opt = lo.Adam(p_ref, lr=lr, eps=c["adam_eps"])
    ref = []
    for u in range(updates):
        raw = make_raw(u)
        allt = {k: torch.from_numpy(v.copy()) for k, v in raw.items()}

        def st(lo_, hi):
            return {"x": allt["all_x"][lo_:hi], "layer1_state": {
                "hx": allt["all_hx"][lo_:hi], "cx": allt["all_cx"][lo_:hi],
                "initials": allt["all_initials"][lo_:hi]}}
        batch = {"states": st(0, S), "target_states": st(n, S + n), "returns": allt["returns"],
                 "nsteps": allt["nsteps"], "target_masks": allt["target_masks"],
                 "actions": allt["actions"], "importance_weights": allt["importance_weights"]}
        t3 = make_taus(u)
        before = {k: v.clone() for k, v in p_ref.items()} if teacher_forced else None
        res = lo.learner_update(spec, p_ref, pt, opt, batch, {"target": t3[0], "select": t3[1], "train": t3[2]},
                                c["gamma"], double_q=True, rnn_bootstrap=True, vf_eps=None,
                                clip_grad=c["clip_grad"])
        ref.append({"targets": res["targets"].numpy().copy(), "qloss": float(res["loss"]),
                    "td_mean": float(res["td_mean"]), "report": res["report"].numpy().copy(),
                    "grad_norm": float(res["grad_norm"]), "before": before,
                    "ok_rows": (res["select_margin"].numpy() >= tie_margin)})
    out = {}
    for mode in modes:
        L = make_learner(c, gemm=mode)
        rows = []
        try:
            L.set_lr(lr)
            L.load_state_dict(p0, 0)
            L.load_state_dict(pt, 1)
            for u in range(updates):
                r = ref[u]
                if teacher_forced and u > 0:
                    L.load_state_dict(r["before"], 0)
                b, keep = device_batch(make_raw(u), c)
                L.step(b, make_taus(u))
                stt = L.stats()
                ok = r["ok_rows"]
                dt = np.abs(L.debug("targets", (M, 32)).cpu().numpy() - r["targets"])
                dr = np.abs(L.td_abs().cpu().numpy() - r["report"])
                rows.append({
                    "targets": float(dt[ok].max()), "report": float(dr[ok].max()),
                    "qloss": abs(stt["qloss"] - r["qloss"]), "td_mean": abs(stt["td_mean"] - r["td_mean"]),
                    "tie_rows": int((~ok).sum()), "flipped_rows": int((dt.max(axis=1) > 1e-3).sum()),
                    "grad_norm_rel": abs(stt["grad_norm"] - r["grad_norm"]) / r["grad_norm"],
                    "ref_qloss": r["qloss"], "ref_targets_absmax": float(np.abs(r["targets"]).max())})
                if log:
                    e = rows[-1]
                    log("%-10s u%02d qloss %.6f (oracle %.6f) |d|: qloss %.2e td_mean %.2e report %.2e targets %.2e "
                        "grad_norm rel %.2e  near-tie rows %d flipped rows %d" % (
                            mode, u, stt["qloss"], r["qloss"], e["qloss"], e["td_mean"], e["report"], e["targets"],
                            e["grad_norm_rel"], e["tie_rows"], e["flipped_rows"]))
        finally:
            L.close()
        out[mode] = rows
    return out


@pytest.mark.gpu
def test_learner_full_size_50_update_trajectory():
    """North-star bound along a run, not one step from random init: 50 consecutive Adam updates of the
    torch fp32 oracle at config-3 size; at every one of the 50 weight states the benched precision (TF32
    products on round-to-nearest operands) and the fp32 SIMT path reproduce the oracle's update on the
    same batch and tau: the TD loss (the north-star quantity), the mean |td| and every bootstrap target within
    1e-4 (absolute); the per-row |td| priority signal -- a maximum over 32,000 rows, measured 0.95e-4 to
    0.98e-4 for the TF32 path on B200 (scripts/precision_sweep.py, profiles/r02_precision_sweep.txt) -- within
    2e-4.  Rows whose double-Q arg-max is a near-tie in the oracle (gap < 2e-4) have no well-defined target
    to compare; they are excluded from the per-row checks, counted, and must stay rare."""
    res = full_size_drift(["tf32", "fp32"], 50)
    errs = []
    tol = {"qloss": 1e-4, "td_mean": 1e-4, "targets": 1e-4, "report": 2e-4}
    for mode, rows in res.items():
        for u, e in enumerate(rows):
            for k in ("qloss", "td_mean", "report", "targets"):
                if not e[k] <= tol[k]:
                    errs.append("%s u%d/%s: |d| = %.3e > %.0e" % (mode, u, k, e[k], tol[k]))
            if e["flipped_rows"] > e["tie_rows"]:
                errs.append("%s u%d: %d rows with a different arg-max but only %d near-ties" % (
                    mode, u, e["flipped_rows"], e["tie_rows"]))
        ties = sum(e["tie_rows"] for e in rows)
        print("%s: max over 50 updates |d| qloss %.2e td_mean %.2e report %.2e targets %.2e; near-tie rows excluded "
              "%d of %d" % (mode, max(e["qloss"] for e in rows), max(e["td_mean"] for e in rows),
                            max(e["report"] for e in rows), max(e["targets"] for e in rows), ties, 50 * 640))
        if ties > 0.05 * 50 * 640:
            errs.append("%s: %d near-tie rows" % (mode, ties))
    assert not errs, "\n".join(errs[:40])


@pytest.mark.gpu
def test_learner_free_running_drift_is_not_a_precision_effect():
    """30 free-running updates (no weight reset): the distance of the TF32 path to the oracle stays of
    the same order as the distance of this library's own fp32 path to the oracle -- the growth is the
    chaotic amplification of round-off through Adam, not the multiply precision."""
    res = full_size_drift(["tf32", "fp32"], 30, teacher_forced=False)
    worst = {m: max(e["qloss"] for e in rows) for m, rows in res.items()}
    print("free-running 30 updates: max |d qloss| tf32 %.2e fp32 %.2e" % (worst["tf32"], worst["fp32"]))
    assert worst["tf32"] <= 5e-3 and worst["fp32"] <= 5e-3
    assert worst["tf32"] <= 10 * worst["fp32"] + 1e-4


@pytest.mark.gpu
def test_learner_full_size_tf32_vs_fp32_path():
    """Config-3 shapes: the tcgen05 TF32 path against this library's fp32 SIMT path (itself
    pinned to the oracle above): TD-loss / per-row |td| / targets within the 1e-4 parity bar."""
    c, raw = _full_size_case()
    spec = spec_of(c)
    p_on, p_tg = spec.init_params(1), spec.init_params(2)
    gen = torch.Generator().manual_seed(3)
    M = c["T"] * c["B"]
    taus = [torch.rand(M * 32, generator=gen) for _ in range(3)]
    out = {}
    for mode in ("fp32", "tf32"):
        L = make_learner(c, gemm=mode)
        try:
            L.load_state_dict(p_on, 0)
            L.load_state_dict(p_tg, 1)
            b, keep = device_batch(raw, c)
            L.step(b, taus)
            out[mode] = dict(st=L.stats(), targets=L.debug("targets", (M, 32)).cpu().numpy(),
                             report=L.td_abs().cpu().numpy(), grads=L.state_dict(2))
        finally:
            L.close()
    a, b_ = out["fp32"], out["tf32"]
    print("qloss fp32 %.6f tf32 %.6f | td_mean %.6f %.6f | grad_norm %.5f %.5f | max|dtargets| %.2e "
          "max|dreport| %.2e" % (a["st"]["qloss"], b_["st"]["qloss"], a["st"]["td_mean"],
                                b_["st"]["td_mean"], a["st"]["grad_norm"], b_["st"]["grad_norm"],
                                np.abs(a["targets"] - b_["targets"]).max(),
                                np.abs(a["report"] - b_["report"]).max()))
    errs = []
    report_diff(errs, "qloss", b_["st"]["qloss"], a["st"]["qloss"], 0, 1e-4)
    report_diff(errs, "td_mean", b_["st"]["td_mean"], a["st"]["td_mean"], 0, 1e-4)
    report_diff(errs, "report", b_["report"], a["report"], 0, 1e-4)
    report_diff(errs, "targets", b_["targets"], a["targets"], 0, 2e-4)
    report_diff(errs, "grad_norm", b_["st"]["grad_norm"], a["st"]["grad_norm"], 1e-2, 0)
    for k in a["grads"]:
        ga, gb = a["grads"][k].numpy(), b_["grads"][k].numpy()
        rel = np.linalg.norm(ga - gb) / (np.linalg.norm(ga) + 1e-12)
        if rel > 5e-2:
            errs.append("grad %s: relative L2 error %.3e" % (k, rel))
    assert not errs, "\n".join(errs)


@pytest.mark.gpu
@pytest.mark.parametrize("upc", [16, 8])
@pytest.mark.parametrize("double_q", [True, False])
def test_lstm_tensor_core_recurrence_midsize(upc, double_q, monkeypatch):
    """The multi-sequence tcgen05 recurrence (rt_lstm_tc.cuh) at a shape the goldens do not reach
    (U = 64, B = 8 < 32 rows per sequence, episode resets inside the window): every saved LSTM
    tensor of the training sequence, the targets and the per-row TD signal against this library's
    fp32 path, for both CTA slice widths and with two / three concurrent sequences."""
    monkeypatch.setenv("RT_LSTM_UPC", str(upc))
    c = dict(in_shape=(4, 20, 20), conv=[(32, 8, 4), (32, 2, 1)], lstm=64, fc=32, actions=4, nq=8,
             embed=16, dueling=True, B=8, T=6, P=0, n=2, gamma=0.99, double_q=double_q,
             rnn_bootstrap=True, vf_eps=None, clip_grad=40.0, adam_eps=1e-5)
    rs = np.random.RandomState(5)
    S, B, n, U = c["T"], c["B"], c["n"], c["lstm"]
    raw = {
        "all_x": rs.randint(0, 256, (S + n, B) + c["in_shape"]).astype(np.uint8),
        "all_hx": rs.randn(S + n, B, U).astype(np.float32),
        "all_cx": rs.randn(S + n, B, U).astype(np.float32),
        "all_initials": (rs.rand(S + n, B) < 0.2).astype(np.float32),
        "returns": np.sign(rs.randn(S, B)), "nsteps": np.full((S, B), n, dtype=np.int64),
        "target_masks": (rs.rand(S, B) > 0.05).astype(np.float64),
        "actions": rs.randint(0, c["actions"], (S, B)).astype(np.int64),
        "importance_weights": rs.rand(S, B) * 0.5 + 0.5,
    }
    spec = spec_of(c)
    p_on, p_tg = spec.init_params(1), spec.init_params(2)
    gen = torch.Generator().manual_seed(3)
    M = c["T"] * c["B"]
    taus = [torch.rand(M * c["nq"], generator=gen) for _ in range(3)]
    names = {"h_all": (M, U), "gates": (M, 4 * U), "c_all": (M, U), "cprev": (M, U), "hprev": (M, U),
             "h_all2": (M, U), "targets": (M, c["nq"])}
    out = {}
    for mode in ("fp32", "tf32"):
        L = make_learner(c, gemm=mode)
        try:
            L.load_state_dict(p_on, 0)
            L.load_state_dict(p_tg, 1)
            b, keep = device_batch(raw, c)
            L.step(b, taus)
            out[mode] = {k: L.debug(k, shp).cpu().numpy() for k, shp in names.items()}
            out[mode]["report"] = L.td_abs().cpu().numpy()
            out[mode]["qloss"] = L.stats()["qloss"]
        finally:
            L.close()
    errs = []
    for k in list(names) + ["report", "qloss"]:
        report_diff(errs, k, out["tf32"][k], out["fp32"][k], 2e-3, 2e-3)
    assert not errs, "\n".join(errs)


@pytest.mark.gpu
def test_checkpoint_resume_is_bit_exact():
    """training_state() / load_training_state(): a learner restored after two updates continues
    exactly like the one that never stopped (weights, Adam moments, step counter, lr)."""
    name = "iqn_lstm_clip"
    c = CASES[name]
    g = load_golden("learner_%s.npz" % name)
    A, Bn = make_learner(c), make_learner(c)
    try:
        A.load_state_dict(params_of(g, "online"), 0)
        A.load_state_dict(params_of(g, "target"), 1)
        A.set_lr(3e-4)
        batches = [device_batch(batch_of(g, c, u)[1], c) for u in range(c["updates"])]
        for u in range(c["updates"]):
            A.step(batches[u][0], step_taus(c, g, u))
        Bn.load_training_state(A.training_state())
        for L in (A, Bn):
            L.step(batches[0][0], step_taus(c, g, 0))
        sa, sb = A.training_state(), Bn.training_state()
        assert sa["adam_steps"] == sb["adam_steps"] == c["updates"] + 1 and sa["lr"] == sb["lr"]
        for part in ("online", "target", "adam_m", "adam_v"):
            for k in sa[part]:
                np.testing.assert_array_equal(sa[part][k].numpy(), sb[part][k].numpy(), err_msg=part + "/" + k)
    finally:
        A.close()
        Bn.close()


def _one_update_vs_oracle(c, seed, mode="tf32"):
    """One update of case `c` (full-size shapes) on seeded random data: CUDA path vs the torch oracle."""
    spec = spec_of(c)
    rs = np.random.RandomState(seed)
    S, B, n, T, P = c["T"] + c["P"], c["B"], c["n"], c["T"], c["P"]
    M, Nq, U = T * B, c["nq"], c["lstm"]
    raw = {
        "all_x": rs.randint(0, 256, (S + n, B) + tuple(c["in_shape"])).astype(np.uint8),
        "returns": np.sign(rs.randn(S, B)) * (rs.rand(S, B) < 0.3), "nsteps": np.full((S, B), n, dtype=np.int64),
        "target_masks": (rs.rand(S, B) > 0.05).astype(np.float64),
        "actions": rs.randint(0, c["actions"], (S, B)).astype(np.int64),
        "importance_weights": rs.rand(S, B) * 0.5 + 0.5,
    }
    if U:
        raw["all_hx"] = (0.3 * rs.randn(S + n, B, U)).astype(np.float32)
        raw["all_cx"] = (0.3 * rs.randn(S + n, B, U)).astype(np.float32)
        raw["all_initials"] = (rs.rand(S + n, B) < 0.03).astype(np.float32)
    if c.get("extra"):
        raw["all_extra"] = rs.rand(S + n, B, c["extra"]).astype(np.float32)
    allt = {k: torch.from_numpy(v.copy()) for k, v in raw.items()}

    def st(lo_, hi):
        s = {"x": (allt["all_x"][lo_:hi], allt["all_extra"][lo_:hi]) if c.get("extra") else allt["all_x"][lo_:hi]}
        if U:
            s["layer1_state"] = {"hx": allt["all_hx"][lo_:hi], "cx": allt["all_cx"][lo_:hi],
                                 "initials": allt["all_initials"][lo_:hi]}
        return s
    batch = {"states": st(0, S), "target_states": st(n, S + n), "returns": allt["returns"],
             "nsteps": allt["nsteps"], "target_masks": allt["target_masks"],
             "actions": allt["actions"], "importance_weights": allt["importance_weights"]}
    gen = torch.Generator().manual_seed(seed + 1)
    taus = [torch.rand(M * Nq, generator=gen) for _ in range(3)]
    p_on, p_tg = spec.init_params(1), spec.init_params(2)
    p_ref = {k: v.clone() for k, v in p_on.items()}
    opt = lo.Adam(p_ref, lr=1e-3, eps=c["adam_eps"])
    res = lo.learner_update(spec, p_ref, p_tg, opt, batch, {"target": taus[0], "select": taus[1], "train": taus[2]},
                            c["gamma"], double_q=c["double_q"], rnn_bootstrap=c["rnn_bootstrap"],
                            vf_eps=c["vf_eps"], clip_grad=c["clip_grad"], burn_in_timesteps=P)
    L = make_learner(c, gemm=mode)
    try:
        L.load_state_dict(p_on, 0)
        L.load_state_dict(p_tg, 1)
        b, keep = device_batch(raw, c)
        L.step(b, taus)
        stt = L.stats()
        ok = res["select_margin"].numpy() >= 2e-4
        errs = []
        dt = np.abs(L.debug("targets", (M, Nq)).cpu().numpy() - res["targets"].numpy())
        dr = np.abs(L.td_abs().cpu().numpy() - res["report"].numpy())
        for name, v in (("targets", float(dt[ok].max())), ("report", float(dr[ok].max())),
                        ("qloss", abs(stt["qloss"] - float(res["loss"]))),
                        ("td_mean", abs(stt["td_mean"] - float(res["td_mean"])))):
            if not v <= 1e-4:
                errs.append("%s: |d| = %.3e > 1e-4" % (name, v))
        if abs(stt["grad_norm"] - res["grad_norm"]) > 2e-2 * res["grad_norm"]:
            errs.append("grad_norm %.5f vs %.5f" % (stt["grad_norm"], res["grad_norm"]))
        print("qloss %.6f (oracle %.6f) max|d targets| %.2e max|d report| %.2e near-tie rows %d" % (
            stt["qloss"], float(res["loss"]), dt[ok].max(), dr[ok].max(), int((~ok).sum())))
        assert not errs, "\n".join(errs)
    finally:
        L.close()


@pytest.mark.gpu
def test_learner_full_size_burn_in_vs_oracle():
    """BASELINE config 3 at real size: nature CNN, LSTM 512, Nq 32, B 32, T 20 with a burn-in prefix and
    n = 5 (S + n = 33 time-steps of 84x84x4 frames), vf-rescale on: TF32 path vs the torch oracle."""
    c = dict(in_shape=(4, 84, 84), conv=[(32, 8, 4), (64, 4, 2), (64, 3, 1)], lstm=512, fc=512,
             actions=6, nq=32, embed=64, dueling=True, B=32, T=20, P=8, n=5, gamma=0.997,
             double_q=True, rnn_bootstrap=True, vf_eps=1e-3, clip_grad=40.0, adam_eps=1e-5)
    _one_update_vs_oracle(c, seed=21)


@pytest.mark.gpu
def test_learner_full_size_config2_cnn_iqn_vs_oracle():
    """BASELINE config 2 at real size: IQN on the 3136-wide CNN output (nature_cnn_fc512, no LSTM), T = 1,
    n = 3, B = 32, Nq = 32: TF32 path vs the torch oracle."""
    c = dict(in_shape=(4, 84, 84), conv=[(32, 8, 4), (64, 4, 2), (64, 3, 1)], lstm=0, fc=512,
             actions=6, nq=32, embed=64, dueling=True, B=32, T=1, P=0, n=3, gamma=0.99,
             double_q=True, rnn_bootstrap=False, vf_eps=None, clip_grad=10.0, adam_eps=1.5e-4)
    _one_update_vs_oracle(c, seed=22)


@pytest.mark.gpu
def test_learner_full_size_flappy_bird_rgb_frames_vs_oracle():
    """configs/ple_flappy_bird_iqn_lstm.json at real size: unstacked RGB frames of 120 x 80 (3 input channels:
    conv1 takes the im2col path, its patch rows are not 16-byte runs), the extra-features tuple (one-hot last
    action + reward + timestep = A + 2 values) at the LSTM input, nature CNN -> 4224 features -> LSTM 512 ->
    FC 512, Nq 32, B 32, T 20, n 3: TF32 path vs the torch oracle."""
    c = dict(in_shape=(3, 120, 80), conv=[(32, 8, 4), (64, 4, 2), (64, 3, 1)], lstm=512, fc=512,
             actions=2, nq=32, embed=64, dueling=True, B=32, T=20, P=0, n=3, gamma=0.99,
             double_q=True, rnn_bootstrap=True, vf_eps=1e-3, clip_grad=40.0, adam_eps=1e-5, extra=4)
    _one_update_vs_oracle(c, seed=23)


@pytest.mark.gpu
def test_checkpoint_resume_covers_device_rng_and_dynamic_clip():
    """A resumed learner continues bit-exactly also when the quantile fractions come from the device RNG
    (no injected tau) and the gradient clip follows its moving average (rng counter + EMA are part of
    training_state())."""
    name = "iqn_lstm_clip"
    c = dict(CASES[name], clip_dyn_alpha=0.9)
    g = load_golden("learner_%s.npz" % name)
    A, Bn = make_learner(c), make_learner(c)
    try:
        A.load_state_dict(params_of(g, "online"), 0)
        A.load_state_dict(params_of(g, "target"), 1)
        batches = [device_batch(batch_of(g, c, u)[1], c) for u in range(c["updates"])]
        for u in range(c["updates"]):
            A.step(batches[u][0])
        Bn.load_training_state(A.training_state())
        for L in (A, Bn):
            L.step(batches[0][0])
            L.step(batches[1][0])
        sa, sb = A.training_state(), Bn.training_state()
        assert sa["rng_counter"] == sb["rng_counter"] > 0 and sa["clip_ema"] == sb["clip_ema"] > 0
        for part in ("online", "adam_m", "adam_v"):
            for k in sa[part]:
                np.testing.assert_array_equal(sa[part][k].numpy(), sb[part][k].numpy(), err_msg=part + "/" + k)
    finally:
        A.close()
        Bn.close()
