"""Seeded replay scenarios (TEST INFRASTRUCTURE — never imported by the product path).

A scenario is a deterministic script of `update` / `get_train_data` /
`update_losses` calls (the History interface, rltime/history/history.py:123,296,332)
over a SyntheticStream.  The same driver is run against
  * the unmodified reference class (oracle/gen_golden.py, build container only),
  * the oracle restatement (oracle/replay_oracle.py),
  * the CUDA-backed drop-in (rltime_b200.history),
and the recorded traces are compared field by field.
"""
import random
import numpy as np

from rltime_b200.synthetic import SyntheticStream

# name -> parameters.  Small frames / LSTM widths keep the golden fixtures small;
# full-size shapes are exercised oracle-vs-CUDA on the GPU box.
SCENARIOS = {
    # R2D2-like PER: overlapping sequences, burn-in prefix, eviction wrap-around
    "per_seq_small": dict(kind="per", size=600, envs=4, T=6, P=3, n=2, B=4, iters=260,
                          alpha=0.9, beta=0.6, frame=(2, 5, 5), units=4, actions=3,
                          done_mode="bernoulli", done_p=0.06, feed="lockstep"),
    # defaults (alpha .6 beta .4, overlap T//2), async/ragged feeding, beta anneal
    "per_async": dict(kind="per", size=500, envs=5, T=4, P=0, n=3, B=6, iters=300,
                      alpha=0.6, beta=0.4, beta_anneal=True, frame=(1, 4, 4), units=3,
                      actions=4, done_mode="bernoulli", done_p=0.1, feed="async"),
    # Rainbow-like: T=1 (n >= S: non-overlapped state stacking), global IS scaling + min tree
    "per_rainbow": dict(kind="per", size=400, envs=3, T=1, P=0, n=3, B=8, iters=300,
                        alpha=0.5, beta=0.4, global_importance_scaling=True,
                        frame=(4, 3, 3), units=0, actions=5, done_mode="bernoulli",
                        done_p=0.08, feed="lockstep"),
    # the bench configuration's shape parameters (T=20, n=2, overlap 10), shrunk frames
    "per_bench_shape": dict(kind="per", size=3000, envs=8, T=20, P=0, n=2, B=8, iters=120,
                            alpha=0.9, beta=0.6, frame=(4, 4, 4), units=4, actions=6,
                            done_mode="periodic", done_period=50, feed="lockstep",
                            feed_steps=(8, 24)),
    # long burn-in, n=5, custom overlap
    "per_burnin": dict(kind="per", size=2500, envs=4, T=8, P=12, n=5, B=5, iters=160,
                       alpha=0.9, beta=0.6, overlap=6, frame=(1, 6, 6), units=5, actions=2,
                       done_mode="bernoulli", done_p=0.03, feed="lockstep",
                       feed_steps=(4, 30)),
    # uniform replay (rltime/history/replay_history.py)
    "uniform_seq": dict(kind="uniform", size=700, envs=4, T=5, P=2, n=2, B=6, iters=200,
                        frame=(2, 4, 4), units=3, actions=3, done_mode="bernoulli",
                        done_p=0.05, feed="lockstep"),
    # online n-step buffer (rltime/history/online_history.py): fixed_target, round-robin over envs
    "online_small": dict(kind="online", size=0, envs=4, T=5, P=0, n=5, B=4, iters=120,
                         frame=(2, 4, 4), units=3, actions=3, done_mode="bernoulli", done_p=0.1,
                         feed="lockstep", feed_steps=(1, 4)),
    "online_async": dict(kind="online", size=0, envs=3, T=4, P=0, n=4, B=5, iters=150,
                         frame=(1, 3, 3), units=0, actions=2, done_mode="bernoulli", done_p=0.15,
                         feed="async", feed_steps=(1, 5), max_delayed_steps=9),
    # uniform replay that shifts sequences off episode boundaries (replay_history.py:142-171)
    "uniform_avoid_crossing": dict(kind="uniform", size=600, envs=4, T=6, P=2, n=2, B=6, iters=200,
                                   frame=(2, 4, 4), units=3, actions=3, done_mode="bernoulli",
                                   done_p=0.08, feed="lockstep", avoid_episode_crossing=True),
    # ONE env holding the whole (small) buffer: the evicted predecessor of the oldest transition still backs
    # its state, i.e. N + 1 live positions of one env (eviction wrap-around every 90 transitions)
    "uniform_single_env": dict(kind="uniform", size=90, envs=1, T=4, P=1, n=2, B=5, iters=160,
                               frame=(1, 3, 3), units=2, actions=3, done_mode="bernoulli", done_p=0.07,
                               feed="lockstep", feed_steps=(1, 9)),
    "per_single_env": dict(kind="per", size=120, envs=1, T=4, P=2, n=2, B=4, iters=160, alpha=0.7, beta=0.5,
                           frame=(1, 3, 3), units=2, actions=3, done_mode="bernoulli", done_p=0.07,
                           feed="lockstep", feed_steps=(1, 9)),
    "uniform_async": dict(kind="uniform", size=300, envs=3, T=1, P=0, n=3, B=7, iters=200,
                          frame=(1, 3, 3), units=0, actions=4, done_mode="bernoulli",
                          done_p=0.1, feed="async"),
}

GAMMA = 0.99


def discount_function(nstep, reward, policy_output):
    """Same arithmetic as MultiStepTrainer._get_discount_function
    (rltime/training/multi_step_trainer.py:70-74)."""
    return (GAMMA ** nstep) * reward


def history_kwargs(p):
    if p["kind"] == "online":
        kw = dict(nstep_target=p["n"], nstep_train=p["T"], prefix_steps=p["P"])
        if "max_delayed_steps" in p:
            kw["max_delayed_steps"] = p["max_delayed_steps"]
        return kw
    kw = dict(size=p["size"], train_frequency=None, nstep_target=p["n"],
              nstep_train=p["T"], prefix_steps=p["P"])
    if p["kind"] == "uniform" and p.get("avoid_episode_crossing"):
        kw["avoid_episode_crossing"] = True
    if p["kind"] == "per":
        for k in ("alpha", "beta", "beta_anneal", "overlap", "max_weight_factor",
                  "global_importance_scaling", "eps"):
            if k in p:
                kw[k] = p[k]
    return kw


def make_stream(p, seed=1):
    return SyntheticStream(num_envs=p["envs"], frame_shape=p["frame"],
                           num_actions=p["actions"], lstm_units=max(p["units"], 1),
                           seed=seed, done_mode=p["done_mode"],
                           done_period=p.get("done_period", 500),
                           done_p=p.get("done_p", 0.01), pool=64,
                           recurrent=p["units"] > 0)


def _to_np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def flatten_train_data(td):
    """Nested train-data dict -> flat {path: ndarray} (drops empty dicts)."""
    out = {}

    def rec(prefix, v):
        if isinstance(v, dict):
            for k, vv in v.items():
                rec(prefix + "/" + k if prefix else k, vv)
        elif isinstance(v, (tuple, list)):
            for i, vv in enumerate(v):
                rec(prefix + "/%d" % i, vv)
        elif v is None:
            return
        else:
            out[prefix] = _to_np(v)
    rec("", td)
    return out


def run_scenario(name, history, get_idxes, get_tree_sum=None, seed=0):
    """Drives `history` through scenario `name`; returns a trace dict
    {field: ndarray stacked over the calls that returned data} plus "none_calls"
    (indices of calls that returned None) and "num_calls".  get_idxes(history) -> list of the prioritization
    indices drawn by the last get_train_data (None for uniform replay)."""
    p = SCENARIOS[name]
    stream = make_stream(p)
    script = np.random.RandomState(1000 + seed)
    random.seed(seed)
    np.random.seed(seed)
    cols = {}          # field -> list of per-call arrays
    none_calls = []
    lo, hi = p.get("feed_steps", (1, 6))
    ncall = 0
    total_iters = p["iters"]
    for it in range(total_iters):
        for _ in range(script.randint(lo, hi)):
            if p["feed"] == "lockstep":
                subset = None
            else:
                k = script.randint(1, p["envs"] + 1)
                subset = script.permutation(p["envs"])[:k]
            history.update(stream.next_samples(subset))
        progress = it / total_iters
        td = history.get_train_data(p["B"], progress)
        ncall += 1
        if td is None:
            none_calls.append(ncall - 1)
            continue
        flat = flatten_train_data(td)
        for k, v in flat.items():
            cols.setdefault(k, []).append(v)
        idx = get_idxes(history)
        if idx is not None:
            cols.setdefault("idxes", []).append(np.asarray(idx, dtype=np.int64))
        if get_tree_sum is not None:
            cols.setdefault("tree_sum", []).append(np.float64(get_tree_sum(history)))
        if p["kind"] == "per":
            li = flat["extra_data/loss_indices"][p["P"]:].reshape(-1, 2)
            losses32 = np.abs(script.randn(li.shape[0])).astype(np.float32)
            if it % 7 == 3:
                losses32[::3] = 0.0
            # canonical-semantics shim: losses widened exactly to fp64
            history.update_losses(li, losses32.astype(np.float64))
    trace = {k: np.stack(v) for k, v in cols.items()}
    trace["none_calls"] = np.asarray(none_calls, dtype=np.int64)
    trace["num_calls"] = np.int64(ncall)
    return trace
