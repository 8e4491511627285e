"""Reference arm: times the UNMODIFIED opherlieber/rltime (installed under baseline/_ref by
baseline/install_ref.sh) on BASELINE.json's metric, through the reference's own public API:

    IQN(logger, actors, model_config, policy_args).train(**training_args)
        (rltime/training/torch/iqn.py, training/policy_trainer.py:284-325,
         training/multi_step_trainer.py:152-379)

with its own PrioritizedReplayHistoryBuffer, IQNPolicy, StateStore and torch ops.  Only the
*environment side* is synthetic: `SyntheticActors` implements the reference's ActingInterface
(rltime/acting/acting_interface.py) and hands out seeded transitions in the acting schema
instead of stepping Atari (there is no gym / ALE in this image; `oracle/stubs/gym` provides the
import-time names only).  The replay is pre-filled through the buffer's own `update()` in the
constructor of a thin subclass (the reference resolves a non-string `history_mode.type` as-is,
rltime/general/type_registry.py:29-39); every method that runs inside the timed region is the
reference's.

Two devices:
  device="cpu"   the reference as shipped on the host cores.  threads=1 is "as shipped"
                 (TorchModel.__init__ calls torch.set_num_threads(1), models/torch/torch_model.py:25);
                 threads=N re-raises the intra-op thread count right after policy creation.
  device="cuda"  the reference's own torch-CUDA path on the B200: policy on cuda
                 (policies/torch/torch_policy.py:44-59) and StateStore("cuda")
                 (general/backend.py:136-153) -- the bar SURVEY.md 2.1 / 8(d) names.

TEST / MEASUREMENT INFRASTRUCTURE: imported only by bench.py's reference legs and tests/.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
STUBS = os.path.join(ROOT, "oracle", "stubs")


def available():
    return os.path.isdir(os.path.join(REF_DIR, "rltime"))


def _import_reference():
    if not available():
        raise RuntimeError("baseline/_ref is missing: run baseline/install_ref.sh in the build container")
    for p in (REF_DIR, STUBS, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import rltime  # noqa: F401
    assert os.path.realpath(os.path.dirname(rltime.__file__)).startswith(os.path.realpath(REF_DIR)), \
        "a different rltime is on the path: %s" % rltime.__file__


class _Timeline:
    """Timestamps of every priority write-back (one per learner update).  `phases` = list of
    (threads, warmup, steps): the measurement harness (not the reference) switches the torch intra-op
    thread count between phases, so one pre-filled replay serves the all-cores and the as-shipped
    1-thread measurement."""

    def __init__(self, device, phases):
        self.cuda = device == "cuda"
        self.marks = []
        self.phases = phases
        self.bounds = []          # (first mark index, last mark index, threads) per phase
        n = 0
        for th, w, k in phases:
            self.bounds.append((n + w - 1, n + w + k - 1, th))
            n += w + k
        self.total = n
        self.starts = {}
        n = 0
        for th, w, k in phases:
            self.starts[n] = th
            n += w + k

    def mark(self):
        i = len(self.marks)
        if self.cuda and any(i in (a, b) for a, b, _ in self.bounds):
            import torch
            torch.cuda.synchronize()
        self.marks.append(time.perf_counter())
        nxt = self.starts.get(i + 1)
        if nxt and not self.cuda:
            import torch
            torch.set_num_threads(nxt)


def make_prefilled_per(stream_factory, prefill, timeline):
    """PrioritizedReplayHistoryBuffer + a constructor-time prefill + an update_losses timestamp."""
    from rltime.history.prioritized_replay_history import PrioritizedReplayHistoryBuffer

    class PrefilledPER(PrioritizedReplayHistoryBuffer):
        def __init__(self, **kwargs):
            super().__init__(**kwargs)
            stream = stream_factory()
            t0 = time.perf_counter()
            fed = 0
            while fed < prefill:
                samples = stream.next_samples()
                super().update(samples)
                fed += len(samples)
            self.train_quota = 0          # the prefill does not count as acting
            self.prefill_s = time.perf_counter() - t0
            self.prefilled = fed
            timeline.hist = self

        def update_losses(self, indices, losses):
            import numpy as np
            # canonical fp64 widening: same shim as the golden generator
            super().update_losses(indices, np.asarray(losses, dtype=np.float64))
            timeline.mark()
    return PrefilledPER


class SyntheticActors:
    """ActingInterface over a seeded synthetic transition stream (no env stepping, no inference)."""

    def __init__(self, stream, threads):
        self.stream = stream
        self.threads = threads
        self.policy = None

    def get_spaces(self):
        import gym
        import numpy as np
        return (gym.spaces.Box(0, 255, self.stream.frame_shape, dtype=np.uint8),
                gym.spaces.Discrete(self.stream.num_actions))

    def get_env_count(self):
        return self.stream.num_envs

    def set_actor_policy(self, policy):
        self.policy = policy
        if self.threads and self.threads > 1:
            import torch
            # "all cores" variant: undo TorchModel.__init__'s torch.set_num_threads(1)
            torch.set_num_threads(self.threads)

    def update_state(self, progress, policy_state=None):
        pass

    def get_samples(self, min_samples):
        out = []
        while len(out) < max(1, min_samples):
            out += self.stream.next_samples()
        return out

    def close(self):
        pass


def run(cfg, device="cpu", threads=1, steps=20, warmup=5, size=None, seed=1, extra_phases=()):
    """Runs W + K learner updates of the unmodified reference; returns a dict with updates/s over
    exactly K updates (wall clock between priority write-backs W and W+K, CUDA-synchronised on both
    sides for device="cuda").  extra_phases: further (threads, warmup, steps) measurements on the
    same pre-filled replay, reported under "phases"."""
    _import_reference()
    import random
    import numpy as np
    import torch
    from rltime.general.config import load_config
    from rltime.training.torch.iqn import IQN
    from rltime_b200.synthetic import SyntheticStream

    size = int(size or cfg["size"])
    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    E, T, B = cfg["envs"], cfg["T"], cfg["B"]
    feed = (B * T) // cfg["train_frequency"]

    def stream_factory():
        return SyntheticStream(num_envs=E, frame_shape=cfg["frame"], num_actions=cfg["A"],
                               lstm_units=cfg["units"], seed=seed, clip_rewards=True,
                               pooled_state=True)
    warmup = max(int(warmup), 1)
    phases = [(threads, warmup, steps)] + [tuple(p) for p in extra_phases]
    timeline = _Timeline(device, phases)
    hist_cls = make_prefilled_per(stream_factory, size, timeline)
    actors = SyntheticActors(SyntheticStream(num_envs=E, frame_shape=cfg["frame"], num_actions=cfg["A"],
                                             lstm_units=cfg["units"], seed=seed + 1, clip_rewards=True,
                                             pooled_state=True), threads)
    model_config = load_config("models/nature_cnn_lstm512_fc512.json")
    policy_args = dict(dueling=True, embedding_dim=64, num_sampling_quantiles=cfg["Nq"],
                       injection_layer=-1, cuda=(device == "cuda"))
    tr = IQN(logger=None, actors=actors, model_config=model_config, policy_args=policy_args)
    total_updates = timeline.total
    t_all = time.perf_counter()
    tr.train(
        total_steps=(total_updates + 1) * feed, log_freq=10 ** 12, target_update_freq=20000,
        clip_rewards=True, gamma=cfg["gamma"], mbatch_size=B, nstep_train=T, nstep_target=cfg["n"],
        lr=3e-4, lr_anneal=True, adam_epsilon=cfg["adam_eps"], double_q=True, rnn_bootstrap=True,
        clip_grad=cfg["clip_grad"], vf_scale_epsilon=None, burn_in_timesteps=cfg.get("P", 0),
        warmup_steps=0,
        history_mode={"type": hist_cls, "args": {
            "size": size, "train_frequency": cfg["train_frequency"], "alpha": cfg["alpha"],
            "beta": cfg["beta"]}})
    if device == "cuda":
        torch.cuda.synchronize()
    wall = time.perf_counter() - t_all
    marks = timeline.marks
    assert len(marks) >= total_updates, "reference ran %d updates, expected %d" % (len(marks), total_updates)
    per_phase = []
    for (a, b, th), (_, w, k_) in zip(timeline.bounds, phases):
        d = marks[b] - marks[a]
        per_phase.append({"threads": th if device == "cpu" else None, "warmup": w, "updates_timed": k_,
                          "updates_per_s": k_ / d, "ms_per_update": 1e3 * d / k_})
    k, dt = steps, marks[timeline.bounds[0][1]] - marks[timeline.bounds[0][0]]
    log = tr.value_log.get()            # resets the 'get'-scoped values: read once
    vals = log.get("train", {})
    timings = log.get("timings_mean_ms", {})
    return {
        "updates_per_s": k / dt, "ms_per_update": 1e3 * dt / k, "updates_timed": k,
        "device": device, "threads": threads if device == "cpu" else None, "phases": per_phase,
        "replay_transitions": int(timeline.hist.prefilled), "prefill_s": round(timeline.hist.prefill_s, 1),
        "wall_s": round(wall, 1), "qloss": float(vals.get("qloss", float("nan"))),
        "timings_mean_ms": {k_: float(v) for k_, v in timings.items()},
        "torch": torch.__version__,
    }


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--size", type=int, default=100_000)
    ap.add_argument("--also-1thread", type=int, default=0, help="extra as-shipped 1-thread phase of this many updates")
    a = ap.parse_args()
    sys.path.insert(0, ROOT)
    from bench import CFG
    extra = [(1, 1, a.also_1thread)] if a.also_1thread else []
    print(json.dumps(run(dict(CFG), a.device, a.threads, a.steps, a.warmup, a.size, extra_phases=extra)))
