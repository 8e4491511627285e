"""Headline benchmark: learner updates/sec on a 1M-transition prioritized sequence replay.

One "step" = one full learner update of BASELINE.json's metric:
  stratified sum-tree draw of B=32 sequences -> n-step assembly + importance weights ->
  gather of the (T+n) x B x (4,84,84) uint8 frame stack (+ LSTM states) ->
  double-Q IQN targets (target + selection forward) -> training forward, quantile-Huber loss,
  backward, grad-norm clip, Adam -> priority write-back.
Workload (config.workload): BASELINE.json configs[2]/[4] shape — Atari IQN+LSTM, T=20, n=2,
B=32, Nq=32, 1M-transition weighted PER (alpha .9, beta .6, overlap 10), synthetic transitions
(SURVEY.md 8d), random-init nature-CNN -> LSTM512 -> FC512 dueling IQN.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`value`  : updates/s with everything resident in HBM (device-timed, CUDA events).
`e2e`    : the same update through the public Python API with HOST buffers: every step
           ingests `train_frequency`-worth of new transitions from pinned host memory (H2D)
           and reads the step's loss back (D2H).
`--impl reference`: the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh)
           through its own IQN.train() / PrioritizedReplayHistoryBuffer on the host cores
           (baseline/ref_arm.py): 1M-transition replay, --steps / --warmup honoured, all host
           threads (`value`) and the as-shipped 1 thread.  Falls back to the oracle port only
           when baseline/_ref is absent.
Extra keys of the GPU line: `value_fast` / `value_fp32` (the other GEMM precisions), `value_long`
(>= 500 steps), `cuda_torch_baseline` (the unmodified reference's own torch-CUDA path on the same
B200), `config3_burnin40` and `config2_cnn_iqn` (BASELINE.json configs[2] / configs[1] lines),
`acting_us_per_step`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CFG = dict(size=1_000_000, envs=32, T=20, P=0, n=2, B=32, Nq=32, A=6, units=512, fc=512,
           frame=(4, 84, 84), conv=[(32, 8, 4), (64, 4, 2), (64, 3, 1)], alpha=0.9, beta=0.6,
           gamma=0.99, clip_grad=40.0, adam_eps=1e-5, train_frequency=4)
FRAME_BYTES = 4 * 84 * 84


def peaks():
    """(HBM GB/s, dense bf16 TFLOP/s burst, sustained, source).  TF32 tensor rate = bf16 / 2.
    Burst = a kernel timed alone (what the per-launch event brackets measure); sustained = a kernel
    inside a long step."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        burst = d.get("bf16_tflops", 1600.0)
        return d.get("hbm_gbs", 6650.0), burst, d.get("bf16_tflops_sustained", burst), "measured"
    return 6650.0, 1600.0, 1400.0, "fallback"


def static_config(world, size, gemm="tf32"):
    """The `config` object: identical in the GPU arm and the reference arm."""
    return {"workload": "atari_iqn_lstm seq-PER: N=%d T=20 n=2 B=32 Nq=32 A=6 nature-CNN-LSTM512-FC512 dueling "
                        "double-Q rnn_bootstrap, weighted PER alpha .9 beta .6 overlap 10, train_frequency 4" % size,
            "replay_total": size, "replay_per_gpu": size // world,
            "global_batch": "%d sequences x 20 steps" % (32 * world),
            "parallelism": "dp%d" % world,
            "l2": "inputs (28 GB frame store) exceed L2; every draw gathers different rows"}


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region: NVML in-process (a sample
    every few milliseconds), nvidia-smi as the fallback (one sample per ~100 ms call)."""

    # nvmlClocksThrottleReason* / nvmlClocksEventReason* bit masks
    _MASKS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
              ("sw_power_cap", 0x4))

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._th = None
        self.source = "nvidia-smi"

    def _nvml(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            return pynvml, h, mx
        except Exception:
            return None

    def _run(self):
        nv = self._nvml()
        if nv is not None:
            pynvml, h, mx = nv
            self.source = "nvml"
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
            while not self._stop.is_set():
                try:
                    sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    try:
                        bits = int(get_reasons(h)) if get_reasons else 0
                    except Exception:
                        bits = 0
                    self.rows.append([str(sm), str(mx)] +
                                     ["Active" if bits & m else "Not Active" for _, m in self._MASKS])
                except Exception:
                    break
                self._stop.wait(0.004)
            if self.rows:
                return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                     "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=5)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


def dominant_roofline(shape, fl, ms_total, n, family_ms, prof_steps, tf32_peak, which, tf32_sustained=None):
    """roofline object of the single GEMM shape that takes the most device time per update."""
    kind, M, N, K, tA, tB = shape
    kinds = {0: "GEMM", 1: "implicit-GEMM conv forward", 2: "implicit-GEMM conv weight gradient",
             3: "implicit-GEMM conv data gradient"}
    key = "%dx%dx%d" % (M, N, K)
    name = "tcgen05 TF32 %s [%d x %d] = [%d x %d] . [%d x %d] (A %s, B %s)" % (
        kinds.get(kind, "GEMM"), M, N, M, K, K, N, "MN-major" if tA else "K-major",
        "K-major" if tB else "MN-major")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(key)
    achieved = fl * n / (ms_total * 1e-3) / 1e12 if ms_total > 0 else None
    return {"kernel": name, "shape": key, "bound": "tensor", "achieved": achieved, "peak": tf32_peak,
            "unit": "TFLOP/s",
            "peak_source": which + " cuBLAS bf16 BURST / 2 (TF32 multiplies at half the bf16 rate; the kernel is "
                                   "timed alone, one launch per event bracket, so the burst figure applies)",
            "frac": achieved / tf32_peak if achieved else None,
            "frac_vs_sustained_peak": achieved / tf32_sustained if achieved and tf32_sustained else None,
            "peak_sustained": tf32_sustained, "traffic": traffic,
            "algorithmic_flops_per_launch": fl, "launches_timed": int(n),
            "us_per_launch": 1e3 * ms_total / max(n, 1),
            "share_of_gemm_time": ms_total / family_ms if family_ms > 0 else None,
            "ms_per_update": ms_total / prof_steps,
            "timing": "CUDA events around each launch, launches issued one by one on a single stream "
                      "(the timed region replays the same kernels from CUDA graphs)"}


# --------------------------------------------------------------------------- GPU arm
def build_history(cfg, device, seed, rank):
    """1M-transition (cfg["size"]) prioritized sequence replay filled with the synthetic stream of
    SURVEY.md 8(d) (generated on the device; the fill is untimed setup)."""
    from rltime_b200.history import DevicePrioritizedReplayHistoryBuffer
    hist = DevicePrioritizedReplayHistoryBuffer(
        size=cfg["size"], train_frequency=None, alpha=cfg["alpha"], beta=cfg["beta"],
        nstep_target=cfg["n"], nstep_train=cfg["T"], prefix_steps=cfg["P"], gamma=cfg["gamma"],
        max_envs=cfg["envs"], device=device)
    E, U, A = cfg["envs"], cfg["units"], cfg["A"]
    g = torch.Generator(device=device).manual_seed(seed + rank)
    pool = torch.randint(0, 255, (256,) + cfg["frame"], dtype=torch.uint8, device=device, generator=g)
    if U:
        hist.set_structure({"x": ("leaf", 0), "layer0_state": {},
                            "layer1_state": {"hx": ("leaf", 1), "cx": ("leaf", 2), "initials": ("leaf", 3)},
                            "layer2_state": {}},
                           {"actions": ("leaf", 0), "qvalues": ("leaf", 1)})
    else:
        hist.set_structure({"x": ("leaf", 0), "layer0_state": {}, "layer1_state": {}},
                           {"actions": ("leaf", 0), "qvalues": ("leaf", 1)})
    steps_total = cfg["size"] // E
    chunk_steps = 256
    rs = np.random.RandomState(seed + rank)
    t0 = time.time()
    for s0 in range(0, steps_total, chunk_steps):
        ns = min(chunk_steps, steps_total - s0)
        m = ns * E
        gidx = torch.arange(s0 * E, s0 * E + m, device=device)
        frames = pool[gidx & 255]
        step = (torch.arange(m, device=device) // E) + s0
        env = torch.arange(m, device=device) % E
        done = ((step + 1 + 37 * env) % 500) == 0
        prev_done = ((step + 37 * env) % 500) == 0
        actions = torch.randint(0, A, (m,), device=device, generator=g)
        qv = torch.randn(m, A, device=device, generator=g)
        reward = np.sign(rs.randn(m))
        leaves = [frames]
        if U:
            leaves += [torch.randn(m, U, device=device, generator=g), torch.randn(m, U, device=device, generator=g),
                       prev_done.float()]
        hist.update_arrays(env.cpu().numpy(), reward, done.cpu().numpy(), leaves, [actions, qv])
    torch.cuda.synchronize(device)
    return hist, time.time() - t0


def build_learner(cfg, device, gemm, seed=0):
    from rltime_b200.learner import DeviceLearner
    from rltime_b200.init import init_params
    U, A = cfg["units"], cfg["A"]
    learner = DeviceLearner(cfg["frame"], cfg["conv"], U, cfg["fc"], A, cfg["Nq"], 64, True,
                            mbatch=cfg["B"], nstep_train=cfg["T"], burn_in=cfg["P"],
                            nstep_target=cfg["n"], gamma=cfg["gamma"], double_q=True,
                            rnn_bootstrap=bool(U), vf_scale_epsilon=None, clip_grad=cfg["clip_grad"],
                            adam_epsilon=cfg["adam_eps"], lr=3e-4, seed=seed, device=device, gemm=gemm)
    learner.load_state_dict(init_params(learner.param_info, U, seed=1), 0)
    learner.load_state_dict(init_params(learner.param_info, U, seed=2), 1)
    return learner


def build_device_workload(cfg, device, seed, rank):
    """(history, learner, fill seconds) of the headline workload (scripts/ use this)."""
    hist, fill_s = build_history(cfg, device, seed, rank)
    return hist, build_learner(cfg, device, cfg.get("gemm", "tf32"), seed), fill_s


def one_update(hist, learner, B, world=1):
    from rltime_b200 import parallel
    td = hist.draw(B, 0.0)          # what IQNTrainer.train calls: get_train_data minus the dict of views
    assert td is not None
    learner.prefetch(hist.last_batch, hist._stream())     # frame conversion behind the gather, on the replay stream
    if world > 1:
        # local gradients -> NCCL sum over NVLink -> identical clip + Adam on every rank
        parallel.data_parallel_step(learner, hist.last_batch, world)
    else:
        learner.step(hist.last_batch)
    hist.update_losses_device(learner.td_abs(), ready=learner.wait_loss)


def timed_updates(hist, learner, B, world, steps, warmup, barrier):
    """W untimed + K timed full updates; CUDA events on the current stream, synchronised on both sides."""
    for _ in range(warmup):
        one_update(hist, learner, B, world)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        one_update(hist, learner, B, world)
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1)


def side_config_line(name, cfg, device, steps, warmup, gemm):
    """One extra BASELINE.json config as its own small line (own replay + learner; N=1 only)."""
    try:
        hist, fill_s = build_history(cfg, device, seed=3, rank=0)
        learner = build_learner(cfg, device, gemm)
        ms = timed_updates(hist, learner, cfg["B"], 1, steps, warmup, lambda: torch.cuda.synchronize(device))
        st = learner.stats()
        S, n = cfg["T"] + cfg["P"], cfg["n"]
        state_bytes = FRAME_BYTES + (2 * cfg["units"] * 4 + 4 if cfg["units"] else 0)
        out = {"workload": name, "value": steps / (ms / 1e3), "unit": "updates/s", "ms_per_step": ms / steps,
               "steps": steps, "warmup": warmup, "replay_transitions": cfg["size"], "gemm": gemm,
               "gather_bytes_per_update": int(2 * (S + n) * cfg["B"] * state_bytes), "last_stats": st}
        learner.close()
        hist.close()
        del learner, hist
        torch.cuda.empty_cache()
        return out
    except Exception as ex:  # noqa: BLE001  (a side line must never take the headline down)
        return {"workload": name, "error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}


def ref_subprocess(device, threads, steps, warmup, size, also_1thread=0, timeout=900):
    """Runs baseline/ref_arm.py (the unmodified reference) in its own process; returns its dict."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_arm.py"), "--device", device,
           "--threads", str(threads), "--steps", str(steps), "--warmup", str(warmup), "--size", str(size),
           "--also-1thread", str(also_1thread)]
    try:
        r = subprocess.run(cmd, capture_output=True, timeout=timeout, cwd=ROOT)
        lines = [l for l in r.stdout.decode().splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": "ref_arm rc=%d: %s" % (r.returncode, r.stderr.decode()[-400:].replace("\n", " | "))}
        return json.loads(lines[-1])
    except Exception as ex:  # noqa: BLE001
        return {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:300])}


def ref_threads():
    # measured in round 1: 128 intra-op threads are slower than 32 on these shapes
    return max(1, min(os.cpu_count() or 1, 32))


def cpu_baseline_from(res, sample):
    if "unavailable" in res:
        return {"value": None, "unit": "updates/s", "cores": None, "kind": "reference", "sample": sample,
                "unavailable": res["unavailable"]}
    one = [p for p in res.get("phases", []) if p.get("threads") == 1]
    return {"value": res["updates_per_s"], "unit": "updates/s", "cores": res["threads"], "kind": "reference",
            "sample": sample + "; %d-transition replay (prefill %.1fs untimed), %d updates timed" % (
                res["replay_transitions"], res["prefill_s"], res["updates_timed"]),
            "as_shipped_1_thread": ({"value": one[0]["updates_per_s"], "updates_timed": one[0]["updates_timed"],
                                     "why": "TorchModel.__init__ calls torch.set_num_threads(1) "
                                            "(rltime/models/torch/torch_model.py:25)"} if one else None),
            "reference_timings_mean_ms": res.get("timings_mean_ms"), "torch": res.get("torch")}


def run_gpu(args):
    from rltime_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        from rltime_b200 import parallel
        parallel.init_process_group("nccl")
    cfg = dict(CFG)
    if args.size:
        cfg["size"] = args.size
    total_size = cfg["size"]
    if world > 1:
        # the 1M-transition replay is SHARDED by env across the GPUs (BASELINE config 4): every rank owns the
        # transitions of its 32 / world envs in a buffer of 1M / world
        assert cfg["envs"] % world == 0
        cfg["size"] = cfg["size"] // world
        cfg["envs"] = cfg["envs"] // world
    cfg["gemm"] = args.gemm
    import random
    random.seed(rank)
    hist, fill_s = build_history(cfg, device, seed=0, rank=rank)
    learner = build_learner(cfg, device, cfg["gemm"])
    lib = _lib.load()
    B = cfg["B"]
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(device)

    if world > 1:
        from rltime_b200 import parallel
        if os.environ.get("RT_DP_LIB", "1") != "0":
            parallel.init_library_comm(learner)      # rt_comm_init: the gradient exchange runs inside the library
        parallel.broadcast_params_(learner)
    for _ in range(warmup):
        one_update(hist, learner, B, world)
    barrier()
    launches0 = lib.rt_launch_count()
    # clocks / throttle reasons are sampled from the start of the timed region to the end of the
    # end-to-end loop (every phase in between keeps the GPU under the same load)
    clk = ClockSampler(local)
    clk.__enter__()
    ms = timed_updates(hist, learner, B, world, args.steps, 0, barrier)
    launches = lib.rt_launch_count() - launches0
    stats = learner.stats()
    # the same loop over >= 500 updates: clock ramp / boost cannot flatter a 30 ms region
    long_steps = max(args.long_steps, args.steps)
    ms_long = timed_updates(hist, learner, B, world, long_steps, 0, barrier)

    # GEMM-shaped launches, live (extra profiled steps AFTER the timed region: the event pairs
    # around ~190 launches per update perturb the step slightly, so they stay out of `value`)
    learner.profile_gemms(True)
    prof_steps = 5
    for _ in range(prof_steps):
        one_update(hist, learner, B, world)
    per_launch = learner.gemm_launches()
    shapes = learner.gemm_shapes()
    gemm_ms, gemm_flops, gemm_n = learner.gemm_time()
    learner.profile_gemms(False)
    # dominant kernel = the GEMM shape with the largest summed device time over the profiled steps.
    # The profiled steps issue every launch one by one, so each bracket also holds the launch latency
    # of its kernel (the timed region replays graphs and does not pay it); the ranking subtracts the
    # live-measured cost of bracketing a near-empty launch so that 19 tiny BPTT products do not outrank
    # the three 50 us hidden-layer GEMMs they are shorter than under ncu.  `achieved` uses raw times.
    x1 = torch.zeros(32, device=device)
    ovh = []
    for _ in range(30):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        x1.add_(1.0)
        a1.record()
        a1.synchronize()
        ovh.append(a0.elapsed_time(a1))
    ovh_ms = sorted(ovh)[len(ovh) // 2]
    groups = {}
    for (fl, t), sh in zip(per_launch, shapes):
        gsum = groups.setdefault(sh, [0.0, 0, fl, 0.0])
        gsum[0] += t
        gsum[1] += 1
        gsum[3] += max(t - ovh_ms, 0.0)
    ranked = sorted(groups.items(), key=lambda kv: -kv[1][3])
    top_shape, (top_ms, top_n, top_fl, _) = ranked[0] if ranked else ((0, 0, 0, 0, 0, 0), (0.0, 0, 0.0, 0.0))
    top_shapes = [{"shape": "%dx%dx%d" % (sh[1], sh[2], sh[3]), "kind": sh[0], "launches_per_update": v[1] / prof_steps,
                   "us_per_launch": 1e3 * v[0] / v[1], "us_per_update": 1e3 * v[0] / prof_steps}
                  for sh, v in ranked[:5]]

    # gather kernel alone (roofline): time draws without the learner
    torch.cuda.synchronize(device)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    for _ in range(5):
        hist.get_train_data(B, 0.0)
    torch.cuda.synchronize(device)
    # the gather kernel timed alone (inside an update it shares the GPU with the previous update's
    # backward pass): CUDA events around each launch on the replay stream, one draw at a time
    hist.profile_gather(True)
    g0.record()
    for _ in range(reps):
        hist.get_train_data(B, 0.0)
        torch.cuda.synchronize(device)
    g1.record()
    torch.cuda.synchronize(device)
    draw_ms = g0.elapsed_time(g1) / reps
    gather_ms_total, gather_n = hist.gather_time()
    hist.profile_gather(False)

    # end to end through the public API with host buffers
    E = cfg["envs"]
    feed = (B * cfg["T"]) // cfg["train_frequency"]          # transitions ingested per update
    feed = (feed // E) * E
    rs = np.random.RandomState(7)
    h_frames = torch.from_numpy(rs.randint(0, 255, (feed,) + cfg["frame"]).astype(np.uint8)).pin_memory()
    h_hx = torch.randn(feed, cfg["units"]).pin_memory()
    h_cx = torch.randn(feed, cfg["units"]).pin_memory()
    h_init = torch.zeros(feed).pin_memory()
    h_act = torch.randint(0, cfg["A"], (feed,)).pin_memory()
    h_qv = torch.randn(feed, cfg["A"]).pin_memory()
    env_ids = np.arange(feed) % E
    reward = np.sign(rs.randn(feed))
    done = np.zeros(feed, dtype=np.uint8)
    h2d = sum(t.numel() * t.element_size() for t in (h_frames, h_hx, h_cx, h_init, h_act, h_qv)) + \
        feed * (8 + 1 + 8 + 4) + B * 8
    d2h = cfg["T"] * B * 4 + B * 4 + 16

    def e2e_update():
        hist.update_arrays(env_ids, reward, done,
                           [h_frames.numpy(), h_hx.numpy(), h_cx.numpy(), h_init.numpy()],
                           [h_act.numpy(), h_qv.numpy()])
        one_update(hist, learner, B, world)
        return learner.loss()["qloss"]       # D2H of the step's loss (final before its backward pass)
    for _ in range(3):
        e2e_update()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_update()
    barrier()
    e2e_s = time.perf_counter() - t0
    clk.__exit__()

    # acting-side inference: actor_predict for E envs at T=1 through the public policy object
    acting = None
    try:
        from rltime_b200.training import DevicePolicy
        pol = DevicePolicy(learner, cfg["A"])
        obs = rs.randint(0, 255, (E,) + cfg["frame"]).astype(np.uint8)
        state = pol.make_input_state(obs, np.ones(E, dtype=bool))
        for _ in range(8):      # both alternating state buffers reach their CUDA graph
            pol.actor_predict(state)
            state = pol.make_input_state(obs, np.zeros(E, dtype=bool))
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(50):
            pol.actor_predict(state)
            state = pol.make_input_state(obs, np.zeros(E, dtype=bool))
        torch.cuda.synchronize(device)
        acting = {"us_per_vector_step": 1e6 * (time.perf_counter() - t0) / 50, "envs": E,
                  "what": "DevicePolicy.actor_predict + make_input_state for %d envs, host observations in, "
                          "actions + q-values out (wall clock)" % E}
    except Exception as ex:  # noqa: BLE001
        acting = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}

    vals = torch.tensor([ms, e2e_s, ms_long], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms, e2e_s, ms_long = float(vals[0]), float(vals[1]), float(vals[2])
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    hbm, bf16_burst, bf16_sus, which = peaks()
    gather_traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        gather_traffic = json.load(open(tpath)).get("gather")
    tf32_peak, tf32_sus = bf16_burst / 2.0, bf16_sus / 2.0
    S, n = cfg["T"] + cfg["P"], cfg["n"]
    state_bytes = FRAME_BYTES + 2 * cfg["units"] * 4 + 4
    gather_bytes = 2 * (S + n) * B * state_bytes        # read once + write once
    dtype_of = {"tf32": "tf32 multiply on round-to-nearest operands / f32 accumulate",
                "tf32_trunc": "tf32 multiply (operands truncated by the tensor core) / f32 accumulate", "fp32": "f32"}
    out = {
        "metric": "learner updates/sec (32x20-step seq batches, 1M prioritized replay)",
        "value": world * args.steps / (ms / 1e3), "unit": "updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype_of[cfg["gemm"]],
        "data": "synthetic",
        "config": static_config(world, total_size),
        "arm": {"gemm": cfg["gemm"], "fill_s": round(fill_s, 1),
                "schedule": "update replayed from CUDA graphs; replay on its own stream (priority write-back, next "
                            "draw and gather overlap the backward pass); weight gradients on a second graph branch",
                "parallelism": "dp%d: replay sharded by env (%d envs, %d transitions per GPU), in-library NCCL "
                               "all-reduce of the flat gradient (rt_learner_step_dp)" % (world, cfg["envs"], cfg["size"]),
                "updates_per_s_counts": "32-sequence batch-equivalents: world x optimizer steps/s"},
        "optimizer_steps_per_s": args.steps / (ms / 1e3),
        "value_long": {"value": world * long_steps / (ms_long / 1e3), "steps": long_steps,
                       "ms_per_step": ms_long / long_steps},
        "e2e": {"value": world * args.steps / e2e_s, "unit": "updates/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": dict(dominant_roofline(top_shape, top_fl, top_ms, top_n, gemm_ms, prof_steps, tf32_peak, which,
                                           tf32_sus),
                         event_bracket_overhead_us=1e3 * ovh_ms, top_shapes=top_shapes),
        "roofline_gemm_family": {"kernel": "tcgen05 GEMM family (k_gemm_tc_p / k_gemm_tc / k_conv_tc_p / k_convdw_tc / "
                               "k_convdx_tc): all GEMM-shaped launches of the update", "bound": "tensor",
                     "achieved": gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None,
                     "peak": tf32_peak, "unit": "TFLOP/s",
                     "peak_source": which + " cuBLAS bf16 burst / 2 (TF32 multiplies at half the bf16 rate)",
                     "frac": (gemm_flops / (gemm_ms * 1e-3) / 1e12 / tf32_peak) if gemm_ms > 0 else None,
                     "traffic": None, "launches_timed": int(gemm_n), "profiled_steps": prof_steps,
                     "gflop_per_update": gemm_flops / prof_steps / 1e9,
                     "gemm_ms_per_update": gemm_ms / prof_steps},
        "roofline_gather": {"kernel": "k_gather_bulk (cp.async.bulk global->shared->global ring), timed alone", "bound": "hbm",
                     "achieved": gather_bytes / (gather_ms_total / max(gather_n, 1) * 1e-3) / 1e9,
                     "peak": hbm, "peak_source": which, "unit": "GB/s",
                     "frac": gather_bytes / (gather_ms_total / max(gather_n, 1) * 1e-3) / 1e9 / hbm,
                     "traffic": gather_traffic, "launches_timed": int(gather_n),
                     "us_per_launch": 1e3 * gather_ms_total / max(gather_n, 1),
                     "algorithmic_bytes": int(gather_bytes),
                     "draw_call_ms_host": draw_ms},
        "acting": acting,
        "last_stats": stats,
    }
    if world == 1 and not args.no_side_lines:
        # the other GEMM precisions on the same replay (short runs): `value` above is the mode that holds the
        # 1e-4 parity bar (tests/test_learner_gpu.py), value_fast drops the operand rounding, value_fp32 is the
        # CUDA-core fp32 path
        learner.close()
        for key, mode, k in (("value_fast", "tf32_trunc", args.steps), ("value_fp32", "fp32", max(3, args.steps // 4))):
            if mode == cfg["gemm"]:
                continue
            try:
                L2 = build_learner(cfg, device, mode)
                ms2 = timed_updates(hist, L2, B, 1, k, 3, barrier)
                out[key] = {"value": k / (ms2 / 1e3), "unit": "updates/s", "gemm": mode, "dtype": dtype_of[mode],
                            "steps": k, "ms_per_step": ms2 / k}
                L2.close()
            except Exception as ex:  # noqa: BLE001
                out[key] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
        hist.close()
        del hist
        torch.cuda.empty_cache()
        side_n = args.side_size
        c3 = dict(cfg, size=side_n, P=40, n=5)
        out["config3_burnin40"] = side_config_line(
            "BASELINE configs[2]: IQN+LSTM R2D2, T=20 burn-in 40 n=5, weighted PER, %d-transition replay" % side_n,
            c3, device, args.steps, 3, cfg["gemm"])
        c2 = dict(cfg, size=side_n, P=0, n=3, T=1, units=0)
        out["config2_cnn_iqn"] = side_config_line(
            "BASELINE configs[1]: Atari 84x84x4 IQN (nature-CNN-FC512, no LSTM), T=1 n=3 B=32 Nq=32, PER, "
            "%d-transition replay" % side_n, c2, device, args.steps, 3, cfg["gemm"])
    if world == 1 and not args.no_cpu_baseline:
        th = ref_threads()
        res = ref_subprocess("cpu", th, 5, 1, args.ref_size, also_1thread=2)
        out["cpu_baseline"] = cpu_baseline_from(
            res, "the unmodified reference (baseline/_ref) through IQN.train(), torch CPU fp32, %d intra-op threads" % th)
        res = ref_subprocess("cuda", 1, args.steps, warmup, args.ref_size)
        if "unavailable" in res:
            out["cuda_torch_baseline"] = {"value": None, "unavailable": res["unavailable"]}
        else:
            out["cuda_torch_baseline"] = {
                "value": res["updates_per_s"], "unit": "updates/s", "ms_per_step": res["ms_per_update"],
                "steps": res["updates_timed"], "kind": "reference",
                "what": "the unmodified reference on this GPU: policy on cuda, StateStore('cuda'), its own Python "
                        "PER buffer on the host (rltime/policies/torch/torch_policy.py:44-59, "
                        "rltime/general/backend.py:136-153)",
                "replay_transitions": res["replay_transitions"],
                "reference_timings_mean_ms": res.get("timings_mean_ms"), "torch": res.get("torch")}
    print(json.dumps(out))


# ----------------------------------------------------------------------- reference arm
def run_reference(args):
    """The unmodified reference on the host cores (baseline/ref_arm.py); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_arm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = dict(CFG)
    size = args.size or args.ref_size
    th = args.ref_threads or ref_threads()
    warmup = max(args.warmup, 1)
    line = {
        "impl": "reference",
        "metric": "learner updates/sec (32x20-step seq batches, 1M prioritized replay)",
        "unit": "updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": static_config(world, size),
    }
    if not ref_arm.available():
        line["unavailable"] = "baseline/_ref missing (run baseline/install_ref.sh in the build container)"
        print(json.dumps(line))
        return
    k1 = max(2, min(args.steps, 5))
    res = ref_arm.run(cfg, "cpu", th, args.steps, warmup, size, extra_phases=[(1, 1, k1)])
    cb = cpu_baseline_from(res, "the unmodified reference (baseline/_ref) through IQN.train(), torch CPU fp32, "
                                "%d intra-op threads" % th)
    line.update({"value": res["updates_per_s"], "ms_per_step": res["ms_per_update"], "cpu_baseline": cb,
                 "e2e": {"value": res["updates_per_s"], "unit": "updates/s", "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=0, help="override replay capacity (debug)")
    ap.add_argument("--ref-size", type=int, default=1_000_000,
                    help="replay transitions of the reference arms (the reference's own Python buffer)")
    ap.add_argument("--ref-threads", type=int, default=0,
                    help="torch intra-op threads for the CPU arm (0 = min(cores, 32); 128 threads are "
                         "slower than 32 on these shapes)")
    ap.add_argument("--long-steps", type=int, default=500)
    ap.add_argument("--side-size", type=int, default=262_144, help="replay capacity of the config-2 / config-3 side lines")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip cpu_baseline and cuda_torch_baseline")
    ap.add_argument("--no-side-lines", action="store_true", help="skip value_fast / value_fp32 / config-2 / config-3")
    ap.add_argument("--gemm", default=os.environ.get("RT_BENCH_GEMM", "tf32"), choices=["tf32", "tf32_trunc", "fp32"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
