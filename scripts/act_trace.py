"""GPU-side cost of one acting step (rt_learner_act at E envs, config-3 model): CUDA-event time of the
whole step (H2D of the observations .. D2H of [q | h | c]) next to the wall clock, with and without the
CUDA graph.  Run under `ncu --metrics gpu__time_duration.sum` with RT_ACT_GRAPH=0 for the per-kernel list."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rltime_b200.learner import DeviceLearner          # noqa: E402
from rltime_b200.training import DevicePolicy          # noqa: E402

E = int(os.environ.get("ACT_ENVS", "32"))
calls = int(os.environ.get("ACT_CALLS", "40"))
L = DeviceLearner((4, 84, 84), [(32, 8, 4), (64, 4, 2), (64, 3, 1)], 512, 512, 18, 32, 64, True, mbatch=32,
                  nstep_train=20, nstep_target=3, double_q=True, rnn_bootstrap=True, gemm="tf32")
pol = DevicePolicy(L, 18)
obs = np.random.RandomState(0).randint(0, 255, (E, 4, 84, 84)).astype(np.uint8)
state = pol.make_input_state(obs, np.ones(E, dtype=bool))
for _ in range(8):
    pol.actor_predict(state)
    state = pol.make_input_state(obs, np.zeros(E, dtype=bool))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gpu_us, t0 = 0.0, time.perf_counter()
for _ in range(calls):
    e0.record()
    pol.actor_predict(state)
    e1.record()
    e1.synchronize()
    gpu_us += 1e3 * e0.elapsed_time(e1)
    state = pol.make_input_state(obs, np.zeros(E, dtype=bool))
wall_us = 1e6 * (time.perf_counter() - t0)
print("envs %d  graph %s  event time %.1f us / step   wall %.1f us / step" %
      (E, os.environ.get("RT_ACT_GRAPH", "1"), gpu_us / calls, wall_us / calls))
L.close()
