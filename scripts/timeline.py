"""Prints clock64 phase timelines of the persistent LSTM kernel (RT_DEBUG_TIMELINE=1)."""
import os
import sys

os.environ["RT_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rltime_b200 import _lib  # noqa: E402
from rltime_b200.init import init_params  # noqa: E402
from rltime_b200.learner import DeviceLearner, batch_from_tensors  # noqa: E402

B, T, n = 32, 20, 2
L = DeviceLearner((4, 84, 84), [(32, 8, 4), (64, 4, 2), (64, 3, 1)], 512, 512, 6, 32, 64, True, mbatch=B,
                  nstep_train=T, nstep_target=n, double_q=True, rnn_bootstrap=True, clip_grad=40.0)
L.load_state_dict(init_params(L.param_info, 512, 1), 0)
L.load_state_dict(init_params(L.param_info, 512, 2), 1)
dev = "cuda"
S = T
b, keep = batch_from_tensors(
    torch.randint(0, 255, (S + n, B, 4, 84, 84), dtype=torch.uint8, device=dev),
    torch.randn(S + n, B, 512, device=dev), torch.randn(S + n, B, 512, device=dev),
    torch.zeros(S + n, B, device=dev), torch.randn(S, B, device=dev, dtype=torch.float64),
    torch.full((S, B), n, device=dev, dtype=torch.int64), torch.ones(S, B, device=dev, dtype=torch.float64),
    torch.randint(0, 6, (S, B), device=dev), torch.ones(S, B, device=dev, dtype=torch.float64), n)
for _ in range(3):
    L.step(b)
torch.cuda.synchronize()
import ctypes as C
p, cnt = C.c_void_p(), C.c_int64()
_lib.check(L._lib.rt_learner_debug_tensor(L._h, b"lstm_dbg", C.byref(p), C.byref(cnt)))
dd = _lib.as_tensor(p.value, (256, 8), "<i8", L.device).cpu().numpy()
for grp in range(2):
    d = dd[64 * grp:64 * grp + T]
    print("tensor-core LSTM kernel, first CTA of weight group %d (%s), cycles per phase" % (
        grp, "target net, 1 sequence" if grp == 0 else "online net, 2 sequences"))
    prev = d[0, 0]
    for t in range(T):
        # stamps: 0 producer saw every CTA's h_{t-1}, 1 last MMA issued, 2 accumulator ready, 5 accumulator
        # in registers, 6 cell math + h stores issued, 4 h_t published, 3 remaining stores issued
        if t in (0, 1, 2, 10, 18, 19):
            print("%2d  publish->seen %6d | TMA+MMA issue %6d | retire %4d | tmem ld %5d | cell+h stores %5d | release %5d | other stores %5d | step %6d" % (
                t, d[t, 0] - prev, d[t, 1] - d[t, 0], d[t, 2] - d[t, 1], d[t, 5] - d[t, 2], d[t, 6] - d[t, 5],
                d[t, 4] - d[t, 6], d[t, 3] - d[t, 4], d[t, 4] - prev))
        prev = d[t, 4]
    print("whole recurrence: %d cycles" % (d[T - 1, 3] - d[0, 0]))
