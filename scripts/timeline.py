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
d = _lib.as_tensor(p.value, (256, 8), "<i8", L.device).cpu().numpy()[:T]
print("LSTM persistent kernel, CTA 0, cycles per phase")
prev = d[0, 0]
for t in range(T):
    print("%2d  xin-issue %5d | h-loads-landed %6d | staged+sync %6d | dots %6d | reduce+cell+stores %6d | "
          "grid-barrier %6d | step %6d" % (t, d[t, 0] - prev, d[t, 1] - d[t, 0], d[t, 2] - d[t, 1],
                                           d[t, 3] - d[t, 2], d[t, 4] - d[t, 3], d[t, 5] - d[t, 4], d[t, 5] - prev))
    prev = d[t, 5]
