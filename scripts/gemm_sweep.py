"""Sweeps tcgen05 GEMM tile configurations on the learner's main shapes (GPU box)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rltime_b200 import _lib  # noqa: E402

lib = _lib.load()
SHAPES = [("fc fwd  NT", 20480, 512, 512, 0, 1), ("fc dX   NN", 20480, 512, 512, 0, 0),
          ("fc dW   TN", 512, 512, 20480, 1, 0), ("ih fwd  NT", 640, 2048, 3136, 0, 1),
          ("ih dW   TN", 2048, 3136, 640, 1, 0), ("q fwd   NT", 20480, 512, 64, 0, 1)]
for name, M, N, K, tA, tB in SHAPES:
    flops = 2.0 * M * N * K
    out = []
    for pers in (1, 0):
        os.environ["RT_TC_PERSISTENT"] = str(pers)
        for bn in (64, 128, 256):
            for st in ((0,) if pers else (3, 4)):
                us = C.c_double()
                rc = lib.rt_gemm_bench(1, M, N, K, tA, tB, bn, st, 20, C.byref(us), 0)
                tag = "P/BN%d" % bn if pers else "BN%d/S%d" % (bn, st)
                out.append("%s %6.1fus %4.0fTF" % (tag, us.value, flops / us.value / 1e6) if rc == 0
                           else "%s ERR" % tag)
    os.environ["RT_TC_PERSISTENT"] = "1"
    us = C.c_double()
    lib.rt_gemm_bench(0, M, N, K, tA, tB, 0, 0, 5, C.byref(us), 0)
    print("%-12s M=%d N=%d K=%d | simt %.0fus | %s" % (name, M, N, K, us.value, " | ".join(out)))
