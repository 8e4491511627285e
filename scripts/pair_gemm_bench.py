"""Times the hidden-layer product (20480 x 1024 x 512, K-major operands) with the single-CTA persistent kernel
and with the CTA-pair kernel (RT_TC_PAIR=0 / 1; read at call time by rt_gemm_bench)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rltime_b200 import _lib  # noqa: E402

lib = _lib.load()
for M, N, K in ((20480, 1024, 512), (20480, 512, 512), (40960, 1024, 512)):
    for pair in ("0", "1"):
        os.environ["RT_TC_PAIR"] = pair
        us = C.c_double()
        rc = lib.rt_gemm_bench(1, M, N, K, 0, 1, 0, 0, 50, C.byref(us), 0)
        if rc:
            print(M, N, K, "pair", pair, "rc", rc, lib.rt_last_error().decode())
        else:
            print("M=%d N=%d K=%d pair=%s  %.1f us  %.0f TFLOP/s" % (M, N, K, pair, us.value, 2.0 * M * N * K / us.value / 1e6))
