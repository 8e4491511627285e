"""tcgen05 TF32 GEMM (rt_gemm_tc.cuh) against numpy on operands that are exact in TF32
(multiples of 1/4 in [-2, 2]): every product and partial sum is exact in fp32, so the result
must match bit for bit whatever the operand layout (K-major / MN-major), tile tails or
split-K path."""
import ctypes as C

import numpy as np
import pytest

# (M, N, K, transA, transB, bias_relu)
CASES = [
    (256, 128, 64, 0, 1, 0), (128, 64, 32, 0, 1, 0), (128, 32, 256, 0, 1, 1),
    (640, 2048, 3136, 0, 1, 1), (1000, 512, 512, 0, 1, 0), (200, 96, 100, 0, 1, 0),
    (51200, 32, 256, 0, 1, 1), (300, 40, 576, 0, 1, 0), (32, 2048, 512, 0, 1, 0),
    # data-gradient form: B stored [K][N]
    (256, 128, 64, 0, 0, 0), (1000, 512, 512, 0, 0, 0), (640, 3136, 2048, 0, 0, 0),
    (200, 96, 100, 0, 0, 0), (32, 512, 2048, 0, 0, 0),
    # weight-gradient form: A stored [K][M], B stored [K][N]
    (128, 128, 64, 1, 0, 0), (512, 512, 2048, 1, 0, 0), (2048, 3136, 640, 1, 0, 0),
    (32, 256, 51200, 1, 0, 0), (64, 576, 6272, 1, 0, 0), (512, 64, 4096, 1, 0, 0),
    (200, 96, 100, 1, 0, 0),
    # > 148 tiles: the persistent, accumulator-double-buffered kernel (k_gemm_tc_p)
    (20480, 512, 512, 0, 1, 1), (20480, 512, 512, 0, 0, 0), (512, 512, 20480, 1, 0, 0),
    (20000, 500, 96, 0, 1, 1), (2048, 3136, 640, 1, 0, 0), (19000, 64, 576, 0, 1, 1),
    # wide K-major products with >= 148 128x256 tiles: the CTA-pair kernel (k_gemm_tc_pair, cta_group::2),
    # full size, a ragged last row tile whose second half is empty, and an odd number of 128-row tiles
    (20480, 1024, 512, 0, 1, 1), (19000, 512, 256, 0, 1, 1), (19328, 512, 512, 0, 1, 0),
]


def run_case(lib, mode, M, N, K, tA, tB, br, seed=0):
    from rltime_b200 import _lib
    rs = np.random.RandomState(seed)
    A = (rs.randint(-8, 9, (M, K)) / 4.0).astype(np.float32)
    B = (rs.randint(-8, 9, (K, N)) / 4.0).astype(np.float32)
    bias = (rs.randint(-8, 9, N) / 4.0).astype(np.float32) if br else None
    want = A.astype(np.float64) @ B.astype(np.float64)
    if br:
        want = np.maximum(want + bias, 0)
    A_st = np.ascontiguousarray(A.T if tA else A)
    B_st = np.ascontiguousarray(B.T if tB else B)
    out = np.empty((M, N), dtype=np.float32)
    rc = lib.rt_gemm_test(mode, M, N, K, tA, tB, A_st.ctypes.data, B_st.ctypes.data,
                          bias.ctypes.data if br else None, 1 if br else 0, out.ctypes.data, 0)
    if rc != 0:
        return "rc=%d %s" % (rc, lib.rt_last_error().decode())
    bad = out.astype(np.float64) != want
    if bad.any():
        idx = np.argwhere(bad)
        return "%d/%d wrong, first at %s got %g want %g, max|d|=%g; wrong rows %s cols %s" % (
            bad.sum(), bad.size, tuple(idx[0]), out[tuple(idx[0])], want[tuple(idx[0])],
            np.abs(out - want).max(), sorted(set(idx[:, 0] // 32))[:8], sorted(set(idx[:, 1] // 32))[:8])
    return None


@pytest.mark.gpu
def test_simt_gemm_exact():
    from rltime_b200 import _lib
    lib = _lib.load()
    errs = []
    for case in CASES[:6] + CASES[9:12] + CASES[14:17]:
        e = run_case(lib, 0, *case)
        if e:
            errs.append("%s: %s" % (case, e))
    assert not errs, "\n".join(errs)


@pytest.mark.gpu
def test_tcgen05_gemm_exact():
    from rltime_b200 import _lib
    lib = _lib.load()
    errs = []
    for case in CASES:
        e = run_case(lib, 1, *case)
        print(case, "OK" if e is None else e)
        if e:
            errs.append("%s: %s" % (case, e))
    assert not errs, "\n".join(errs)
