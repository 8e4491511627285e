import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name)) as d:
        return {k: d[k] for k in d.files}


def assert_trace_equal(got, want, float_exact=True, rtol=0.0, skip=()):
    """Field-by-field comparison of two scenario traces (oracle/scenario.py)."""
    assert int(got["num_calls"]) == int(want["num_calls"])
    np.testing.assert_array_equal(got["none_calls"], want["none_calls"])
    for k in want:
        if k in skip or k in ("num_calls", "none_calls"):
            continue
        assert k in got, "missing field %s" % k
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, (k, g.shape, w.shape)
        if np.issubdtype(w.dtype, np.floating) and not float_exact:
            np.testing.assert_allclose(g.astype(np.float64), w.astype(np.float64),
                                       rtol=rtol, atol=0, err_msg=k)
        else:
            # integer / byte / index fields, and fp64 fields pinned bit-exact
            if not np.array_equal(g.astype(w.dtype), w):
                bad = np.argwhere(g.astype(w.dtype) != w)
                raise AssertionError("field %s differs at %d positions, first %s: got %r want %r"
                                     % (k, len(bad), bad[0], g[tuple(bad[0])], w[tuple(bad[0])]))
