"""The one-launch BPTT recurrence (rt_bptt.cuh, the default since round 2) against the stepwise
recurrence it replaced (RT_BPTT_PERSISTENT=0), at the full config-3 LSTM width."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


def _grads(env):
    import torch
    from rltime_b200 import _lib
    from rltime_b200.init import init_params
    from rltime_b200.learner import DeviceLearner, batch_from_tensors
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        B, T, n, U = 32, 6, 2, 512
        L = DeviceLearner((4, 84, 84), [(32, 8, 4), (64, 4, 2), (64, 3, 1)], U, 512, 6, 8, 64, True, mbatch=B,
                          nstep_train=T, nstep_target=n, double_q=True, rnn_bootstrap=True, clip_grad=40.0, seed=3)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    L.load_state_dict(init_params(L.param_info, U, 1), 0)
    L.load_state_dict(init_params(L.param_info, U, 2), 1)
    g = torch.Generator(device="cuda").manual_seed(5)
    dev = "cuda"
    b, keep = batch_from_tensors(
        torch.randint(0, 255, (T + n, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
        torch.randn(T + n, B, U, device=dev, generator=g), torch.randn(T + n, B, U, device=dev, generator=g),
        (torch.rand(T + n, B, device=dev, generator=g) < 0.1).float(),
        torch.randn(T, B, device=dev, generator=g, dtype=torch.float64),
        torch.full((T, B), n, device=dev, dtype=torch.int64), torch.ones(T, B, device=dev, dtype=torch.float64),
        torch.randint(0, 6, (T, B), device=dev, generator=g), torch.ones(T, B, device=dev, dtype=torch.float64), n)
    taus = [torch.rand(T * B * 8, generator=torch.Generator().manual_seed(k)) for k in range(3)]
    L.compute_grads(b, taus)
    torch.cuda.synchronize()
    out = L.flat(_lib.RT_BUF_GRAD).cpu().numpy().copy()
    L.close()
    return out


def test_persistent_bptt_matches_stepwise_bptt():
    """rt_bptt.cuh (RT_BPTT_PERSISTENT=1): same gradients as the stepwise recurrence (both TF32
    products with fp32 accumulation; only the summation order over k differs)."""
    ref = _grads({"RT_BPTT_PERSISTENT": "0", "RT_GRAPHS": "0"})
    got = _grads({"RT_BPTT_PERSISTENT": "1", "RT_GRAPHS": "0"})
    scale = np.abs(ref).max()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-4 * scale)
