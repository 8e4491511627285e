"""World-size-2 gloo test (CPU): the data-parallel recipe of rltime_b200/parallel.py — local
gradients of the per-shard mean loss, summed by all-reduce and scaled by 1/world — equals the
single-process gradient of the global-batch mean loss (oracle learner, torch fp32)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_case():
    from oracle import learner_oracle as lo
    spec = lo.ModelSpec((2, 12, 12), [(4, 4, 2)], 8, 8, 3, 4, 4, True)
    rs = np.random.RandomState(0)
    T, B, n = 3, 4, 1
    allx = torch.from_numpy(rs.randint(0, 256, (T + n, B, 2, 12, 12)).astype(np.uint8))
    allh = torch.from_numpy(rs.randn(T + n, B, 8).astype(np.float32))
    allc = torch.from_numpy(rs.randn(T + n, B, 8).astype(np.float32))
    alli = torch.from_numpy((rs.rand(T + n, B) < 0.2).astype(np.float32))
    targets = torch.from_numpy(rs.randn(T, B, 4).astype(np.float32))
    actions = torch.from_numpy(rs.randint(0, 3, (T, B)))
    weights = torch.from_numpy(rs.rand(T, B))
    taus = torch.rand(T, B, 4, generator=torch.Generator().manual_seed(1))
    return spec, (allx, allh, allc, alli, targets, actions, weights, taus), (T, B, n)


def _grads(spec, params, data, cols):
    """Gradient of the mean IQN loss over the batch columns `cols` (time-major rows)."""
    from oracle import learner_oracle as lo
    allx, allh, allc, alli, targets, actions, weights, taus = data
    T = targets.shape[0]
    sel = lambda x: x[:T][:, cols].reshape((-1,) + x.shape[2:])
    states = {"x": sel(allx), "layer1_state": {"hx": sel(allh), "cx": sel(allc), "initials": sel(alli)}}
    leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss, _, _ = lo.iqn_loss(spec, leaf, states, sel(targets), sel(actions), sel(weights), T,
                             sel(taus).reshape(-1))
    loss.backward()
    return torch.cat([leaf[k].grad.reshape(-1) for k in sorted(leaf)])


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    from rltime_b200 import parallel
    r, w = parallel.init_process_group("gloo")
    assert (r, w) == (rank, world)
    spec, data, (T, B, n) = _make_case()
    params = spec.init_params(3)
    cols = [b for b in range(B) if parallel.env_rank(b, world) == rank]   # shard by env/column
    g = _grads(spec, params, data, cols)
    parallel.allreduce_sum_(g)
    g *= 1.0 / world
    if rank == 0:
        torch.save(g, out)
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gradient_mean(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    sys.path.insert(0, ROOT)
    spec, data, (T, B, n) = _make_case()
    want = _grads(spec, spec.init_params(3), data, list(range(B)))
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-7)


def test_shard_helpers():
    from rltime_b200 import parallel
    samples = [{"env_id": e} for e in range(7)]
    parts = [parallel.shard_samples(samples, r, 3) for r in range(3)]
    assert sorted(s["env_id"] for p in parts for s in p) == list(range(7))
    assert [s["env_id"] for s in parts[1]] == [1, 4]


class _FakeLearner:
    """Host-side stand-in with the DeviceLearner surface data_parallel_step drives."""

    def __init__(self, rank, n=96, first=32):
        self.device = torch.device("cpu")
        self.g = torch.zeros(n)
        self.rank, self.first, self.applied = rank, first, []

    def compute_grads(self, batch, taus=None):
        self.g.copy_(torch.arange(self.g.numel(), dtype=torch.float32) * (self.rank + 1) + batch)

    def flat(self):
        return self.g

    def wait_late_grads(self, stream_ptr=None):
        return self.first, self.g.numel() - self.first

    def apply_grads(self, scale):
        self.applied.append((scale, self.g.clone()))


def _dp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from rltime_b200 import parallel
    parallel.init_process_group("gloo")
    L = _FakeLearner(rank)
    # CPU tensors: the single all-reduce path (the two-bucket path needs CUDA streams and is
    # checked on GPUs by scripts/dist_check.py)
    parallel.data_parallel_step(L, 0.5, world, overlap=False)
    if rank == 0:
        torch.save(L.applied, out)
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_data_parallel_step(tmp_path):
    out = str(tmp_path / "a.pt")
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    (scale, g), = torch.load(out)
    assert scale == 0.5
    want = torch.arange(96, dtype=torch.float32) * 3 + 1.0      # ranks contribute x1 and x2, batch term twice
    np.testing.assert_array_equal(g.numpy(), want.numpy())
