"""Multi-GPU check (run under torchrun, one rank per GPU): after data-parallel updates every
replica holds bit-identical weights, and the all-reduced gradient equals the sum of the ranks'
local gradients."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rltime_b200 import _lib, parallel  # noqa: E402
from rltime_b200.init import init_params  # noqa: E402
from rltime_b200.learner import DeviceLearner, batch_from_tensors  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
rank, world = parallel.init_process_group("nccl")
dev = torch.device("cuda", local)
B, T, n, U = 8, 6, 2, 64
L = DeviceLearner((4, 84, 84), [(32, 8, 4), (64, 4, 2), (64, 3, 1)], U, 128, 6, 8, 64, True, mbatch=B,
                  nstep_train=T, nstep_target=n, double_q=True, rnn_bootstrap=True, clip_grad=40.0,
                  device=dev)
L.load_state_dict(init_params(L.param_info, U, 1), 0)
L.load_state_dict(init_params(L.param_info, U, 2), 1)
parallel.broadcast_params_(L)
g = torch.Generator(device=dev).manual_seed(100 + rank)     # different data per rank
S = T
for step in range(3):
    b, keep = batch_from_tensors(
        torch.randint(0, 255, (S + n, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
        torch.randn(S + n, B, U, device=dev, generator=g), torch.randn(S + n, B, U, device=dev, generator=g),
        torch.zeros(S + n, B, device=dev), torch.randn(S, B, device=dev, generator=g, dtype=torch.float64),
        torch.full((S, B), n, device=dev, dtype=torch.int64), torch.ones(S, B, device=dev, dtype=torch.float64),
        torch.randint(0, 6, (S, B), device=dev, generator=g), torch.ones(S, B, device=dev, dtype=torch.float64), n)
    L.compute_grads(b)
    local_grad = L.flat().clone()
    gathered = [torch.empty_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    parallel.allreduce_sum_(L.flat())
    want = torch.stack(gathered).sum(0)
    err = (L.flat() - want).abs().max().item() / (want.abs().max().item() + 1e-12)
    assert err < 1e-5, "all-reduced gradient differs from the sum of local gradients: %g" % err
    L.apply_grads(1.0 / world)
# overlapped (two-bucket) reduction == single all-reduce, bit for bit
L2 = DeviceLearner((4, 84, 84), [(32, 8, 4), (64, 4, 2), (64, 3, 1)], U, 128, 6, 8, 64, True, mbatch=B,
                   nstep_train=T, nstep_target=n, double_q=True, rnn_bootstrap=True, clip_grad=40.0,
                   device=dev)
L2.load_training_state(L.training_state())
# third replica: the whole exchange inside the library (rt_comm_init + rt_learner_step_dp, own NCCL communicator)
L3 = DeviceLearner((4, 84, 84), [(32, 8, 4), (64, 4, 2), (64, 3, 1)], U, 128, 6, 8, 64, True, mbatch=B,
                   nstep_train=T, nstep_target=n, double_q=True, rnn_bootstrap=True, clip_grad=40.0,
                   device=dev)
L3.load_training_state(L.training_state())
parallel.init_library_comm(L3)
os.environ["RT_DP_LIB"] = "1"
for step in range(4):
    b, keep = batch_from_tensors(
        torch.randint(0, 255, (S + n, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
        torch.randn(S + n, B, U, device=dev, generator=g), torch.randn(S + n, B, U, device=dev, generator=g),
        torch.zeros(S + n, B, device=dev), torch.randn(S, B, device=dev, generator=g, dtype=torch.float64),
        torch.full((S, B), n, device=dev, dtype=torch.int64), torch.ones(S, B, device=dev, dtype=torch.float64),
        torch.randint(0, 6, (S, B), device=dev, generator=g), torch.ones(S, B, device=dev, dtype=torch.float64), n)
    taus = [torch.rand(T * B * 8, generator=torch.Generator().manual_seed(1000 * step + 10 * rank + k)) for k in range(3)]
    parallel.data_parallel_step(L, b, world, taus=taus, overlap=True)
    parallel.data_parallel_step(L2, b, world, taus=taus, overlap=False)
    parallel.data_parallel_step(L3, b, world, taus=taus)          # -> L3.step_dp
torch.cuda.synchronize()
d12 = (L.flat(_lib.RT_BUF_ONLINE) - L2.flat(_lib.RT_BUF_ONLINE)).abs().max().item()
d13 = (L.flat(_lib.RT_BUF_ONLINE) - L3.flat(_lib.RT_BUF_ONLINE)).abs().max().item()
assert d13 == 0.0 if world == 2 else d13 < 1e-6, "in-library data-parallel step changed the weights: %g" % d13
# two addends commute exactly; with more ranks NCCL's reduction order depends on the message size
assert d12 == 0.0 if world == 2 else d12 < 1e-6, "overlapped all-reduce changed the weights: %g" % d12
w = L.flat(_lib.RT_BUF_ONLINE).clone()
lo, hi = w.clone(), w.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
diff = (hi - lo).abs().max().item()
st = L.stats()
if rank == 0:
    print("dist_check world=%d: replicas identical (max spread %g), grad-sum rel err %.2e, overlapped == plain "
          "all-reduce (max diff %g), rt_learner_step_dp == torch.distributed path (max diff %g), qloss %.5f "
          "grad_norm %.5f" % (world, diff, err, d12, d13, st["qloss"], st["grad_norm"]))
assert diff == 0.0, "replicas diverged: %g" % diff
dist.barrier()
dist.destroy_process_group()
