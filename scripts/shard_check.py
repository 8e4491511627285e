"""Sharded prioritized replay check (SURVEY.md 8e "sampling parity across shards"), one rank per GPU under
torchrun.  Envs are dealt to ranks by env_id % world (rltime_b200.parallel.env_rank).
 (i)  every shard is bit-identical -- sampled prioritization indices, loss indices, returns, every state
      byte -- to a single-buffer instance (the CPU oracle restatement of the reference buffer) fed only that
      shard's envs, over a scripted run with priority write-backs;
 (ii) the importance weights of every rank are normalised by the maximum over ALL shards
      (prioritized_replay_history.py:347-354 applied to the union): weights * local_max / global_max.
TEST INFRASTRUCTURE (imports oracle/)."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import replay_oracle as ro  # noqa: E402
from oracle import scenario as sc  # noqa: E402
from rltime_b200 import parallel  # noqa: E402
from rltime_b200.history import DevicePrioritizedReplayHistoryBuffer  # noqa: E402
from rltime_b200.synthetic import SyntheticStream  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
rank, world = parallel.init_process_group("nccl")
dev = torch.device("cuda", local)
E, T, P, n, B = 8, 6, 2, 2, 4
kw = dict(size=400, train_frequency=None, alpha=0.9, beta=0.6, nstep_target=n, nstep_train=T, prefix_steps=P)
dev_hist = DevicePrioritizedReplayHistoryBuffer(**kw, discount_function=sc.discount_function, max_envs=E, device=dev)
dev_hist.global_weight_max = lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX)
cpu_hist = ro.PrioritizedReplayOracle(**kw, discount_function=sc.discount_function)
stream = SyntheticStream(num_envs=E, frame_shape=(2, 5, 5), num_actions=3, lstm_units=4, seed=1,
                         done_mode="bernoulli", done_p=0.05, pool=64)
script = np.random.RandomState(7)
checked = 0
for it in range(120):
    for _ in range(script.randint(1, 5)):
        mine = parallel.shard_samples(stream.next_samples(), rank, world)
        dev_hist.update(mine)
        cpu_hist.update(mine)
    random.seed(1000 * it + rank)
    got = dev_hist.get_train_data(B, 0.3)
    random.seed(1000 * it + rank)
    want = cpu_hist.get_train_data(B, 0.3)
    # every rank must take part in the collective of a draw: shards fill at the same pace here
    assert (got is None) == (want is None)
    if got is None:
        continue
    fg, fw = sc.flatten_train_data(got), sc.flatten_train_data(want)
    for k in fw:
        if k == "extra_data/importance_weights":
            continue
        assert np.array_equal(fg[k], np.asarray(fw[k])), "rank %d iter %d field %s differs" % (rank, it, k)
    assert dev_hist.last_sampled_idxes == [int(i) for i in cpu_hist.last_sampled_idxes]
    # (ii) global normalisation: the oracle's weights are divided by ITS batch max w_max_local; with
    # w_raw = weights * w_max_local the global form is w_raw / max over ranks(w_max_local)
    w_local = np.asarray(fw["extra_data/importance_weights"], dtype=np.float64)
    wmax = torch.tensor([cpu_hist.last_weight_max], dtype=torch.float64, device=dev)
    gmax = wmax.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    want_w = w_local * float(wmax) / float(gmax)
    np.testing.assert_allclose(fg["extra_data/importance_weights"], want_w, rtol=1e-12, atol=0)
    li = fw["extra_data/loss_indices"][P:].reshape(-1, 2)
    losses = np.abs(script.randn(li.shape[0])).astype(np.float32).astype(np.float64)
    dev_hist.update_losses(li, losses)
    cpu_hist.update_losses(li, losses)
    checked += 1
tot = torch.tensor([checked], device=dev)
dist.all_reduce(tot)
if rank == 0:
    print("shard parity ok: world=%d, %d draws checked bit-exact per shard, weights normalised by the global max" % (
        world, int(tot)))
dist.barrier()
dist.destroy_process_group()
