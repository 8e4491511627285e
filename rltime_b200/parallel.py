"""Data-parallel plumbing: one process per GPU, replay sharded by env id, one
gradient all-reduce per update.

The reference has no multi-GPU path; trajectories are independent per env_id
(rltime/history/history.py:50-55), so each rank owns the transitions of its envs, draws its
own B sequences and trains a replica.  Replay data never crosses GPUs; the only exchange is
the sum of the flat fp32 gradient buffer (8.1 M floats = 32.5 MB at config 3) over
NCCL / NVLink, after which every rank applies the identical clip + Adam with
grad_scale = 1 / world_size (the mean over the global batch of equal-sized shards).
"""
import os


def env_rank(env_id, world_size):
    """Owner rank of an env (contiguous ids interleave across ranks)."""
    return int(env_id) % world_size


def shard_samples(samples, rank, world_size):
    """Keeps the acting samples whose env belongs to this rank."""
    return [s for s in samples if env_rank(s["env_id"], world_size) == rank]


def init_process_group(backend=None):
    """Initialises torch.distributed from the torchrun environment (RANK / WORLD_SIZE /
    MASTER_ADDR / MASTER_PORT).  Returns (rank, world_size)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend, **kw)
    return rank, world


def allreduce_sum_(flat_grad):
    """In-place sum of the flat gradient over all ranks (no-op when not distributed)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad


_comm_streams = {}


def init_library_comm(learner):
    """Creates the learner's own NCCL communicator (rt_comm_init): rank 0 draws the 128-byte id, the
    process group carries it to the other ranks -- the only thing the host language has to do."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == 0:
        ident = torch.tensor(list(learner.comm_unique_id()), dtype=torch.uint8)
    else:
        ident = torch.zeros(128, dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        ident = ident.to(learner.device)
    dist.broadcast(ident, src=0)
    learner.comm_init(bytes(ident.cpu().tolist()), rank, world)


def data_parallel_step(learner, batch, world_size, taus=None, overlap=True):
    """compute local gradients -> all-reduce (sum) -> clip + Adam on the mean gradient.

    The gradient is reduced in two buckets: everything but the conv parameters (final before the
    conv backward starts, rt_learner_wait_late_grads) on a communication stream while the conv
    backward still runs, then the small conv bucket on the caller's stream."""
    import torch
    if getattr(learner, "world", 1) == world_size and world_size > 1 and os.environ.get("RT_DP_LIB", "1") != "0":
        learner.step_dp(batch, taus)      # the whole exchange inside the library (rt_learner_step_dp)
        return
    if os.environ.get("RT_DP_OVERLAP") == "0":
        overlap = False
    learner.compute_grads(batch, taus)
    if world_size > 1:
        g = learner.flat()
        first, count = learner.wait_late_grads()
        if overlap and 0 < first < g.numel():
            import ctypes as C
            comm = _comm_streams.get(learner.device)
            if comm is None:
                comm = _comm_streams[learner.device] = torch.cuda.Stream(learner.device)
            learner.wait_late_grads(C.c_void_p(comm.cuda_stream))
            with torch.cuda.stream(comm):
                allreduce_sum_(g[first:first + count])
            allreduce_sum_(g[:first])
            torch.cuda.current_stream(learner.device).wait_stream(comm)
        else:
            allreduce_sum_(g)
    learner.apply_grads(1.0 / world_size)


def broadcast_params_(learner, src=0):
    """Makes every replica start from rank `src`'s weights (online and target nets)."""
    import torch.distributed as dist
    from . import _lib
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if getattr(learner, "world", 1) == dist.get_world_size():
            learner.comm_broadcast_params(src)
            return
        for which in (_lib.RT_BUF_ONLINE, _lib.RT_BUF_TARGET):
            dist.broadcast(learner.flat(which), src=src)
        learner.params_changed()
