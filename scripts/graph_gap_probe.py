"""Probe: spacing of dependent short kernels launched eagerly vs replayed from a CUDA graph."""
import torch

x = torch.randn(32 * 2048, device="cuda")
y = torch.empty_like(x)


def chain(n):
    for _ in range(n):
        torch.mul(x, 1.0001, out=y)
        torch.add(y, 0.5, out=x)


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


N = 100
eager = timeit(lambda: chain(N))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    chain(3)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        chain(N)
graph = timeit(g.replay)
print("2x%d dependent elementwise kernels on 64K floats: eager %.1f us (%.2f us/kernel), graph %.1f us (%.2f us/kernel)"
      % (N, eager, eager / (2 * N), graph, graph / (2 * N)))
