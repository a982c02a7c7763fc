"""Experiment: how much of the training step is launch overhead?  Times the config-2 step eagerly and replayed from a
CUDA graph captured around ``NavCMTAgent.train_step`` (same kernels, same buffers; the dropout seed and the optimiser
step count are frozen inside the graph, so this is a timing probe, not a training mode).
Usage (GPU box): python tools/graph_probe.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench_train import TrainWorkload

dev = torch.device("cuda", 0)
wl = TrainWorkload(0, 1)
wl.setup_gpu(dev)
for _ in range(3):
    wl.step()
torch.cuda.synchronize()


def timed(fn, n=10):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    t0 = time.perf_counter()
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    return sum(a.elapsed_time(b) for a, b in evs) / n, wall


e_ms, e_wall = timed(wl.step)
print(f"eager : {e_ms:.2f} ms/step (device), {e_wall:.2f} ms wall")
t0 = time.perf_counter()
wl.step()
cpu_launch = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize()
print(f"host time to ENQUEUE one step: {cpu_launch:.2f} ms")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
try:
    with torch.cuda.stream(s):
        wl.step()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            wl.step()
    torch.cuda.synchronize()
    g_ms, g_wall = timed(g.replay)
    print(f"graph : {g_ms:.2f} ms/step (device), {g_wall:.2f} ms wall   -> launch overhead in the eager step ~ {e_ms - g_ms:.2f} ms")
except Exception as ex:                                   # capture can fail on an unsupported call: say which
    print("graph capture failed:", repr(ex)[:400])
