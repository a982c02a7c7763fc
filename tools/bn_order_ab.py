"""Experiment: traversal order of the elementwise BatchNorm passes (avdn_bn_set_order, csrc/trunk.cu) inside the
config-2 training step.  One process, one box: the masks are timed interleaved, twice, so that clock drift under the
power cap shows up as the spread between the two rounds rather than as a difference between masks.
Usage (GPU box): python tools/bn_order_ab.py [steps] [rounds] [mask,mask,...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench_train import TrainWorkload          # noqa: E402
from avdn_b200 import _lib                     # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
MASKS = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 8, 9, 11, 13, 15, 10, 12]
dev = torch.device("cuda", 0)
wl = TrainWorkload(0, 1)
wl.setup_gpu(dev)
h = _lib.lib()
for _ in range(4):
    wl.step()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(n):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        flush.zero_()
        a.record(); wl.step(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2], t[0]


def bn_ms():
    agg = wl.profile_step()
    out = {}
    for name, (cnt, ms, _fl) in agg.items():
        if "bn_" in name:
            out[name] = round(ms, 3)
    out["total"] = round(sum(out.values()), 3)
    return out


res = {m: [] for m in MASKS}
for rnd in range(ROUNDS):
    for m in MASKS:
        h.avdn_bn_set_order(m)
        wl.step()                               # one untimed step under the new order
        med, best = timed(steps)
        res[m].append((round(med, 3), round(best, 3)))
        print(f"round {rnd} order {m:2d}: median {med:.3f} ms  best {best:.3f} ms", flush=True)
prof = {}
for m in MASKS:
    h.avdn_bn_set_order(m)
    wl.step()
    prof[m] = bn_ms()
    print(f"order {m:2d} BatchNorm calls (CUDA events, one instrumented step): {prof[m]}", flush=True)
h.avdn_bn_set_order(0)
print(json.dumps({"steps": steps, "median_best_ms": {str(k): v for k, v in res.items()},
                  "bn_calls_ms": {str(k): v for k, v in prof.items()}}))
