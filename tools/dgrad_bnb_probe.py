"""Timing of one data-gradient convolution with and without the fused BatchNorm-backward reductions
(avdn_gemm_core.bnb_*), plus the clock64 trace of epilogue thread 0 of CTA 0 for the fused launch.
Usage (GPU box): python tools/dgrad_bnb_probe.py [case ...]   cases: L13 L12 L14 L2 L3 L6"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {   # name: (H of the dgrad output, Cin (= output channels of the dgrad), Cout, k, stride, accumulate)
    "L13": (28, 256, 128, 1, 1, 1), "L12": (56, 128, 256, 3, 2, 0), "L14": (28, 128, 256, 3, 1, 0),
    "L2": (112, 64, 32, 1, 1, 1), "L3": (112, 32, 64, 3, 1, 0), "L6": (56, 128, 64, 1, 1, 1),
}
N = int(os.environ.get("PROBE_N", "640"))
trace = os.environ.get("PROBE_TRACE", "1") != "0"
buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
if trace:
    os.environ["AVDN_GEMM_DBG_BUF"] = str(buf.data_ptr())
from avdn_b200 import gemm as G


def run(name):
    H, Cin, Cout, k, s, acc = CASES[name]
    dev = "cuda"
    Ho = H // s
    dz = torch.randn(N, Ho, Ho, Cout, device=dev).bfloat16()
    w_d = (torch.randn(Cin, k * k * Cout, device=dev) * 0.05).bfloat16()
    dx = torch.zeros(N, H, H, Cin, device=dev, dtype=torch.bfloat16)
    z = torch.randn(N, H, H, Cin, device=dev).bfloat16()
    sc, sh, mu = (torch.randn(Cin, device=dev) for _ in range(3))
    sums = torch.zeros(2 * Cin, dtype=torch.float64, device=dev)
    kw = dict(N=N, H=H, W=H, Cin=Cin, Cout=Cout, k=k, stride=s, accumulate=acc)
    for fused in (False, True):
        plans = G.plan_conv_dgrad(dz, w_d, dx, bnb=(z, sc, sh, mu, sums, 0.01) if fused else None, **kw)
        for _ in range(2):
            for p in plans: p.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            for p in plans: p.run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gb = (dz.numel() + dx.numel() * (1 + acc + fused)) * 2 / 1e9
        kp = plans[0]
        print(f"{name} fused={int(fused)} acc={acc}: {ms:7.3f} ms  ({gb:.2f} GB -> {gb/ms:.2f} TB/s)  launches={len(plans)}")
        if fused and trace:
            buf.zero_()
            plans[-1].run(); torch.cuda.synchronize()
            e = buf.cpu()[2048:2048 + 16 * 24].view(-1, 16)
            print("  tile: wait tfull | ->slab written+bar (6) | ->phase2 start (8) | loads issued (9) | batch0 done (10) | all rows (11) | end (12) | tile end (7)")
            for i in range(2, 20):
                r = [int(x) for x in e[i]]
                if r[0] == 0: break
                print(f"  {i:3d} tfull+{r[1]-r[0]:6d} p1+{r[6]-r[1]:5d} ->8+{r[8]-r[6]:5d} issue+{r[9]-r[8]:5d} batch0+{r[10]-r[9]:6d} rest+{r[11]-r[10]:6d} end+{r[12]-r[11]:5d} | tile total {r[7]-r[0]:7d}")


for name in (sys.argv[1:] or ["L13", "L12", "L14", "L2"]):
    run(name)
