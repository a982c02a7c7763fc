"""Runs single conv GEMM launches of the training step's shapes (for `ncu --set full -k regex:gemm_kernel`):
  L14 fwd   3x3 128->256 @28x28   (tensor-bound, the bulk of the step)      gemm_kernel<128,0,0,2,64>
  L3  fwd   3x3  32->64  @112x112 (thin: L2->SM bandwidth / latency bound)  gemm_kernel<64,0,0,1,32>
  L13 fwd   1x1 256->128 @28x28   (HBM-bound)
  L14 wgrad                        gemm_kernel<128,1,1,2,64>
Usage (GPU box): python tools/gemm_cases.py [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avdn_b200 import gemm as G

N = int(os.environ.get("PROBE_N", "640"))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda"


def conv_case(H, Cin, Cout, k, s, wgrad=False):
    x = torch.randn(N, H, H, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, k * k * Cin, device=dev) * 0.05).bfloat16()
    z = torch.empty(N, H // s, H // s, Cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    if wgrad:
        dz = torch.randn(N, H // s, H // s, Cout, device=dev).bfloat16()
        dw = torch.zeros(Cout, k * k * Cin, device=dev)
        return G.plan_conv_wgrad(dz, x, dw, N=N, H=H, W=H, Cin=Cin, Cout=Cout, k=k, stride=s)
    return G.plan_conv_fwd(x, w, z, N=N, H=H, W=H, Cin=Cin, Cout=Cout, k=k, stride=s, stats=st)


cases = [("L14 fwd 3x3 128->256 @28", conv_case(28, 128, 256, 3, 1)), ("L3 fwd 3x3 32->64 @112", conv_case(112, 32, 64, 3, 1)),
         ("L1 fwd 3x3 32->64 s2 @224", conv_case(224, 32, 64, 3, 2)), ("L2 fwd 1x1 64->32 @112", conv_case(112, 64, 32, 1, 1)),
         ("L7 fwd 3x3 64->128 @56", conv_case(56, 64, 128, 3, 1)),
         ("L13 fwd 1x1 256->128 @28", conv_case(28, 256, 128, 1, 1)), ("L14 wgrad", conv_case(28, 128, 256, 3, 1, wgrad=True))]
for name, p in cases:
    for _ in range(reps):
        p.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); p.run(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms  {p.flops / e0.elapsed_time(e1) / 1e9:.0f} TFLOP/s")

# the same 3x3 32->64 layer at 112x112 through the halo-tile kernel (csrc/conv3_halo.cu)
from avdn_b200 import _lib
x = torch.randn(N, 112, 112, 32, device=dev).bfloat16()
w = (torch.randn(64, 288, device=dev) * 0.05).bfloat16()
z = torch.empty(N, 112, 112, 64, device=dev, dtype=torch.bfloat16)
st = torch.zeros(128, dtype=torch.float64, device=dev)
run = lambda: _lib.call("avdn_conv3x3_thin_fwd", _lib.ptr(x), _lib.ptr(w), _lib.ptr(z), N, 112, 112, 32, 64, _lib.ptr(st))
for _ in range(reps):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
fl = 2.0 * N * 112 * 112 * 64 * 288
print(f"L3 fwd 3x3 32->64 @112 halo tiles: {e0.elapsed_time(e1):.3f} ms  {fl / e0.elapsed_time(e1) / 1e9:.0f} TFLOP/s")

# its data gradient: implicit GEMM vs halo tiles
from avdn_b200 import gemm as G2
dzt = torch.randn(N, 112, 112, 64, device=dev).bfloat16()
wdt = (torch.randn(32, 576, device=dev) * 0.05).bfloat16()
dxt = torch.empty(N, 112, 112, 32, device=dev, dtype=torch.bfloat16)
plans = G2.plan_conv_dgrad(dzt, wdt, dxt, N=N, H=112, W=112, Cin=32, Cout=64, k=3, stride=1)
def run_g():
    for pl in plans:
        pl.run()
run_h = lambda: _lib.call("avdn_conv3x3_thin_dgrad", _lib.ptr(dzt), _lib.ptr(wdt), _lib.ptr(dxt), N, 112, 112, 32, 64)
for name, fn in (("L3 dgrad implicit GEMM", run_g), ("L3 dgrad halo tiles", run_h)):
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms  {fl / e0.elapsed_time(e1) / 1e9:.0f} TFLOP/s")

# and its weight gradient: pixel-pair implicit GEMM vs halo tiles
dwf = torch.zeros(2 * 64, 3 * 3 * 2 * 32, device=dev)          # the pixel-pair layout of the engine (dwf_shape)
try:
    pw = G2.plan_conv_wgrad_pairs(dzt, x, dwf, N=N, H=112, W=112, Cin=32, Cout=64, k=3, stride=1)
    run_wg = pw.run
except Exception as e:          # layout of the pair buffer is the engine's business; time the halo kernel alone then
    print("pair-wgrad plan not built here:", e); run_wg = None
dwt = torch.zeros(64, 32, 3, 3, device=dev)
run_wh = lambda: _lib.call("avdn_conv3x3_thin_wgrad", _lib.ptr(dzt), _lib.ptr(x), _lib.ptr(dwt), N, 112, 112, 32, 64)
for name, fn in (("L3 wgrad pixel-pair implicit GEMM", run_wg), ("L3 wgrad halo tiles", run_wh)):
    if fn is None:
        continue
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms  {fl / e0.elapsed_time(e1) / 1e9:.0f} TFLOP/s")
