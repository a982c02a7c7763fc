"""Timing experiments on the conv GEMM: per-layer times for different k-blocks-per-stage / slab counts
(AVDN_GEMM_KPS, AVDN_GEMM_SLABS) and clock64 traces of the MMA-issuing thread and of epilogue thread 0
of CTA 0 (PROBE_TRACE=L2|L3|L13|L14; the kernel writes the trace when AVDN_GEMM_DBG_BUF is set).
Usage (GPU box): [AVDN_GEMM_KPS=k] [AVDN_GEMM_SLABS=n] [PROBE_TRACE=L14] python tools/gemm_epilogue_probe.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avdn_b200 import gemm as G

def bench(tag, N, H, Cin, Cout, k, s, stats=False, reps=20):
    dev = "cuda"
    x = torch.randn(N, H, H, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, k * k * Cin, device=dev) * 0.05).bfloat16()
    z = torch.empty(N, H // s, H // s, Cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * Cout, dtype=torch.float64, device=dev) if stats else None
    p = G.plan_conv_fwd(x, w, z, N=N, H=H, W=H, Cin=Cin, Cout=Cout, k=k, stride=s, stats=st)
    for _ in range(2):
        p.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        p.run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2 * N * (H // s) ** 2 * Cout * k * k * Cin
    print(f"{tag:28s} kps={os.environ.get('AVDN_GEMM_KPS','auto')} slabs={os.environ.get('AVDN_GEMM_SLABS','auto')} stats={int(stats)}  {ms:7.3f} ms  {fl/ms/1e9:7.0f} TF/s")

N = int(os.environ.get("PROBE_N", "640"))
if os.environ.get("PROBE_TRACE"):
    buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
    os.environ["AVDN_GEMM_DBG_BUF"] = str(buf.data_ptr())
    which = os.environ["PROBE_TRACE"]
    if which == "L3":
        bench("L3  3x3 32->64 @112", N, 112, 32, 64, 3, 1, reps=1)
    elif which == "L2":
        bench("L2  1x1 64->32 @112", N, 112, 64, 32, 1, 1, reps=1)
    elif which == "L13":
        bench("L13 1x1 256->128 @28", N, 28, 256, 128, 1, 1, reps=1)
    else:
        bench("L14 3x3 128->256 @28", N, 28, 128, 256, 3, 1, reps=1)
    e = buf.cpu()[2048:2048 + 16 * 40].view(-1, 16)      # 16 clock slots per tile (8..12: fused BN-backward phase)
    eb = int(e[0, 0])
    print("epilogue (thread 0) per tile: t_wait_start | +tfull | +wait_group | +bar1 | slab0: +ld/cvt/sts | +fence | +bar2 | tile_end")
    for i in range(2, 30):
        r = [int(x) for x in e[i]]
        print(f"{i:3d} start={r[0]-eb:8d} tfull+{r[1]-r[0]:5d} wg+{r[2]-r[1]:5d} bar1+{r[3]-r[2]:5d} tmemwait+{r[4]-r[3]:5d} cvt/sts+{r[5]-r[4]:5d} fence/bar2+{r[6]-r[5]:5d} rest+{r[7]-r[6]:6d}  total={r[7]-r[0]:6d}")
    t = buf.cpu().view(-1, 4)
    base = int(t[0, 0])
    prev = base
    for i in range(0, 120):
        a, b, c, kb = [int(x) for x in t[i]]
        print(f"{i:3d} kb={kb:2d} start={a-base:7d} wait={b-a:5d} issue+commit={c-b:5d}  gap_from_prev_end={a-prev:5d}")
        prev = c
    sys.exit(0)
bench("L14 3x3 128->256 @28", N, 28, 128, 256, 3, 1)
bench("L14 3x3 128->256 @28", N, 28, 128, 256, 3, 1, stats=True)
bench("L13 1x1 256->128 @28", N, 28, 256, 128, 1, 1)
bench("L3  3x3 32->64 @112", N, 112, 32, 64, 3, 1)
bench("L2  1x1 64->32 @112", N, 112, 64, 32, 1, 1)
bench("L39 3x3 256->512 @14", N, 14, 256, 512, 3, 1)
