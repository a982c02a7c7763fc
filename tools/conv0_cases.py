"""Block 0 of the trunk alone (for `ncu --set full -k regex:conv0`): the recompute-path kernels at the bench shape.
Usage (GPU box): python tools/conv0_cases.py [N=640]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avdn_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 640
H = W = 224
dev = "cuda"
call, ptr = _lib.call, _lib.ptr
x = torch.zeros(N, H, W, 4, device=dev, dtype=torch.bfloat16)
x[..., :3] = torch.randn(N, H, W, 3, device=dev).to(torch.bfloat16)
w = torch.randn(32, 3, 3, 3, device=dev) * 0.2
sc, sh, mu, rs = (torch.rand(32, device=dev) + 0.5 for _ in range(4))
a = torch.empty(N, H, W, 32, device=dev, dtype=torch.bfloat16)
da = torch.randn(N, H, W, 32, device=dev).to(torch.bfloat16)
sums = torch.zeros(128, dtype=torch.float64, device=dev)
zw, gw = torch.zeros(864, device=dev), torch.zeros(864, device=dev)
xs9 = torch.zeros(36, dtype=torch.float64, device=dev)
mask = torch.zeros(N * H * W, dtype=torch.int32, device=dev)
dw, dg, db = torch.zeros(32, 3, 3, 3, device=dev), torch.zeros(32, device=dev), torch.zeros(32, device=dev)


def run():
    call("avdn_conv0_fwd_stats", ptr(x), ptr(w), N, H, W, ptr(sums), ptr(zw), ptr(xs9))
    call("avdn_conv0_fwd_apply", ptr(x), ptr(w), ptr(sc), ptr(sh), 0.01, ptr(a), ptr(mask), N, H, W)
    call("avdn_conv0_bwd", ptr(x), ptr(w), ptr(da), ptr(mask), ptr(sc), ptr(sh), ptr(mu), ptr(rs), 0.01, N, H, W, ptr(zw), ptr(xs9),
         ptr(sums), ptr(gw), ptr(dw), ptr(dg), ptr(db))


for path in (0, 1):
    _lib.lib().avdn_conv0_set_tensor_path(path)
    run(); torch.cuda.synchronize()
    for rep in range(2):
        _lib.PROFILE = []
        run(); torch.cuda.synchronize()
        for name, e0, e1, fl, nb in _lib.PROFILE:
            print(f"{'tcgen05 ' if path else 'mma.sync'} {name}: {e0.elapsed_time(e1):.3f} ms")
        _lib.PROFILE = None
_lib.lib().avdn_conv0_set_tensor_path(1)
