"""GPU: the tcgen05 GEMM / implicit-GEMM conv primitive against fp32 torch
references on the same bf16-rounded inputs (tolerance: bf16 output rounding,
fp32 accumulation => 1e-2 relative as north_star states for bf16)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G(built_lib):
    from avdn_b200 import gemm
    return gemm


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def _close(out, ref, rtol=1e-2):
    out, ref = out.float(), ref.float()
    err = (out - ref).abs().max().item()
    tol = rtol * ref.abs().max().item() + 1e-6
    assert err <= tol, f"max err {err} > tol {tol}"


@pytest.mark.parametrize("M,N,K,bn", [(300, 200, 520, 128), (128, 64, 64, 64), (1000, 512, 768, 256),
                                      (77, 270, 64, 128), (256, 2304, 768, 256)])
def test_plain_kmajor(G, M, N, K, bn):
    a, b = _rand(M, K, seed=1), _rand(N, K, seed=2)
    out = G.gemm(a, b, bn=bn, out_dtype=torch.float32)
    _close(out, a.float() @ b.float().t(), rtol=2e-3)
    out16 = G.gemm(a, b, bn=bn)
    _close(out16, a.float() @ b.float().t())


def test_plain_epilogues(G):
    M, N, K = 260, 200, 192
    a, b = _rand(M, K, seed=3), _rand(N, K, seed=4)
    bias = torch.randn(N, device="cuda")
    ref = torch.relu(0.5 * (a.float() @ b.float().t()) + bias)
    _close(G.gemm(a, b, bias=bias, relu=True, alpha=0.5, out_dtype=torch.float32), ref, rtol=2e-3)
    base = torch.randn(M, N, device="cuda")
    out = base.clone()
    G.gemm(a, b, out=out, accumulate=1)
    _close(out, base + a.float() @ b.float().t(), rtol=2e-3)
    base16 = _rand(M, N, seed=9)
    out16 = base16.clone()
    G.gemm(a, b, out=out16, accumulate=1)
    _close(out16, base16.float() + a.float() @ b.float().t())


@pytest.mark.parametrize("a_mn,b_mn", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("bn", [64, 128, 256])
def test_plain_mn_major(G, a_mn, b_mn, bn):
    M, N, K = 200, 320, 333 + 3        # K multiple of 8 for the pitch; ragged tiles
    K = 336
    a, b = _rand(M, K, seed=5), _rand(N, K, seed=6)
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    out = G.gemm(A, B, a_mn=a_mn, b_mn=b_mn, bn=bn, out_dtype=torch.float32)
    _close(out, a.float() @ b.float().t(), rtol=2e-3)


def test_split_k_atomic(G):
    M, N, K = 136, 104, 4096
    a, b = _rand(M, K, seed=7), _rand(N, K, seed=8)
    out = G.gemm(a.t().contiguous(), b.t().contiguous(), a_mn=True, b_mn=True, split_k=7, accumulate=2,
                 out_dtype=torch.float32)
    _close(out, a.float() @ b.float().t(), rtol=2e-3)


def test_batched_attention_shapes(G):
    """QK^T and PV with the ET strides: qkv [B,S,3*768], heads of 64, S=270."""
    B, S, H, dh = 2, 270, 12, 64
    qkv = _rand(B, S, 3 * H * dh, seed=11)
    Sp = 272
    scores = torch.zeros(B, H, S, Sp, device="cuda", dtype=torch.float32)
    p = G.plan_plain(M=S, N=S, K=dh, a_ptr=qkv.data_ptr(), lda=3 * H * dh, a_mn=False,
                     b_ptr=qkv.data_ptr() + H * dh * 2, ldb=3 * H * dh, b_mn=False, out=scores, ldc=Sp,
                     alpha=0.125, batch0=H, batch1=B, a_bs=(dh, S * 3 * H * dh), b_bs=(dh, S * 3 * H * dh),
                     out_bs=(S * Sp, H * S * Sp), keep=(qkv,))
    p.run()
    q = qkv[:, :, :H * dh].view(B, S, H, dh).permute(0, 2, 1, 3).float()
    k = qkv[:, :, H * dh:2 * H * dh].view(B, S, H, dh).permute(0, 2, 1, 3).float()
    v = qkv[:, :, 2 * H * dh:].view(B, S, H, dh).permute(0, 2, 1, 3).float()
    ref = 0.125 * q @ k.transpose(-1, -2)
    _close(scores[..., :S], ref, rtol=2e-3)
    assert (scores[..., S:] == 0).all()
    # PV: P [B,H,S,Sp] bf16 (K-major over keys), V MN-major
    P = torch.softmax(ref, -1).to(torch.bfloat16)
    Pp = torch.zeros(B, H, S, Sp, device="cuda", dtype=torch.bfloat16)
    Pp[..., :S] = P
    o = torch.zeros(B, S, H * dh, device="cuda", dtype=torch.bfloat16)
    p2 = G.plan_plain(M=S, N=dh, K=S, a_ptr=Pp.data_ptr(), lda=Sp, a_mn=False,
                      b_ptr=qkv.data_ptr() + 2 * H * dh * 2, ldb=3 * H * dh, b_mn=True, out=o, ldc=H * dh,
                      batch0=H, batch1=B, a_bs=(S * Sp, H * S * Sp), b_bs=(dh, S * 3 * H * dh),
                      out_bs=(dh, S * H * dh), bn=64, keep=(Pp, qkv))
    p2.run()
    ref_o = (P.float() @ v).permute(0, 2, 1, 3).reshape(B, S, H * dh)
    _close(o, ref_o)


def _conv_ref(x_nhwc, w, stride):
    k = w.shape[-1]
    return torch.nn.functional.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w.float(), stride=stride,
                                      padding=(k - 1) // 2)


@pytest.mark.parametrize("N,H,Cin,Cout,k,stride", [
    (4, 14, 128, 256, 3, 1), (3, 28, 64, 128, 3, 2), (5, 7, 256, 128, 1, 1), (2, 56, 64, 64, 3, 1),
    (20, 7, 128, 256, 3, 1), (2, 14, 256, 512, 3, 2), (1, 112, 64, 64, 3, 2)])
def test_conv_fwd_dgrad_wgrad(G, N, H, Cin, Cout, k, stride):
    W = H
    x = _rand(N, H, W, Cin, seed=21)
    w = _rand(Cout, Cin, k, k, seed=22, scale=0.05)
    Ho, Wo = H // stride, W // stride
    # forward
    w_f = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous()
    z = torch.empty(N, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    G.plan_conv_fwd(x, w_f, z, N=N, H=H, W=W, Cin=Cin, Cout=Cout, k=k, stride=stride).run()
    ref = _conv_ref(x, w, stride)
    _close(z.permute(0, 3, 1, 2), ref)
    # dgrad
    dz = _rand(N, Ho, Wo, Cout, seed=23)
    w_d = w.permute(1, 2, 3, 0).reshape(Cin, k * k * Cout).contiguous()
    dx = torch.full((N, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    for p in G.plan_conv_dgrad(dz, w_d, dx, N=N, H=H, W=W, Cin=Cin, Cout=Cout, k=k, stride=stride):
        p.run()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().requires_grad_(True)
    out = torch.nn.functional.conv2d(xr, wr, stride=stride, padding=(k - 1) // 2)
    out.backward(dz.float().permute(0, 3, 1, 2))
    _close(dx.permute(0, 3, 1, 2), xr.grad)
    # dgrad accumulate
    dx2 = dx.clone()
    for p in G.plan_conv_dgrad(dz, w_d, dx2, N=N, H=H, W=W, Cin=Cin, Cout=Cout, k=k, stride=stride, accumulate=1):
        p.run()
    _close(dx2.permute(0, 3, 1, 2), 2 * xr.grad, rtol=2e-2)
    # wgrad
    dw = torch.zeros(Cout, k * k * Cin, device="cuda", dtype=torch.float32)
    G.plan_conv_wgrad(dz, x, dw, N=N, H=H, W=W, Cin=Cin, Cout=Cout, k=k, stride=stride).run()
    ref_dw = wr.grad.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin)
    _close(dw, ref_dw, rtol=3e-3)


@pytest.mark.parametrize("N,H,Cin,Cout,k,stride,acc", [
    (4, 14, 128, 256, 3, 1, 0),      # 128-column tiles, CTA pairs
    (5, 28, 256, 128, 1, 1, 1),      # 1x1 into a residual source: old dA read, added in fp32, stored
    (3, 28, 64, 128, 3, 2, 0),       # stride 2: four parity launches add into one sums buffer
    (3, 56, 32, 64, 3, 1, 0),        # 32-channel gradient (64-byte slab rows)
    (2, 56, 64, 32, 1, 1, 1),        # 32-channel dZ operand (32-element k-blocks), accumulate
    (20, 7, 128, 256, 3, 1, 0),      # 7x7 boxes spanning images, 98 of 128 tile rows used
    (5, 7, 512, 1024, 3, 1, 1),      # odd tile count under CTA pairs (phantom tile)
    (2, 112, 32, 64, 3, 2, 0)])      # first stride-2 block: 224x224x32 gradient
def test_conv_dgrad_fused_bn_backward_sums(G, N, H, Cin, Cout, k, stride, acc):
    """avdn_gemm_core.bnb_*: the dgrad epilogue reduces sum g and sum g*(z-mean), g = dA*leaky'(z*scale+shift), of the
    block that produced the layer input -- checked against torch on the dA the launch actually stored (bit-exact
    inputs, fp32 partial sums => 1e-4), and dA itself against the unfused launch."""
    W = H
    Ho, Wo = H // stride, W // stride
    dz = _rand(N, Ho, Wo, Cout, seed=31)
    w = _rand(Cout, Cin, k, k, seed=32, scale=0.05)
    w_d = w.permute(1, 2, 3, 0).reshape(Cin, k * k * Cout).contiguous()
    z = _rand(N, H, W, Cin, seed=33)
    g = torch.Generator(device="cuda").manual_seed(34)
    scale = torch.randn(Cin, device="cuda", generator=g)
    shift = torch.randn(Cin, device="cuda", generator=g) * 0.5
    mean = torch.randn(Cin, device="cuda", generator=g) * 0.3
    slope = 0.01
    base = _rand(N, H, W, Cin, seed=35) if acc else torch.full((N, H, W, Cin), float("nan"), device="cuda",
                                                                 dtype=torch.bfloat16)
    kw = dict(N=N, H=H, W=W, Cin=Cin, Cout=Cout, k=k, stride=stride, accumulate=acc)
    dx_ref = base.clone()
    for p in G.plan_conv_dgrad(dz, w_d, dx_ref, **kw):
        p.run()
    dx = base.clone()
    sums = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
    plans = G.plan_conv_dgrad(dz, w_d, dx, bnb=(z, scale, shift, mean, sums, slope), **kw)
    assert all(p.desc.core.bnb_z for p in plans)
    for p in plans:
        p.run()
    torch.cuda.synchronize()
    assert torch.isfinite(dx.float()).all()
    # the fused launch adds old + new in fp32 and rounds once; the TMA reduce-add rounds twice
    _close(dx, dx_ref, rtol=1e-2 if acc else 0.0)
    da, zf = dx.double(), z.double()
    # z*scale + shift in float64 is exact up to one rounding: its sign is the sign of the kernel's fp32 fma
    y = zf * scale.double() + shift.double()
    gg = torch.where(y > 0, da, (dx.float() * slope).double())
    hh = gg * (zf - mean.double())
    s1, s2 = gg.sum(dim=(0, 1, 2)), hh.sum(dim=(0, 1, 2))
    got = sums.view(2, Cin).clone()
    tol1 = 2e-4 * gg.abs().sum(dim=(0, 1, 2)) + 1e-6
    tol2 = 2e-4 * hh.abs().sum(dim=(0, 1, 2)) + 1e-6
    assert ((got[0] - s1).abs() <= tol1).all(), ((got[0] - s1).abs() / tol1).max().item()
    assert ((got[1] - s2).abs() <= tol2).all(), ((got[1] - s2).abs() / tol2).max().item()
    if not acc:
        # the caller owns the zeroing: a second run adds the same sums again (and they are reproducible)
        for p in plans:
            p.run()
        torch.cuda.synchronize()
        assert torch.allclose(sums.view(2, Cin), 2 * got, rtol=1e-12, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,W", [(1, 16, 8), (2, 32, 24), (3, 112, 112)])
def test_conv3x3_thin_halo_matches_implicit_gemm(built_lib, N, H, W):
    """avdn_conv3x3_thin_fwd (halo-tile reuse: three 8x18-pixel boxes per 8x16 output tile, resident filters) against the
    general implicit-GEMM launch of the same 3x3 32->64 layer and against fp32 torch: same bf16 operands, fp32
    accumulation, bf16 outputs, f64 statistics of the rounded outputs."""
    from avdn_b200 import _lib, gemm as G
    call, ptr = _lib.call, _lib.ptr
    Cin, Cout = 32, 64
    assert _lib.lib().avdn_conv3x3_thin_supported(H, W, Cin, Cout) == 1
    assert _lib.lib().avdn_conv3x3_thin_supported(H + 1, W, Cin, Cout) == 0
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, 9 * Cin, device="cuda", generator=g) * 0.06).bfloat16()      # [co][tap][ci]
    z_ref = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    st_ref = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    G.plan_conv_fwd(x, w, z_ref, N=N, H=H, W=W, Cin=Cin, Cout=Cout, k=3, stride=1, stats=st_ref).run()
    z = torch.full_like(z_ref, float("nan"))
    st = torch.full((2 * Cout,), 7.0, dtype=torch.float64, device="cuda")
    call("avdn_conv3x3_thin_fwd", ptr(x), ptr(w), ptr(z), N, H, W, Cin, Cout, ptr(st))
    torch.cuda.synchronize()
    wt = w.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).contiguous()
    zt = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1)
    assert torch.isfinite(z.float()).all()
    d = (z.float() - zt).abs()
    assert (d <= 2.0 ** -7 * zt.abs() + 1e-3).all(), d.max()          # one bf16 step of the fp32 result
    assert (z.float() - z_ref.float()).abs().max() <= 2.0 ** -7 * zt.abs().max()
    assert (z != z_ref).float().mean() < 1e-3                            # same operands, same accumulation depth
    zs = z.double().view(-1, Cout)
    assert torch.allclose(st[:Cout], zs.sum(0), rtol=1e-5, atol=1e-3 * (N * H * W) ** 0.5)
    assert torch.allclose(st[Cout:], (zs * zs).sum(0), rtol=1e-4)
    assert torch.allclose(st, st_ref, rtol=1e-4, atol=1e-2 * (N * H * W) ** 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,W", [(1, 16, 8), (2, 32, 24), (2, 112, 112)])
def test_conv3x3_thin_halo_dgrad_matches_implicit_gemm(built_lib, N, H, W):
    """avdn_conv3x3_thin_dgrad (the data gradient of the 3x3 32->64 layer as a halo-tile convolution of dz with the
    mirrored filters) against the implicit-GEMM dgrad launch and fp32 torch autograd."""
    from avdn_b200 import _lib, gemm as G
    call, ptr = _lib.call, _lib.ptr
    Cin, Cout = 32, 64
    g = torch.Generator(device="cuda").manual_seed(9)
    dz = torch.randn(N, H, W, Cout, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * 0.06).bfloat16()
    wd = w.permute(1, 2, 3, 0).reshape(Cin, 9 * Cout).contiguous()                      # [ci][tap][co]
    dx_ref = torch.empty(N, H, W, Cin, device="cuda", dtype=torch.bfloat16)
    for pl in G.plan_conv_dgrad(dz, wd, dx_ref, N=N, H=H, W=W, Cin=Cin, Cout=Cout, k=3, stride=1):
        pl.run()
    dx = torch.full_like(dx_ref, float("nan"))
    call("avdn_conv3x3_thin_dgrad", ptr(dz), ptr(wd), ptr(dx), N, H, W, Cin, Cout)
    torch.cuda.synchronize()
    xt = torch.zeros(N, Cin, H, W, device="cuda", requires_grad=True)
    torch.nn.functional.conv2d(xt, w.float(), padding=1).backward(dz.float().permute(0, 3, 1, 2))
    ref = xt.grad.permute(0, 2, 3, 1)
    assert torch.isfinite(dx.float()).all()
    d = (dx.float() - ref).abs()
    assert (d <= 2.0 ** -7 * ref.abs() + 2e-3).all(), d.max()
    assert (dx != dx_ref).float().mean() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,W", [(1, 16, 8), (2, 32, 24), (2, 112, 112)])
def test_conv3x3_thin_halo_wgrad_vs_torch(built_lib, N, H, W):
    """avdn_conv3x3_thin_wgrad (pixels as K: the dz tile and the three horizontal-tap boxes of x read MN-major, the boxes
    as ONE N = 96 operand, three accumulators alive for the whole kernel) against fp32 torch autograd; accumulates."""
    from avdn_b200 import _lib
    call, ptr = _lib.call, _lib.ptr
    Cin, Cout = 32, 64
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).bfloat16()
    dz = torch.randn(N, H, W, Cout, device="cuda", generator=g).bfloat16()
    dw = torch.full((Cout, Cin, 3, 3), 0.5, device="cuda")
    call("avdn_conv3x3_thin_wgrad", ptr(dz), ptr(x), ptr(dw), N, H, W, Cin, Cout)
    torch.cuda.synchronize()
    wt = torch.zeros(Cout, Cin, 3, 3, device="cuda", requires_grad=True)
    torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1).backward(dz.float().permute(0, 3, 1, 2))
    err = ((dw - 0.5 - wt.grad).norm() / wt.grad.norm()).item()
    assert err < 2e-3, err
