"""Host-side planning for the tcgen05 tensor-core primitive (``avdn_gemm_*``).

A :class:`GemmPlan` owns the ctypes mirror of ``avdn_gemm_desc``, the host plan
blob (encoded TMA descriptors + launch geometry) and references to the tensors
whose device pointers are baked into it.  Plans are built once per (layer,
shape) and re-run every step; nothing here touches the data.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

i32, i64, f32, u32 = C.c_int32, C.c_int64, C.c_float, C.c_uint32

PLAIN, CONV, WGRAD = 0, 1, 2
DT_BF16, DT_F32 = 0, 1


class Tap(C.Structure):
    _fields_ = [("map", i32), ("d1", i32), ("d2", i32), ("bk", i32)]


class Operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dim", i64 * 4), ("stride", i64 * 4), ("box", i32 * 4)]


class GemmCore(C.Structure):
    _fields_ = [("mode", i32), ("M", i32), ("N", i32), ("num_kb", i32), ("split_k", i32),
                ("batch0", i32), ("batch1", i32), ("b_batched", i32), ("cblocks", i32), ("n_taps", i32),
                ("taps", Tap * 9),
                ("tiles_w", i32), ("tiles_h", i32), ("tiles_n", i32),
                ("box_w", i32), ("box_h", i32), ("box_n", i32),
                ("valid_w", i32), ("valid_h", i32), ("valid_n", i32),
                ("out_H", i32), ("out_W", i32), ("out_sh", i32), ("out_sw", i32), ("out_oh", i32), ("out_ow", i32),
                ("out_dtype", i32), ("accumulate", i32), ("relu", i32), ("alpha", f32),
                ("tx_bytes", u32), ("pad_", i32),
                ("ldc", i64), ("out_bs0", i64), ("out_bs1", i64),
                ("out", C.c_void_p), ("bias", C.c_void_p), ("relu_mask", C.c_void_p), ("stats", C.c_void_p),
                ("col_scale", C.c_void_p), ("col_shift", C.c_void_p), ("residual", C.c_void_p),
                ("leaky_slope", f32), ("pad2_", i32),
                ("bnb_z", C.c_void_p), ("bnb_scale", C.c_void_p), ("bnb_shift", C.c_void_p),
                ("bnb_mean", C.c_void_p), ("bnb_sums", C.c_void_p),
                ("n_classes", i32), ("cls_tap0", i32 * 5), ("cls_oh", i32 * 4), ("cls_ow", i32 * 4), ("pad4_", i32 * 2)]


class GemmDesc(C.Structure):
    _fields_ = [("core", GemmCore), ("bn", i32), ("a_mn", i32), ("b_mn", i32), ("n_a", i32), ("n_b", i32),
                ("grid_m", i32), ("grid_n", i32), ("grid_z", i32), ("ctas", i32), ("bk", i32),
                ("a", Operand * 4), ("b", Operand * 4)]




def _cdiv(a, b):
    return (a + b - 1) // b


# CTA pairs (tcgen05 cta_group::2) wherever the tile shape allows it; AVDN_GEMM_CTAS=1 forces
# single-CTA tiles (A/B comparison in the benchmarks).
PAIR_DEFAULT = int(os.environ.get("AVDN_GEMM_CTAS", "2"))


def pick_ctas(bn, grid_m, ctas=None):
    if ctas is None:
        ctas = PAIR_DEFAULT
    return 2 if (ctas == 2 and bn >= 128 and grid_m >= 2) else 1


def operand(ptr, dims, strides, box):
    o = Operand()
    o.ptr = ptr
    for i in range(4):
        o.dim[i], o.stride[i], o.box[i] = int(dims[i]), int(strides[i]), int(box[i])
    return o


def _dt(t):
    if t.dtype == torch.bfloat16:
        return DT_BF16
    if t.dtype == torch.float32:
        return DT_F32
    raise TypeError(f"unsupported output dtype {t.dtype}")


class GemmPlan:
    def __init__(self, desc: GemmDesc, keep=(), flops=0, tag="gemm"):
        h = _lib.lib()
        self.flops, self.tag = flops, tag
        n = h.avdn_gemm_plan_bytes()
        self.desc = desc
        self.buf = C.create_string_buffer(n)
        self.keep = keep
        _lib.check(h.avdn_gemm_plan(C.byref(desc), self.buf, n), "avdn_gemm_plan")
        self._run = h.avdn_gemm_run

    def run(self):
        if _lib.PROFILE is not None:
            _lib.profile_record(self.tag, lambda: _lib.check(self._run(self.buf, _lib.stream_ptr()), "avdn_gemm_run"),
                                self.flops)
            return
        _lib.check(self._run(self.buf, _lib.stream_ptr()), "avdn_gemm_run")


BN_MAX = int(os.environ.get("AVDN_GEMM_BN_MAX", "128"))      # 128: two co-resident CTAs per SM beat one 256-wide tile
# data gradient of a stride-2 convolution: 1 = one launch with the four output parities as tile classes
# (avdn_gemm_core.n_classes), 0 = four launches that each re-read dZ
DGRAD_CLASSES = os.environ.get("AVDN_DGRAD_CLASSES", "1") != "0"


def pick_bn(N):
    if N <= 32:
        return 32
    if N <= 64:
        return 64
    if N <= 128 or N % 256 != 0 or BN_MAX < 256:
        return 128
    return 256


def plan_plain(*, M, N, K, a_ptr, lda, a_mn, b_ptr, ldb, b_mn, out, ldc, bias=None, relu=False, alpha=1.0,
               accumulate=0, batch0=1, batch1=1, a_bs=(0, 0), b_bs=None, out_bs=(0, 0), split_k=1, bn=None,
               relu_mask=None, out_ptr=None, keep=(), tag="gemm_plain", ctas=None):
    """D[M,N] = alpha * A.B^T (+bias)(relu).  Operands are bf16.

    K-major operand: stored ``[rows][K]`` with row pitch ``ld``; MN-major operand:
    stored ``[K][rows]`` with pitch ``ld`` (i.e. the transpose, contiguous along
    the M/N index).  ``a_bs``/``b_bs``/``out_bs`` are the element strides of the two
    batch dims; ``b_bs=None`` shares B across the batch.
    """
    bn = bn or max(64, pick_bn(N))
    d = GemmDesc()
    d.bk = 64
    c = d.core
    c.mode, c.M, c.N = PLAIN, M, N
    c.num_kb = _cdiv(K, 64)
    c.split_k, c.batch0, c.batch1 = split_k, batch0, batch1
    c.b_batched = 0 if b_bs is None else 1
    c.out_dtype, c.accumulate, c.relu, c.alpha = _dt(out), accumulate, int(relu), alpha
    c.ldc, c.out_bs0, c.out_bs1 = ldc, out_bs[0], out_bs[1]
    c.out = out_ptr if out_ptr is not None else out.data_ptr()
    c.bias = bias.data_ptr() if bias is not None else None
    c.relu_mask = relu_mask.data_ptr() if relu_mask is not None else None
    d.bn, d.a_mn, d.b_mn, d.n_a, d.n_b = bn, int(a_mn), int(b_mn), 1, 1
    safe = lambda s, fallback: s if s else fallback
    fa = lda * (K if a_mn else M)
    if a_mn:
        d.a[0] = operand(a_ptr, (M, K, batch0, batch1), (1, lda, safe(a_bs[0], fa), safe(a_bs[1], fa)), (64, 64, 1, 1))
    else:
        d.a[0] = operand(a_ptr, (K, M, batch0, batch1), (1, lda, safe(a_bs[0], fa), safe(a_bs[1], fa)), (64, 128, 1, 1))
    bb = b_bs or (0, 0)
    fb = ldb * (K if b_mn else N)
    nb0, nb1 = (batch0, batch1) if b_bs is not None else (1, 1)
    if b_mn:
        d.b[0] = operand(b_ptr, (N, K, nb0, nb1), (1, ldb, safe(bb[0], fb), safe(bb[1], fb)), (64, 64, 1, 1))
    else:
        d.b[0] = operand(b_ptr, (K, N, nb0, nb1), (1, ldb, safe(bb[0], fb), safe(bb[1], fb)), (64, bn, 1, 1))
    d.grid_m, d.grid_n, d.grid_z = _cdiv(M, 128), _cdiv(N, bn), batch0 * batch1 * split_k
    d.ctas = pick_ctas(bn, d.grid_m, ctas)
    return GemmPlan(d, keep=keep + (out, bias, relu_mask), flops=2 * M * N * K * batch0 * batch1, tag=tag)


def conv_box(W, H, N, rows):
    """Choose the (bw, bh, bn) TMA box of ``rows`` (128 for fwd/dgrad tiles, 64 for
    wgrad k-steps) pixels that wastes the fewest rows on a W x H x N grid."""
    best = None
    for bw in range(1, min(W, rows) + 1):
        if bw > 256:
            break
        tw = _cdiv(W, bw)
        for bh in range(1, min(H, rows // bw) + 1):
            bnn = min(rows // (bw * bh), N, 256)
            if bnn < 1:
                continue
            th, tn = _cdiv(H, bh), _cdiv(N, bnn)
            tiles = tw * th * tn
            # fewer tiles = less waste; tie-break on wider boxes (longer contiguous runs)
            key = (tiles, -bw, -bh)
            if best is None or key < best[0]:
                best = (key, (bw, bh, bnn))
    return best[1]


def wgrad_box(W, H, N):
    """Box of EXACTLY 64 pixels for a wgrad k-step (the unused part of a k-step would
    otherwise multiply stale shared memory).  Box dims may exceed the tensor extent:
    TMA zero-fills out-of-bounds elements of both operands."""
    best = None
    for lw in range(7):
        for lh in range(7 - lw):
            bw, bh, bnn = 1 << lw, 1 << lh, 1 << (6 - lw - lh)
            tiles = _cdiv(W, bw) * _cdiv(H, bh) * _cdiv(N, bnn)
            key = (tiles, -bw, -bh)
            if best is None or key < best[0]:
                best = (key, (bw, bh, bnn))
    return best[1]


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16, bias=None, relu=False,
         alpha=1.0, accumulate=0, split_k=1, bn=None):
    """Convenience one-shot 2-D GEMM on contiguous bf16 matrices (tests, small ops).
    ``a`` is ``[M,K]`` (or ``[K,M]`` if ``a_mn``), ``b`` is ``[N,K]`` (or ``[K,N]`` if ``b_mn``)."""
    _lib.require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous()
    K, M = (a.shape[0], a.shape[1]) if a_mn else (a.shape[1], a.shape[0])
    Kb, N = (b.shape[0], b.shape[1]) if b_mn else (b.shape[1], b.shape[0])
    assert K == Kb
    if out is None:
        out = (torch.zeros if accumulate == 2 else torch.empty)((M, N), dtype=out_dtype, device=a.device)
    p = plan_plain(M=M, N=N, K=K, a_ptr=a.data_ptr(), lda=a.shape[1], a_mn=a_mn, b_ptr=b.data_ptr(),
                   ldb=b.shape[1], b_mn=b_mn, out=out, ldc=out.shape[1], bias=bias, relu=relu, alpha=alpha,
                   accumulate=accumulate, split_k=split_k, bn=bn, keep=(a, b))
    p.run()
    return out


# ======================================================================= conv
def _act_operand(ptr, C_, W, H, N, box, *, sw=1, sh=1, Wfull=None, Hfull=None):
    """NHWC activation view as a rank-4 operand (C, W, H, N).  ``sw``/``sh`` = 2
    selects one parity view of a stride-2 layer (``ptr`` already offset)."""
    Wf, Hf = Wfull or W, Hfull or H
    return operand(ptr, (C_, W, H, N), (1, sw * C_, sh * Wf * C_, Hf * Wf * C_), box)


def conv_taps_fwd(k, stride, cin):
    """(view, dW, dH, k-offset) per filter tap for a forward conv with padding (k-1)//2."""
    taps = []
    pad = (k - 1) // 2
    for kh in range(k):
        for kw in range(k):
            bk = (kh * k + kw) * cin
            if stride == 1:
                taps.append((0, kw - pad, kh - pad, bk))
            else:            # stride 2, k = 3, pad = 1: input coord 2*o + kk - 1
                ph, d2 = ((1, -1), (0, 0), (1, 0))[kh]
                pw, d1 = ((1, -1), (0, 0), (1, 0))[kw]
                taps.append((ph * 2 + pw, d1, d2, bk))
    return taps


def _fill_taps(core, taps):
    core.n_taps = len(taps)
    for i, (m, d1, d2, bk) in enumerate(taps):
        core.taps[i].map, core.taps[i].d1, core.taps[i].d2, core.taps[i].bk = m, d1, d2, bk


def _x_views(x, Cin, W, H, N, stride, box):
    """The A (fwd) / B (wgrad) views of the layer input: 1 view for stride 1, the 4
    parity views for stride 2."""
    if stride == 1:
        return [_act_operand(x.data_ptr(), Cin, W, H, N, box)]
    assert W % 2 == 0 and H % 2 == 0
    views = []
    for ph in range(2):
        for pw in range(2):
            ptr = x.data_ptr() + (ph * W + pw) * Cin * 2
            views.append(_act_operand(ptr, Cin, W // 2, H // 2, N, box, sw=2, sh=2, Wfull=W, Hfull=H))
    return views


def plan_conv_fwd(x, w_f, z, *, N, H, W, Cin, Cout, k, stride, bn=None, flops=None, stats=None, ctas=None,
                  affine=None, residual=None):
    """z[N,Ho,Wo,Cout] = conv(x[N,H,W,Cin], w) ; ``w_f`` is ``[Cout, k*k*Cin]`` bf16
    (tap-major, channel-minor).  Channels are multiples of 64.
    ``affine=(scale, shift, slope)`` (fp32 ``[Cout]`` tensors): the epilogue writes
    ``leaky_slope(conv*scale + shift) (+ residual)`` instead -- eval-mode BatchNorm + LeakyReLU +
    shortcut folded into the convolution."""
    assert (Cin % 64 == 0 or Cin == 32) and (Cout % 64 == 0 or Cout == 32)
    Ho, Wo = H // stride, W // stride
    bn = bn or pick_bn(Cout)
    bk = 64 if Cin % 64 == 0 else 32           # 32-channel input: 64-byte rows, 64-byte swizzle
    assert not (bk == 32 and bn != 64), "32-channel input with a 32-channel output is not planned"
    bw, bh, bnn = conv_box(Wo, Ho, N, 128)
    d = GemmDesc()
    d.bk = bk
    c = d.core
    c.mode, c.M, c.N = CONV, 0, Cout
    taps = conv_taps_fwd(k, stride, Cin)
    _fill_taps(c, taps)
    c.cblocks = Cin // bk
    c.num_kb = len(taps) * c.cblocks
    c.split_k, c.batch0, c.batch1 = 1, 1, 1
    c.tiles_w, c.tiles_h, c.tiles_n = _cdiv(Wo, bw), _cdiv(Ho, bh), _cdiv(N, bnn)
    c.box_w, c.box_h, c.box_n = bw, bh, bnn
    c.valid_w, c.valid_h, c.valid_n = Wo, Ho, N
    c.out_H, c.out_W, c.out_sh, c.out_sw, c.out_oh, c.out_ow = Ho, Wo, 1, 1, 0, 0
    c.out_dtype, c.accumulate, c.relu, c.alpha = _dt(z), 0, 0, 1.0
    c.ldc = Cout
    c.out = z.data_ptr()
    d.bn, d.a_mn, d.b_mn = bn, 0, 0
    views = _x_views(x, Cin, W, H, N, stride, (bk, bw, bh, bnn))
    d.n_a, d.n_b = len(views), 1
    for i, v in enumerate(views):
        d.a[i] = v
    Kt = k * k * Cin
    d.b[0] = operand(w_f.data_ptr(), (Kt, Cout, 1, 1), (1, Kt, Kt * Cout, Kt * Cout), (bk, bn, 1, 1))
    d.grid_m, d.grid_n, d.grid_z = c.tiles_w * c.tiles_h * c.tiles_n, _cdiv(Cout, bn), 1
    d.ctas = pick_ctas(bn, d.grid_m, ctas)
    if stats is not None:        # fused BatchNorm statistics: f64 [2, Cout]
        assert stats.dtype == torch.float64 and stats.numel() >= 2 * Cout
        c.stats = stats.data_ptr()
    if affine is not None:
        sc, sh, slope = affine
        assert stats is None and sc.dtype == torch.float32 and sh.dtype == torch.float32
        assert sc.numel() >= Cout and sh.numel() >= Cout
        c.col_scale, c.col_shift, c.leaky_slope = sc.data_ptr(), sh.data_ptr(), float(slope)
        if residual is not None:
            assert residual.dtype == torch.bfloat16 and residual.shape == z.shape and residual.is_contiguous()
            c.residual = residual.data_ptr()
    else:
        assert residual is None
    return GemmPlan(d, keep=(x, w_f, z, stats, affine, residual), flops=flops if flops is not None else 2 * N * Ho * Wo * Cout * k * k * Cin,
                    tag="gemm_conv_fwd")


def plan_conv_dgrad(dz, w_d, dx, *, N, H, W, Cin, Cout, k, stride, accumulate=0, bn=None, flops=None, ctas=None,
                    bnb=None):
    """dx[N,H,W,Cin] (+)= conv_transpose(dz[N,Ho,Wo,Cout], w); ``w_d`` is
    ``[Cin, k*k*Cout]`` bf16 (tap-major, out-channel-minor).  Returns a list of plans
    (4 output-parity plans for stride 2).

    ``bnb=(z, scale, shift, mean, sums, slope)``: ``dx`` is the FINAL gradient of the activation
    ``leaky(z*scale + shift)`` of the block that produced the layer input, and the epilogue adds that block's
    BatchNorm-backward reductions into ``sums`` (f64 ``[2, Cin]``, zeroed by the caller); see
    ``avdn_gemm_core.bnb_*``."""
    assert (Cin % 64 == 0 or Cin == 32) and (Cout % 64 == 0 or Cout == 32)
    Ho, Wo = H // stride, W // stride
    bn = bn or pick_bn(Cin)
    bk = 64 if Cout % 64 == 0 else 32
    assert not (bk == 32 and bn != 64)
    pad = (k - 1) // 2
    Kt = k * k * Cout
    plans = []
    # dX[2i+a, 2j+b] = sum over taps with kh = a+1 (mod 2): a=0 -> kh=1 (d=0); a=1 -> kh=0 (d=+1), kh=2 (d=0)
    sel = {0: [(1, 0)], 1: [(0, 1), (2, 0)]}
    parities = [(0, 0)] if stride == 1 else [(0, 0), (0, 1), (1, 0), (1, 1)]
    one_launch = stride == 2 and DGRAD_CLASSES
    if one_launch:
        parities = [None]                      # one plan: the four output parities are classes of the same launch
    for par in parities:
        classes = None
        if stride == 1:
            taps = [(0, pad - kw, pad - kh, (kh * k + kw) * Cout) for kh in range(k) for kw in range(k)]
            gw, gh = W, H
            pa = pb = 0
        elif one_launch:
            taps, classes = [], []
            for (ca, cb) in [(0, 0), (0, 1), (1, 0), (1, 1)]:
                classes.append((len(taps), ca, cb))
                taps += [(0, dw, dh, (kh * k + kw) * Cout) for (kh, dh) in sel[ca] for (kw, dw) in sel[cb]]
            gw, gh = Wo, Ho
            pa = pb = 0
        else:
            pa, pb = par
            taps = [(0, dw, dh, (kh * k + kw) * Cout) for (kh, dh) in sel[pa] for (kw, dw) in sel[pb]]
            gw, gh = Wo, Ho
        bw, bh, bnn = conv_box(gw, gh, N, 128)
        d = GemmDesc()
        d.bk = bk
        c = d.core
        c.mode, c.M, c.N = CONV, 0, Cin
        _fill_taps(c, taps)
        c.cblocks = Cout // bk
        c.num_kb = len(taps) * c.cblocks
        if classes is not None:
            c.n_classes = len(classes)
            for i, (t0, ca, cb) in enumerate(classes):
                c.cls_tap0[i], c.cls_oh[i], c.cls_ow[i] = t0, ca, cb
            c.cls_tap0[len(classes)] = len(taps)
        c.split_k, c.batch0, c.batch1 = 1, 1, 1
        c.tiles_w, c.tiles_h, c.tiles_n = _cdiv(gw, bw), _cdiv(gh, bh), _cdiv(N, bnn)
        c.box_w, c.box_h, c.box_n = bw, bh, bnn
        c.valid_w, c.valid_h, c.valid_n = gw, gh, N
        c.out_H, c.out_W = H, W
        c.out_sh = c.out_sw = stride
        c.out_oh, c.out_ow = pa, pb
        c.out_dtype, c.accumulate, c.relu, c.alpha = _dt(dx), accumulate, 0, 1.0
        c.ldc = Cin
        c.out = dx.data_ptr()
        d.bn, d.a_mn, d.b_mn, d.n_a, d.n_b = bn, 0, 0, 1, 1
        d.a[0] = _act_operand(dz.data_ptr(), Cout, Wo, Ho, N, (bk, bw, bh, bnn))
        d.b[0] = operand(w_d.data_ptr(), (Kt, Cin, 1, 1), (1, Kt, Kt * Cin, Kt * Cin), (bk, bn, 1, 1))
        d.grid_m, d.grid_n, d.grid_z = c.tiles_w * c.tiles_h * c.tiles_n, _cdiv(Cin, bn), 1
        d.ctas = pick_ctas(bn, d.grid_m, ctas)
        if bnb is not None:
            bz, bsc, bsh, bmu, bsums, slope = bnb
            assert bz.dtype == torch.bfloat16 and bz.shape == dx.shape and bz.is_contiguous()
            assert bsums.dtype == torch.float64 and bsums.numel() >= 2 * Cin
            assert all(t.dtype == torch.float32 and t.numel() >= Cin for t in (bsc, bsh, bmu))
            c.bnb_z, c.bnb_scale, c.bnb_shift = bz.data_ptr(), bsc.data_ptr(), bsh.data_ptr()
            c.bnb_mean, c.bnb_sums, c.leaky_slope = bmu.data_ptr(), bsums.data_ptr(), float(slope)
        fl = flops if flops is not None else 2 * N * Ho * Wo * Cout * k * k * Cin
        plans.append(GemmPlan(d, keep=(dz, w_d, dx, bnb), flops=fl // len(parities), tag="gemm_conv_dgrad"))
    return plans


def plan_conv_wgrad(dz, x, dw, *, N, H, W, Cin, Cout, k, stride, split_k=None, bn=None, sms=148, flops=None,
                    ctas=None):
    """dw[Cout, k*k*Cin] (fp32, atomically accumulated -- zero it first) +=
    sum_pixels dz[pix, co] * x[pix + tap, ci]."""
    assert Cin % 64 == 0 and Cout % 64 == 0 and dw.dtype == torch.float32
    Ho, Wo = H // stride, W // stride
    bn = bn or pick_bn(Cin)
    bw, bh, bnn = wgrad_box(Wo, Ho, N)
    d = GemmDesc()
    d.bk = 64
    c = d.core
    c.mode, c.M, c.N = WGRAD, Cout, Cin
    taps = conv_taps_fwd(k, stride, Cin)
    _fill_taps(c, taps)
    c.tiles_w, c.tiles_h, c.tiles_n = _cdiv(Wo, bw), _cdiv(Ho, bh), _cdiv(N, bnn)
    c.box_w, c.box_h, c.box_n = bw, bh, bnn
    c.num_kb = c.tiles_w * c.tiles_h * c.tiles_n
    gm, gn = _cdiv(Cout, 128), _cdiv(Cin, bn)
    if split_k is None:
        split_k = max(1, min(c.num_kb, (4 * sms) // max(1, gm * gn * len(taps))))
    c.split_k, c.batch0, c.batch1 = split_k, 1, 1
    c.out_dtype, c.accumulate, c.relu, c.alpha = DT_F32, 2, 0, 1.0
    c.ldc = k * k * Cin
    c.out = dw.data_ptr()
    d.bn, d.a_mn, d.b_mn = bn, 1, 1
    d.a[0] = _act_operand(dz.data_ptr(), Cout, Wo, Ho, N, (64, bw, bh, bnn))
    views = _x_views(x, Cin, W, H, N, stride, (64, bw, bh, bnn))
    d.n_a, d.n_b = 1, len(views)
    for i, v in enumerate(views):
        d.b[i] = v
    d.grid_m, d.grid_n, d.grid_z = gm, gn, len(taps) * split_k
    d.ctas = pick_ctas(bn, gm, ctas)
    return GemmPlan(d, keep=(dz, x, dw), flops=flops if flops is not None else 2 * N * Ho * Wo * Cout * k * k * Cin,
                    tag="gemm_conv_wgrad")


def plan_conv_wgrad_pairs(dz, x, dwp, *, N, H, W, Cin, Cout, k, stride, split_k=None, sms=148, flops=None, ctas=None):
    """Weight gradient of a conv whose input and/or output has 32 channels, on PIXEL-PAIR views: two
    adjacent pixels (2 x 32 channels = 128 bytes) are one operand row, so both MN-major operands keep
    64-element rows.  ``dwp`` (fp32, zero it first) receives blocks that
    ``avdn_unpack_conv_wgrad_pairs`` folds into the [Cout,Cin,k,k] gradient:

    * stride 1: both operands paired along W -> D'[(a,co)][blk(kh,s)][(b,ci)], s = pair shift in {-1,0,1}
    * stride 2: X paired (pair index = output pixel), dZ as is -> D'[co][blk(kh,s)][(b,ci)], s in {-1,0}
    """
    assert W % 2 == 0 and dwp.dtype == torch.float32
    pad = (k - 1) // 2
    Np = 2 * Cin
    d = GemmDesc()
    d.bk = 64
    c = d.core
    if stride == 1:
        Wp = W // 2
        Mp = 2 * Cout
        shifts = (-1, 0, 1) if k == 3 else (0,)
        taps = [(0, s, kh - pad, (kh * len(shifts) + i) * Np) for kh in range(k) for i, s in enumerate(shifts)]
        bw, bh, bnn = wgrad_box(Wp, H, N)
        a_op = operand(dz.data_ptr(), (Mp, Wp, H, N), (1, Mp, W * Cout, H * W * Cout), (64, bw, bh, bnn))
        b_ops = [operand(x.data_ptr(), (Np, Wp, H, N), (1, Np, W * Cin, H * W * Cin), (64, bw, bh, bnn))]
        gw, gh = Wp, H
    else:
        assert k == 3 and H % 2 == 0
        Wp, Hp = W // 2, H // 2
        Mp = Cout
        taps = []
        for kh in range(3):
            ph, dh = ((1, -1), (0, 0), (1, 0))[kh]
            for i, s in enumerate((-1, 0)):
                taps.append((ph, s, dh, (kh * 2 + i) * Np))
        bw, bh, bnn = wgrad_box(Wp, Hp, N)
        a_op = operand(dz.data_ptr(), (Cout, Wp, Hp, N), (1, Cout, Wp * Cout, Hp * Wp * Cout), (64, bw, bh, bnn))
        b_ops = [operand(x.data_ptr() + ph * W * Cin * 2, (Np, Wp, Hp, N), (1, Np, 2 * W * Cin, H * W * Cin),
                         (64, bw, bh, bnn)) for ph in range(2)]
        gw, gh = Wp, Hp
    bn = 64 if Np <= 64 else 128
    c.mode, c.M, c.N = WGRAD, Mp, Np
    _fill_taps(c, taps)
    c.tiles_w, c.tiles_h, c.tiles_n = _cdiv(gw, bw), _cdiv(gh, bh), _cdiv(N, bnn)
    c.box_w, c.box_h, c.box_n = bw, bh, bnn
    c.num_kb = c.tiles_w * c.tiles_h * c.tiles_n
    gm, gn = _cdiv(Mp, 128), _cdiv(Np, bn)
    if split_k is None:
        split_k = max(1, min(c.num_kb, (4 * sms) // max(1, gm * gn * len(taps))))
    c.split_k, c.batch0, c.batch1 = split_k, 1, 1
    c.out_dtype, c.accumulate, c.relu, c.alpha = DT_F32, 2, 0, 1.0
    c.ldc = len(taps) * Np
    assert dwp.numel() >= Mp * c.ldc
    c.out = dwp.data_ptr()
    d.bn, d.a_mn, d.b_mn = bn, 1, 1
    d.a[0] = a_op
    d.n_a, d.n_b = 1, len(b_ops)
    for i, v in enumerate(b_ops):
        d.b[i] = v
    d.grid_m, d.grid_n, d.grid_z = gm, gn, len(taps) * split_k
    d.ctas = pick_ctas(bn, gm, ctas)
    return GemmPlan(d, keep=(dz, x, dwp), flops=flops if flops is not None else 2 * N * (H // stride) * (W // stride) * Cout * k * k * Cin,
                    tag="gemm_conv_wgrad")
