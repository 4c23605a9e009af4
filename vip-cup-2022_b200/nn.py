"""Layer-level host functions over the C ABI (include/vipcup.h).  torch tensors are only handles: every function
passes ``data_ptr()``s and the current CUDA stream to libvipcup.so; there is no torch arithmetic on this path."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import VipError

BF16 = torch.bfloat16
ACT = {None: 0, "none": 0, "linear": 0, "relu": 1, "gelu": 2, "sigmoid": 3, "swish": 4}


def _p(t):
    return None if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, name, dtype=BF16):
    if t is None:
        return
    if not t.is_cuda or not t.is_contiguous() or t.dtype != dtype:
        raise VipError(f"{name} must be a contiguous CUDA {dtype} tensor")


STATS = torch.int64   # row statistics records [M, 3] and pooled sums are 64-bit fixed point (csrc/stats.cuh)


def row_stats_buffer(*lead, device):
    """Zeroed row statistics records: int64 [*lead, 3] = (sum (v - p), sum (v - p)^2 in 36.28 fixed point, bits of p)."""
    return zero_(torch.empty((*lead, 3), dtype=STATS, device=device))


def lo_plane(m, n, device):
    """Uninitialised low plane for an [m, n] two-plane tensor (opaque blocked layout, see include/vipcup.h)."""
    return torch.empty(((m + 31) // 32 * 32, n), dtype=BF16, device=device)


def lo_plane_to_rows(lo, m):
    """Blocked low plane -> ordinary [m, n] row-major tensor (tests / debugging)."""
    mp, n = lo.shape
    return lo.view(mp // 32, n // 8, 32, 8).permute(0, 2, 1, 3).reshape(mp, n)[:m]


def lo_plane_from_rows(x):
    """[m, n] row-major bf16 -> blocked low plane."""
    m, n = x.shape
    mp = (m + 31) // 32 * 32
    xp = torch.zeros((mp, n), dtype=x.dtype, device=x.device)
    xp[:m] = x
    return xp.view(mp // 32, 32, n // 8, 8).permute(0, 2, 1, 3).contiguous().view(mp, n)


def finalize_stats(records, cols, eps=1e-5, out=None):
    """int64 [M, 3] row statistics records -> f32 [M, 2] (mean, 1 / sigma): the ``ln_stats`` operand of a folded LayerNorm."""
    _chk(records, "records", STATS)
    m = records.numel() // 3
    if out is None:
        out = torch.empty((m, 2), dtype=torch.float32, device=records.device)
    _chk(out, "out", torch.float32)
    _lib.check(_lib.lib().vip_row_stats_finalize(_p(records), m, int(cols), float(eps), _p(out), _st()), "vip_row_stats_finalize")
    return out


def _epilogue(out, bias=None, act=None, colscale=None, residual=None, ln_stats=None, ln_colsum=None,
              row_stats=None, gap=None, gap_rows=0, row_gate=None, gate_rows=0, residual_lo=None, out_lo=None,
              row_pivot=None):
    _chk(residual, "residual"), _chk(residual_lo, "residual_lo"), _chk(out_lo, "out_lo")
    for name, t in (("bias", bias), ("colscale", colscale), ("ln_colsum", ln_colsum), ("row_gate", row_gate),
                    ("ln_stats", ln_stats), ("row_pivot", row_pivot)):
        _chk(t, name, torch.float32)
    for name, t in (("row_stats", row_stats), ("gap", gap)):
        _chk(t, name, STATS)
    if ln_stats is not None and ln_stats.numel() != 2 * out.shape[0]:
        raise VipError("ln_stats must be f32 [M, 2] (mean, 1 / sigma)")
    if row_stats is not None and row_stats.numel() != 3 * out.shape[0]:
        raise VipError("row_stats must be int64 [M, 3] row statistics records")
    e = _lib.Epilogue()
    e.bias, e.act, e.colscale = _p(bias), ACT[act], _p(colscale)
    e.residual, e.ldr = _p(residual), (0 if residual is None else residual.stride(0))
    e.out, e.ldc = _p(out), out.stride(0)
    e.out_dtype = _lib.VIP_DTYPE_BF16 if out.dtype == BF16 else _lib.VIP_DTYPE_F32
    e.ln_stats, e.ln_colsum = _p(ln_stats), _p(ln_colsum)
    e.row_stats, e.gap, e.gap_rows = _p(row_stats), _p(gap), int(gap_rows)
    e.row_gate, e.gate_rows = _p(row_gate), int(gate_rows)
    e.residual_lo, e.out_lo, e.row_pivot = _p(residual_lo), _p(out_lo), _p(row_pivot)
    return e


def gemm(a, w, bias=None, act=None, colscale=None, residual=None, out=None, out_dtype=BF16, **fused):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T)  (persistent tcgen05 kernel, fp32 accumulation).  ``fused`` takes the
    vip_epilogue_t extras: ln_stats / ln_colsum / ln_cols / ln_eps (LayerNorm folded into this contraction), row_stats
    (statistics of the output rows for the next folded LayerNorm), gap / gap_rows (global-average-pool partial sums)."""
    _chk(a, "a"), _chk(w, "w")
    m, k = a.shape
    n = w.shape[0]
    if w.shape[1] != k:
        raise VipError(f"gemm: K mismatch {a.shape} x {w.shape}")
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    e = _epilogue(out, bias, act, colscale, residual, **fused)
    rc = _lib.lib().vip_gemm_bf16_ex(_p(a), a.stride(0), _p(w), w.stride(0), m, n, k, e, _st())
    _lib.check(rc, "vip_gemm_bf16_ex")
    return out


def gemm_grouped(a, w, rows_per_group, bias=None, act=None, residual=None, out=None, **fused):
    """``gemm`` with one weight matrix per group of ``rows_per_group`` rows of ``a`` (a multiple of 128):
    w is [groups * N, K], group i of the rows is contracted with w[i * N : (i + 1) * N]."""
    _chk(a, "a"), _chk(w, "w")
    m, k = a.shape
    groups = (m + rows_per_group - 1) // rows_per_group
    n = w.shape[0] // groups
    if w.shape[1] != k or n * groups != w.shape[0]:
        raise VipError(f"gemm_grouped: {tuple(a.shape)} x {tuple(w.shape)} with {groups} groups")
    if out is None:
        out = torch.empty((m, n), dtype=BF16, device=a.device)
    e = _epilogue(out, bias, act, None, residual, **fused)
    rc = _lib.lib().vip_gemm_grouped_bf16(_p(a), a.stride(0), _p(w), w.stride(0), m, n, k, rows_per_group, e, _st())
    _lib.check(rc, "vip_gemm_grouped_bf16")
    return out


def scale_weights(w, gate):
    """[G * N, K] bf16 = w[n, k] * gate[g, k]: a per-image input-channel scale folded into copies of the weights."""
    _chk(w, "w"), _chk(gate, "gate", torch.float32)
    n, k = w.shape
    g = gate.shape[0]
    out = torch.empty((g * n, k), dtype=BF16, device=w.device)
    _lib.check(_lib.lib().vip_scale_weights_bf16(_p(w), w.stride(0), _p(gate), g, n, k, _p(out), _st()), "vip_scale_weights_bf16")
    return out


def conv2d(x, w, bias=None, ksize=1, stride=1, pad=0, act=None, residual=None, out_hw=None, **fused):
    """x bf16 [N,H,W,C]; w bf16 [Cout, Kp] with K order (r,s,c). Returns bf16 [N,Ho,Wo,Cout].
    1x1 stride 1: plain GEMM on the NHWC activation; C % 8 == 0: implicit GEMM (im2col-mode TMA, nothing materialised);
    otherwise (the 3-channel network input): explicit im2col matrix + GEMM.  ``out_hw`` (explicit-im2col path only):
    output size for asymmetric padding -- ``pad`` is then the top / left padding, whatever the output size needs beyond
    the image at the bottom / right reads as zero (TF 'SAME')."""
    _chk(x, "x"), _chk(w, "w")
    n, h, wd, c = x.shape
    ho = (h + 2 * pad - ksize) // stride + 1
    wo = (wd + 2 * pad - ksize) // stride + 1
    if out_hw is not None:
        if c % 8 == 0 and w.shape[1] == ksize * ksize * c:
            raise VipError("conv2d: out_hw (asymmetric padding) is only supported on the explicit-im2col path")
        ho, wo = out_hw
    cout = w.shape[0]
    if fused.get("gap") is not None and not fused.get("gap_rows"):
        fused["gap_rows"] = ho * wo
    res2 = None if residual is None else residual.view(n * ho * wo, -1)
    if ksize == 1 and stride == 1 and pad == 0 and w.shape[1] == c:
        y = gemm(x.view(n * h * wd, c), w, bias=bias, act=act, residual=res2, **fused)
    elif c % 8 == 0 and w.shape[1] == ksize * ksize * c:
        y = torch.empty((n * ho * wo, cout), dtype=BF16, device=x.device)
        e = _epilogue(y, bias, act, None, res2, **fused)
        rc = _lib.lib().vip_conv2d_bf16(_p(x), n, h, wd, c, _p(w), w.stride(0), cout, ksize, stride, pad, e, _st())
        _lib.check(rc, "vip_conv2d_bf16")
    else:
        kp = w.shape[1]
        a = torch.empty((n * ho * wo, kp), dtype=BF16, device=x.device)
        rc = _lib.lib().vip_im2col_bf16(_p(x), n, h, wd, c, ksize, stride, pad, ho, wo, _p(a), kp, _st())
        _lib.check(rc, "vip_im2col_bf16")
        y = gemm(a, w, bias=bias, act=act, residual=res2, **fused)
    return y.view(n, ho, wo, cout)


def conv2d_grouped(x, ws, biases, ksize=3, stride=1, pad=1, act=None):
    """Conv2D(groups=G) as G implicit-GEMM convolutions on channel slices: x bf16 [N,H,W,C]; ws[g] bf16 [Cout_g, k*k*C/G]
    (K order r,s,c); biases[g] f32 [Cout_g] | None.  Returns bf16 [N,Ho,Wo,sum Cout_g]."""
    _chk(x, "x")
    n, h, wd, c = x.shape
    g = len(ws)
    cg = c // g
    ho = (h + 2 * pad - ksize) // stride + 1
    wo = (wd + 2 * pad - ksize) // stride + 1
    cout = sum(w.shape[0] for w in ws)
    y = torch.empty((n * ho * wo, cout), dtype=BF16, device=x.device)
    xe, co0 = x.element_size(), 0
    for gi, w in enumerate(ws):
        _chk(w, "w")
        e = _epilogue(y[:, co0: co0 + w.shape[0]], None if biases is None else biases[gi], act)
        e.out, e.ldc = y.data_ptr() + co0 * xe, cout
        rc = _lib.lib().vip_conv2d_slice_bf16(x.data_ptr() + gi * cg * xe, n, h, wd, cg, c, _p(w), w.stride(0), w.shape[0], ksize,
                                              stride, pad, e, _st())
        _lib.check(rc, "vip_conv2d_slice_bf16")
        co0 += w.shape[0]
    return y.view(n, ho, wo, cout)


def split_attention2(x, logits):
    """Radix-2 split attention: x bf16 [N,H,W,2F], logits f32 [N,2F] -> bf16 [N,H,W,F] (r-softmax weighted sum of the halves)."""
    _chk(x, "x"), _chk(logits, "logits", torch.float32)
    n, h, w, c2 = x.shape
    out = torch.empty((n, h, w, c2 // 2), dtype=BF16, device=x.device)
    _lib.check(_lib.lib().vip_split_attention2_bf16(_p(x), _p(logits), _p(out), n, h * w, c2 // 2, _st()), "vip_split_attention2_bf16")
    return out


def avgpool3s2(x):
    """ZeroPadding2D(1) + AveragePooling2D(3, strides=2), divisor 9."""
    _chk(x, "x")
    n, h, w, c = x.shape
    out = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), dtype=BF16, device=x.device)
    _lib.check(_lib.lib().vip_avgpool3s2_bf16(_p(x), _p(out), n, h, w, c, _st()), "vip_avgpool3s2_bf16")
    return out


def act_scale(x, act=None, scale=1.0):
    """act(x) * scale elementwise (bf16): the NFNet pre-activation swish(x) * beta."""
    _chk(x, "x")
    out = torch.empty_like(x)
    _lib.check(_lib.lib().vip_act_scale_bf16(_p(x), _p(out), x.numel(), ACT[act], float(scale), _st()), "vip_act_scale_bf16")
    return out


def eca_gate(gap, w, hw, out_scale=1.0):
    """ECA gate f32 [N,C] from fixed-point pooled sums (int64 [N,C]) and the Conv1D taps w f32 [k]."""
    _chk(gap, "gap", STATS), _chk(w, "w", torch.float32)
    n, c = gap.shape
    gate = torch.empty((n, c), dtype=torch.float32, device=gap.device)
    _lib.check(_lib.lib().vip_eca_gate_f32(_p(gap), _p(w), _p(gate), n, c, w.numel(), 1.0 / hw, float(out_scale), _st()),
               "vip_eca_gate_f32")
    return gate


def pair_rows_weights(w, bias):
    """For ``conv2d_paired``: [Cout, 32] (K padded to 32) -> block-diagonal bf16 [2 Cout, 64] and the duplicated bias."""
    cout, kp = w.shape
    w2 = torch.zeros((2 * cout, 2 * kp), dtype=w.dtype, device=w.device)
    w2[:cout, :kp] = w
    w2[cout:, kp:] = w
    return w2.contiguous(), (None if bias is None else torch.cat([bias, bias]).contiguous())


def conv2d_paired(x, w_pair, bias_pair, ksize, stride, pad, act=None):
    """Small-Cin convolution (the 3-channel network input): explicit im2col rows of 32, then TWO output pixels per GEMM row
    (the [M, 32] matrix viewed as [M/2, 64], block-diagonal weights): full 64-wide k-blocks and half as many tiles."""
    _chk(x, "x"), _chk(w_pair, "w_pair")
    n, h, wd, c = x.shape
    ho = (h + 2 * pad - ksize) // stride + 1
    wo = (wd + 2 * pad - ksize) // stride + 1
    kp, cout = w_pair.shape[1] // 2, w_pair.shape[0] // 2
    m = n * ho * wo
    if m % 2:
        raise VipError("conv2d_paired needs an even number of output pixels")
    a = torch.empty((m, kp), dtype=BF16, device=x.device)
    _lib.check(_lib.lib().vip_im2col_bf16(_p(x), n, h, wd, c, ksize, stride, pad, ho, wo, _p(a), kp, _st()), "vip_im2col_bf16")
    y = gemm(a.view(m // 2, 2 * kp), w_pair, bias=bias_pair, act=act)
    return y.view(n, ho, wo, cout)


def zero_(t):
    """cudaMemsetAsync(0) on the current stream (accumulators of the fused row-statistics / pooling epilogues)."""
    _lib.check(_lib.lib().vip_memset_async(_p(t), 0, t.numel() * t.element_size(), _st()), "vip_memset_async")
    return t


def avgpool2_same(x):
    _chk(x, "x")
    n, h, w, c = x.shape
    out = torch.empty((n, (h + 1) // 2, (w + 1) // 2, c), dtype=BF16, device=x.device)
    _lib.check(_lib.lib().vip_avgpool2_same_bf16(_p(x), n, h, w, c, _p(out), _st()), "vip_avgpool2_same_bf16")
    return out


def global_avgpool(x, want_bf16=True, want_f32=False):
    """x bf16 [N,...,C] -> ([N,C] bf16 | None, [N,C] f32 | None)"""
    _chk(x, "x")
    n, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * c)
    ob = torch.empty((n, c), dtype=BF16, device=x.device) if want_bf16 else None
    of = torch.empty((n, c), dtype=torch.float32, device=x.device) if want_f32 else None
    _lib.check(_lib.lib().vip_global_avgpool_bf16(_p(x), n, hw, c, _p(ob), _p(of), _st()), "vip_global_avgpool_bf16")
    return ob, of


def scale_add_act(y, gate=None, shortcut=None, act=None, out=None):
    _chk(y, "y"), _chk(shortcut, "shortcut"), _chk(gate, "gate", torch.float32)
    n, c = y.shape[0], y.shape[-1]
    hw = y.numel() // (n * c)
    if out is None:
        out = torch.empty_like(y)
    _lib.check(_lib.lib().vip_scale_add_act_bf16(_p(y), _p(gate), _p(shortcut), _p(out), n, hw, c, ACT[act], _st()),
               "vip_scale_add_act_bf16")
    return out


def layernorm(x, gamma, beta, eps=1e-5, ln_next=None, next_eps=1e-5):
    """``ln_next`` (f32 [M, 2]) receives (mean, 1 / sqrt(var + next_eps)) of the output rows: the ``ln_stats`` of a
    LayerNorm folded into the contraction that consumes the output."""
    _chk(x, "x"), _chk(gamma, "gamma", torch.float32), _chk(beta, "beta", torch.float32)
    _chk(ln_next, "ln_next", torch.float32)
    c = x.shape[-1]
    out = torch.empty_like(x)
    _lib.check(_lib.lib().vip_layernorm_bf16(_p(x), _p(gamma), _p(beta), _p(out), _p(ln_next), float(next_eps),
                                             x.numel() // c, c, eps, _st()), "vip_layernorm_bf16")
    return out


def dwconv3x3(x, w, gelu=False, gap=None):
    """gap: int64 [N,C] zeroed fixed-point accumulator that receives the per-image channel sums of the output (fused SE
    squeeze)."""
    _chk(x, "x"), _chk(w, "w", torch.float32), _chk(gap, "gap", STATS)
    n, h, wd, c = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.lib().vip_dwconv3x3_bf16(_p(x), _p(w), _p(out), _p(gap), n, h, wd, c, int(gelu), _st()),
               "vip_dwconv3x3_bf16")
    return out


DW_ACT = {None: 0, "none": 0, "swish": 1, "gelu": 2, "relu": 3}


def dwconv(x, w, bias=None, ksize=3, stride=1, pad=None, act=None, gap=None):
    """Depthwise ksize x ksize convolution on NHWC bf16.  w f32 [k, k, C], bias f32 [C] | None.  ``pad`` = (top, left,
    bottom, right) explicit zero padding (default: symmetric k // 2); gap as for :func:`dwconv3x3`."""
    _chk(x, "x"), _chk(w, "w", torch.float32), _chk(bias, "bias", torch.float32), _chk(gap, "gap", STATS)
    n, h, wd, c = x.shape
    pt, pl, pb, pr = pad if pad is not None else (ksize // 2,) * 4
    ho, wo = (h + pt + pb - ksize) // stride + 1, (wd + pl + pr - ksize) // stride + 1
    out = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    _lib.check(_lib.lib().vip_dwconv_bf16(_p(x), _p(w), _p(bias), _p(out), _p(gap), n, h, wd, c, ksize, stride, pt, pl, ho, wo,
                                          DW_ACT[act], _st()), "vip_dwconv_bf16")
    return out


def layernorm_f32(x, gamma, beta, eps=1e-5):
    _chk(x, "x", torch.float32), _chk(gamma, "gamma", torch.float32), _chk(beta, "beta", torch.float32)
    m, c = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.lib().vip_layernorm_f32(_p(x), _p(gamma), _p(beta), _p(out), m, c, float(eps), _st()), "vip_layernorm_f32")
    return out


def maxpool3s2(x):
    _chk(x, "x")
    n, h, w, c = x.shape
    out = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), dtype=BF16, device=x.device)
    _lib.check(_lib.lib().vip_maxpool3s2_bf16(_p(x), _p(out), n, h, w, c, _st()), "vip_maxpool3s2_bf16")
    return out


def pad_crop(x, ho, wo, top=0, left=0, out=None):
    """out[n, y, x] = x[n, y - top, x - left] (zero outside): FitWindow padding / crop of an NHWC tensor [N,H,W,C] (bf16
    activations or int64 [N,H,W,3] row statistics records)."""
    if not x.is_cuda or not x.is_contiguous() or x.dim() != 4:
        raise VipError("pad_crop: x must be a contiguous CUDA [N,H,W,C] tensor")
    n, h, w, c = x.shape
    bpp = c * x.element_size()
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().vip_pad_crop(_p(x), n, h, w, bpp, _p(out), ho, wo, top, left, _st()), "vip_pad_crop")
    return out


def window_attention(qkv, q_global, rel_bias, B, H, W, C, ws, heads):
    _chk(qkv, "qkv"), _chk(q_global, "q_global"), _chk(rel_bias, "rel_bias", torch.float32)
    out = torch.empty((B * H * W, C), dtype=BF16, device=qkv.device)
    _lib.check(_lib.lib().vip_window_attention_bf16(_p(qkv), _p(q_global), _p(rel_bias), _p(out), B, H, W, C, ws, heads,
                                                    _st()), "vip_window_attention_bf16")
    return out


def head(feat_f32, w, b, sigmoid_head: bool, acc=None, acc_weight=1.0):
    """feat f32 [N,C]; w f32 [C,k]; returns probs f32 [N,k]; optionally acc[n] += acc_weight * P(synthetic) (f64)."""
    _chk(feat_f32, "feat", torch.float32), _chk(w, "w", torch.float32), _chk(b, "b", torch.float32)
    _chk(acc, "acc", torch.float64)
    n, c = feat_f32.shape
    k = w.shape[1]
    probs = torch.empty((n, k), dtype=torch.float32, device=feat_f32.device)
    _lib.check(_lib.lib().vip_head_f32(_p(feat_f32), _p(w), _p(b), _p(probs), _p(acc), float(acc_weight), n, c, k,
                                       int(sigmoid_head), _st()), "vip_head_f32")
    return probs


def cast_bf16(x_f32):
    _chk(x_f32, "x", torch.float32)
    out = torch.empty(x_f32.shape, dtype=BF16, device=x_f32.device)
    _lib.check(_lib.lib().vip_cast_f32_bf16(_p(x_f32), _p(out), x_f32.numel(), _st()), "vip_cast_f32_bf16")
    return out


def to_bf16(a, device):
    """numpy / host f32 array -> bf16 tensor on ``device``: the rounding runs in vip_cast_f32_bf16 on a CUDA device (weight
    packing at load time); on the CPU (shape-only model objects of the CPU tests) torch's own cast is the stand-in."""
    import numpy as np

    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device).contiguous()
    return cast_bf16(t) if t.is_cuda else t.to(BF16)


MLP_FUSED_SHAPES = {(96, 192), (64, 192)}


def mlp_fused(x, ln_stats, w1, colsum1, bias1, w2, bias2, next_eps=1e-5, ln_next=None, x_lo=None, want_lo=False):
    """x + fc2(gelu(fc1(LayerNorm(x)))) in one kernel (hidden activations stay on the SM); arguments as for the two
    ``gemm`` calls it replaces: w1 [hidden, C] gamma-scaled + colsum1 + bias1 (LayerNorm folded), w2 [C, hidden] + bias2.
    Only for (C, hidden) in MLP_FUSED_SHAPES.  ``x_lo`` / ``want_lo``: low planes of the two-plane residual stream;
    with want_lo returns (out, out_lo)."""
    _chk(x, "x"), _chk(w1, "w1"), _chk(w2, "w2"), _chk(ln_stats, "ln_stats", torch.float32), _chk(x_lo, "x_lo")
    _chk(ln_next, "ln_next", torch.float32)
    m, c = x.shape
    hidden = w1.shape[0]
    out = torch.empty_like(x)
    out_lo = lo_plane(m, c, x.device) if want_lo else None
    rc = _lib.lib().vip_mlp_fused_bf16(_p(x), _p(x_lo), m, c, hidden, _p(ln_stats), float(next_eps), _p(w1), w1.stride(0),
                                       _p(colsum1), _p(bias1), _p(w2), w2.stride(0), _p(bias2), _p(out), _p(out_lo),
                                       _p(ln_next), _st())
    _lib.check(rc, "vip_mlp_fused_bf16")
    return (out, out_lo) if want_lo else out


def scale_cast_bf16(x_fx, scale):
    """bf16(x * scale) of fixed-point pooled sums (int64, csrc/stats.cuh): the SE squeeze mean as a GEMM operand."""
    _chk(x_fx, "x", STATS)
    out = torch.empty(x_fx.shape, dtype=BF16, device=x_fx.device)
    _lib.check(_lib.lib().vip_scale_cast_fx_bf16(_p(x_fx), float(scale), _p(out), x_fx.numel(), _st()),
               "vip_scale_cast_fx_bf16")
    return out
