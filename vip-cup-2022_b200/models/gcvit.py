"""GCViT forward on the B200 kernels.  Mirrors ``models/gcvit`` of the reference (constructors gcvit.py:128-181, config
table gcvit.py:9-42): Keras-named / Keras-layout weights in, bf16 packed GEMM operands inside.

Layer -> kernel:
  Dense qkv / proj / fc1 / fc2, Conv 1x1     tcgen05 GEMM; bias, exact GELU, layer-scale gamma and the residual add are
                                             fused in the epilogue                                   (nn.gemm)
  Conv 3x3 stride 2 (stem proj, reduction)   im2col view + tcgen05 GEMM                               (nn.conv2d)
  WindowAttention (local / global query)     vip_window_attention_bf16, window partition/reverse folded into addressing
  LayerNormalization                         vip_layernorm_bf16
  DepthwiseConv 3x3 + GELU, SE, MaxPool      vip_dwconv3x3_bf16, vip_global_avgpool + 2 GEMMs, vip_maxpool3s2_bf16
  head                                       vip_layernorm -> vip_global_avgpool (f32) -> vip_head_f32
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import nn

CONFIGS = {  # models/gcvit/models/gcvit.py:9-42
    "xxtiny": dict(window_size=(7, 7, 14, 7), dim=64, depths=(2, 2, 6, 2), num_heads=(2, 4, 8, 16), mlp_ratio=3.0, layer_scale=None),
    "xtiny": dict(window_size=(7, 7, 14, 7), dim=64, depths=(3, 4, 6, 5), num_heads=(2, 4, 8, 16), mlp_ratio=3.0, layer_scale=None),
    "tiny": dict(window_size=(7, 7, 14, 7), dim=64, depths=(3, 4, 19, 5), num_heads=(2, 4, 8, 16), mlp_ratio=3.0, layer_scale=None),
    "small": dict(window_size=(7, 7, 14, 7), dim=96, depths=(3, 4, 19, 5), num_heads=(3, 6, 12, 24), mlp_ratio=2.0, layer_scale=1e-5),
    "base": dict(window_size=(7, 7, 14, 7), dim=128, depths=(3, 4, 19, 5), num_heads=(4, 8, 16, 32), mlp_ratio=2.0, layer_scale=1e-5),
}
KEEP_DIMS = [(False, False, False), (False, False), (True,), (True,)]  # gcvit.py:70
LN_EPS = 1e-5


FUSED_MLP = os.environ.get("VIP_FUSED_MLP", "1") != "0"   # level-0 MLPs through vip_mlp_fused_bf16
TWO_PLANE = os.environ.get("VIP_TWO_PLANE", "1") != "0"   # hi + lo bf16 planes for the block residual stream


def _pad_rows(a, mult=32):  # [N, K] -> N rounded up to `mult` with zero rows
    n = (a.shape[0] + mult - 1) // mult * mult
    if n == a.shape[0]:
        return a
    return np.concatenate([a, np.zeros((n - a.shape[0], a.shape[1]), a.dtype)], 0)


def _pad_cols(a, mult=32):
    k = (a.shape[1] + mult - 1) // mult * mult
    if k == a.shape[1]:
        return a
    return np.concatenate([a, np.zeros((a.shape[0], k - a.shape[1]), a.dtype)], 1)


class GCViT:
    def __init__(self, variant="tiny", input_shape=(224, 224, 3), num_classes=2, head_act="softmax", first_strides=2,
                 device="cuda"):
        if variant not in CONFIGS:
            raise ValueError(f"unknown GCViT variant {variant}")
        if head_act not in ("softmax", "sigmoid"):
            raise ValueError("head_act must be 'softmax' or 'sigmoid'")
        self.cfg, self.variant = CONFIGS[variant], variant
        self.input_shape, self.num_classes, self.head_act = tuple(input_shape), num_classes, head_act
        self.first_strides, self.device = first_strides, torch.device(device)
        self.name = f"GCViT{variant.capitalize()}"
        self.p = None

    # ---- weight packing ------------------------------------------------------------------------------------------
    def _bf(self, a):
        return nn.to_bf16(a, self.device)

    def _f32(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device).contiguous()

    def _dense(self, W, name, bias=True):
        k = np.asarray(W[name + "/kernel"], np.float32)
        k = k.reshape(-1, k.shape[-1])
        return self._bf(k.T), (self._f32(W[name + "/bias"]) if bias else None)

    def _dense_scaled(self, W, name, scale_name):
        """Dense followed by the layer-scale gamma (block.py:41-56,79-80), folded into the weights: g * (x W + b)."""
        k = np.asarray(W[name + "/kernel"], np.float32)
        k = k.reshape(-1, k.shape[-1])
        b = np.asarray(W[name + "/bias"], np.float32)
        if scale_name in W:
            g = np.asarray(W[scale_name], np.float32)
            k, b = k * g[None, :], b * g
        return self._bf(k.T), self._f32(b)

    def _dense_ln(self, W, name, norm):
        """Dense preceded by LayerNormalization ``norm``, packed for the folded-LN contraction (vip_epilogue_t.ln_stats):
        LN(x) W + b = rstd (x (gamma W) - mean colsum(gamma W)) + (beta W + b).  Returns (weights bf16 [out,in], bias',
        colsum of the bf16-rounded scaled weights)."""
        k = np.asarray(W[name + "/kernel"], np.float32)
        k = k.reshape(-1, k.shape[-1])                                   # (in, out)
        gamma, beta = np.asarray(W[norm + "/gamma"], np.float32), np.asarray(W[norm + "/beta"], np.float32)
        wg = self._bf((k * gamma[:, None]).T)                            # [out, in], rounded by vip_cast_f32_bf16
        # column sums of the ROUNDED weights (the algebra stays exact): bf16 -> f32 is a 16-bit shift, done on the host copy
        wr = wg.cpu().view(torch.int16).numpy().astype(np.uint16).astype(np.uint32) << 16
        colsum = self._f32(wr.view(np.float32).sum(1, dtype=np.float32))
        # beta W with a fixed summation order (numpy's pairwise f32 sum, like the f32 dot products of the Keras layer): a BLAS
        # matrix-vector product rounds differently with the number of host threads (torchrun sets OMP_NUM_THREADS=1), which
        # made 1-GPU and 2-GPU runs differ in the last bits
        bw = (beta[:, None] * k).sum(0, dtype=np.float32)
        bias = self._f32(bw + np.asarray(W[name + "/bias"], np.float32))
        return wg, bias, colsum

    def _conv3(self, W, name, bias=False):
        k = np.asarray(W[name + "/kernel"], np.float32)  # (3,3,Cin,Cout)
        w2 = _pad_cols(k.reshape(-1, k.shape[3]).T, 8)
        return self._bf(w2), (self._f32(W[name + "/bias"]) if bias else None)

    def _mbconv(self, W, n):
        c = W[n + "/conv/0/depthwise_kernel"].shape[2]
        fc0 = np.asarray(W[n + "/conv/2/fc/0/kernel"], np.float32)      # (C, C/4)
        fc2 = np.asarray(W[n + "/conv/2/fc/2/kernel"], np.float32)      # (C/4, C)
        return dict(dw=self._f32(np.asarray(W[n + "/conv/0/depthwise_kernel"]).reshape(3, 3, c)),
                    fc0=self._bf(_pad_rows(fc0.T)),                      # [hid_p, C]
                    fc2=self._bf(_pad_cols(fc2.T)),                      # [C, hid_p]
                    pw=self._bf(np.asarray(W[n + "/conv/3/kernel"], np.float32).reshape(c, c).T))

    def _reduce(self, W, n):
        d = self._mbconv(W, n)
        d.update(n1=(self._f32(W[n + "/norm1/gamma"]), self._f32(W[n + "/norm1/beta"])),
                 n2=(self._f32(W[n + "/norm2/gamma"]), self._f32(W[n + "/norm2/beta"])),
                 red=self._conv3(W, n + "/reduction")[0])
        return d

    def weight_shapes(self) -> dict:
        """Keras weight inventory (name -> shape): embedding.py:15, feature.py:20-22,57-59,87-101,128-139,
        attention.py:24-33, block.py:28-56, gcvit.py:79,88."""
        cfg, s = self.cfg, {}

        def lnorm(n, c):
            s[n + "/gamma"], s[n + "/beta"] = (c,), (c,)

        def mb(n, c):
            s[n + "/conv/0/depthwise_kernel"] = (3, 3, c, 1)
            s[n + "/conv/2/fc/0/kernel"], s[n + "/conv/2/fc/2/kernel"] = (c, int(c * 0.25)), (int(c * 0.25), c)
            s[n + "/conv/3/kernel"] = (1, 1, c, c)

        def reduce(n, c, keep):
            lnorm(n + "/norm1", c)
            mb(n, c)
            s[n + "/reduction/kernel"] = (3, 3, c, c if keep else 2 * c)
            lnorm(n + "/norm2", c if keep else 2 * c)

        c = cfg["dim"]
        s["patch_embed/proj/kernel"], s["patch_embed/proj/bias"] = (3, 3, 3, c), (c,)
        reduce("patch_embed/conv_down", c, True)
        for i, depth in enumerate(cfg["depths"]):
            ws, heads, hidden = cfg["window_size"][i], cfg["num_heads"][i], int(c * cfg["mlp_ratio"])
            for k in range(len(KEEP_DIMS[i])):
                mb(f"levels/{i}/q_global_gen/to_q_global/{k}", c)
            for j in range(depth):
                n, nq = f"levels/{i}/blocks/{j}", (2 if j % 2 else 3)
                lnorm(n + "/norm1", c)
                lnorm(n + "/norm2", c)
                s[n + "/attn/qkv/kernel"], s[n + "/attn/qkv/bias"] = (c, nq * c), (nq * c,)
                s[n + "/attn/relative_position_bias_table"] = ((2 * ws - 1) ** 2, heads)
                s[n + "/attn/proj/kernel"], s[n + "/attn/proj/bias"] = (c, c), (c,)
                s[n + "/mlp/fc1/kernel"], s[n + "/mlp/fc1/bias"] = (c, hidden), (hidden,)
                s[n + "/mlp/fc2/kernel"], s[n + "/mlp/fc2/bias"] = (hidden, c), (c,)
                if cfg["layer_scale"] is not None:
                    s[n + "/gamma1"], s[n + "/gamma2"] = (c,), (c,)
            if i < 3:
                reduce(f"levels/{i}/downsample", c, False)
                c *= 2
        lnorm("norm", c)
        s["head/kernel"], s["head/bias"] = (c, self.num_classes), (self.num_classes,)
        return s

    def init_random(self, seed=0):
        """Random initialisation like the Keras constructor (glorot-scale kernels, unit LayerNorm, zero biases)."""
        rng = np.random.default_rng(seed)
        W = {}
        for name, shp in self.weight_shapes().items():
            leaf = name.rsplit("/", 1)[1]
            if leaf in ("kernel", "depthwise_kernel"):
                fan_in = int(np.prod(shp[:-1])) if leaf == "kernel" else 9
                W[name] = (rng.standard_normal(shp) * np.sqrt(1.0 / fan_in)).astype(np.float32)
            elif leaf == "gamma":
                W[name] = np.ones(shp, np.float32)
            elif leaf in ("gamma1", "gamma2"):
                W[name] = np.full(shp, 0.1, np.float32)
            elif leaf == "relative_position_bias_table":
                W[name] = (rng.standard_normal(shp) * 0.02).astype(np.float32)
            else:
                W[name] = np.zeros(shp, np.float32)
        return self.load_weights(W)

    def load_weights(self, W: dict):
        cfg, p = self.cfg, {}
        p["proj"] = self._conv3(W, "patch_embed/proj", bias=True)
        p["proj_pair"] = nn.pair_rows_weights(*p["proj"])
        p["conv_down"] = self._reduce(W, "patch_embed/conv_down")
        for i, depth in enumerate(cfg["depths"]):
            ws, heads = cfg["window_size"][i], cfg["num_heads"][i]
            for k in range(len(KEEP_DIMS[i])):
                p[f"q{i}_{k}"] = self._mbconv(W, f"levels/{i}/q_global_gen/to_q_global/{k}")
            for j in range(depth):
                n = f"levels/{i}/blocks/{j}"
                # [(2ws-1)^2, heads] -> [heads, (2ws-1)^2]; the kernel gathers by relative position (attention.py:39-50)
                table = np.asarray(W[n + "/attn/relative_position_bias_table"], np.float32).T
                p[f"b{i}_{j}"] = dict(
                    n1=(self._f32(W[n + "/norm1/gamma"]), self._f32(W[n + "/norm1/beta"])),
                    n2=(self._f32(W[n + "/norm2/gamma"]), self._f32(W[n + "/norm2/beta"])),
                    qkv=self._dense_ln(W, n + "/attn/qkv", n + "/norm1"),
                    proj=self._dense_scaled(W, n + "/attn/proj", n + "/gamma1"),
                    fc1=self._dense_ln(W, n + "/mlp/fc1", n + "/norm2"),
                    fc2=self._dense_scaled(W, n + "/mlp/fc2", n + "/gamma2"),
                    rel=self._f32(table))
            if i < 3:
                p[f"down{i}"] = self._reduce(W, f"levels/{i}/downsample")
        p["norm"] = (self._f32(W["norm/gamma"]), self._f32(W["norm/beta"]))
        p["head_w"], p["head_b"] = self._f32(W["head/kernel"]), self._f32(W["head/bias"])
        self.p = p
        return self

    # ---- layers --------------------------------------------------------------------------------------------------
    def _mb_apply(self, x, d):
        """x + [pad1 -> DW3x3 -> GELU -> SE -> Conv1x1](x)   (feature.py:105-109,144-150)"""
        b, h, w, c = x.shape
        gap = nn.zero_(torch.empty((b, c), dtype=nn.STATS, device=x.device))
        y = nn.dwconv3x3(x, d["dw"], gelu=True, gap=gap)        # SE squeeze accumulated by the same kernel
        pooled = nn.scale_cast_bf16(gap, 1.0 / (h * w))
        hid = nn.gemm(pooled, d["fc0"], act="gelu")
        gate = nn.gemm(hid, d["fc2"], act="sigmoid", out_dtype=torch.float32)
        if (h * w) % 128 == 0 and d["pw"].shape[1] == c:
            # the gate scales the input channels of the 1x1 convolution: fold it into per-image copies of the (C x C)
            # weights instead of a read-modify-write pass over y (whole 128-row tiles per image required)
            wg = nn.scale_weights(d["pw"], gate)
            return nn.gemm_grouped(y.view(-1, c), wg, h * w, residual=x.view(-1, c)).view(b, h, w, c)
        y = nn.scale_add_act(y, gate, None, out=y)
        return nn.gemm(y.view(-1, c), d["pw"], residual=x.view(-1, c)).view(b, h, w, c)

    def _reduce_apply(self, x, d, stride, ln_next=None):
        x = nn.layernorm(x, *d["n1"], eps=LN_EPS)
        x = self._mb_apply(x, d)
        x = nn.conv2d(x, d["red"], None, ksize=3, stride=stride, pad=1)
        return nn.layernorm(x, *d["n2"], eps=LN_EPS, ln_next=ln_next, next_eps=LN_EPS)

    def _block(self, x, x_lo, ln_in, d, heads, ws, q_global, ln_mid, ln_out, rec):
        """GCViTBlock (block.py:60-81).  Both LayerNorms are folded into the contraction that consumes them: ``ln_in`` holds
        (mean, 1 / sigma) of the rows of x; the proj / fc2 epilogues accumulate row statistics records of their outputs
        (``rec`` [2, tokens, 3], integer atomics) that vip_row_stats_finalize turns into ``ln_mid`` / ``ln_out``; the fused
        MLP kernel owns whole rows and writes ``ln_out`` itself.  The residual stream
        is carried in two bf16 planes (x = hi + lo, 16 mantissa bits): the hi plane is the operand of qkv / fc1, the lo
        plane only meets the proj / fc2 epilogues, so 2 x depth bf16 roundings of the running sum do not pile up
        (SURVEY.md 7 "fp32 residual stream if needed"; measured: 1.5e-2 -> 4e-3 logit error on GCViT-small)."""
        b, h, w, c = x.shape
        x2 = x.view(-1, c)
        wq, bq, cq = d["qkv"]
        qkv = nn.gemm(x2, wq, bias=bq, ln_stats=ln_in, ln_colsum=cq)
        a = nn.window_attention(qkv, q_global, d["rel"], b, h, w, c, ws, heads)
        lo1 = nn.lo_plane(x2.shape[0], c, x2.device) if TWO_PLANE else None
        x2 = nn.gemm(a, *d["proj"], residual=x2, row_stats=rec[0], residual_lo=x_lo, out_lo=lo1, row_pivot=ln_in)
        nn.finalize_stats(rec[0], c, LN_EPS, out=ln_mid)
        w1, b1, c1 = d["fc1"]
        if FUSED_MLP and (c, w1.shape[0]) in nn.MLP_FUSED_SHAPES:
            # narrow levels: both contractions in one kernel, the [tokens, hidden] tensor never reaches HBM
            res = nn.mlp_fused(x2, ln_mid, w1, c1, b1, *d["fc2"], next_eps=LN_EPS, ln_next=ln_out, x_lo=lo1, want_lo=TWO_PLANE)
            x2, lo2 = res if TWO_PLANE else (res, None)
        else:
            hdn = nn.gemm(x2, w1, bias=b1, act="gelu", ln_stats=ln_mid, ln_colsum=c1)
            lo2 = nn.lo_plane(x2.shape[0], c, x2.device) if TWO_PLANE else None
            x2 = nn.gemm(hdn, *d["fc2"], residual=x2, row_stats=rec[1], residual_lo=lo1, out_lo=lo2, row_pivot=ln_mid)
            nn.finalize_stats(rec[1], c, LN_EPS, out=ln_out)
        return x2.view(b, h, w, c), lo2

    def features(self, x, taps=None):
        p, cfg = self.p, self.cfg
        if p is None:
            raise RuntimeError("load_weights() first")
        if (x.shape[0] * ((x.shape[1] - 1) // 2 + 1) * ((x.shape[2] - 1) // 2 + 1)) % 2 == 0 and p["proj"][0].shape[1] == 32:
            x = nn.conv2d_paired(x, *p["proj_pair"], ksize=3, stride=2, pad=1)
        else:
            x = nn.conv2d(x, *p["proj"], ksize=3, stride=2, pad=1)
        nimg = x.shape[0]

        def level_stats(i, tokens):
            """Per level: (mean, 1 / sigma) slots [1 + 2 depth, tokens, 2] and one zeroed arena of row statistics records
            [2 depth, tokens, 3] for the epilogues that accumulate them."""
            depth = cfg["depths"][i]
            return (torch.empty((1 + 2 * depth, tokens, 2), dtype=torch.float32, device=x.device),
                    nn.row_stats_buffer(2 * depth, tokens, device=x.device))

        def fit(v, ws):   # FitWindow (feature.py:240-249): the size a map is padded to, and the top / left share of the pad
            vp = (v + ws - 1) // ws * ws
            return vp, (vp - v) // 2

        def reduce_into_level(x_in, d, stride, i):
            """ReduceSize whose output feeds level i: returns (x, level statistics arena); ln[0] = (mean, 1 / sigma) of x's
            rows, laid out for the FitWindow-padded map when the level pads (padded rows are all zero: the folded
            LayerNorm then yields beta W + b for them, exactly what LayerNorm of a zero row gives)."""
            ho, wo = (x_in.shape[1] + 2 - 3) // stride + 1, (x_in.shape[2] + 2 - 3) // stride + 1
            ws = cfg["window_size"][i]
            (hp, _), (wp, _) = fit(ho, ws), fit(wo, ws)
            arena = level_stats(i, nimg * hp * wp)
            if (hp, wp) == (ho, wo):
                return self._reduce_apply(x_in, d, stride, ln_next=arena[0][0]), arena
            ln_in = torch.empty((nimg * ho * wo, 2), dtype=torch.float32, device=x_in.device)
            x_out = self._reduce_apply(x_in, d, stride, ln_next=ln_in)
            nn.pad_crop(ln_in.view(nimg, ho, wo, 2), hp, wp, fit(ho, ws)[1], fit(wo, ws)[1], out=arena[0][0].view(nimg, hp, wp, 2))
            return x_out, arena

        x, st = reduce_into_level(x, p["conv_down"], self.first_strides, 0)
        if taps is not None:
            taps["stem"] = x
        for i, depth in enumerate(cfg["depths"]):
            ws, heads = cfg["window_size"][i], cfg["num_heads"][i]
            b, h, w, c = x.shape
            (hp, top), (wp, left) = fit(h, ws), fit(w, ws)
            padded = (hp, wp) != (h, w)
            if padded:
                # FitWindow: zero padding on both sides; the padded tokens take part in attention like any other token, and the
                # level's output is the TOP-LEFT h x w corner of the padded map (level.py:49,61 -- mirrored as it is)
                x = nn.pad_crop(x, hp, wp, top, left)
            q = x
            for k, keep in enumerate(KEEP_DIMS[i]):
                q = self._mb_apply(q, p[f"q{i}_{k}"])
                if not keep:
                    q = nn.maxpool3s2(q)
            if q.shape[1] != ws or q.shape[2] != ws:
                raise nn.VipError(f"GCViT level {i}: global query map {q.shape[1]}x{q.shape[2]} does not match window {ws} "
                                  f"(input {self.input_shape}; attention.py:65 has the same requirement)")
            q = q.view(b, ws * ws, c)
            x_lo = None   # the level starts from a LayerNorm output: low plane = 0
            ln, rec = st
            for j in range(depth):
                x, x_lo = self._block(x, x_lo, ln[2 * j], p[f"b{i}_{j}"], heads, ws, q if j % 2 else None, ln[2 * j + 1],
                                      ln[2 * j + 2], rec[2 * j: 2 * j + 2])
            if padded:
                x = nn.pad_crop(x, h, w, 0, 0)
            if i < 3:
                x, st = reduce_into_level(x, p[f"down{i}"], 2, i + 1)
            if taps is not None:
                taps[f"level{i}"] = x
        return x

    def forward(self, x, acc=None, acc_weight=1.0, taps=None):
        if x.dtype == torch.float32:
            x = nn.cast_bf16(x)
        f = self.features(x, taps)
        f = nn.layernorm(f, *self.p["norm"], eps=LN_EPS)
        _, feat = nn.global_avgpool(f, want_bf16=False, want_f32=True)
        if taps is not None:
            taps["feat"] = feat
        return nn.head(feat, self.p["head_w"], self.p["head_b"], self.head_act == "sigmoid", acc, acc_weight)

    __call__ = forward


def GCViTXXTiny(**kw):
    return GCViT("xxtiny", **kw)


def GCViTXTiny(**kw):
    return GCViT("xtiny", **kw)


def GCViTTiny(**kw):
    return GCViT("tiny", **kw)


def GCViTSmall(**kw):
    return GCViT("small", **kw)


def GCViTBase(**kw):
    return GCViT("base", **kw)
