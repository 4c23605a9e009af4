"""EfficientNetV2-T / EfficientNetV1-B4 forward on the B200 kernels.  Mirrors the builder of the reference
(``models/keras_cv_attention_models/efficientnet/efficientnet_v2.py:47-193`` and ``efficientnet_v1.py:9-36,68-73``):
Keras-named / Keras-layout weights in, BatchNorm folded into bf16 GEMM / f32 depthwise weights at load.

Layer -> kernel:
  stem Conv 3x3 stride 2 (3 input channels)   explicit im2col rows + tcgen05 GEMM, BN bias + swish in the epilogue
  Fused-MBConv 3x3 (stride 1 | 2)             implicit-GEMM convolution (im2col-mode TMA), swish epilogue
  1x1 expand / project / post convolutions    tcgen05 GEMM on the NHWC activation; the block's shortcut add in the epilogue
  DepthwiseConv k x k + BN + swish            vip_dwconv_bf16 (TF 'SAME' asymmetric padding for V1), SE squeeze fused
  SE (biased 1x1 convs, swish / sigmoid)      fixed-point pooled sums -> two small GEMMs -> vip_scale_add_act_bf16
  head GAP -> Dense                           vip_global_avgpool (f32) -> vip_head_f32
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import nn


def make_divisible(vv, divisor=4, min_value=None, limit_round_down=0.9):  # common_layers.py:398-406
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(vv + divisor / 2) // divisor * divisor)
    if new_v < limit_round_down * vv:
        new_v += divisor
    return new_v


def config(variant):
    """Builder arguments of the two registry members (efficientnet_v2.py:268-275; efficientnet_v1.py:9-36,68-73)."""
    if variant == "v2t":
        return dict(expands=[1, 4, 4, 4, 6, 6], out_channels=[24, 40, 48, 104, 128, 208], depthes=[2, 4, 4, 6, 9, 14],
                    strides=[1, 2, 2, 2, 1, 2], se_ratios=[0, 0, 0, 0.25, 0.25, 0.25], kernel_sizes=[3] * 6,
                    first_conv_filter=24, output_conv_filter=1024, is_torch_mode=True)
    if variant == "v1b4":
        width, depth = 1.4, 1.8
        return dict(expands=[1, 6, 6, 6, 6, 6, 6], out_channels=[c * width for c in [16, 24, 40, 80, 112, 192, 320]],
                    depthes=[int(math.ceil(np.float32(d) * np.float32(depth))) for d in [1, 2, 2, 3, 3, 4, 1]],
                    strides=[1, 2, 2, 2, 1, 2, 1], se_ratios=[0.25] * 7, kernel_sizes=[3, 3, 5, 3, 5, 5, 3],
                    first_conv_filter=32 * width, output_conv_filter=1280 * width, is_torch_mode=False)
    raise ValueError(f"unknown EfficientNet variant {variant}")


def plan(variant):
    """The block list of the builder loop (efficientnet_v2.py:160-181)."""
    cfg = config(variant)
    stem = make_divisible(cfg["first_conv_filter"], 8)
    blocks, pre = [], stem
    for sid, (e, oc, d, s, se, k) in enumerate(zip(cfg["expands"], cfg["out_channels"], cfg["depthes"], cfg["strides"],
                                                   cfg["se_ratios"], cfg["kernel_sizes"])):
        out = make_divisible(oc, 8)
        for bid in range(d):
            stride = s if bid == 0 else 1
            hidden = make_divisible(pre * e, 8)
            red = make_divisible(hidden * (se / e), 1, limit_round_down=0.9) if se > 0 else 0
            blocks.append(dict(name=f"stack_{sid}_block{bid}_", cin=pre, cout=out, hidden=hidden, stride=stride, expand=e,
                               kernel=k, fused=(se == 0), se=red, shortcut=(out == pre and stride == 1)))
            pre = out
    return dict(stem=stem, blocks=blocks, post=make_divisible(cfg["output_conv_filter"], 8), last=pre,
                torch_mode=cfg["is_torch_mode"], bn_eps=1e-5 if cfg["is_torch_mode"] else 1e-3)


def same_pad(n, k, s):
    """TF 'SAME': (before, after) zero padding."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _up(v, m):
    return (v + m - 1) // m * m


class EfficientNet:
    def __init__(self, variant="v2t", input_shape=(200, 200, 3), num_classes=2, classifier_activation="softmax",
                 first_strides=2, device="cuda"):
        if classifier_activation not in ("softmax", "sigmoid"):
            raise ValueError("classifier_activation must be 'softmax' or 'sigmoid'")
        self.variant, self.plan = variant, plan(variant)
        self.input_shape, self.num_classes, self.head_act = tuple(input_shape), num_classes, classifier_activation
        self.first_strides, self.device = first_strides, torch.device(device)
        self.name = {"v2t": "EfficientNetV2T", "v1b4": "EfficientNetV1B4"}[variant]
        self.p = None

    def weight_shapes(self) -> dict:
        """Keras weight inventory (layer names of efficientnet_v2.py:66-108,151-186, common_layers.py:203-209,238-248,323-328)."""
        p, s = self.plan, {}

        def bnorm(n, c):
            for q in ("gamma", "beta", "moving_mean", "moving_variance"):
                s[f"{n}bn/{q}"] = (c,)

        s["stem_conv/kernel"] = (3, 3, 3, p["stem"])
        bnorm("stem_", p["stem"])
        for b in p["blocks"]:
            n = b["name"]
            if b["expand"] != 1:
                k = 3 if b["fused"] else 1
                s[n + "sortcut_conv/kernel"] = (k, k, b["cin"], b["hidden"])
                bnorm(n + "sortcut_", b["hidden"])
            if not b["fused"]:
                s[n + "MB_dw_/depthwise_kernel"] = (b["kernel"], b["kernel"], b["hidden"], 1)
                bnorm(n + "MB_dw_", b["hidden"])
            if b["se"] > 0:
                s[n + "se_1_conv/kernel"], s[n + "se_1_conv/bias"] = (1, 1, b["hidden"], b["se"]), (b["se"],)
                s[n + "se_2_conv/kernel"], s[n + "se_2_conv/bias"] = (1, 1, b["se"], b["hidden"]), (b["hidden"],)
            if b["fused"] and b["expand"] == 1:
                s[n + "fu_conv/kernel"] = (3, 3, b["cin"], b["cout"])
                bnorm(n + "fu_", b["cout"])
            else:
                s[n + "MB_pw_conv/kernel"] = (1, 1, b["hidden"], b["cout"])
                bnorm(n + "MB_pw_", b["cout"])
        s["post_conv/kernel"] = (1, 1, p["last"], p["post"])
        bnorm("post_", p["post"])
        s["predictions/kernel"], s["predictions/bias"] = (p["post"], self.num_classes), (self.num_classes,)
        return s

    def init_random(self, seed=0):
        rng = np.random.default_rng(seed)
        W = {}
        for name, shp in self.weight_shapes().items():
            leaf = name.rsplit("/", 1)[1]
            if leaf in ("kernel", "depthwise_kernel"):
                fan_in = int(np.prod(shp[:-1])) if leaf == "kernel" else shp[0] * shp[1]
                W[name] = (rng.standard_normal(shp) * np.sqrt(2.0 / fan_in)).astype(np.float32)
            elif leaf in ("gamma", "moving_variance"):
                W[name] = np.ones(shp, np.float32) * (0.3 if name.endswith("MB_pw_bn/gamma") or name.endswith("fu_bn/gamma") else 1.0)
            else:
                W[name] = np.zeros(shp, np.float32)
        return self.load_weights(W)

    # ---- weight packing ------------------------------------------------------------------------------------------
    def _f32(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device).contiguous()

    def _bf(self, a):
        return nn.cast_bf16(self._f32(a)) if self.device.type == "cuda" else self._f32(a).to(torch.bfloat16)

    def _bn_scale(self, W, n):
        g, b = np.asarray(W[n + "bn/gamma"], np.float32), np.asarray(W[n + "bn/beta"], np.float32)
        m, v = np.asarray(W[n + "bn/moving_mean"], np.float32), np.asarray(W[n + "bn/moving_variance"], np.float32)
        s = g / np.sqrt(v + np.float32(self.plan["bn_eps"]))
        return s, b - m * s

    def _conv_bn(self, W, n):
        """conv (kh,kw,Cin,Cout) + BN -> bf16 [Cout, Kp] (K order r,s,c; K rounded up to 8) and f32 bias."""
        k = np.asarray(W[n + "conv/kernel"], np.float32)
        s, bias = self._bn_scale(W, n)
        w2 = (k * s[None, None, None, :]).reshape(-1, k.shape[3]).T
        wp = np.zeros((w2.shape[0], _up(w2.shape[1], 8)), np.float32)
        wp[:, : w2.shape[1]] = w2
        return self._bf(wp), self._f32(bias)

    def load_weights(self, W: dict):
        p = {"stem": self._conv_bn(W, "stem_")}
        for b in self.plan["blocks"]:
            n, d = b["name"], {}
            if b["expand"] != 1:
                d["expand"] = self._conv_bn(W, n + "sortcut_")
            if not b["fused"]:
                s, bias = self._bn_scale(W, n + "MB_dw_")
                k = np.asarray(W[n + "MB_dw_/depthwise_kernel"], np.float32)[:, :, :, 0]
                d["dw"] = (self._f32(k * s[None, None, :]), self._f32(bias))
            if b["se"] > 0:
                rp = _up(b["se"], 32)     # reduction width padded with zero rows / columns to a GEMM-friendly size
                k1 = np.asarray(W[n + "se_1_conv/kernel"], np.float32)[0, 0]              # (hidden, red)
                k2 = np.asarray(W[n + "se_2_conv/kernel"], np.float32)[0, 0]              # (red, hidden)
                w1, b1 = np.zeros((rp, b["hidden"]), np.float32), np.zeros((rp,), np.float32)
                w1[: b["se"]], b1[: b["se"]] = k1.T, np.asarray(W[n + "se_1_conv/bias"], np.float32)
                w2 = np.zeros((b["hidden"], rp), np.float32)
                w2[:, : b["se"]] = k2.T
                d["se1"], d["se2"] = (self._bf(w1), self._f32(b1)), (self._bf(w2), self._f32(W[n + "se_2_conv/bias"]))
            d["out"] = self._conv_bn(W, n + ("fu_" if b["fused"] and b["expand"] == 1 else "MB_pw_"))
            p[n] = d
        p["post"] = self._conv_bn(W, "post_")
        p["head_w"], p["head_b"] = self._f32(W["predictions/kernel"]), self._f32(W["predictions/bias"])
        self.p = p
        return self

    # ---- forward ---------------------------------------------------------------------------------------------------
    def _block(self, x, b, d):
        """inverted_residual_block (efficientnet_v2.py:47-108)."""
        tm, inp = self.plan["torch_mode"], x
        nimg = x.shape[0]
        if b["fused"] and b["expand"] != 1:
            x = nn.conv2d(x, *d["expand"], ksize=3, stride=b["stride"], pad=1, act="swish")
        elif b["expand"] != 1:
            x = nn.conv2d(x, *d["expand"], act="swish")
        if not b["fused"]:
            k, s = b["kernel"], b["stride"]
            if tm:
                pad = (k // 2,) * 4
            else:
                (pt, pb), (pl, pr) = same_pad(x.shape[1], k, s), same_pad(x.shape[2], k, s)
                pad = (pt, pl, pb, pr)
            gap = nn.zero_(torch.empty((nimg, b["hidden"]), dtype=nn.STATS, device=x.device)) if b["se"] > 0 else None
            x = nn.dwconv(x, *d["dw"], ksize=k, stride=s, pad=pad, act="swish", gap=gap)
            if b["se"] > 0:
                # se_module (common_layers.py:311-332): mean -> 1x1 conv + bias, swish -> 1x1 conv + bias, sigmoid -> multiply
                pooled = nn.scale_cast_bf16(gap, 1.0 / (x.shape[1] * x.shape[2]))
                hid = nn.gemm(pooled, *d["se1"], act="swish")
                gate = nn.gemm(hid, *d["se2"], act="sigmoid", out_dtype=torch.float32)
                x = nn.scale_add_act(x, gate, None, out=x)
        if b["fused"] and b["expand"] == 1:
            # the activation comes BEFORE the shortcut add (swish(BN(conv)) + x): the order of the contraction's epilogue
            return nn.conv2d(x, *d["out"], ksize=3, stride=b["stride"], pad=1, act="swish", residual=inp if b["shortcut"] else None)
        return nn.conv2d(x, *d["out"], residual=inp if b["shortcut"] else None)

    def features(self, x, taps=None):
        p, pl = self.p, self.plan
        if p is None:
            raise RuntimeError("load_weights() first")
        s = self.first_strides
        if pl["torch_mode"]:
            x = nn.conv2d(x, *p["stem"], ksize=3, stride=s, pad=1, act="swish")
        else:
            (pt, _), (pleft, _) = same_pad(x.shape[1], 3, s), same_pad(x.shape[2], 3, s)
            if pt != pleft:
                raise nn.VipError("non-square TF 'SAME' stem padding is not supported")
            x = nn.conv2d(x, *p["stem"], ksize=3, stride=s, pad=pt, act="swish", out_hw=(-(-x.shape[1] // s), -(-x.shape[2] // s)))
        if taps is not None:
            taps["stem"] = x
        last = None
        for b in pl["blocks"]:
            stack = b["name"].split("_")[1]
            if taps is not None and last is not None and stack != last:
                taps[f"stack{last}"] = x
            last = stack
            x = self._block(x, b, p[b["name"]])
        if taps is not None:
            taps[f"stack{last}"] = x
        return nn.conv2d(x, *p["post"], act="swish")

    def forward(self, x, acc=None, acc_weight=1.0, taps=None):
        if x.dtype == torch.float32:
            x = nn.cast_bf16(x)
        f = self.features(x, taps)
        _, feat = nn.global_avgpool(f, want_bf16=False, want_f32=True)
        if taps is not None:
            taps["feat"] = feat
        return nn.head(feat, self.p["head_w"], self.p["head_b"], self.head_act == "sigmoid", acc, acc_weight)

    __call__ = forward
