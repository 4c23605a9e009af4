"""CPU ORACLE (test infrastructure, not shipped) -- EfficientNetV2-T / EfficientNetV1-B4 forward in PyTorch fp32, restating
``models/keras_cv_attention_models/efficientnet`` of the reference on Keras-layout, Keras-named weights.

Reference -> here:
  EfficientNetV2 builder        efficientnet_v2.py:111-193                       -> :func:`plan`, :func:`forward`
  inverted_residual_block       efficientnet_v2.py:47-108 (fused / MBConv, SE)   -> :func:`block`
  EfficientNetV2T               efficientnet_v2.py:268-275 (is_torch_mode: ZeroPadding + 'VALID', BN eps 1e-5)
  EfficientNetV1 / V1B4         efficientnet_v1.py:9-36, 68-73 (width 1.4, depth 1.8; TF 'SAME' padding, BN eps 1e-3)
  conv2d_no_bias, batchnorm_with_activation, se_module, make_divisible, output_block
                                common_layers.py:230-248, 190-212, 311-332, 398-406, 271-283

Weights: Conv2D ``kernel`` (kh,kw,Cin,Cout) [+ ``bias`` in the SE convs], DepthwiseConv2D ``depthwise_kernel`` (kh,kw,C,1),
BatchNormalization ``gamma,beta,moving_mean,moving_variance``, Dense ``kernel`` (in,out) + ``bias``.  Parity status: unpinned
against real Keras (TensorFlow is not installable offline); known answers = the parameter counts of the kecam model table
(EfficientNetV2T 13.6 M, EfficientNetV1B4 19.3 M with the 1000-class head) and the stage shapes."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def make_divisible(vv, divisor=4, min_value=None, limit_round_down=0.9):  # common_layers.py:398-406
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(vv + divisor / 2) // divisor * divisor)
    if new_v < limit_round_down * vv:
        new_v += divisor
    return new_v


def config(variant):
    """Builder arguments of the two registry members (efficientnet_v2.py:268-275, efficientnet_v1.py:9-36,68-73)."""
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
    raise ValueError(variant)


def plan(variant):
    """The block list the builder loop produces (efficientnet_v2.py:160-181): dicts with name, cin, cout, hidden, stride,
    expand, kernel, fused, se (reduction channels or 0), shortcut."""
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


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).float()


def conv(x, W, name, k, stride, torch_mode):
    """conv2d_no_bias(padding='same' for k > 1): ZeroPadding(k // 2) + 'VALID' in torch mode, TF 'SAME' otherwise.  NCHW."""
    w = _t(W[name + "conv/kernel"]).permute(3, 2, 0, 1).contiguous()
    if k > 1:
        if torch_mode:
            x = F.pad(x, (k // 2,) * 4)
        else:
            (pt, pb), (pl, pr) = same_pad(x.shape[2], k, stride), same_pad(x.shape[3], k, stride)
            x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, w, stride=stride)


def bn(x, W, name, eps, act):
    g, b = _t(W[name + "bn/gamma"]), _t(W[name + "bn/beta"])
    m, v = _t(W[name + "bn/moving_mean"]), _t(W[name + "bn/moving_variance"])
    x = (x - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + eps) * g[None, :, None, None] + b[None, :, None, None]
    return F.silu(x) if act else x


def block(x, W, b, torch_mode, eps):
    n, inp = b["name"], x
    if b["fused"] and b["expand"] != 1:
        x = bn(conv(x, W, n + "sortcut_", 3, b["stride"], torch_mode), W, n + "sortcut_", eps, True)
    elif b["expand"] != 1:
        x = bn(conv(x, W, n + "sortcut_", 1, 1, torch_mode), W, n + "sortcut_", eps, True)
    if not b["fused"]:
        k, s = b["kernel"], b["stride"]
        dw = _t(W[n + "MB_dw_/depthwise_kernel"]).permute(2, 3, 0, 1).contiguous()
        if torch_mode:
            xp = F.pad(x, (k // 2,) * 4)
        else:
            (pt, pb), (pl, pr) = same_pad(x.shape[2], k, s), same_pad(x.shape[3], k, s)
            xp = F.pad(x, (pl, pr, pt, pb))
        x = bn(F.conv2d(xp, dw, stride=s, groups=x.shape[1]), W, n + "MB_dw_", eps, True)
    if b["se"] > 0:
        se = x.mean(dim=(2, 3))
        se = F.silu(se @ _t(W[n + "se_1_conv/kernel"])[0, 0] + _t(W[n + "se_1_conv/bias"]))
        se = torch.sigmoid(se @ _t(W[n + "se_2_conv/kernel"])[0, 0] + _t(W[n + "se_2_conv/bias"]))
        x = x * se[:, :, None, None]
    if b["fused"] and b["expand"] == 1:
        x = bn(conv(x, W, n + "fu_", 3, b["stride"], torch_mode), W, n + "fu_", eps, True)
    else:
        x = bn(conv(x, W, n + "MB_pw_", 1, 1, torch_mode), W, n + "MB_pw_", eps, False)
    return inp + x if b["shortcut"] else x


def forward(x_nhwc, W, variant="v2t", head_act="softmax", return_logits=False, first_strides=2, taps=None):
    p = plan(variant)
    tm, eps = p["torch_mode"], p["bn_eps"]
    with torch.no_grad():
        x = _t(x_nhwc).permute(0, 3, 1, 2)
        x = bn(conv(x, W, "stem_", 3, first_strides, tm), W, "stem_", eps, True)
        if taps is not None:
            taps["stem"] = x.permute(0, 2, 3, 1).numpy().copy()
        last_stack = None
        for b in p["blocks"]:
            stack = b["name"].split("_")[1]
            if taps is not None and last_stack is not None and stack != last_stack:
                taps[f"stack{last_stack}"] = x.permute(0, 2, 3, 1).numpy().copy()
            last_stack = stack
            x = block(x, W, b, tm, eps)
        if taps is not None:
            taps[f"stack{last_stack}"] = x.permute(0, 2, 3, 1).numpy().copy()
        x = bn(conv(x, W, "post_", 1, 1, tm), W, "post_", eps, True)
        feat = x.mean(dim=(2, 3))
        if taps is not None:
            taps["feat"] = feat.numpy().copy()
        logits = feat @ _t(W["predictions/kernel"]) + _t(W["predictions/bias"])
        if return_logits:
            return logits.numpy()
        return (torch.softmax(logits, -1) if head_act == "softmax" else torch.sigmoid(logits)).numpy()


def weight_shapes(variant="v2t", num_classes=2) -> dict:
    p, s = plan(variant), {}

    def bnorm(n, c):
        for q in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"{n}bn/{q}"] = (c,)

    s["stem_conv/kernel"] = (3, 3, 3, p["stem"])
    bnorm("stem_", p["stem"])
    for b in p["blocks"]:
        n = b["name"]
        if b["fused"] and b["expand"] != 1:
            s[n + "sortcut_conv/kernel"] = (3, 3, b["cin"], b["hidden"])
            bnorm(n + "sortcut_", b["hidden"])
        elif b["expand"] != 1:
            s[n + "sortcut_conv/kernel"] = (1, 1, b["cin"], b["hidden"])
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
    s["predictions/kernel"], s["predictions/bias"] = (p["post"], num_classes), (num_classes,)
    return s


def random_weights(variant="v2t", num_classes=2, seed=0) -> dict:
    """Seeded, non-degenerate weights: fan-in scaled kernels, non-trivial BatchNorm statistics, damped projection BNs (the
    last BN of every residual branch) so that activations do not blow up with depth (cf. oracle/resnet_rs.py)."""
    rng = np.random.default_rng(seed)
    nblocks = len(plan(variant)["blocks"])
    W = {}
    for name, shp in weight_shapes(variant, num_classes).items():
        leaf = name.rsplit("/", 1)[1]
        if name == "predictions/kernel":
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            W[name] = rng.uniform(-lim, lim, shp).astype(np.float32)
        elif leaf == "depthwise_kernel":
            W[name] = (rng.standard_normal(shp) * np.sqrt(2.0 / (shp[0] * shp[1]))).astype(np.float32)
        elif leaf == "kernel":
            W[name] = (rng.standard_normal(shp) * np.sqrt(2.0 / np.prod(shp[:-1]))).astype(np.float32)
        elif leaf == "gamma":
            W[name] = rng.uniform(0.6, 1.4, shp).astype(np.float32)
        elif leaf == "beta":
            W[name] = (rng.standard_normal(shp) * 0.2).astype(np.float32)
        elif leaf == "moving_mean":
            W[name] = (rng.standard_normal(shp) * 0.1).astype(np.float32)
        elif leaf == "moving_variance":
            W[name] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        elif leaf == "bias":
            W[name] = (rng.standard_normal(shp) * 0.1).astype(np.float32)
        else:
            raise KeyError(name)
    damp = 1.5 / np.sqrt(nblocks)
    for name in W:
        if name.endswith("MB_pw_bn/gamma") or name.endswith("fu_bn/gamma"):
            W[name] *= damp
    return W


def param_count(W, include_head=True):
    return int(sum(v.size for k, v in W.items() if include_head or not k.startswith("predictions/")))
