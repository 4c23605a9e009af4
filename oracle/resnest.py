"""CPU ORACLE (test infrastructure, not shipped) -- ResNeSt-50 forward in PyTorch fp32, restating
``models/keras_cv_attention_models/resnest/resnest.py`` + ``aotnet/aotnet.py`` of the reference on Keras-named weights.

Reference -> here:
  ResNest50 / ResNest          resnest.py:69-77 (AotNet with stem_type 'deep', attn_types 'sa', bn_after_attn False,
                               shortcut_type 'avg', num_blocks [3,4,6,3], stem_width 64)
  AotNet, aot_stem, deep_stem  aotnet.py:284-377, 235-242, 264-281 (3 x conv3x3, BN + ReLU, ZeroPadding + MaxPool 3/2)
  aot_stack / aot_block        aotnet.py:195-232, 137-192 (conv shortcut: AvgPool 'SAME' + 1x1 conv + BN; deep branch; '3_' BN;
                               add; ReLU)
  deep_branch / attn_block     aotnet.py:117-134, 30-97 (1x1 conv + BN + ReLU -> split attention -> 1x1 conv)
  split_attention_conv2d       resnest.py:27-66 (radix-2: one 3x3 conv per input half, BN + ReLU, pooled sum -> 1x1 + bias ->
                               BN + ReLU -> 1x1 + bias -> r-softmax over the radix (16-24) -> weighted sum; stride-2 blocks end
                               with ZeroPadding(1) + AveragePooling 3x3 stride 2)
  conv2d_no_bias               common_layers.py:230-248 (torch padding: ZeroPadding(k // 2) + 'VALID')

Parity status: unpinned against real Keras (TensorFlow is not installable offline); known answers = the parameter count of
the kecam model table (ResNest50 27.6 M) and the stage shapes."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

NUM_BLOCKS, OUT_CHANNELS, STRIDES, STEM_WIDTH, RADIX, BN_EPS = [3, 4, 6, 3], [256, 512, 1024, 2048], [1, 2, 2, 2], 64, 2, 1e-5


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).float()


def conv(x, W, name, k, stride=1, bias=False):
    """conv2d_no_bias with torch padding (or a biased Conv2D for the attention 1x1s); NCHW."""
    w = _t(W[name + "conv/kernel"]).permute(3, 2, 0, 1).contiguous()
    if k > 1:
        x = F.pad(x, (k // 2,) * 4)
    return F.conv2d(x, w, _t(W[name + "conv/bias"]) if bias else None, stride=stride)


def bn(x, W, name, relu):
    g, b = _t(W[name + "bn/gamma"]), _t(W[name + "bn/beta"])
    m, v = _t(W[name + "bn/moving_mean"]), _t(W[name + "bn/moving_variance"])
    x = (x - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + BN_EPS) * g[None, :, None, None] + b[None, :, None, None]
    return F.relu(x) if relu else x


def avgpool_same(x, s):
    h, w = x.shape[2], x.shape[3]
    ph, pw = (-h) % s, (-w) % s
    ones = torch.ones((1, 1, h, w))
    return F.avg_pool2d(F.pad(x, (0, pw, 0, ph)), s, s) / F.avg_pool2d(F.pad(ones, (0, pw, 0, ph)), s, s)


def split_attention(x, W, name, filters, stride):
    cin = x.shape[1]
    halves = torch.split(x, cin // RADIX, dim=1)
    logits = torch.cat([conv(halves[i], W, f"{name}1_g{i + 1}_", 3) for i in range(RADIX)], dim=1)
    logits = bn(logits, W, name + "1_", True)
    gap = sum(torch.split(logits, filters, dim=1)).mean(dim=(2, 3), keepdim=True)
    att = bn(conv(gap, W, name + "2_", 1, bias=True), W, name + "2_", True)
    att = conv(att, W, name + "3_", 1, bias=True)                                   # [N, radix * filters, 1, 1]
    att = torch.softmax(att.reshape(-1, RADIX, filters), dim=1).reshape(-1, RADIX * filters, 1, 1)
    out = sum(torch.split(att * logits, filters, dim=1))
    if stride > 1:
        out = F.avg_pool2d(F.pad(out, (1, 1, 1, 1)), 3, 2)                          # zeros count: divisor 9
    return out


def block(x, W, name, filters, stride, conv_shortcut):
    hidden = filters // 4
    if conv_shortcut:
        sc = avgpool_same(x, stride) if stride > 1 else x
        sc = bn(conv(sc, W, name + "shortcut_", 1), W, name + "shortcut_", False)
    else:
        sc = x
    d = bn(conv(x, W, name + "deep_1_", 1), W, name + "deep_1_", True)
    d = split_attention(d, W, name + "deep_2_sa_", hidden, stride)
    d = bn(conv(d, W, name + "deep_3_", 1), W, name + "3_", False)
    return F.relu(sc + d)


def forward(x_nhwc, W, head_act="softmax", return_logits=False, first_strides=2, taps=None):
    with torch.no_grad():
        x = _t(x_nhwc).permute(0, 3, 1, 2)
        x = bn(conv(x, W, "stem_1_", 3, first_strides), W, "stem_1_", True)
        x = bn(conv(x, W, "stem_2_", 3), W, "stem_2_", True)
        x = bn(conv(x, W, "stem_3_", 3), W, "stem_", True)
        x = F.max_pool2d(F.pad(x, (1, 1, 1, 1)), 3, 2)
        if taps is not None:
            taps["stem"] = x.permute(0, 2, 3, 1).numpy().copy()
        cin = STEM_WIDTH
        for sid, (nb, oc, st) in enumerate(zip(NUM_BLOCKS, OUT_CHANNELS, STRIDES)):
            for bid in range(nb):
                x = block(x, W, f"stack{sid + 1}_block{bid + 1}_", oc, st if bid == 0 else 1, bid == 0 and (st != 1 or cin != oc))
                cin = oc
            if taps is not None:
                taps[f"stack{sid + 1}"] = x.permute(0, 2, 3, 1).numpy().copy()
        feat = x.mean(dim=(2, 3))
        if taps is not None:
            taps["feat"] = feat.numpy().copy()
        logits = feat @ _t(W["predictions/kernel"]) + _t(W["predictions/bias"])
        if return_logits:
            return logits.numpy()
        return (torch.softmax(logits, -1) if head_act == "softmax" else torch.sigmoid(logits)).numpy()


def weight_shapes(num_classes=2) -> dict:
    s = {}

    def bnorm(n, c):
        for q in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"{n}bn/{q}"] = (c,)

    s["stem_1_conv/kernel"], s["stem_2_conv/kernel"], s["stem_3_conv/kernel"] = (3, 3, 3, 32), (3, 3, 32, 32), (3, 3, 32, 64)
    bnorm("stem_1_", 32), bnorm("stem_2_", 32), bnorm("stem_", 64)
    cin = STEM_WIDTH
    for sid, (nb, oc, st) in enumerate(zip(NUM_BLOCKS, OUT_CHANNELS, STRIDES)):
        hidden = oc // 4
        for bid in range(nb):
            n = f"stack{sid + 1}_block{bid + 1}_"
            if bid == 0 and (st != 1 or cin != oc):
                s[n + "shortcut_conv/kernel"] = (1, 1, cin, oc)
                bnorm(n + "shortcut_", oc)
            s[n + "deep_1_conv/kernel"] = (1, 1, cin, hidden)
            bnorm(n + "deep_1_", hidden)
            sa = n + "deep_2_sa_"
            for g in range(RADIX):
                s[f"{sa}1_g{g + 1}_conv/kernel"] = (3, 3, hidden // RADIX, hidden)
            bnorm(sa + "1_", hidden * RADIX)
            inter = max(hidden * RADIX // 4, 32)
            s[sa + "2_conv/kernel"], s[sa + "2_conv/bias"] = (1, 1, hidden, inter), (inter,)
            bnorm(sa + "2_", inter)
            s[sa + "3_conv/kernel"], s[sa + "3_conv/bias"] = (1, 1, inter, hidden * RADIX), (hidden * RADIX,)
            s[n + "deep_3_conv/kernel"] = (1, 1, hidden, oc)
            bnorm(n + "3_", oc)
            cin = oc
    s["predictions/kernel"], s["predictions/bias"] = (cin, num_classes), (num_classes,)
    return s


def random_weights(num_classes=2, seed=0) -> dict:
    rng = np.random.default_rng(seed)
    W = {}
    for name, shp in weight_shapes(num_classes).items():
        leaf = name.rsplit("/", 1)[1]
        if name == "predictions/kernel":
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            W[name] = rng.uniform(-lim, lim, shp).astype(np.float32)
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
    # well-conditioned residual branches: damp the last BN of every block ('3_' BN, zero-gamma initialised in the reference)
    damp = 0.5 / np.sqrt(sum(NUM_BLOCKS))
    for name in W:
        if name.endswith("_3_bn/gamma") and "sa_" not in name:
            W[name] *= damp
        if name.endswith("_3_bn/beta") and "sa_" not in name:
            W[name] *= 0.25
    return W


def param_count(W, include_head=True, trainable_only=False):
    return int(sum(v.size for k, v in W.items() if (include_head or not k.startswith("predictions/"))
                   and not (trainable_only and "moving_" in k)))
