"""CPU ORACLE (test infrastructure, not shipped) -- ECA-NFNet-L0 forward in PyTorch fp32, restating
``models/keras_cv_attention_models/nfnets/nfnets.py`` of the reference on Keras-layout, Keras-named weights.

Reference -> here:
  ScaledStandardizedConv2D   nfnets.py:42-81 (weight standardisation over HWI, * gain * gamma, biased Conv2D)  -> :func:`std_conv`
  block / stack / stem       nfnets.py:116-191                                                                    -> :func:`block`, :func:`forward`
  NormFreeNet, NormFreeNet_Light, ECA_NFNetL0   nfnets.py:194-269, 304-320 (channel_ratio .25, group_size 64, torch padding,
                             no zero-init gain, gamma_in_act=False => conv gamma 1.7881 (swish), activation gamma 1)
  eca_module                 common_layers.py:335-353 (GAP -> zero pad -> Conv1D k over channels -> sigmoid)      -> :func:`eca`

Weights: ScaledStandardizedConv2D ``kernel`` (kh,kw,Cin/groups,Cout), ``bias`` (Cout), ``gain`` (Cout); Conv1D ``kernel``
(k,1,1); Dense ``kernel`` (in,out) + ``bias``.  Parity status: unpinned against real Keras (TensorFlow is not installable
offline); known answers = the parameter count of the kecam model table (ECA_NFNetL0 24.14 M) and the stage shapes."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

SWISH_GAMMA = 1.7881293296813965     # NON_LINEAR_GAMMA["swish"], nfnets.py:35
NUM_BLOCKS, OUT_CHANNELS, STRIDES = [1, 2, 6, 3], [256, 512, 1536, 1536], [1, 2, 2, 2]
CHANNEL_RATIO, GROUP_SIZE, ALPHA, STEM_WIDTH, FEATURES, ATTN_GAIN, STD_EPS = 0.25, 64, 0.2, 128, 2304, 2.0, 1e-5


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).float()


def standardized_kernel(W, name, gamma=SWISH_GAMMA):
    """(kernel - mean) * rsqrt(max(var * fan_in, eps)) * gain * gamma, moments over (kh, kw, Cin) per output channel."""
    k = _t(W[name + "conv/kernel"])
    mean, var = k.mean(dim=(0, 1, 2), keepdim=True), k.var(dim=(0, 1, 2), keepdim=True, unbiased=False)
    fan_in = float(k.shape[0] * k.shape[1] * k.shape[2])
    scale = torch.rsqrt(torch.clamp(var * fan_in, min=STD_EPS)) * (_t(W[name + "conv/gain"]) * gamma)
    return (k - mean) * scale


def std_conv(x, W, name, k, stride=1, groups=1):
    """std_conv2d_with_init: torch padding (ZeroPadding k // 2 + 'VALID') for k > 1; NCHW."""
    w = standardized_kernel(W, name).permute(3, 2, 0, 1).contiguous()
    if k > 1:
        x = F.pad(x, (k // 2,) * 4)
    return F.conv2d(x, w, _t(W[name + "conv/bias"]), stride=stride, groups=groups)


def eca(x, W, name):
    c = x.shape[1]
    kernel = _t(W[name + "conv1d/kernel"]).reshape(1, 1, -1)
    pad = kernel.shape[-1] // 2
    m = x.mean(dim=(2, 3))
    g = torch.sigmoid(F.conv1d(F.pad(m, (pad, pad))[:, None, :], kernel)[:, 0, :])
    assert g.shape[1] == c
    return x * g[:, :, None, None]


def avgpool_same(x, s):
    """AvgPool2D(s, strides=s, padding='SAME'): windows at the border average the valid inputs only."""
    h, w = x.shape[2], x.shape[3]
    ph, pw = (-h) % s, (-w) % s
    ones = torch.ones((1, 1, h, w))
    num = F.avg_pool2d(F.pad(x, (0, pw, 0, ph)), s, s) * (s * s)
    den = F.avg_pool2d(F.pad(ones, (0, pw, 0, ph)), s, s) * (s * s)
    return num / den


def block(x, W, name, filters, beta, stride):
    hidden = int(filters * CHANNEL_RATIO)
    groups = hidden // GROUP_SIZE
    preact = F.silu(x) * beta
    if stride > 1 or x.shape[1] != filters:
        sc = avgpool_same(preact, stride) if stride > 1 else preact
        sc = std_conv(sc, W, name + "shortcut_", 1)
    else:
        sc = x
    d = F.silu(std_conv(preact, W, name + "deep_1_", 1))
    d = F.silu(std_conv(d, W, name + "deep_2_", 3, stride, groups))
    d = F.silu(std_conv(d, W, name + "deep_3_", 3, 1, groups))
    d = std_conv(d, W, name + "deep_4_", 1)
    d = eca(d, W, name + "eca_") * ATTN_GAIN
    return sc + d * ALPHA


def betas():
    """Per-block beta (nfnets.py:246-255, 170-178): beta_list[i] = (1 + alpha^2 i)^-1/2, the first block of a stack takes the
    last beta of the previous stack."""
    beta_list = [(1 + ALPHA ** 2 * i) ** -0.5 for i in range(max(NUM_BLOCKS) + 1)]
    out, pre = [], 1.0
    for nb in NUM_BLOCKS:
        b = beta_list[: nb + 1]
        b[0] = pre
        out.append(b[:nb])
        pre = b[-1]
    return out


def forward(x_nhwc, W, head_act="softmax", return_logits=False, first_strides=2, taps=None):
    with torch.no_grad():
        x = _t(x_nhwc).permute(0, 3, 1, 2)
        x = F.silu(std_conv(x, W, "stem_1_", 3, first_strides))
        x = F.silu(std_conv(x, W, "stem_2_", 3, 1))
        x = F.silu(std_conv(x, W, "stem_3_", 3, 1))
        x = std_conv(x, W, "stem_4_", 3, 2)
        if taps is not None:
            taps["stem"] = x.permute(0, 2, 3, 1).numpy().copy()
        for sid, (nb, oc, st, bs) in enumerate(zip(NUM_BLOCKS, OUT_CHANNELS, STRIDES, betas())):
            for bid in range(nb):
                x = block(x, W, f"stack{sid + 1}_block{bid + 1}_", oc, bs[bid], st if bid == 0 else 1)
            if taps is not None:
                taps[f"stack{sid + 1}"] = x.permute(0, 2, 3, 1).numpy().copy()
        x = F.silu(std_conv(x, W, "post_", 1))
        feat = x.mean(dim=(2, 3))
        if taps is not None:
            taps["feat"] = feat.numpy().copy()
        logits = feat @ _t(W["predictions/kernel"]) + _t(W["predictions/bias"])
        if return_logits:
            return logits.numpy()
        return (torch.softmax(logits, -1) if head_act == "softmax" else torch.sigmoid(logits)).numpy()


def weight_shapes(num_classes=2) -> dict:
    s = {}

    def sconv(n, k, cin, cout, groups=1):
        s[n + "conv/kernel"], s[n + "conv/bias"], s[n + "conv/gain"] = (k, k, cin // groups, cout), (cout,), (cout,)

    for i, (ci, co) in enumerate(((3, 16), (16, 32), (32, 64), (64, 128)), 1):
        sconv(f"stem_{i}_", 3, ci, co)
    cin = STEM_WIDTH
    for sid, (nb, oc, st) in enumerate(zip(NUM_BLOCKS, OUT_CHANNELS, STRIDES)):
        hidden = int(oc * CHANNEL_RATIO)
        g = hidden // GROUP_SIZE
        for bid in range(nb):
            n = f"stack{sid + 1}_block{bid + 1}_"
            stride = st if bid == 0 else 1
            if stride > 1 or cin != oc:
                sconv(n + "shortcut_", 1, cin, oc)
            sconv(n + "deep_1_", 1, cin, hidden)
            sconv(n + "deep_2_", 3, hidden, hidden, g)
            sconv(n + "deep_3_", 3, hidden, hidden, g)
            sconv(n + "deep_4_", 1, hidden, oc)
            s[n + "eca_conv1d/kernel"] = (5, 1, 1)
            cin = oc
    sconv("post_", 1, cin, FEATURES)
    s["predictions/kernel"], s["predictions/bias"] = (FEATURES, num_classes), (num_classes,)
    return s


def random_weights(num_classes=2, seed=0) -> dict:
    """Seeded weights: normal kernels (the standardisation fixes their scale), gains ~ U(0.6, 1.4) with the last conv of each
    residual branch damped (1.5 / sqrt(#blocks)), small biases, ECA taps ~ N(0, 1)."""
    rng = np.random.default_rng(seed)
    W = {}
    for name, shp in weight_shapes(num_classes).items():
        leaf = name.rsplit("/", 1)[1]
        if name == "predictions/kernel":
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            W[name] = rng.uniform(-lim, lim, shp).astype(np.float32)
        elif leaf == "kernel":
            W[name] = rng.standard_normal(shp).astype(np.float32)
        elif leaf == "gain":
            W[name] = (rng.uniform(0.6, 1.4, shp) * (1.5 / np.sqrt(sum(NUM_BLOCKS)) if "deep_4_" in name else 1.0)).astype(np.float32)
        elif leaf == "bias":
            W[name] = (rng.standard_normal(shp) * 0.1).astype(np.float32)
        else:
            raise KeyError(name)
    return W


def param_count(W, include_head=True):
    return int(sum(v.size for k, v in W.items() if include_head or not k.startswith("predictions/")))
