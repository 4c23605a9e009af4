"""CPU ORACLE (test infrastructure, not shipped) -- ResNet-RS forward in PyTorch fp32, restating
``models/resnet_rs/resnet_rs_model.py`` of the reference line by line on Keras-layout weights.

Reference -> here:
  Conv2DFixedPadding  resnet_rs_model.py:64-84  + fixed_padding model_utils.py:22-46   -> :func:`conv_fixed`
  STEM                resnet_rs_model.py:87-142                                        -> :func:`stem`
  SE                  resnet_rs_model.py:145-183                                       -> :func:`se`
  BottleneckBlock     resnet_rs_model.py:186-282                                       -> :func:`bottleneck`
  BlockGroup          resnet_rs_model.py:285-326, BLOCK_ARGS block_args.py:1-44        -> :func:`forward`
  head                resnet_rs_model.py:468-476                                       -> :func:`forward`

Weights: dict name -> numpy array in Keras layout: conv ``kernel`` (kh,kw,Cin,Cout), BatchNormalization
``gamma,beta,moving_mean,moving_variance``, Dense ``kernel`` (in,out) + ``bias``.  Parity status: unpinned against real
Keras (TensorFlow is not installable offline; the reference ships no checkpoints or golden outputs); the known-answer
tests are the parameter counts of SURVEY.md 8c and the stage shapes of Appendix B.2.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BLOCK_ARGS = {  # block_args.py:1-44 (input_filters, num_repeats)
    50: [(64, 3), (128, 4), (256, 6), (512, 3)],
    101: [(64, 3), (128, 4), (256, 23), (512, 3)],
    152: [(64, 3), (128, 8), (256, 36), (512, 3)],
    200: [(64, 3), (128, 24), (256, 36), (512, 3)],
}
BN_EPS = 1e-5


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).float()


def _conv_w(k):  # (kh,kw,I,O) -> (O,I,kh,kw)
    return _t(k).permute(3, 2, 0, 1).contiguous()


def conv_fixed(x, kernel, stride):
    """Conv2DFixedPadding: stride 1 -> 'same'; stride > 1 -> explicit symmetric zero pad (k-1)//2.. then 'valid'."""
    w = _conv_w(kernel)
    k = w.shape[-1]
    if stride > 1:
        pb = (k - 1) // 2
        pe = (k - 1) - pb
        x = F.pad(x, (pb, pe, pb, pe))
        return F.conv2d(x, w, stride=stride)
    return F.conv2d(x, w, stride=1, padding=(k - 1) // 2)


def bn(x, W, prefix):
    g, b = _t(W[prefix + "/gamma"]), _t(W[prefix + "/beta"])
    m, v = _t(W[prefix + "/moving_mean"]), _t(W[prefix + "/moving_variance"])
    return (x - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + BN_EPS) * g[None, :, None, None] \
        + b[None, :, None, None]


def stem(x, W, first_strides=2):
    for i, s in ((1, first_strides), (2, 1), (3, 1), (4, 2)):
        x = conv_fixed(x, W[f"stem_conv_{i}/kernel"], s)
        x = torch.relu(bn(x, W, f"stem_batch_norm_{i}"))
    return x


def se(x, W, name):
    p = x.mean(dim=(2, 3), keepdim=True)
    p = torch.relu(F.conv2d(p, _conv_w(W[name + "se_reduce/kernel"]), _t(W[name + "se_reduce/bias"])))
    p = torch.sigmoid(F.conv2d(p, _conv_w(W[name + "se_expand/kernel"]), _t(W[name + "se_expand/bias"])))
    return x * p


def avg_pool_same_2x2(x):
    """AveragePooling2D(2, 2, 'same'): pad bottom/right, divide by the number of VALID elements."""
    h, w = x.shape[-2:]
    xp = F.pad(x, (0, w % 2, 0, h % 2))
    s = F.avg_pool2d(xp, 2, 2) * 4.0
    ones = F.pad(torch.ones(1, 1, h, w), (0, w % 2, 0, h % 2))
    cnt = F.avg_pool2d(ones, 2, 2) * 4.0
    return s / cnt


def bottleneck(x, W, name, filters, strides, use_projection):
    shortcut = x
    if use_projection:
        if strides == 2:
            shortcut = avg_pool_same_2x2(x)
            shortcut = conv_fixed(shortcut, W[name + "projection_conv/kernel"], 1)
        else:
            shortcut = conv_fixed(x, W[name + "projection_conv/kernel"], strides)
        shortcut = bn(shortcut, W, name + "projection_batch_norm")
    y = torch.relu(bn(conv_fixed(x, W[name + "conv_1/kernel"], 1), W, name + "batch_norm_1"))
    y = torch.relu(bn(conv_fixed(y, W[name + "conv_2/kernel"], strides), W, name + "batch_norm_2"))
    y = bn(conv_fixed(y, W[name + "conv_3/kernel"], 1), W, name + "batch_norm_3")
    y = se(y, W, name)
    return torch.relu(y + shortcut)


def forward(x_nhwc: np.ndarray, W: dict, depth: int = 50, head_act: str = "softmax", return_logits: bool = False,
            first_strides: int = 2, taps: dict | None = None):
    """x_nhwc float32 [N,H,W,3] in [0,1] -> probabilities [N,k] (or logits)."""
    with torch.no_grad():
        x = _t(x_nhwc).permute(0, 3, 1, 2).contiguous()
        x = stem(x, W, first_strides)
        if taps is not None:
            taps["stem"] = x.permute(0, 2, 3, 1).numpy().copy()
        for gi, (filters, reps) in enumerate(BLOCK_ARGS[depth]):
            for bi in range(reps):
                x = bottleneck(x, W, f"c{gi + 2}_block_{bi}_", filters, (1 if gi == 0 else 2) if bi == 0 else 1, bi == 0)
            if taps is not None:
                taps[f"c{gi + 2}"] = x.permute(0, 2, 3, 1).numpy().copy()
        feat = x.mean(dim=(2, 3))
        logits = feat @ _t(W["predictions/kernel"]) + _t(W["predictions/bias"])
        if taps is not None:
            taps["feat"] = feat.numpy().copy()
        if return_logits:
            return logits.numpy()
        if head_act == "softmax":
            return torch.softmax(logits, dim=-1).numpy()
        if head_act == "sigmoid":
            return torch.sigmoid(logits).numpy()
        raise ValueError(head_act)


# ---- shared random-init weights (Keras names / layouts), identical file for oracle and CUDA path --------------
def weight_shapes(depth: int = 50, num_classes: int = 2) -> dict:
    s = {}

    def conv(name, kh, ci, co):
        s[name + "/kernel"] = (kh, kh, ci, co)

    def bnorm(name, c):
        for p in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"{name}/{p}"] = (c,)

    for i, (ci, co) in enumerate(((3, 32), (32, 32), (32, 64), (64, 64)), 1):
        conv(f"stem_conv_{i}", 3, ci, co)
        bnorm(f"stem_batch_norm_{i}", co)
    cin = 64
    for gi, (f, reps) in enumerate(BLOCK_ARGS[depth]):
        for bi in range(reps):
            n = f"c{gi + 2}_block_{bi}_"
            if bi == 0:
                conv(n + "projection_conv", 1, cin, 4 * f)
                bnorm(n + "projection_batch_norm", 4 * f)
            conv(n + "conv_1", 1, cin, f)
            bnorm(n + "batch_norm_1", f)
            conv(n + "conv_2", 3, f, f)
            bnorm(n + "batch_norm_2", f)
            conv(n + "conv_3", 1, f, 4 * f)
            bnorm(n + "batch_norm_3", 4 * f)
            s[n + "se_reduce/kernel"] = (1, 1, 4 * f, f)
            s[n + "se_reduce/bias"] = (f,)
            s[n + "se_expand/kernel"] = (1, 1, f, 4 * f)
            s[n + "se_expand/bias"] = (4 * f,)
            cin = 4 * f
    s["predictions/kernel"] = (cin, num_classes)
    s["predictions/bias"] = (num_classes,)
    return s


def random_weights(depth: int = 50, num_classes: int = 2, seed: int = 0) -> dict:
    """Seeded weights with the reference's initialiser families (VarianceScaling fan_in convs, resnet_rs_model.py:80)
    but NON-trivial BatchNorm statistics and a wider head, so that outputs are not degenerate."""
    rng = np.random.default_rng(seed)
    W = {}
    for name, shp in weight_shapes(depth, num_classes).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "kernel" and len(shp) == 4:
            fan_in = shp[0] * shp[1] * shp[2]
            W[name] = (rng.standard_normal(shp) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        elif leaf == "kernel":  # Dense head: Keras default glorot-uniform
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            W[name] = rng.uniform(-lim, lim, shp).astype(np.float32)
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
    # identity-ish residual branches: damp the last BN of every block so activations do not blow up with depth
    damp = 1.5 / np.sqrt(sum(r for _, r in BLOCK_ARGS[depth]))
    for name in W:
        if name.endswith("batch_norm_3/gamma"):
            W[name] *= damp
    return W


def param_count(W: dict, include_head: bool = True) -> int:
    return int(sum(v.size for k, v in W.items() if include_head or not k.startswith("predictions/")))


def calibrate_head(W: dict, feat: np.ndarray, head_kernel: str = "predictions/kernel", head_bias: str = "predictions/bias",
                   seed: int = 0, target_std: float = 1.5, direction: str = "random") -> dict:
    """Random-init backbones give almost image-independent logits ('random-init degeneracy').
    Re-draw the head so that logits are centred and spread (std ``target_std``) over the calibration features
    ``feat`` [n, C]: probabilities then fall on both sides of the 0.487 threshold and label agreement means something.
    ``direction``: "random" = a Gaussian read-out direction; "pca" = the direction along which the calibration features
    vary most from image to image (what a trained head would pick up), i.e. the largest image-dependent signal relative
    to the backbone's numerical noise.  Returns a copy of ``W`` with the new head."""
    rng = np.random.default_rng(seed)
    k = W[head_kernel].shape[1]
    mu = feat.mean(axis=0)
    d = rng.standard_normal((feat.shape[1], k)).astype(np.float64)
    if direction == "pca":
        _, _, vt = np.linalg.svd((feat - mu).astype(np.float64), full_matrices=False)
        d[:, -1] = vt[0] * np.sqrt(feat.shape[1])          # class 1 (or the single sigmoid unit) reads the first component
        if k > 1:
            d[:, :-1] *= 0.0
    z = (feat - mu) @ d
    spread = (z[:, 0] - z[:, 1]).std() if k > 1 else z[:, 0].std()
    kern = d * (target_std / max(spread, 1e-12))
    out = dict(W)
    out[head_kernel] = kern.astype(np.float32)
    out[head_bias] = (-(mu @ kern)).astype(np.float32)
    return out


def center_head(W: dict, feat: np.ndarray, thr: float = 0.487, head_kernel: str = "predictions/kernel",
                head_bias: str = "predictions/bias") -> dict:
    """Shift the head bias (no amplification) so that the median P(synthetic) over the calibration features sits on the
    decision threshold: both labels occur and 'identical labels' is not vacuous."""
    out = dict(W)
    z = feat @ W[head_kernel] + W[head_bias]
    b = np.array(W[head_bias], dtype=np.float32)
    target = float(np.log(thr / (1 - thr)))
    if z.shape[1] > 1:                       # P(syn) = 1 - softmax(z)[0]; for k = 2 this is sigmoid(z1 - z0)
        b[1] += target - float(np.median(z[:, 1] - z[:, 0]))
    else:
        b[0] += target - float(np.median(z[:, 0]))
    out[head_bias] = b
    return out
