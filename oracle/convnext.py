"""CPU ORACLE (test infrastructure, not shipped) -- ConvNeXt forward in PyTorch fp32, restating
``models/tfimm/architectures/convnext.py`` of the reference on Keras-layout, Keras-named weights.

Reference -> here:
  ConvNeXtConfig / registry   convnext.py:39-138, 440-470, 611-620 (convnext_tiny_in22k) -> CONFIGS
  ConvNeXt.forward_features   convnext.py:373-405 (stem conv kernel ``patch_size`` = 4, stride ``first_down * 2`` = 2 --
                              the author's modification, :320-327 --, 'valid' padding, LayerNorm eps 1e-6) -> :func:`forward`
  ConvNeXtStage               convnext.py:232-296 (LayerNorm + Conv 2x2 stride 2 'valid' downsample)      -> :func:`stage`
  ConvNeXtBlock               convnext.py:147-229 (ZeroPadding 3 -> DepthwiseConv 7x7 + bias -> LayerNorm -> MLP ->
                              * gamma -> + shortcut; DropPath is the identity at inference)                -> :func:`block`
  MLP                         models/tfimm/layers/transformers.py:176-214 (Dense, exact-erf GELU, Dense)   -> :func:`block`
  head                        convnext.py:427-438 (GlobalAveragePooling -> LayerNorm -> Dense)            -> :func:`forward`

Weights: dict name -> numpy array, Keras layouts: Conv2D ``kernel`` (kh,kw,Cin,Cout), DepthwiseConv2D
``depthwise_kernel`` (kh,kw,C,1), Dense ``kernel`` (in,out), LayerNormalization ``gamma,beta``, block ``gamma`` (C).
Parity status: unpinned against real Keras (TensorFlow is not installable offline); known answers = the parameter count
of the timm/tfimm model card (28.6 M with the 1000-class head, 27.8 M backbone) and the stage shapes."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

CONFIGS = {  # convnext.py:440-470 (tiny / small / base), shared by the *_in22k / *_in22ft1k registrations
    "tiny": dict(embed_dim=(96, 192, 384, 768), nb_blocks=(3, 3, 9, 3)),
    "small": dict(embed_dim=(96, 192, 384, 768), nb_blocks=(3, 3, 27, 3)),
    "base": dict(embed_dim=(128, 256, 512, 1024), nb_blocks=(3, 3, 27, 3)),
}
PATCH, MLP_RATIO, LN_EPS = 4, 4.0, 1e-6


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).float()


def ln(x, W, name):
    return F.layer_norm(x, (x.shape[-1],), _t(W[name + "/gamma"]), _t(W[name + "/beta"]), LN_EPS)


def conv_valid(x, W, name, stride):  # NHWC, 'valid' padding, bias
    w = _t(W[name + "/kernel"]).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x.permute(0, 3, 1, 2), w, _t(W[name + "/bias"]), stride=stride).permute(0, 2, 3, 1)


def block(x, W, name):
    c = x.shape[-1]
    dw = _t(W[name + "/conv_dw/depthwise_kernel"]).permute(2, 3, 0, 1).contiguous()        # (7,7,C,1) -> (C,1,7,7)
    y = F.conv2d(F.pad(x.permute(0, 3, 1, 2), (3, 3, 3, 3)), dw, _t(W[name + "/conv_dw/bias"]), groups=c).permute(0, 2, 3, 1)
    y = ln(y, W, name + "/norm")
    y = F.gelu(y @ _t(W[name + "/mlp/fc1/kernel"]) + _t(W[name + "/mlp/fc1/bias"]))
    y = y @ _t(W[name + "/mlp/fc2/kernel"]) + _t(W[name + "/mlp/fc2/bias"])
    return x + y * _t(W[name + "/gamma"])


def stage(x, W, j, nb_blocks):
    if j > 0:
        x = ln(x, W, f"stages/{j}/downsample/0")
        x = conv_valid(x, W, f"stages/{j}/downsample/1", 2)
    for i in range(nb_blocks):
        x = block(x, W, f"stages/{j}/blocks/{i}")
    return x


def forward(x_nhwc, W, variant="tiny", head_act="softmax", return_logits=False, first_down=1, taps=None):
    cfg = CONFIGS[variant]
    with torch.no_grad():
        x = conv_valid(_t(x_nhwc), W, "stem/0", first_down * 2)
        x = ln(x, W, "stem/1")
        if taps is not None:
            taps["stem"] = x.numpy().copy()
        for j, nb in enumerate(cfg["nb_blocks"]):
            x = stage(x, W, j, nb)
            if taps is not None:
                taps[f"stage{j}"] = x.numpy().copy()
        feat = ln(x.mean(dim=(1, 2)), W, "head/norm")
        if taps is not None:
            taps["feat"] = feat.numpy().copy()
        logits = feat @ _t(W["head/fc/kernel"]) + _t(W["head/fc/bias"])
        if return_logits:
            return logits.numpy()
        return (torch.softmax(logits, -1) if head_act == "softmax" else torch.sigmoid(logits)).numpy()


def weight_shapes(variant="tiny", num_classes=2) -> dict:
    cfg, s = CONFIGS[variant], {}

    def lnorm(n, c):
        s[n + "/gamma"], s[n + "/beta"] = (c,), (c,)

    d = cfg["embed_dim"]
    s["stem/0/kernel"], s["stem/0/bias"] = (PATCH, PATCH, 3, d[0]), (d[0],)
    lnorm("stem/1", d[0])
    for j, nb in enumerate(cfg["nb_blocks"]):
        c = d[j]
        if j > 0:
            lnorm(f"stages/{j}/downsample/0", d[j - 1])
            s[f"stages/{j}/downsample/1/kernel"], s[f"stages/{j}/downsample/1/bias"] = (2, 2, d[j - 1], c), (c,)
        for i in range(nb):
            n = f"stages/{j}/blocks/{i}"
            s[n + "/conv_dw/depthwise_kernel"], s[n + "/conv_dw/bias"] = (7, 7, c, 1), (c,)
            lnorm(n + "/norm", c)
            h = int(MLP_RATIO * c)
            s[n + "/mlp/fc1/kernel"], s[n + "/mlp/fc1/bias"] = (c, h), (h,)
            s[n + "/mlp/fc2/kernel"], s[n + "/mlp/fc2/bias"] = (h, c), (c,)
            s[n + "/gamma"] = (c,)
    lnorm("head/norm", d[-1])
    s["head/fc/kernel"], s["head/fc/bias"] = (d[-1], num_classes), (num_classes,)
    return s


def random_weights(variant="tiny", num_classes=2, seed=0) -> dict:
    """Seeded, non-degenerate weights (cf. oracle/gcvit.py): fan-in scaled kernels, LayerNorm gamma ~ U(0.6, 1.4), small
    biases, and layer-scale gammas around 0.5 / sqrt(#blocks) instead of the 1e-6 initialiser so that the blocks
    contribute without blowing up the residual stream."""
    rng = np.random.default_rng(seed)
    nblocks = sum(CONFIGS[variant]["nb_blocks"])
    W = {}
    for name, shp in weight_shapes(variant, num_classes).items():
        leaf = name.rsplit("/", 1)[1]
        if name == "head/fc/kernel":
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            W[name] = rng.uniform(-lim, lim, shp).astype(np.float32)
        elif leaf == "depthwise_kernel":
            W[name] = (rng.standard_normal(shp) * np.sqrt(1.5 / 49.0)).astype(np.float32)
        elif leaf == "kernel":
            W[name] = (rng.standard_normal(shp) * np.sqrt(1.5 / np.prod(shp[:-1]))).astype(np.float32)
        elif leaf == "gamma" and "/blocks/" in name and name.split("/")[-2].isdigit():   # stages/j/blocks/i/gamma: layer scale
            W[name] = (rng.uniform(0.3, 0.7, shp) * 1.5 / np.sqrt(nblocks)).astype(np.float32)
        elif leaf == "gamma":
            W[name] = rng.uniform(0.6, 1.4, shp).astype(np.float32)
        elif leaf in ("beta", "bias"):
            W[name] = (rng.standard_normal(shp) * 0.1).astype(np.float32)
        else:
            raise KeyError(name)
    return W


def param_count(W, include_head=True):
    return int(sum(v.size for k, v in W.items() if include_head or not k.startswith("head/fc")))
