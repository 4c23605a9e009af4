"""ConvNeXt forward on the B200 kernels.  Mirrors ``models/tfimm/architectures/convnext.py`` of the reference (config
39-138, registry entries 440-620 incl. ``convnext_tiny_in22k``, the ckpts.json member; the author's stride-2 4x4 stem,
320-327): Keras-named / Keras-layout weights in, bf16 packed GEMM operands inside.

Layer -> kernel:
  stem Conv 4x4 stride 2 'valid' + bias      explicit im2col rows (3 input channels) + tcgen05 GEMM        (nn.conv2d)
  LayerNormalization eps 1e-6                vip_layernorm_bf16
  DepthwiseConv 7x7 (pad 3) + bias           vip_dwconv_bf16
  MLP fc1 + exact GELU, fc2                  tcgen05 GEMMs; the block's layer scale gamma is folded into fc2, the
                                             residual add (two-plane stream, see models/gcvit.py) sits in its epilogue
  downsample LayerNorm + Conv 2x2 stride 2   vip_layernorm_bf16 + implicit-GEMM convolution (im2col-mode TMA)
  head GAP -> LayerNorm -> Dense             vip_global_avgpool (f32) -> vip_layernorm_f32 -> vip_head_f32
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import nn

CONFIGS = {  # convnext.py:440-470
    "tiny": dict(embed_dim=(96, 192, 384, 768), nb_blocks=(3, 3, 9, 3)),
    "small": dict(embed_dim=(96, 192, 384, 768), nb_blocks=(3, 3, 27, 3)),
    "base": dict(embed_dim=(128, 256, 512, 1024), nb_blocks=(3, 3, 27, 3)),
}
PATCH, MLP_RATIO, LN_EPS = 4, 4.0, 1e-6
TWO_PLANE = os.environ.get("VIP_TWO_PLANE", "1") != "0"


class ConvNeXt:
    def __init__(self, variant="tiny", input_shape=(200, 200, 3), num_classes=2, head_act="softmax", first_down=1, device="cuda"):
        if variant not in CONFIGS:
            raise ValueError(f"unknown ConvNeXt variant {variant}")
        if head_act not in ("softmax", "sigmoid"):
            raise ValueError("head_act must be 'softmax' or 'sigmoid'")
        self.cfg, self.variant = CONFIGS[variant], variant
        self.input_shape, self.num_classes, self.head_act = tuple(input_shape), num_classes, head_act
        self.first_down, self.device = first_down, torch.device(device)
        self.name = f"convnext_{variant}"
        self.p = None

    def weight_shapes(self) -> dict:
        """Keras weight inventory (name -> shape): convnext.py:192-229 (block), 258-267 (downsample), 320-327 (stem), 432-438."""
        cfg, s = self.cfg, {}

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
        s["head/fc/kernel"], s["head/fc/bias"] = (d[-1], self.num_classes), (self.num_classes,)
        return s

    def init_random(self, seed=0):
        """Random initialisation like the Keras constructor (truncated-normal-scale kernels, unit LayerNorm, zero biases);
        the layer scale starts at 0.1 instead of 1e-6 so that the blocks contribute."""
        rng = np.random.default_rng(seed)
        W = {}
        for name, shp in self.weight_shapes().items():
            leaf = name.rsplit("/", 1)[1]
            if leaf in ("kernel", "depthwise_kernel"):
                fan_in = int(np.prod(shp[:-1])) if leaf == "kernel" else 49
                W[name] = (rng.standard_normal(shp) * np.sqrt(1.0 / fan_in)).astype(np.float32)
            elif leaf == "gamma" and "/blocks/" in name and name.split("/")[-2].isdigit():
                W[name] = np.full(shp, 0.1, np.float32)
            elif leaf == "gamma":
                W[name] = np.ones(shp, np.float32)
            else:
                W[name] = np.zeros(shp, np.float32)
        return self.load_weights(W)

    # ---- weight packing ------------------------------------------------------------------------------------------
    def _f32(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device).contiguous()

    def _bf(self, a):
        return nn.cast_bf16(self._f32(a)) if self.device.type == "cuda" else self._f32(a).to(torch.bfloat16)

    def _conv(self, W, name):
        """(kh,kw,Cin,Cout) -> bf16 [Cout, Kp] (K order r,s,c; K rounded up to 8), f32 bias."""
        k = np.asarray(W[name + "/kernel"], np.float32)
        w2 = k.reshape(-1, k.shape[3]).T
        kp = (w2.shape[1] + 7) // 8 * 8
        wp = np.zeros((w2.shape[0], kp), np.float32)
        wp[:, : w2.shape[1]] = w2
        return self._bf(wp), self._f32(W[name + "/bias"])

    def load_weights(self, W: dict):
        cfg, p = self.cfg, {}
        p["stem"] = self._conv(W, "stem/0")
        p["stem_ln"] = (self._f32(W["stem/1/gamma"]), self._f32(W["stem/1/beta"]))
        for j, nb in enumerate(cfg["nb_blocks"]):
            if j > 0:
                p[f"down{j}_ln"] = (self._f32(W[f"stages/{j}/downsample/0/gamma"]), self._f32(W[f"stages/{j}/downsample/0/beta"]))
                p[f"down{j}"] = self._conv(W, f"stages/{j}/downsample/1")
            for i in range(nb):
                n = f"stages/{j}/blocks/{i}"
                c = W[n + "/gamma"].shape[0]
                g = np.asarray(W[n + "/gamma"], np.float32)
                k2 = np.asarray(W[n + "/mlp/fc2/kernel"], np.float32) * g[None, :]      # layer scale folded: g * (h W2 + b2)
                p[f"b{j}_{i}"] = dict(
                    dw=self._f32(np.asarray(W[n + "/conv_dw/depthwise_kernel"]).reshape(7, 7, c)),
                    dw_b=self._f32(W[n + "/conv_dw/bias"]),
                    ln=(self._f32(W[n + "/norm/gamma"]), self._f32(W[n + "/norm/beta"])),
                    fc1=(self._bf(np.asarray(W[n + "/mlp/fc1/kernel"], np.float32).T), self._f32(W[n + "/mlp/fc1/bias"])),
                    fc2=(self._bf(k2.T), self._f32(np.asarray(W[n + "/mlp/fc2/bias"], np.float32) * g)))
        p["head_ln"] = (self._f32(W["head/norm/gamma"]), self._f32(W["head/norm/beta"]))
        p["head_w"], p["head_b"] = self._f32(W["head/fc/kernel"]), self._f32(W["head/fc/bias"])
        self.p = p
        return self

    # ---- forward ---------------------------------------------------------------------------------------------------
    def _block(self, x, x_lo, d):
        """ConvNeXtBlock (convnext.py:220-229): x + gamma * MLP(LN(DWConv7x7(x))).  The residual stream is carried as hi + lo
        bf16 planes through the fc2 epilogue (models/gcvit.py:_block explains why); the depthwise convolution reads hi."""
        b, h, w, c = x.shape
        y = nn.dwconv(x, d["dw"], d["dw_b"], ksize=7, stride=1)
        y = nn.layernorm(y, *d["ln"], eps=LN_EPS)
        hdn = nn.gemm(y.view(-1, c), *d["fc1"], act="gelu")
        lo = nn.lo_plane(b * h * w, c, x.device) if TWO_PLANE else None
        out = nn.gemm(hdn, *d["fc2"], residual=x.view(-1, c), residual_lo=x_lo, out_lo=lo)
        return out.view(b, h, w, c), lo

    def features(self, x, taps=None):
        p, cfg = self.p, self.cfg
        if p is None:
            raise RuntimeError("load_weights() first")
        x = nn.conv2d(x, *p["stem"], ksize=PATCH, stride=2 * self.first_down, pad=0)
        x = nn.layernorm(x, *p["stem_ln"], eps=LN_EPS)
        if taps is not None:
            taps["stem"] = x
        for j, nb in enumerate(cfg["nb_blocks"]):
            if j > 0:
                x = nn.layernorm(x, *p[f"down{j}_ln"], eps=LN_EPS)
                x = nn.conv2d(x, *p[f"down{j}"], ksize=2, stride=2, pad=0)
            x_lo = None
            for i in range(nb):
                x, x_lo = self._block(x, x_lo, p[f"b{j}_{i}"])
            if taps is not None:
                taps[f"stage{j}"] = x
        return x

    def forward(self, x, acc=None, acc_weight=1.0, taps=None):
        if x.dtype == torch.float32:
            x = nn.cast_bf16(x)
        f = self.features(x, taps)
        _, pooled = nn.global_avgpool(f, want_bf16=False, want_f32=True)
        feat = nn.layernorm_f32(pooled, *self.p["head_ln"], eps=LN_EPS)
        if taps is not None:
            taps["feat"] = feat
        return nn.head(feat, self.p["head_w"], self.p["head_b"], self.head_act == "sigmoid", acc, acc_weight)

    __call__ = forward
