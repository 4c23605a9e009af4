"""ECA-NFNet-L0 forward on the B200 kernels.  Mirrors ``models/keras_cv_attention_models/nfnets/nfnets.py`` of the reference
(ECA_NFNetL0, 316-320 -> NormFreeNet_Light 304-306 -> NormFreeNet 194-269): Keras-named / Keras-layout weights in.

Load-time folds: ScaledStandardizedConv2D (nfnets.py:42-81) standardises its kernel on every call -- (w - mean) *
rsqrt(max(var * fan_in, eps)) * gain * gamma over (kh, kw, Cin) -- which is a constant at inference: done once in fp32,
then bf16.  The block's attention gain (2.0) and alpha (0.2) go into the ECA gate.

Layer -> kernel:
  stem conv 1 (3 input channels)               explicit im2col rows + tcgen05 GEMM, bias + swish epilogue
  std convs 3x3 (stem 2-4), 1x1 (deep_1/4, shortcut, post)   implicit-GEMM conv / plain GEMM, bias (+ swish) epilogue
  grouped 3x3 (deep_2/3, 64 channels per group)               one implicit-GEMM conv per group on a channel slice
  pre-activation swish(x) * beta                vip_act_scale_bf16
  AvgPool2D(2, 'SAME') shortcut                 vip_avgpool2_same_bf16
  ECA                                           pooled sums fused in deep_4's epilogue -> vip_eca_gate_f32 -> vip_scale_add_act_bf16
  head GAP -> Dense                             vip_global_avgpool (f32) -> vip_head_f32
"""
from __future__ import annotations

import numpy as np
import torch

from .. import nn

SWISH_GAMMA = 1.7881293296813965     # NON_LINEAR_GAMMA["swish"], nfnets.py:35 (gamma_in_act=False: the convs carry it)
NUM_BLOCKS, OUT_CHANNELS, STRIDES = [1, 2, 6, 3], [256, 512, 1536, 1536], [1, 2, 2, 2]
CHANNEL_RATIO, GROUP_SIZE, ALPHA, STEM_WIDTH, FEATURES, ATTN_GAIN, STD_EPS = 0.25, 64, 0.2, 128, 2304, 2.0, 1e-5


def _betas():
    """nfnets.py:246-255, 170-178."""
    beta_list = [(1 + ALPHA ** 2 * i) ** -0.5 for i in range(max(NUM_BLOCKS) + 1)]
    out, pre = [], 1.0
    for nb in NUM_BLOCKS:
        b = beta_list[: nb + 1]
        b[0] = pre
        out.append(b[:nb])
        pre = b[-1]
    return out


class ECANFNetL0:
    def __init__(self, input_shape=(200, 200, 3), num_classes=2, classifier_activation="softmax", first_strides=2, device="cuda"):
        if classifier_activation not in ("softmax", "sigmoid"):
            raise ValueError("classifier_activation must be 'softmax' or 'sigmoid'")
        self.input_shape, self.num_classes, self.head_act = tuple(input_shape), num_classes, classifier_activation
        self.first_strides, self.device = first_strides, torch.device(device)
        self.name = "ECA_NFNetL0"
        self.p = None

    def weight_shapes(self) -> dict:
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
                if (st if bid == 0 else 1) > 1 or cin != oc:
                    sconv(n + "shortcut_", 1, cin, oc)
                sconv(n + "deep_1_", 1, cin, hidden)
                sconv(n + "deep_2_", 3, hidden, hidden, g)
                sconv(n + "deep_3_", 3, hidden, hidden, g)
                sconv(n + "deep_4_", 1, hidden, oc)
                s[n + "eca_conv1d/kernel"] = (5, 1, 1)
                cin = oc
        sconv("post_", 1, cin, FEATURES)
        s["predictions/kernel"], s["predictions/bias"] = (FEATURES, self.num_classes), (self.num_classes,)
        return s

    def init_random(self, seed=0):
        rng = np.random.default_rng(seed)
        W = {}
        for name, shp in self.weight_shapes().items():
            leaf = name.rsplit("/", 1)[1]
            if leaf == "kernel":
                W[name] = rng.standard_normal(shp).astype(np.float32) * (0.05 if name.startswith("predictions") else 1.0)
            elif leaf == "gain":
                W[name] = np.full(shp, 0.4 if "deep_4_" in name else 1.0, np.float32)
            else:
                W[name] = np.zeros(shp, np.float32)
        return self.load_weights(W)

    # ---- weight packing ------------------------------------------------------------------------------------------
    def _f32(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device).contiguous()

    def _bf(self, a):
        return nn.cast_bf16(self._f32(a)) if self.device.type == "cuda" else self._f32(a).to(torch.bfloat16)

    def _std_conv(self, W, n, groups=1):
        """Standardised kernel (nfnets.py:64-70) -> per group bf16 [Cout_g, Kp] (K order r,s,c, K rounded up to 8) + f32 bias."""
        k = np.asarray(W[n + "conv/kernel"], np.float64)                       # (kh, kw, Cin/groups, Cout)
        mean, var = k.mean(axis=(0, 1, 2), keepdims=True), k.var(axis=(0, 1, 2), keepdims=True)
        fan_in = k.shape[0] * k.shape[1] * k.shape[2]
        scale = 1.0 / np.sqrt(np.maximum(var * fan_in, STD_EPS)) * (np.asarray(W[n + "conv/gain"], np.float64) * SWISH_GAMMA)
        ks = ((k - mean) * scale).astype(np.float32)
        bias = np.asarray(W[n + "conv/bias"], np.float32)
        cout = ks.shape[3]
        cg = cout // groups
        ws, bs = [], []
        for g in range(groups):
            w2 = ks[:, :, :, g * cg: (g + 1) * cg].reshape(-1, cg).T               # [Cout_g, kh*kw*Cin_g]
            kp = (w2.shape[1] + 7) // 8 * 8
            wp = np.zeros((cg, kp), np.float32)
            wp[:, : w2.shape[1]] = w2
            ws.append(self._bf(wp))
            bs.append(self._f32(bias[g * cg: (g + 1) * cg]))
        return (ws, bs) if groups > 1 else (ws[0], bs[0])

    def load_weights(self, W: dict):
        p = {f"stem{i}": self._std_conv(W, f"stem_{i}_") for i in range(1, 5)}
        cin = STEM_WIDTH
        for sid, (nb, oc, st) in enumerate(zip(NUM_BLOCKS, OUT_CHANNELS, STRIDES)):
            hidden = int(oc * CHANNEL_RATIO)
            g = hidden // GROUP_SIZE
            for bid in range(nb):
                n = f"stack{sid + 1}_block{bid + 1}_"
                d = {}
                if (st if bid == 0 else 1) > 1 or cin != oc:
                    d["shortcut"] = self._std_conv(W, n + "shortcut_")
                d["deep1"], d["deep4"] = self._std_conv(W, n + "deep_1_"), self._std_conv(W, n + "deep_4_")
                d["deep2"], d["deep3"] = self._std_conv(W, n + "deep_2_", g), self._std_conv(W, n + "deep_3_", g)
                d["eca"] = self._f32(np.asarray(W[n + "eca_conv1d/kernel"]).reshape(-1))
                p[n] = d
                cin = oc
        p["post"] = self._std_conv(W, "post_")
        p["head_w"], p["head_b"] = self._f32(W["predictions/kernel"]), self._f32(W["predictions/bias"])
        self.p = p
        return self

    # ---- forward ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _conv3(x, wb, stride, act, groups):
        if groups > 1:
            return nn.conv2d_grouped(x, wb[0], wb[1], ksize=3, stride=stride, pad=1, act=act)
        return nn.conv2d(x, wb[0], wb[1], ksize=3, stride=stride, pad=1, act=act)

    def _block(self, x, d, filters, beta, stride):
        """nfnets.py:116-168."""
        n, h, w, cin = x.shape
        groups = int(filters * CHANNEL_RATIO) // GROUP_SIZE
        preact = nn.act_scale(x, "swish", beta)
        if "shortcut" in d:
            sc = nn.avgpool2_same(preact) if stride > 1 else preact
            sc = nn.conv2d(sc, *d["shortcut"])
        else:
            sc = x
        y = nn.conv2d(preact, *d["deep1"], act="swish")
        y = self._conv3(y, d["deep2"], stride, "swish", groups)
        y = self._conv3(y, d["deep3"], 1, "swish", groups)
        gap = nn.zero_(torch.empty((n, filters), dtype=nn.STATS, device=x.device))
        y = nn.conv2d(y, *d["deep4"], gap=gap)                 # the ECA squeeze is accumulated by this epilogue
        gate = nn.eca_gate(gap, d["eca"], y.shape[1] * y.shape[2], out_scale=ATTN_GAIN * ALPHA)
        return nn.scale_add_act(y, gate, sc)                   # shortcut + alpha * 2 * eca(y)

    def features(self, x, taps=None):
        p = self.p
        if p is None:
            raise RuntimeError("load_weights() first")
        x = nn.conv2d(x, *p["stem1"], ksize=3, stride=self.first_strides, pad=1, act="swish")
        x = nn.conv2d(x, *p["stem2"], ksize=3, stride=1, pad=1, act="swish")
        x = nn.conv2d(x, *p["stem3"], ksize=3, stride=1, pad=1, act="swish")
        x = nn.conv2d(x, *p["stem4"], ksize=3, stride=2, pad=1)
        if taps is not None:
            taps["stem"] = x
        for sid, (nb, oc, st, bs) in enumerate(zip(NUM_BLOCKS, OUT_CHANNELS, STRIDES, _betas())):
            for bid in range(nb):
                x = self._block(x, p[f"stack{sid + 1}_block{bid + 1}_"], oc, bs[bid], st if bid == 0 else 1)
            if taps is not None:
                taps[f"stack{sid + 1}"] = x
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
