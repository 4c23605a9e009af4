"""ResNeSt-50 forward on the B200 kernels.  Mirrors ``models/keras_cv_attention_models/resnest/resnest.py`` (ResNest50,
76-77) over ``aotnet/aotnet.py`` (AotNet 284-377, aot_block 137-192, deep stem 235-242) of the reference: Keras-named /
Keras-layout weights in, BatchNorm folded into the bf16 GEMM weights at load.

Layer -> kernel:
  deep stem: 3 x conv3x3 (+BN+ReLU), ZeroPad + MaxPool 3/2    explicit im2col (3 channels) / implicit-GEMM conv; vip_maxpool3s2_bf16
  1x1 convs (deep_1, deep_3, shortcut)                         tcgen05 GEMM; '3_' BN + shortcut add + ReLU in deep_3's epilogue
  split attention (resnest.py:27-66)                           one implicit-GEMM 3x3 conv per radix half (channel slices, BN +
                                                               ReLU epilogue) -> vip_global_avgpool -> two small GEMMs (the sum
                                                               over the radix is folded into the first one's weights) ->
                                                               vip_split_attention2_bf16 (r-softmax + weighted sum)
  stride-2 blocks: ZeroPad + AvgPool 3x3/2; AvgPool 'SAME' shortcut   vip_avgpool3s2_bf16, vip_avgpool2_same_bf16
  head GAP -> Dense                                            vip_global_avgpool (f32) -> vip_head_f32
"""
from __future__ import annotations

import numpy as np
import torch

from .. import nn

NUM_BLOCKS, OUT_CHANNELS, STRIDES, STEM_WIDTH, RADIX, BN_EPS = [3, 4, 6, 3], [256, 512, 1024, 2048], [1, 2, 2, 2], 64, 2, 1e-5


class ResNeSt50:
    def __init__(self, input_shape=(200, 200, 3), num_classes=2, classifier_activation="softmax", first_strides=2, device="cuda"):
        if classifier_activation not in ("softmax", "sigmoid"):
            raise ValueError("classifier_activation must be 'softmax' or 'sigmoid'")
        self.input_shape, self.num_classes, self.head_act = tuple(input_shape), num_classes, classifier_activation
        self.first_strides, self.device = first_strides, torch.device(device)
        self.name = "ResNest50"
        self.p = None

    def weight_shapes(self) -> dict:
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
        s["predictions/kernel"], s["predictions/bias"] = (cin, self.num_classes), (self.num_classes,)
        return s

    def init_random(self, seed=0):
        rng = np.random.default_rng(seed)
        W = {}
        for name, shp in self.weight_shapes().items():
            leaf = name.rsplit("/", 1)[1]
            if leaf == "kernel":
                W[name] = (rng.standard_normal(shp) * np.sqrt(2.0 / np.prod(shp[:-1]))).astype(np.float32)
            elif leaf in ("gamma", "moving_variance"):
                W[name] = np.ones(shp, np.float32) * (0.2 if name.endswith("_3_bn/gamma") and "sa_" not in name else 1.0)
            else:
                W[name] = np.zeros(shp, np.float32)
        return self.load_weights(W)

    # ---- weight packing ------------------------------------------------------------------------------------------
    def _f32(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device).contiguous()

    def _bf(self, a):
        return nn.cast_bf16(self._f32(a)) if self.device.type == "cuda" else self._f32(a).to(torch.bfloat16)

    @staticmethod
    def _bn(W, n):
        g, b = np.asarray(W[n + "bn/gamma"], np.float32), np.asarray(W[n + "bn/beta"], np.float32)
        m, v = np.asarray(W[n + "bn/moving_mean"], np.float32), np.asarray(W[n + "bn/moving_variance"], np.float32)
        s = g / np.sqrt(v + np.float32(BN_EPS))
        return s, b - m * s

    def _pack(self, k, scale, bias):
        """(kh,kw,Cin,Cout) kernel * per-output scale -> bf16 [Cout, Kp] (K order r,s,c; K rounded up to 8), f32 bias."""
        w2 = (np.asarray(k, np.float32) * scale[None, None, None, :]).reshape(-1, k.shape[3]).T
        wp = np.zeros((w2.shape[0], (w2.shape[1] + 7) // 8 * 8), np.float32)
        wp[:, : w2.shape[1]] = w2
        return self._bf(wp), self._f32(bias)

    def _conv_bn(self, W, conv, bnorm):
        s, b = self._bn(W, bnorm)
        return self._pack(W[conv + "conv/kernel"], s, b)

    def load_weights(self, W: dict):
        p = {"stem1": self._conv_bn(W, "stem_1_", "stem_1_"), "stem2": self._conv_bn(W, "stem_2_", "stem_2_"),
             "stem3": self._conv_bn(W, "stem_3_", "stem_")}
        cin = STEM_WIDTH
        for sid, (nb, oc, st) in enumerate(zip(NUM_BLOCKS, OUT_CHANNELS, STRIDES)):
            hidden = oc // 4
            for bid in range(nb):
                n, d = f"stack{sid + 1}_block{bid + 1}_", {}
                if bid == 0 and (st != 1 or cin != oc):
                    d["shortcut"] = self._conv_bn(W, n + "shortcut_", n + "shortcut_")
                d["deep1"] = self._conv_bn(W, n + "deep_1_", n + "deep_1_")
                sa = n + "deep_2_sa_"
                s1, b1 = self._bn(W, sa + "1_")                                     # BN over the concatenated radix outputs
                ws, bs = [], []
                for g in range(RADIX):
                    wg, bg = self._pack(W[f"{sa}1_g{g + 1}_conv/kernel"], s1[g * hidden:(g + 1) * hidden], b1[g * hidden:(g + 1) * hidden])
                    ws.append(wg), bs.append(bg)
                d["sa1"] = (ws, bs)
                # attention MLP on the pooled SUM of the radix splits: W2 (g1 + g2) = [W2 W2] [g1; g2] -> duplicate along K
                s2, b2 = self._bn(W, sa + "2_")
                k2 = np.asarray(W[sa + "2_conv/kernel"], np.float32)[0, 0] * s2[None, :]           # (hidden, inter)
                d["sa2"] = (self._bf(np.concatenate([k2, k2], 0).T),                                # [inter, 2 hidden]
                            self._f32(np.asarray(W[sa + "2_conv/bias"], np.float32) * s2 + b2))
                d["sa3"] = (self._bf(np.asarray(W[sa + "3_conv/kernel"], np.float32)[0, 0].T), self._f32(W[sa + "3_conv/bias"]))
                d["deep3"] = self._conv_bn(W, n + "deep_3_", n + "3_")
                p[n] = d
                cin = oc
        p["head_w"], p["head_b"] = self._f32(W["predictions/kernel"]), self._f32(W["predictions/bias"])
        self.p, self._ones = p, {}
        return self

    # ---- forward ---------------------------------------------------------------------------------------------------
    def _unit_gate(self, nimg, c, device):
        """relu((acc + bias) * 1 + shortcut): the activation after the add is the SE-tail epilogue with a gate of ones."""
        key = (nimg, c)
        if key not in self._ones:
            self._ones[key] = torch.ones((nimg, c), dtype=torch.float32, device=device)
        return self._ones[key]

    def _split_attention(self, x, d, stride):
        """split_attention_conv2d (resnest.py:27-66), radix 2."""
        n = x.shape[0]
        logits = nn.conv2d_grouped(x, *d["sa1"], ksize=3, stride=1, pad=1, act="relu")       # [N,H,W,2 hidden]
        pooled, _ = nn.global_avgpool(logits, want_bf16=True, want_f32=False)                # [N, 2 hidden]
        att = nn.gemm(pooled, *d["sa2"], act="relu")
        att = nn.gemm(att, *d["sa3"], out_dtype=torch.float32)                               # [N, 2 hidden] radix logits
        out = nn.split_attention2(logits, att)
        return nn.avgpool3s2(out) if stride > 1 else out

    def _block(self, x, d, filters, stride):
        """aot_block (aotnet.py:137-192) with the ResNeSt options."""
        if "shortcut" in d:
            sc = nn.avgpool2_same(x) if stride > 1 else x
            sc = nn.conv2d(sc, *d["shortcut"])
        else:
            sc = x
        y = nn.conv2d(x, *d["deep1"], act="relu")
        y = self._split_attention(y, d, stride)
        n, h, w, _ = y.shape
        return nn.conv2d(y, *d["deep3"], act="relu", residual=sc, row_gate=self._unit_gate(n, filters, x.device), gate_rows=h * w)

    def features(self, x, taps=None):
        p = self.p
        if p is None:
            raise RuntimeError("load_weights() first")
        x = nn.conv2d(x, *p["stem1"], ksize=3, stride=self.first_strides, pad=1, act="relu")
        x = nn.conv2d(x, *p["stem2"], ksize=3, stride=1, pad=1, act="relu")
        x = nn.conv2d(x, *p["stem3"], ksize=3, stride=1, pad=1, act="relu")
        x = nn.maxpool3s2(x)
        if taps is not None:
            taps["stem"] = x
        for sid, (nb, oc, st) in enumerate(zip(NUM_BLOCKS, OUT_CHANNELS, STRIDES)):
            for bid in range(nb):
                x = self._block(x, p[f"stack{sid + 1}_block{bid + 1}_"], oc, st if bid == 0 else 1)
            if taps is not None:
                taps[f"stack{sid + 1}"] = x
        return x

    def forward(self, x, acc=None, acc_weight=1.0, taps=None):
        if x.dtype == torch.float32:
            x = nn.cast_bf16(x)
        f = self.features(x, taps)
        _, feat = nn.global_avgpool(f, want_bf16=False, want_f32=True)
        if taps is not None:
            taps["feat"] = feat
        return nn.head(feat, self.p["head_w"], self.p["head_b"], self.head_act == "sigmoid", acc, acc_weight)

    __call__ = forward
