"""ResNet-RS forward on the B200 kernels.  Mirrors the constructors of the reference
(``models/resnet_rs/__init__.py:3-11``, ``resnet_rs_model.py:329-567``): same layer names, same Keras weight layout in,
BatchNorm folded into the bf16 GEMM weights at load time.

Layer -> kernel:
  Conv2DFixedPadding 1x1            tcgen05 GEMM directly on the NHWC activation            (nn.gemm)
  Conv2DFixedPadding 3x3 (s1 / s2)  im2col view + tcgen05 GEMM, BN bias + ReLU in the epilogue (nn.conv2d)
  AveragePooling2D 'same'           vip_avgpool2_same_bf16
  SE                                vip_global_avgpool -> two batch-wide tcgen05 GEMMs (bias+relu / bias+sigmoid)
  excite + Add + ReLU               vip_scale_add_act_bf16
  head                              vip_global_avgpool (f32) -> vip_head_f32 (Dense + softmax|sigmoid [+ ensemble acc])
"""
from __future__ import annotations

import numpy as np
import torch

from .. import nn

BLOCK_ARGS = {  # models/resnet_rs/block_args.py:1-44
    50: [(64, 3), (128, 4), (256, 6), (512, 3)],
    101: [(64, 3), (128, 4), (256, 23), (512, 3)],
    152: [(64, 3), (128, 8), (256, 36), (512, 3)],
    200: [(64, 3), (128, 24), (256, 36), (512, 3)],
}
BN_EPS = 1e-5


def _fold_conv_bn(W, conv, bnorm, device):
    """(kh,kw,Cin,Cout) kernel + BN stats -> bf16 [Cout, Kp] (K order r,s,c; Kp = K rounded up to 8) and f32 bias."""
    k = np.asarray(W[conv + "/kernel"], dtype=np.float32)
    g, b = np.asarray(W[bnorm + "/gamma"], np.float32), np.asarray(W[bnorm + "/beta"], np.float32)
    m, v = np.asarray(W[bnorm + "/moving_mean"], np.float32), np.asarray(W[bnorm + "/moving_variance"], np.float32)
    s = g / np.sqrt(v + BN_EPS)
    co = k.shape[3]
    w2 = (k * s[None, None, None, :]).reshape(-1, co).T  # [Cout, K]
    kk = w2.shape[1]
    kp = (kk + 7) // 8 * 8
    wp = np.zeros((co, kp), np.float32)
    wp[:, :kk] = w2
    return (nn.to_bf16(wp, device),
            torch.from_numpy(b - m * s).to(device).contiguous())


def _pair_columns(w, bias, cin, cout):
    """3x3 stride-1 'same' convolution on [H, W, cin] restated on horizontal pixel PAIRS: input [H, W/2, 2 cin], output
    [H, W/2, 2 cout] (the same memory as [H, W, cout]).  With cin = 32 a 64-channel k-block of the implicit GEMM is then
    full instead of half zero-filled and the number of output tiles halves, which is what bounds these memory-heavy
    stem layers (stem_conv_2/3, resnet_rs_model.py:104-123).  w: bf16-to-be f32 [cout, 9*cin] in (r, s, c) order."""
    w = w.reshape(cout, 3, 3, cin)
    wp = np.zeros((2, cout, 3, 3, 2, cin), np.float32)            # [dxo, co, r, sxs, half, ci]
    for dxo in range(2):
        for s in range(3):                                        # original tap s - 1 in {-1, 0, 1}
            off = dxo + s - 1                                     # input pixel offset from the pair's first pixel
            sxs, half = off // 2, off % 2                         # pair offset in {-1, 0, 1}, pixel inside that pair
            wp[dxo, :, :, sxs + 1, half, :] = w[:, :, s, :]
    return wp.reshape(2 * cout, 9 * 2 * cin), np.concatenate([bias, bias])


def _dense(W, name, device):
    k = np.asarray(W[name + "/kernel"], np.float32)
    k = k.reshape(-1, k.shape[-1])  # (1,1,Cin,Cout) -> (Cin,Cout)
    return (nn.to_bf16(k.T, device),
            torch.from_numpy(np.asarray(W[name + "/bias"], np.float32)).to(device).contiguous())


def _se_reduce_on_conv2(W, n, device):
    """The SE squeeze of a bottleneck pools the conv_3 + BN output (resnet_rs_model.py:149,253-262).  conv_3 is 1x1 and BN is
    affine at inference, so GAP(BN(conv_3(y2))) = W3' mean(y2) + b3': the gate can be computed from the pooled conv_2
    output BEFORE conv_3 runs and applied inside conv_3's epilogue.  Returns se_reduce composed with that affine map:
    bf16 [f, f] weights and f32 bias acting on mean(y2)."""
    k3 = np.asarray(W[n + "conv_3/kernel"], np.float64)[0, 0]                      # (f, 4f)
    g, b = np.asarray(W[n + "batch_norm_3/gamma"], np.float64), np.asarray(W[n + "batch_norm_3/beta"], np.float64)
    m, v = np.asarray(W[n + "batch_norm_3/moving_mean"], np.float64), np.asarray(W[n + "batch_norm_3/moving_variance"], np.float64)
    s = g / np.sqrt(v + BN_EPS)
    w3, b3 = k3 * s[None, :], b - m * s                                             # y3 = y2 @ w3 + b3
    k1 = np.asarray(W[n + "se_reduce/kernel"], np.float64)[0, 0]                    # (4f, f)
    # float64 products: independent of the host thread count up to ~1e-16, far below the bf16 / f32 roundings that follow
    wc = w3 @ k1                                                                    # (f, f): mean(y2) -> hidden pre-activation
    bc = (b3[:, None] * k1).sum(0) + np.asarray(W[n + "se_reduce/bias"], np.float64)
    return (nn.to_bf16(wc.T.astype(np.float32), device),
            torch.from_numpy(bc.astype(np.float32)).to(device).contiguous())


class ResNetRS:
    def __init__(self, depth=50, input_shape=(200, 200, 3), classes=2, classifier_activation="softmax", first_strides=2,
                 device="cuda"):
        if depth not in BLOCK_ARGS:
            raise ValueError(f"ResNetRS depth {depth} not supported")
        if classifier_activation not in ("softmax", "sigmoid"):
            raise ValueError("classifier_activation must be 'softmax' or 'sigmoid'")
        self.depth, self.input_shape, self.classes = depth, tuple(input_shape), classes
        self.head_act, self.first_strides = classifier_activation, first_strides
        self.device = torch.device(device)
        self.name = f"ResNetRS{depth}"
        self.p = None

    def weight_shapes(self) -> dict:
        """Keras weight inventory (name -> shape) of this architecture (resnet_rs_model.py:64-183, 468-476)."""
        s = {}

        def bnorm(n, c):
            for q in ("gamma", "beta", "moving_mean", "moving_variance"):
                s[f"{n}/{q}"] = (c,)

        for i, (ci, co) in enumerate(((3, 32), (32, 32), (32, 64), (64, 64)), 1):
            s[f"stem_conv_{i}/kernel"] = (3, 3, ci, co)
            bnorm(f"stem_batch_norm_{i}", co)
        cin = 64
        for gi, (f, reps) in enumerate(BLOCK_ARGS[self.depth]):
            for bi in range(reps):
                n = f"c{gi + 2}_block_{bi}_"
                if bi == 0:
                    s[n + "projection_conv/kernel"] = (1, 1, cin, 4 * f)
                    bnorm(n + "projection_batch_norm", 4 * f)
                for j, (kk, a, b) in enumerate(((1, cin, f), (3, f, f), (1, f, 4 * f)), 1):
                    s[n + f"conv_{j}/kernel"] = (kk, kk, a, b)
                    bnorm(n + f"batch_norm_{j}", b)
                s[n + "se_reduce/kernel"], s[n + "se_reduce/bias"] = (1, 1, 4 * f, f), (f,)
                s[n + "se_expand/kernel"], s[n + "se_expand/bias"] = (1, 1, f, 4 * f), (4 * f,)
                cin = 4 * f
        s["predictions/kernel"], s["predictions/bias"] = (cin, self.classes), (self.classes,)
        return s

    def init_random(self, seed=0):
        """``weights=None`` of the Keras constructor: fan-in scaled normal kernels, identity BatchNorm, zero biases."""
        rng = np.random.default_rng(seed)
        W = {}
        for name, shp in self.weight_shapes().items():
            leaf = name.rsplit("/", 1)[1]
            if leaf == "kernel":
                W[name] = (rng.standard_normal(shp) * np.sqrt(2.0 / np.prod(shp[:-1]))).astype(np.float32)
            elif leaf in ("gamma", "moving_variance"):
                W[name] = np.ones(shp, np.float32) * (0.3 if name.endswith("batch_norm_3/gamma") else 1.0)
            else:
                W[name] = np.zeros(shp, np.float32)
        return self.load_weights(W)

    # Keras-named weights in
    def load_weights(self, W: dict):
        d, p = self.device, {}
        for i in range(1, 5):
            p[f"stem{i}"] = _fold_conv_bn(W, f"stem_conv_{i}", f"stem_batch_norm_{i}", d)
        p["stem1_pair"] = nn.pair_rows_weights(*p["stem1"])
        for i in (2, 3):   # 32-channel inputs: pixel-pair form (see _pair_columns); used when the map width is even
            wt, bias = p[f"stem{i}"]
            w2, b2 = _pair_columns(wt.float().cpu().numpy(), bias.cpu().numpy(), 32, wt.shape[0])
            p[f"stem{i}_pair"] = (nn.to_bf16(w2, d), torch.from_numpy(b2).to(d).contiguous())
        for gi, (f, reps) in enumerate(BLOCK_ARGS[self.depth]):
            for bi in range(reps):
                n = f"c{gi + 2}_block_{bi}_"
                if bi == 0:
                    p[n + "proj"] = _fold_conv_bn(W, n + "projection_conv", n + "projection_batch_norm", d)
                for j in (1, 2, 3):
                    p[n + f"conv{j}"] = _fold_conv_bn(W, n + f"conv_{j}", n + f"batch_norm_{j}", d)
                p[n + "se1"] = _se_reduce_on_conv2(W, n, d)
                p[n + "se2"] = _dense(W, n + "se_expand", d)
        p["head_w"] = torch.from_numpy(np.asarray(W["predictions/kernel"], np.float32)).to(d).contiguous()
        p["head_b"] = torch.from_numpy(np.asarray(W["predictions/bias"], np.float32)).to(d).contiguous()
        self.p = p
        return self

    def _bottleneck(self, x, n, strides, use_projection, gap):
        p = self.p
        shortcut = x
        if use_projection:
            s_in = nn.avgpool2_same(x) if strides == 2 else x
            shortcut = nn.conv2d(s_in, *p[n + "proj"])
        y = nn.conv2d(x, *p[n + "conv1"], act="relu")
        # conv_2 + BN + ReLU; its epilogue also accumulates the per-image channel sums the SE gate needs
        y = nn.conv2d(y, *p[n + "conv2"], ksize=3, stride=strides, pad=1, act="relu", gap=gap, gap_rows=None)
        b, h, w, f = y.shape
        pooled = nn.scale_cast_bf16(gap, 1.0 / (h * w))
        hid = nn.gemm(pooled, *p[n + "se1"], act="relu")
        gate = nn.gemm(hid, *p[n + "se2"], act="sigmoid", out_dtype=torch.float32)
        # conv_3 + BN, excite, + shortcut, ReLU in one epilogue (resnet_rs_model.py:253-280)
        return nn.conv2d(y, *p[n + "conv3"], act="relu", residual=shortcut, row_gate=gate, gate_rows=h * w)

    def features(self, x, taps=None):
        """x bf16 [N,H,W,3] -> bf16 [N,h,w,2048]"""
        p = self.p
        if p is None:
            raise RuntimeError("load_weights() first")
        for i, s in ((1, self.first_strides), (2, 1), (3, 1), (4, 2)):
            n, h, w, c = x.shape
            if i == 1 and c == 3 and (n * ((h - 1) // s + 1) * ((w - 1) // s + 1)) % 2 == 0:
                x = nn.conv2d_paired(x, *p["stem1_pair"], ksize=3, stride=s, pad=1, act="relu")
            elif f"stem{i}_pair" in p and w % 2 == 0 and c == 32:
                y = nn.conv2d(x.view(n, h, w // 2, 2 * c), *p[f"stem{i}_pair"], ksize=3, stride=1, pad=1, act="relu")
                x = y.view(n, h, w, -1)
            else:
                x = nn.conv2d(x, *p[f"stem{i}"], ksize=3, stride=s, pad=1, act="relu")
        if taps is not None:
            taps["stem"] = x
        nimg = x.shape[0]
        nblocks = sum(r for _, r in BLOCK_ARGS[self.depth])
        gaps = nn.zero_(torch.empty((nblocks, nimg, 512), dtype=nn.STATS, device=x.device))  # one memset (fixed-point sums)
        k = 0
        for gi, (f, reps) in enumerate(BLOCK_ARGS[self.depth]):
            for bi in range(reps):
                gap = gaps[k].view(-1)[: nimg * f].view(nimg, f)
                k += 1
                x = self._bottleneck(x, f"c{gi + 2}_block_{bi}_", (1 if gi == 0 else 2) if bi == 0 else 1, bi == 0, gap)
            if taps is not None:
                taps[f"c{gi + 2}"] = x
        return x

    def forward(self, x, acc=None, acc_weight=1.0, taps=None):
        """x: bf16 (or f32) [N,H,W,3] in [0,1] on the device -> probabilities f32 [N,classes]."""
        if x.dtype == torch.float32:
            x = nn.cast_bf16(x)
        f = self.features(x, taps)
        _, feat = nn.global_avgpool(f, want_bf16=False, want_f32=True)
        if taps is not None:
            taps["feat"] = feat
        return nn.head(feat, self.p["head_w"], self.p["head_b"], self.head_act == "sigmoid", acc, acc_weight)

    __call__ = forward


def ResNetRS50(**kw):
    return ResNetRS(50, **kw)


def ResNetRS101(**kw):
    return ResNetRS(101, **kw)


def ResNetRS152(**kw):
    return ResNetRS(152, **kw)


def ResNetRS200(**kw):
    return ResNetRS(200, **kw)
