"""CPU ORACLE (test infrastructure, not shipped) -- GCViT forward in PyTorch fp32, restating ``models/gcvit`` of the
reference on Keras-layout, Keras-named weights.

Reference -> here:
  GCViT / NAME2CONFIG       models/gcvit/models/gcvit.py:9-125          -> CONFIGS, :func:`forward`
  Stem                      layers/embedding.py:7-29                    -> :func:`stem`
  ReduceSize, SE            layers/feature.py:81-120, 46-78             -> :func:`reduce_size`, :func:`se`
  FeatExtract, GlobalQueryGen  layers/feature.py:123-188                -> :func:`feat_extract`
  GCViTLevel, FitWindow     layers/level.py:46-67, feature.py:234-256   -> :func:`level`
  GCViTBlock, Mlp           layers/block.py:60-81, feature.py:8-43      -> :func:`block`
  WindowAttention           layers/attention.py:39-83                   -> :func:`window_attention`, :func:`relative_position_index`
  window_partition/reverse  layers/window.py:3-14

Pitfalls mirrored: FeatExtract max-pool sees explicit ZERO padding; ReduceSize residual is taken after
norm1; q_global is neither normalised nor projected and is shared by the odd blocks of a level; GELU is the exact erf
form; LayerNorm eps 1e-5; SE Dense layers have no bias.  Parity status: unpinned against real Keras (no TensorFlow
offline); known answers = parameter counts of the vendored doc table and stage shapes (Appendix B.1).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

CONFIGS = {  # gcvit.py:9-42
    "xxtiny": dict(window_size=(7, 7, 14, 7), dim=64, depths=(2, 2, 6, 2), num_heads=(2, 4, 8, 16), mlp_ratio=3.0, layer_scale=None),
    "xtiny": dict(window_size=(7, 7, 14, 7), dim=64, depths=(3, 4, 6, 5), num_heads=(2, 4, 8, 16), mlp_ratio=3.0, layer_scale=None),
    "tiny": dict(window_size=(7, 7, 14, 7), dim=64, depths=(3, 4, 19, 5), num_heads=(2, 4, 8, 16), mlp_ratio=3.0, layer_scale=None),
    "small": dict(window_size=(7, 7, 14, 7), dim=96, depths=(3, 4, 19, 5), num_heads=(3, 6, 12, 24), mlp_ratio=2.0, layer_scale=1e-5),
    "base": dict(window_size=(7, 7, 14, 7), dim=128, depths=(3, 4, 19, 5), num_heads=(4, 8, 16, 32), mlp_ratio=2.0, layer_scale=1e-5),
}
KEEP_DIMS = [(False, False, False), (False, False), (True,), (True,)]  # gcvit.py:70
LN_EPS = 1e-5


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).float()


def _conv_w(k):
    return _t(k).permute(3, 2, 0, 1).contiguous()


def ln(x, W, name):  # x [..., C]
    return F.layer_norm(x, (x.shape[-1],), _t(W[name + "/gamma"]), _t(W[name + "/beta"]), LN_EPS)


def gelu(x):
    return F.gelu(x)  # exact erf form


def se(x, W, name):  # x NHWC
    p = x.mean(dim=(1, 2))
    p = gelu(p @ _t(W[name + "/fc/0/kernel"]))
    p = torch.sigmoid(p @ _t(W[name + "/fc/2/kernel"]))
    return x * p[:, None, None, :]


def mbconv(x, W, name):
    """pad1 -> DepthwiseConv 3x3 valid -> GELU -> SE -> Conv1x1 (no bias), NHWC in/out (feature.py:92-97,132-137)."""
    c = x.shape[-1]
    xc = x.permute(0, 3, 1, 2)
    dw = _t(W[name + "/conv/0/depthwise_kernel"]).permute(2, 3, 0, 1).contiguous()  # (kh,kw,C,1) -> (C,1,kh,kw)
    y = F.conv2d(F.pad(xc, (1, 1, 1, 1)), dw, groups=c).permute(0, 2, 3, 1)
    y = se(gelu(y), W, name + "/conv/2")
    return y @ _t(W[name + "/conv/3/kernel"]).reshape(c, c)


def conv3x3_pad1(x, kernel, stride, bias=None):  # NHWC, explicit zero pad 1 + 'valid'
    xc = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1))
    y = F.conv2d(xc, _conv_w(kernel), None if bias is None else _t(bias), stride=stride)
    return y.permute(0, 2, 3, 1)


def reduce_size(x, W, name, first_strides=2):
    x = ln(x, W, name + "/norm1")
    x = x + mbconv(x, W, name)
    x = conv3x3_pad1(x, W[name + "/reduction/kernel"], first_strides)
    return ln(x, W, name + "/norm2")


def feat_extract(x, W, name, keep_dim):
    x = x + mbconv(x, W, name)
    if not keep_dim:
        xc = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1))  # zeros take part in the max
        x = F.max_pool2d(xc, 3, 2).permute(0, 2, 3, 1)
    return x


def stem(x, W, first_strides=2):
    x = conv3x3_pad1(x, W["patch_embed/proj/kernel"], 2, W["patch_embed/proj/bias"])
    return reduce_size(x, W, "patch_embed/conv_down", first_strides)


def relative_position_index(ws: int) -> np.ndarray:
    """attention.py:39-50: index [N,N] into the ((2ws-1)^2, heads) table."""
    coords = np.stack(np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij")).reshape(2, -1)
    rel = coords[:, :, None] - coords[:, None, :]
    return (rel[0] + ws - 1) * (2 * ws - 1) + (rel[1] + ws - 1)


def window_partition(x, ws):
    b, h, w, c = x.shape
    x = x.reshape(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws * ws, c)


def window_reverse(win, ws, b, h, w, c):
    x = win.reshape(b, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(b, h, w, c)


def window_attention(xw, W, name, heads, ws, q_global=None):
    """xw [B_, N, C]; q_global [B, ws, ws, C] or None."""
    b_, n, c = xw.shape
    hd = c // heads
    qkv = xw @ _t(W[name + "/qkv/kernel"]) + _t(W[name + "/qkv/bias"])
    if q_global is not None:
        kv = qkv.reshape(b_, n, 2, heads, hd).permute(2, 0, 3, 1, 4)
        k, v = kv[0], kv[1]
        b = q_global.shape[0]
        q = q_global.reshape(b, n, c).repeat_interleave(b_ // b, dim=0)      # tf.repeat(q_global, B_//B, axis=0)
        q = q.reshape(b_, n, heads, hd).permute(0, 2, 1, 3)
    else:
        t = qkv.reshape(b_, n, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = t[0], t[1], t[2]
    q = q * (hd ** -0.5)
    attn = q @ k.transpose(-1, -2)
    table = _t(W[name + "/relative_position_bias_table"])                    # [(2ws-1)^2, heads]
    idx = torch.from_numpy(relative_position_index(ws).reshape(-1))
    bias = table[idx].reshape(n, n, heads).permute(2, 0, 1)
    attn = torch.softmax(attn + bias[None], dim=-1)
    out = (attn @ v).permute(0, 2, 1, 3).reshape(b_, n, c)
    return out @ _t(W[name + "/proj/kernel"]) + _t(W[name + "/proj/bias"])


def block(x, W, name, heads, ws, q_global):
    b, h, w, c = x.shape
    t = window_partition(ln(x, W, name + "/norm1"), ws)
    t = window_attention(t, W, name + "/attn", heads, ws, q_global)
    t = window_reverse(t, ws, b, h, w, c)
    g1 = _t(W[name + "/gamma1"]) if name + "/gamma1" in W else 1.0
    g2 = _t(W[name + "/gamma2"]) if name + "/gamma2" in W else 1.0
    x = x + t * g1
    m = ln(x, W, name + "/norm2")
    m = gelu(m @ _t(W[name + "/mlp/fc1/kernel"]) + _t(W[name + "/mlp/fc1/bias"]))
    m = m @ _t(W[name + "/mlp/fc2/kernel"]) + _t(W[name + "/mlp/fc2/bias"])
    return x + g2 * m


def level(x, W, li, depth, heads, ws, keep_dims, downsample):
    b, h, w, c = x.shape
    hp, wp = (ws - h % ws) % ws, (ws - w % ws) % ws
    x = F.pad(x, (0, 0, wp // 2, wp // 2 + wp % 2, hp // 2, hp // 2 + hp % 2))     # FitWindow
    q = x
    for k, keep in enumerate(keep_dims):
        q = feat_extract(q, W, f"levels/{li}/q_global_gen/to_q_global/{k}", keep)
    for i in range(depth):
        x = block(x, W, f"levels/{li}/blocks/{i}", heads, ws, q if i % 2 else None)
    x = x[:, :h, :w, :]
    if downsample:
        x = reduce_size(x, W, f"levels/{li}/downsample")
    return x


def forward(x_nhwc, W, variant="tiny", head_act="softmax", return_logits=False, first_strides=2, taps=None):
    cfg = CONFIGS[variant]
    with torch.no_grad():
        x = stem(_t(x_nhwc), W, first_strides)
        if taps is not None:
            taps["stem"] = x.numpy().copy()
        for i, d in enumerate(cfg["depths"]):
            x = level(x, W, i, d, cfg["num_heads"][i], cfg["window_size"][i], KEEP_DIMS[i], i < 3)
            if taps is not None:
                taps[f"level{i}"] = x.numpy().copy()
        x = ln(x, W, "norm")
        feat = x.mean(dim=(1, 2))
        if taps is not None:
            taps["feat"] = feat.numpy().copy()
        logits = feat @ _t(W["head/kernel"]) + _t(W["head/bias"])
        if return_logits:
            return logits.numpy()
        return (torch.softmax(logits, -1) if head_act == "softmax" else torch.sigmoid(logits)).numpy()


# ---- weight inventory (Keras names / layouts) and seeded random init -------------------------------
def weight_shapes(variant="tiny", num_classes=2) -> dict:
    cfg = CONFIGS[variant]
    s = {}

    def lnorm(n, c):
        s[n + "/gamma"] = (c,)
        s[n + "/beta"] = (c,)

    def mb(n, c):
        s[n + "/conv/0/depthwise_kernel"] = (3, 3, c, 1)
        s[n + "/conv/2/fc/0/kernel"] = (c, int(c * 0.25))
        s[n + "/conv/2/fc/2/kernel"] = (int(c * 0.25), c)
        s[n + "/conv/3/kernel"] = (1, 1, c, c)

    def reduce(n, c, keep):
        lnorm(n + "/norm1", c)
        mb(n, c)
        s[n + "/reduction/kernel"] = (3, 3, c, c if keep else 2 * c)
        lnorm(n + "/norm2", c if keep else 2 * c)

    d0 = cfg["dim"]
    s["patch_embed/proj/kernel"] = (3, 3, 3, d0)
    s["patch_embed/proj/bias"] = (d0,)
    reduce("patch_embed/conv_down", d0, True)
    c = d0
    for i, depth in enumerate(cfg["depths"]):
        ws, heads = cfg["window_size"][i], cfg["num_heads"][i]
        for k in range(len(KEEP_DIMS[i])):
            mb(f"levels/{i}/q_global_gen/to_q_global/{k}", c)
        hidden = int(c * cfg["mlp_ratio"])
        for j in range(depth):
            n = f"levels/{i}/blocks/{j}"
            lnorm(n + "/norm1", c)
            nq = 2 if j % 2 else 3
            s[n + "/attn/qkv/kernel"] = (c, nq * c)
            s[n + "/attn/qkv/bias"] = (nq * c,)
            s[n + "/attn/relative_position_bias_table"] = ((2 * ws - 1) ** 2, heads)
            s[n + "/attn/proj/kernel"] = (c, c)
            s[n + "/attn/proj/bias"] = (c,)
            lnorm(n + "/norm2", c)
            s[n + "/mlp/fc1/kernel"] = (c, hidden)
            s[n + "/mlp/fc1/bias"] = (hidden,)
            s[n + "/mlp/fc2/kernel"] = (hidden, c)
            s[n + "/mlp/fc2/bias"] = (c,)
            if cfg["layer_scale"] is not None:
                s[n + "/gamma1"] = (c,)
                s[n + "/gamma2"] = (c,)
        # every level constructs a downsample ReduceSize (level.py:40); the last level never calls it
        reduce(f"levels/{i}/downsample", c, False)
        if i < 3:
            c *= 2
    lnorm("norm", c)
    s["head/kernel"] = (c, num_classes)
    s["head/bias"] = (num_classes,)
    return s


def random_weights(variant="tiny", num_classes=2, seed=0) -> dict:
    """Seeded, non-degenerate weights: glorot-like kernels, LayerNorm gamma ~ U(0.6,1.4), small biases, and (small/base)
    layer-scale gammas drawn around 0.5 instead of the 1e-5 initialiser so that the blocks contribute."""
    rng = np.random.default_rng(seed)
    W = {}
    for name, shp in weight_shapes(variant, num_classes).items():
        leaf = name.rsplit("/", 1)[1]
        if name == "head/kernel":  # Keras default glorot-uniform
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            W[name] = rng.uniform(-lim, lim, shp).astype(np.float32)
        elif leaf in ("kernel", "depthwise_kernel"):
            fan_in = int(np.prod(shp[:-1])) if leaf == "kernel" else 9
            W[name] = (rng.standard_normal(shp) * np.sqrt(1.5 / fan_in)).astype(np.float32)
        elif leaf == "gamma":
            W[name] = rng.uniform(0.6, 1.4, shp).astype(np.float32)
        elif leaf in ("gamma1", "gamma2"):
            W[name] = rng.uniform(0.3, 0.7, shp).astype(np.float32)
        elif leaf in ("beta", "bias"):
            W[name] = (rng.standard_normal(shp) * 0.1).astype(np.float32)
        elif leaf == "relative_position_bias_table":
            W[name] = (rng.standard_normal(shp) * 0.5).astype(np.float32)
        else:
            raise KeyError(name)
    # well-conditioned residual branches (trained nets are): damp the branch outputs by ~1/sqrt(#blocks)
    damp = 1.5 / np.sqrt(sum(CONFIGS[variant]["depths"]))
    for name in W:
        if name.endswith("attn/proj/kernel") or name.endswith("mlp/fc2/kernel"):
            W[name] *= damp
    return W


def param_count(W, include_head=True, include_unused=True):
    tot = 0
    for k, v in W.items():
        if not include_head and k.startswith("head/"):
            continue
        if not include_unused and k.startswith("levels/3/downsample"):
            continue
        tot += v.size
    return int(tot)
