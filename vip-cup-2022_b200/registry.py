"""Model registry of the reference: ``ckpts/ckpts.json`` entries ``[base_dir, [H, W], idx]`` with
``base_dir = "<Arch>-<H>x<W>"`` (main.py:43-56, 186-197), mapped onto the B200 model classes.

Checkpoint format: the reference loads full Keras models from ``ckpts/<base_dir>/ckpt/*.h5`` (each file = one fold,
main.py:187-192).  Both formats are read here: Keras HDF5 files through the pure-Python reader ``h5lite`` (no h5py /
libhdf5 offline; the architecture comes from the directory name, only the weights are taken from the file) and ``*.npz``
files holding the same Keras-named, Keras-layout weight arrays plus optional ``__num_classes__`` / ``__head_act__`` entries
(``tests/tools/make_random_ckpts.py`` writes those for the synthetic runs)."""
from __future__ import annotations

import os
from glob import glob

import numpy as np

NAME2BS = {  # main.py:43-56
    "convnext_large_384_in22ft1k-200x200": 16, "convnext_large_in22ft1k-200x200": 16,
    "convnext_base_384_in22ft1k-200x200": 32, "HorNetBase-200x200": 32, "EfficientNetV2M-200x200": 64,
    "convnext_base_in22k-200x200": 32, "ECA_NFNetL2-200x200": 32, "GCViTBase-224x224": 48, "ResNest200-200x200": 64,
    "EfficientNetV2L-200x200": 32, "ResNetRS200-200x200": 32, "ResNet200D-200x200": 32,
}


def arch_of(model_name: str) -> str:
    return model_name.rsplit("-", 1)[0]


def constructors():
    from .models import convnext, efficientnet, gcvit, nfnet, resnest, resnet_rs

    table = {f"ResNetRS{d}": (lambda d=d, **kw: resnet_rs.ResNetRS(d, **kw)) for d in resnet_rs.BLOCK_ARGS}
    for v in gcvit.CONFIGS:
        pretty = {"xxtiny": "XXTiny", "xtiny": "XTiny", "tiny": "Tiny", "small": "Small", "base": "Base"}[v]
        table[f"GCViT{pretty}"] = (lambda v=v, **kw: gcvit.GCViT(v, **kw))
    # tfimm registry names (models/tfimm/architectures/convnext.py:440-620): convnext_<size>[_384][_in22k | _in22ft1k]
    for v in convnext.CONFIGS:
        for suffix in ("", "_in22k", "_in22ft1k", "_384_in22ft1k"):
            table[f"convnext_{v}{suffix}"] = (lambda v=v, **kw: convnext.ConvNeXt(v, **kw))
    # keras_cv_attention_models constructors (efficientnet_v2.py:268-275, efficientnet_v1.py:68-73)
    table["EfficientNetV2T"] = lambda **kw: efficientnet.EfficientNet("v2t", **kw)
    table["EfficientNetV1B4"] = lambda **kw: efficientnet.EfficientNet("v1b4", **kw)
    table["ECA_NFNetL0"] = lambda **kw: nfnet.ECANFNetL0(**kw)                        # nfnets/nfnets.py:316-320
    table["ResNest50"] = lambda **kw: resnest.ResNeSt50(**kw)                          # resnest/resnest.py:76-77
    return table


def supported_archs():
    return sorted(constructors())


def create_model(model_name, dim, num_classes=2, head_act="softmax", device="cuda"):
    arch = arch_of(model_name)
    table = constructors()
    if arch not in table:
        raise ValueError(f"no B200 implementation for architecture {arch!r} (model {model_name}); built: "
                         f"{supported_archs()} -- the other ckpts.json backbones are listed under 'next' in DESIGN.md")
    if arch.startswith("ResNetRS"):
        return table[arch](input_shape=(dim[0], dim[1], 3), classes=num_classes, classifier_activation=head_act,
                           device=device)
    if arch.startswith("EfficientNet") or arch.startswith("ECA_NFNet") or arch.startswith("ResNest"):
        return table[arch](input_shape=(dim[0], dim[1], 3), num_classes=num_classes, classifier_activation=head_act,
                           device=device)
    return table[arch](input_shape=(dim[0], dim[1], 3), num_classes=num_classes, head_act=head_act, device=device)


def scan_checkpoints(model_dir, ckpt_cfg_path):
    """main.py:186-197: returns [[paths], dim, idx] per registry entry; raises ValueError('no model found for :', ...)."""
    import json

    out = []
    for base_dir, dim, idx in json.load(open(ckpt_cfg_path, "r")):
        paths = sorted(glob(os.path.join(model_dir, base_dir, "ckpt", "*h5")))        # main.py:187
        if not paths:
            paths = sorted(glob(os.path.join(model_dir, base_dir, "ckpt", "*.npz")))
        if paths:
            out.append([paths, dim, idx])
        else:
            raise ValueError("no model found for :", base_dir)
    return out


def load_checkpoint(path):
    """One fold -> ({Keras weight name: ndarray}, meta).  ``.h5`` / ``.hdf5``: Keras HDF5 weights (model.save or
    save_weights), names with the ':0' suffix dropped; ``.npz``: the same dictionary saved with numpy."""
    if path.endswith((".h5", ".hdf5")):
        from . import h5lite

        return h5lite.load_keras_weights(path), {"num_classes": None, "head_act": None}
    z = np.load(path, allow_pickle=False)
    W = {k: z[k] for k in z.files if not k.startswith("__")}
    meta = {"num_classes": int(z["__num_classes__"]) if "__num_classes__" in z.files else None,
            "head_act": str(z["__head_act__"]) if "__head_act__" in z.files else None}
    return W, meta


def resolve_weight_names(W, expected):
    """Maps checkpoint names onto the names a model asks for.  Keras prefixes weight names with the name scopes of the
    enclosing (sub-)models (``<model>/<layer>/kernel``), which depend on how the checkpoint's model object was built; a
    wanted name is matched exactly or as the unique key that ends with ``/<name>``.  Raises KeyError listing what is
    missing or ambiguous."""
    if all(k in W for k in expected):
        return W
    if expected:   # one common scope prefix for every weight (the usual case: "<model name>/")
        for k in W:
            if k.endswith("/" + expected[0]):
                prefix = k[: -len(expected[0])]
                if all(prefix + name in W for name in expected):
                    return {**W, **{name: W[prefix + name] for name in expected}}
    out, bad = dict(W), []
    for name in expected:
        if name in W:
            continue
        hits = [k for k in W if k.endswith("/" + name)]
        if len(hits) == 1:
            out[name] = W[hits[0]]
        else:
            bad.append(f"{name} ({'missing' if not hits else 'ambiguous: ' + ', '.join(hits[:3])})")
    if bad:
        raise KeyError("checkpoint does not provide: " + "; ".join(bad[:8]) + (f" ... (+{len(bad) - 8})" if len(bad) > 8 else ""))
    return out
