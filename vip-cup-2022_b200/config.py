"""Attribute-bag configuration, same surface as the reference's ``utils/config.py:4-32``."""
from __future__ import annotations


class Config:
    """``Config({...})`` -> attributes (utils/config.py:4-6)."""

    def __init__(self, data=None):
        self.__dict__.update(**(data or {}))


def dict2cfg(cfg_dict):
    cfg = Config(cfg_dict)
    if hasattr(cfg, "class_labels") and hasattr(cfg, "class_names"):
        cfg.label2name = dict(zip(cfg.class_labels, cfg.class_names))
    return cfg


def cfg2dict(cfg):
    return {k: v for k, v in dict(vars(cfg)).items() if "__" not in k}
