"""CPU known-answer tests of the model oracles (parameter counts and stage shapes of SURVEY.md 8c / Appendix B)."""
import numpy as np


def test_resnet_rs_param_counts():
    from oracle import resnet_rs as R

    assert R.param_count(R.random_weights(50, 2), include_head=False) == 33_696_288   # 33.70 M
    assert abs(R.param_count(R.random_weights(101, 2), include_head=False) - 61.7e6) < 0.05e6


def test_resnet_rs_stage_shapes():
    from oracle import resnet_rs as R

    x = np.random.default_rng(0).random((1, 200, 200, 3), dtype=np.float32)
    taps = {}
    p = R.forward(x, R.random_weights(50, 2), 50, taps=taps)
    assert taps["stem"].shape == (1, 50, 50, 64)
    assert taps["c2"].shape == (1, 50, 50, 256) and taps["c3"].shape == (1, 25, 25, 512)
    assert taps["c4"].shape == (1, 13, 13, 1024) and taps["c5"].shape == (1, 7, 7, 2048)
    assert p.shape == (1, 2) and abs(p.sum() - 1) < 1e-5


def test_gcvit_param_counts_match_vendored_doc_table():
    """models/keras_cv_attention_models/gcvit/__init__.py:37-43 (1000-class head): 12.0 / 28.2 / 51.1 M."""
    from oracle import gcvit as G

    for variant, millions in (("xxtiny", 12.0), ("tiny", 28.2), ("small", 51.1)):
        n = G.param_count(G.random_weights(variant, 1000), include_unused=False)
        assert abs(n / 1e6 - millions) < 0.06, (variant, n)


def test_gcvit_stage_shapes():
    from oracle import gcvit as G

    x = np.random.default_rng(0).random((1, 224, 224, 3), dtype=np.float32)
    taps = {}
    p = G.forward(x, G.random_weights("xxtiny", 2), "xxtiny", taps=taps)
    assert taps["stem"].shape == (1, 56, 56, 64) and taps["level0"].shape == (1, 28, 28, 128)
    assert taps["level1"].shape == (1, 14, 14, 256) and taps["level2"].shape == (1, 7, 7, 512)
    assert taps["level3"].shape == (1, 7, 7, 512) and p.shape == (1, 2)


def test_convnext_param_count_and_shapes():
    """tfimm / timm model card: convnext_tiny 28.59 M parameters with the 1000-class head; the author's stride-2 4x4 stem
    (models/tfimm/architectures/convnext.py:320-327) turns 200x200 inputs into 99 / 49 / 24 / 12-pixel stages."""
    from oracle import convnext as C

    W = C.random_weights("tiny", 1000)
    assert C.param_count(W) == 28_589_128
    x = np.random.default_rng(0).random((1, 200, 200, 3), dtype=np.float32)
    taps = {}
    p = C.forward(x, C.random_weights("tiny", 2), "tiny", taps=taps)
    assert taps["stem"].shape == (1, 99, 99, 96) and taps["stage1"].shape == (1, 49, 49, 192)
    assert taps["stage2"].shape == (1, 24, 24, 384) and taps["stage3"].shape == (1, 12, 12, 768) and p.shape == (1, 2)


def test_efficientnet_param_counts_and_shapes():
    """kecam model table (efficientnet/__init__.py:80,156): EfficientNetV2T 13.6 M, EfficientNetV1B4 19.3 M trainable
    parameters (the official EfficientNet-B4 count is 19 341 616); stage shapes of SURVEY.md B.3."""
    from oracle import efficientnet as E

    for variant, trainable in (("v2t", 13_649_388), ("v1b4", 19_341_616)):
        W = E.random_weights(variant, 1000)
        assert sum(a.size for k, a in W.items() if "moving_" not in k) == trainable
    x = np.random.default_rng(0).random((1, 200, 200, 3), dtype=np.float32)
    taps = {}
    E.forward(x, E.random_weights("v2t", 2), "v2t", taps=taps)
    assert taps["stem"].shape == (1, 100, 100, 24) and taps["stack2"].shape == (1, 25, 25, 48)
    assert taps["stack4"].shape == (1, 13, 13, 128) and taps["stack5"].shape == (1, 7, 7, 208)
    x = np.random.default_rng(0).random((1, 224, 224, 3), dtype=np.float32)
    taps = {}
    E.forward(x, E.random_weights("v1b4", 2), "v1b4", taps=taps)
    assert taps["stem"].shape == (1, 112, 112, 48) and taps["stack2"].shape == (1, 28, 28, 56)
    assert taps["stack6"].shape == (1, 7, 7, 448) and taps["feat"].shape == (1, 1792)


def test_eca_nfnet_l0_param_count_and_shapes():
    """kecam model table (nfnets/__init__.py:82): ECA_NFNetL0 24.14 M parameters; stage shapes of SURVEY.md B.3."""
    from oracle import nfnet as N

    assert N.param_count(N.random_weights(1000)) == 24_143_872 or abs(N.param_count(N.random_weights(1000)) / 1e6 - 24.14) < 0.01
    x = np.random.default_rng(0).random((1, 200, 200, 3), dtype=np.float32)
    taps = {}
    p = N.forward(x, N.random_weights(2), taps=taps)
    assert taps["stem"].shape == (1, 50, 50, 128) and taps["stack1"].shape == (1, 50, 50, 256)
    assert taps["stack2"].shape == (1, 25, 25, 512) and taps["stack3"].shape == (1, 13, 13, 1536)
    assert taps["stack4"].shape == (1, 7, 7, 1536) and taps["feat"].shape == (1, 2304) and p.shape == (1, 2)


def test_resnest50_param_count_and_shapes():
    """ResNeSt-50: 27.48 M trainable parameters (the published count of the architecture); stage shapes of SURVEY.md B.3."""
    from oracle import resnest as R

    assert abs(R.param_count(R.random_weights(1000), trainable_only=True) / 1e6 - 27.48) < 0.01
    x = np.random.default_rng(0).random((1, 200, 200, 3), dtype=np.float32)
    taps = {}
    p = R.forward(x, R.random_weights(2), taps=taps)
    assert taps["stem"].shape == (1, 50, 50, 64) and taps["stack1"].shape == (1, 50, 50, 256)
    assert taps["stack2"].shape == (1, 25, 25, 512) and taps["stack3"].shape == (1, 13, 13, 1024)
    assert taps["stack4"].shape == (1, 7, 7, 2048) and p.shape == (1, 2)
