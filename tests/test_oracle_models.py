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
