"""GCViT forward on the B200 kernels versus the fp32 PyTorch-CPU oracle (same random-init weights).
Tolerances: see tests/test_resnet_rs_gpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant,head,seed,hw", [
    ("tiny", "softmax", 3, 224), ("small", "sigmoid", 5, 224), ("small", "softmax", 3, 224), ("xxtiny", "softmax", 5, 224),
    # 200x200 input: 50 / 25 / 13-pixel maps are not window multiples -> FitWindow padding (feature.py:234-256) to
    # 56 / 28 / 14, attention over the zero-padded tokens, top-left crop after the blocks (level.py:49,61)
    ("xxtiny", "softmax", 5, 200), ("tiny", "softmax", 7, 200),
])
def test_gcvit_matches_oracle(cuda_device, variant, head, seed, hw):
    import torch

    from oracle import gcvit as G
    from oracle import preprocess as P
    from test_resnet_rs_gpu import check_against_oracle
    from vipcup_b200.models import GCViT

    k = 2 if head == "softmax" else 1
    W = G.random_weights(variant, k, seed=seed)
    x = np.stack([P.decode_to_float(P.synth_image(i), hw, hw) for i in range(8)])
    ref_taps = {}
    ref = G.forward(x, W, variant, head_act=head, taps=ref_taps)
    model = GCViT(variant, input_shape=(hw, hw, 3), num_classes=k, head_act=head, device=cuda_device).load_weights(W)
    taps = {}
    got = model(torch.from_numpy(x).to(cuda_device), taps=taps)
    torch.cuda.synchronize()
    check_against_oracle(ref, ref_taps, got, taps, W["head/kernel"], W["head/bias"],
                         ("stem", "level0", "level1", "level2", "level3"), logit_tol=1e-2)
