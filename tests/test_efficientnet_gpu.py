"""EfficientNetV2-T / EfficientNetV1-B4 (kecam builders, ckpts.json members of the reference) on the B200 kernels versus the
fp32 PyTorch-CPU oracle (oracle/efficientnet.py) on the same seeded weights."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch,variant,hw,head,seed", [
    ("EfficientNetV2T", "v2t", 200, "softmax", 1), ("EfficientNetV2T", "v2t", 200, "sigmoid", 2),
    ("EfficientNetV1B4", "v1b4", 224, "softmax", 1), ("EfficientNetV1B4", "v1b4", 200, "softmax", 3),
])
def test_efficientnet_matches_oracle(cuda_device, arch, variant, hw, head, seed):
    import torch

    from oracle import efficientnet as E
    from oracle import preprocess as P
    from test_resnet_rs_gpu import check_against_oracle
    from vipcup_b200 import registry

    k = 2 if head == "softmax" else 1
    W = E.random_weights(variant, k, seed=seed)
    x = np.stack([P.decode_to_float(P.synth_image(i), hw, hw) for i in range(6)])
    ref_taps = {}
    ref = E.forward(x, W, variant, head_act=head, taps=ref_taps)
    model = registry.create_model(f"{arch}-{hw}x{hw}", (hw, hw), num_classes=k, head_act=head, device=cuda_device)
    model.load_weights(W)
    taps = {}
    got = model(torch.from_numpy(x).to(cuda_device), taps=taps)
    torch.cuda.synchronize()
    stages = [s for s in ref_taps if s != "feat"]
    check_against_oracle(ref, ref_taps, got, taps, W["predictions/kernel"], W["predictions/bias"], stages, logit_tol=1e-2)
