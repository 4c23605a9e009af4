"""ResNeSt-50 (kecam, a ckpts.json member of the reference) on the B200 kernels versus the fp32 PyTorch-CPU oracle
(oracle/resnest.py) on the same seeded weights, plus its two small kernels against torch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_split_attention_and_avgpool3(cuda_device):
    import torch
    import torch.nn.functional as F

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(4)
    n, h, f = 3, 13, 128
    x = torch.randn((n, h, h, 2 * f), generator=g).to(torch.bfloat16).to(cuda_device)
    lg = torch.randn((n, 2 * f), generator=g).to(cuda_device)
    out = nn.split_attention2(x, lg)
    a = torch.softmax(lg.view(n, 2, f), dim=1)
    ref = a[:, 0, None, None, :] * x[..., :f].float() + a[:, 1, None, None, :] * x[..., f:].float()
    assert torch.allclose(out.float(), ref, rtol=2 ** -7, atol=2e-3)
    p = nn.avgpool3s2(out)
    pref = F.avg_pool2d(F.pad(out.float().permute(0, 3, 1, 2), (1, 1, 1, 1)), 3, 2).permute(0, 2, 3, 1)
    assert p.shape == pref.shape and torch.allclose(p.float(), pref, rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("hw,head,seed", [(200, "softmax", 1), (200, "sigmoid", 2), (224, "softmax", 5)])
def test_resnest50_matches_oracle(cuda_device, hw, head, seed):
    import torch

    from oracle import preprocess as P
    from oracle import resnest as R
    from test_resnet_rs_gpu import check_against_oracle
    from vipcup_b200 import registry

    k = 2 if head == "softmax" else 1
    W = R.random_weights(k, seed=seed)
    x = np.stack([P.decode_to_float(P.synth_image(i), hw, hw) for i in range(6)])
    ref_taps = {}
    ref = R.forward(x, W, head_act=head, taps=ref_taps)
    model = registry.create_model(f"ResNest50-{hw}x{hw}", (hw, hw), num_classes=k, head_act=head, device=cuda_device)
    model.load_weights(W)
    taps = {}
    got = model(torch.from_numpy(x).to(cuda_device), taps=taps)
    torch.cuda.synchronize()
    check_against_oracle(ref, ref_taps, got, taps, W["predictions/kernel"], W["predictions/bias"],
                         ("stem", "stack1", "stack2", "stack3", "stack4"), logit_tol=1e-2)
