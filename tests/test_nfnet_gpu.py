"""ECA-NFNet-L0 (kecam, a ckpts.json member of the reference) on the B200 kernels versus the fp32 PyTorch-CPU oracle
(oracle/nfnet.py) on the same seeded weights, plus the grouped-convolution slices and the ECA gate against torch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("groups,c,h,stride", [(2, 128, 25, 1), (6, 384, 13, 2), (2, 128, 50, 2), (1, 64, 9, 1)])
def test_grouped_conv_matches_torch(cuda_device, groups, c, h, stride):
    import torch
    import torch.nn.functional as F

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(groups * 10 + c)
    n, cg = 3, c // groups
    x = torch.randn((n, h, h, c), generator=g).to(torch.bfloat16).to(cuda_device)
    wk = (torch.randn((3, 3, cg, c), generator=g) / (3 * cg ** 0.5)).to(torch.bfloat16)          # Keras (kh,kw,Cin/g,Cout)
    bias = torch.randn((c,), generator=g).to(cuda_device)
    ws = [wk[:, :, :, i * cg:(i + 1) * cg].reshape(-1, cg).t().contiguous().to(cuda_device) for i in range(groups)]
    bs = [bias[i * cg:(i + 1) * cg].contiguous() for i in range(groups)]
    y = nn.conv2d_grouped(x, ws, bs, ksize=3, stride=stride, pad=1, act="swish") if groups > 1 else \
        nn.conv2d(x, ws[0], bs[0], ksize=3, stride=stride, pad=1, act="swish")
    ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), wk.float().permute(3, 2, 0, 1).to(cuda_device), bias, stride=stride,
                          padding=1, groups=groups)).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    err = (y.float() - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -7 + 5e-3).all(), err.max().item()


def test_eca_gate_and_act_scale(cuda_device):
    import torch
    import torch.nn.functional as F

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(3)
    n, c, hw = 4, 512, 625
    sums = (torch.randn((n, c), generator=g) * 100).to(cuda_device)
    gap = torch.round(sums.double() * 2 ** 28).to(torch.int64)
    w = torch.randn((5,), generator=g).to(cuda_device)
    gate = nn.eca_gate(gap, w, hw, out_scale=0.4)
    ref = 0.4 * torch.sigmoid(F.conv1d(F.pad(sums / hw, (2, 2))[:, None, :], w.view(1, 1, 5))[:, 0, :])
    assert torch.allclose(gate, ref, rtol=1e-4, atol=1e-5)
    x = torch.randn((3, 7, 7, 64), generator=g).to(torch.bfloat16).to(cuda_device)
    y = nn.act_scale(x, "swish", 0.9)
    assert torch.allclose(y.float(), (F.silu(x.float()) * 0.9), rtol=2 ** -7, atol=1e-3)


@pytest.mark.parametrize("hw,head,seed", [(200, "softmax", 1), (200, "sigmoid", 2), (224, "softmax", 3)])
def test_eca_nfnet_l0_matches_oracle(cuda_device, hw, head, seed):
    import torch

    from oracle import nfnet as N
    from oracle import preprocess as P
    from test_resnet_rs_gpu import check_against_oracle
    from vipcup_b200 import registry

    k = 2 if head == "softmax" else 1
    W = N.random_weights(k, seed=seed)
    x = np.stack([P.decode_to_float(P.synth_image(i), hw, hw) for i in range(6)])
    ref_taps = {}
    ref = N.forward(x, W, head_act=head, taps=ref_taps)
    model = registry.create_model(f"ECA_NFNetL0-{hw}x{hw}", (hw, hw), num_classes=k, head_act=head, device=cuda_device)
    model.load_weights(W)
    taps = {}
    got = model(torch.from_numpy(x).to(cuda_device), taps=taps)
    torch.cuda.synchronize()
    check_against_oracle(ref, ref_taps, got, taps, W["predictions/kernel"], W["predictions/bias"],
                         ("stem", "stack1", "stack2", "stack3", "stack4"), logit_tol=1e-2)
