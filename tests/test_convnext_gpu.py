"""ConvNeXt (tfimm ``convnext_tiny_in22k``, a ckpts.json member of the reference) on the B200 kernels versus the fp32
PyTorch-CPU oracle (oracle/convnext.py) on the same seeded weights, plus the depthwise K x K kernel against torch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,stride,c,h,w,pad,act", [
    (7, 1, 96, 33, 29, None, None), (7, 1, 384, 24, 24, None, "gelu"), (5, 1, 64, 20, 21, None, "swish"),
    (5, 2, 56, 28, 28, (1, 1, 2, 2), "swish"),      # TF 'SAME' at stride 2: asymmetric padding (efficientnet_v2.py:80-85)
    (3, 2, 48, 25, 25, (1, 1, 1, 1), "relu"), (3, 1, 128, 7, 7, None, None), (3, 2, 32, 112, 112, (0, 0, 1, 1), "swish"),
])
def test_dwconv_kxk_matches_torch(cuda_device, k, stride, c, h, w, pad, act):
    import torch
    import torch.nn.functional as F

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(k * 100 + c)
    n = 3
    x = torch.randn((n, h, w, c), generator=g).to(torch.bfloat16).to(cuda_device)
    wt = (torch.randn((k, k, c), generator=g) / k).to(cuda_device)
    bias = torch.randn((c,), generator=g).to(cuda_device)
    gap = nn.zero_(torch.empty((n, c), dtype=torch.int64, device=cuda_device))
    y = nn.dwconv(x, wt, bias, ksize=k, stride=stride, pad=pad, act=act, gap=gap)
    pt, pl, pb, pr = pad if pad is not None else (k // 2,) * 4
    xin = F.pad(x.float().permute(0, 3, 1, 2), (pl, pr, pt, pb))
    ref = F.conv2d(xin, wt.permute(2, 0, 1)[:, None], bias, stride=stride, groups=c)
    ref = {None: lambda t: t, "gelu": F.gelu, "swish": F.silu, "relu": F.relu}[act](ref).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    err = (y.float() - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -8 + 2e-3).all(), err.max().item()
    assert torch.allclose(gap.double() * 2.0 ** -28, y.double().sum((1, 2)), rtol=1e-6, atol=1e-4)


def test_layernorm_f32(cuda_device):
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(1)
    x = (torch.randn((37, 768), generator=g) * 3 + 1).to(cuda_device)
    gamma, beta = (torch.rand(768, generator=g) + 0.5).to(cuda_device), torch.randn(768, generator=g).to(cuda_device)
    got = nn.layernorm_f32(x, gamma, beta, eps=1e-6)
    ref = torch.nn.functional.layer_norm(x, (768,), gamma, beta, 1e-6)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("seed,hw,head", [(1, 200, "softmax"), (2, 200, "sigmoid"), (3, 224, "softmax")])
def test_convnext_tiny_matches_oracle(cuda_device, seed, hw, head):
    import torch

    from oracle import convnext as C
    from oracle import preprocess as P
    from test_resnet_rs_gpu import check_against_oracle
    from vipcup_b200 import registry

    k = 2 if head == "softmax" else 1
    W = C.random_weights("tiny", k, seed=seed)
    x = np.stack([P.decode_to_float(P.synth_image(i), hw, hw) for i in range(6)])
    ref_taps = {}
    ref = C.forward(x, W, "tiny", head_act=head, taps=ref_taps)
    model = registry.create_model(f"convnext_tiny_in22k-{hw}x{hw}", (hw, hw), num_classes=k, head_act=head, device=cuda_device)
    model.load_weights(W)
    taps = {}
    got = model(torch.from_numpy(x).to(cuda_device), taps=taps)
    torch.cuda.synchronize()
    check_against_oracle(ref, ref_taps, got, taps, W["head/fc/kernel"], W["head/fc/bias"],
                         ("stem", "stage0", "stage1", "stage2", "stage3"), logit_tol=1e-2)
