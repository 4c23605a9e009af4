"""Implicit-GEMM convolution (im2col-mode TMA feeding tcgen05, nothing materialised) against torch.nn.functional.conv2d
in fp32 on the same bf16-rounded operands.  Shapes follow SURVEY.md Appendix B (ResNet-RS stem / c2..c5 3x3 convs with
the odd 25 -> 13 -> 7 maps, GCViT reductions incl. the 96-channel small variant)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,h,c,cout,stride", [
    (2, 50, 64, 64, 1), (3, 50, 128, 128, 2), (2, 25, 256, 256, 2), (5, 13, 512, 512, 2), (4, 7, 512, 512, 1),
    (2, 100, 32, 64, 1), (1, 100, 64, 64, 2), (2, 56, 96, 192, 2), (3, 28, 192, 384, 2), (9, 14, 256, 512, 2),
    (2, 112, 64, 64, 2),
])
def test_conv3x3_matches_torch(cuda_device, n, h, c, cout, stride):
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(n * 1000 + h * 10 + c + stride)
    x = (torch.randn(n, h, h, c, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    k = (torch.randn(3, 3, c, cout, generator=g) * (1.0 / (9 * c)) ** 0.5).to(torch.bfloat16)   # Keras (kh,kw,Cin,Cout)
    w = k.reshape(-1, cout).t().contiguous().to(cuda_device)                                       # [Cout, (r,s,c)]
    bias = torch.randn(cout, generator=g).to(cuda_device)
    out = nn.conv2d(x, w, bias, ksize=3, stride=stride, pad=1, act="relu")
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), k.float().permute(3, 2, 0, 1).to(cuda_device), bias,
                                     stride=stride, padding=1)
    ref = torch.relu(ref).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def test_conv3x3_patch_tiles_in_child_process(cuda_device):
    """The tiled-4-D ("patch" tile) feed of the implicit GEMM is selected by VIP_CONV_PATCH=1, read once per process:
    run three shapes of the main test in a child process with it set."""
    import os
    import subprocess
    import sys

    if os.environ.get("VIP_CONV_PATCH") == "1":
        pytest.skip("already in the child")
    env = dict(os.environ, VIP_CONV_PATCH="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", __file__, "-k", "matches_torch and (2-50-64 or 2-25-256 or 2-56-96)"],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
