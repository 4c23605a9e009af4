"""tcgen05 GEMM core versus a plain PyTorch fp32 reference of the same contraction (bf16 inputs, fp32 accumulate)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k", [
    (128, 32, 64), (128, 128, 64), (256, 256, 256), (1000, 64, 288), (625 * 3, 512, 1152), (49 * 5, 2048, 512),
    (130, 192 + 64, 96), (4096, 768, 256), (77, 32, 8), (20000, 64, 576),
])
@pytest.mark.parametrize("epi", ["plain", "bias_relu_res", "gelu_f32", "colscale_res"])
def test_gemm_matches_torch(cuda_device, m, n, k, epi):
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(m * 31 + n * 7 + k)
    a = (torch.randn(m, k, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    b = (torch.randn(n, k, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    bias = torch.randn(n, generator=g).to(cuda_device)
    res = torch.randn(m, n, generator=g).to(torch.bfloat16).to(cuda_device)
    ref = a.float() @ b.float().t()
    if epi == "plain":
        out = nn.gemm(a, b)
    elif epi == "bias_relu_res":
        out = nn.gemm(a, b, bias=bias, act="relu", residual=res)
        ref = torch.relu(ref + bias) + res.float()
    elif epi == "colscale_res":
        cs = torch.rand(n, generator=g).to(cuda_device)
        out = nn.gemm(a, b, bias=bias, colscale=cs, residual=res)
        ref = (ref + bias) * cs + res.float()
    else:
        out = nn.gemm(a, b, bias=bias, act="gelu", out_dtype=torch.float32)
        ref = torch.nn.functional.gelu(ref + bias)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    tol = 2e-2 * max(1.0, ref.abs().max().item()) if out.dtype == torch.bfloat16 else 1e-3 * max(1.0, k ** 0.5)
    assert err <= tol, f"max abs err {err} > {tol}"
