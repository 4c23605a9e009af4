"""vip_window_attention_bf16 (tensor-core window attention, models/gcvit/layers/attention.py:52-83 with
window_partition/reverse of layers/window.py:3-14 folded in) against a plain PyTorch fp32 restatement of the same op on
the same bf16-rounded inputs.  Tolerance: bf16 output rounding (half an ulp = 2^-9 relative) + bf16 P operand ->
8e-3 x max(1, max |reference|)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel_index(ws):  # attention.py:39-50
    coords = np.stack(np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij")).reshape(2, -1)
    rel = coords[:, :, None] - coords[:, None, :]
    return (rel[0] + ws - 1) * (2 * ws - 1) + (rel[1] + ws - 1)


def _ref(qkv, qg, table, B, H, W, C, ws, heads):
    import torch

    N, hd = ws * ws, 32
    parts = qkv.shape[-1] // C
    x = qkv.float().view(B, H // ws, ws, W // ws, ws, parts * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, N, parts, heads, hd)
    x = x.permute(2, 0, 3, 1, 4)                       # [parts, B_, heads, N, hd]
    if qg is None:
        q, k, v = x[0], x[1], x[2]
    else:
        k, v = x[0], x[1]
        nw = k.shape[0] // B
        q = qg.float().view(B, 1, N, heads, hd).expand(B, nw, N, heads, hd).reshape(-1, N, heads, hd).permute(0, 2, 1, 3)
    attn = (q * hd ** -0.5) @ k.transpose(-1, -2)
    bias = table.float()[:, torch.from_numpy(_rel_index(ws).reshape(-1)).to(table.device)].view(heads, N, N)
    attn = torch.softmax(attn + bias[None], -1)
    o = (attn @ v).permute(0, 2, 1, 3).reshape(-1, N, C)          # [B_, N, C]
    o = o.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B * H * W, C)
    return o


@pytest.mark.parametrize("impl", ["ws", "mma"])
@pytest.mark.parametrize("B,H,ws,heads,glob", [(2, 14, 7, 2, False), (3, 21, 7, 4, True), (2, 14, 14, 8, False),
                                               (1, 14, 14, 3, True), (5, 7, 7, 16, False), (1, 56, 7, 2, True), (3, 14, 7, 3, False),
                                               (2, 14, 14, 3, True), (1, 7, 7, 1, False)])
def test_window_attention_matches_torch(cuda_device, B, H, ws, heads, glob, impl):
    """Both implementations (attention_ws.cu persistent tcgen05 -- the default --, attention.cu mma.sync); the library reads
    VIP_ATTN_IMPL once per process, so the non-default kernel is exercised in a child process.  Odd head counts and odd
    window counts cover the half-empty head pair and the half-empty two-window tile of the persistent kernel."""
    import os

    if impl != os.environ.get("VIP_ATTN_IMPL", "ws"):
        import subprocess
        import sys

        env = dict(os.environ, VIP_ATTN_IMPL=impl)
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", __file__, "-k",
                            f"{B}-{H}-{ws}-{heads}-{glob}-{impl}"], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-3000:]
        return
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(B * 100 + H + heads)
    C, W, N = heads * 32, H, ws * ws
    parts = 2 if glob else 3
    qkv = (torch.randn((B * H * W, parts * C), generator=g) * 1.5).to(torch.bfloat16).to(cuda_device)
    qg = (torch.randn((B, N, C), generator=g) * 1.5).to(torch.bfloat16).to(cuda_device) if glob else None
    table = (torch.randn((heads, (2 * ws - 1) ** 2), generator=g) * 0.5).to(cuda_device)
    got = nn.window_attention(qkv, qg, table, B, H, W, C, ws, heads).float()
    ref = _ref(qkv, qg, table, B, H, W, C, ws, heads)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    assert torch.isfinite(got).all()
    tol = 8e-3 * max(1.0, ref.abs().max().item())
    assert err < tol, f"max abs err {err} > {tol}"


def test_mma_fallback_through_its_real_trigger(cuda_device):
    """ws 14 with 64 heads: 64 x 729 relative-position entries do not fit next to the operand ring in shared memory, the
    tcgen05 kernel declines and the call is served by the mma.sync kernel -- same numbers, and the reason is left in
    vip_last_error()."""
    import torch

    from vipcup_b200 import _lib, nn

    B, H, ws, heads = 1, 14, 14, 64
    g = torch.Generator(device="cpu").manual_seed(7)
    C, N = heads * 32, ws * ws
    qkv = (torch.randn((B * H * H, 3 * C), generator=g) * 1.5).to(torch.bfloat16).to(cuda_device)
    table = (torch.randn((heads, (2 * ws - 1) ** 2), generator=g) * 0.5).to(cuda_device)
    got = nn.window_attention(qkv, None, table, B, H, H, C, ws, heads).float()
    ref = _ref(qkv, None, table, B, H, H, C, ws, heads)
    torch.cuda.synchronize()
    assert b"mma.sync kernel used" in _lib.lib().vip_last_error()
    assert (got - ref).abs().max().item() < 8e-3 * max(1.0, ref.abs().max().item())
