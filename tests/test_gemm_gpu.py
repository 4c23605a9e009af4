"""tcgen05 GEMM core versus a plain PyTorch fp32 reference of the same contraction (bf16 inputs, fp32 accumulate)."""
import pytest

pytestmark = pytest.mark.gpu

FX = 2.0 ** -28   # fixed-point unit of the cross-kernel accumulators (csrc/stats.cuh)


def decode_stats(rec, cols):
    """int64 [M, 3] row statistics records -> (mean, biased variance) per row, float64."""
    import torch

    piv = (rec[:, 2] & 0xFFFFFFFF).to(torch.int32).view(torch.float32).double()
    d = rec[:, 0].double() * FX / cols
    return piv + d, rec[:, 1].double() * FX / cols - d * d


def moments(x, eps=1e-5):
    """f32 [M, 2] (mean, 1 / sigma) of the rows of x: the ln_stats operand of a folded LayerNorm."""
    import torch

    xf = x.double()
    return torch.stack([xf.mean(1), 1.0 / torch.sqrt(xf.var(1, unbiased=False) + eps)], 1).float().contiguous()


@pytest.mark.parametrize("m,n,k", [
    (128, 32, 64), (128, 128, 64), (256, 256, 256), (1000, 64, 288), (625 * 3, 512, 1152), (49 * 5, 2048, 512),
    (130, 192 + 64, 96), (4096, 768, 256), (77, 32, 8), (20000, 64, 576),
    # K > 256 and N >= 128: deep configuration, incl. ragged M / N / K tails (the same shapes run as two-CTA
    # cta_group::2 tiles in test_gemm_pair_mode_selected below)
    (5000, 384, 768), (300, 256, 512), (33000, 1152, 384), (257, 128, 320), (1024, 1000, 520),
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


@pytest.mark.parametrize("m,n,k", [(300, 64, 64), (5000, 256, 128), (1000, 384, 96), (777, 768, 256)])
def test_gemm_folded_layernorm_and_row_stats(cuda_device, m, n, k):
    """LayerNorm folded into the contraction (ln_stats / ln_colsum) and the row statistics a producer GEMM emits for it."""
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    # producer: x = a0 @ w0^T (bf16) with fused row statistics
    a0 = (torch.randn(m, 64, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    w0 = (torch.randn(k, 64, generator=g) * 0.3).to(torch.bfloat16).to(cuda_device)
    b0 = (torch.randn(k, generator=g) * 2.0).to(cuda_device)          # DC offset: mean removal must really happen
    stats = nn.row_stats_buffer(m, device=cuda_device)
    x = nn.gemm(a0, w0, bias=b0, row_stats=stats)
    xs = x.float()
    torch.cuda.synchronize()
    mean, var = decode_stats(stats, k)
    assert torch.allclose(mean, xs.double().mean(1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(var, xs.double().var(1, unbiased=False), rtol=1e-4, atol=1e-5)
    # consumer: LN(x) * gamma + beta, then @ W + b, as one contraction on the raw x
    gamma = (1.0 + 0.2 * torch.randn(k, generator=g)).to(cuda_device)
    beta = (0.3 * torch.randn(k, generator=g)).to(cuda_device)
    wt = (torch.randn(n, k, generator=g) * 0.2).to(cuda_device)
    bias = torch.randn(n, generator=g).to(cuda_device)
    wg = (wt * gamma[None, :]).to(torch.bfloat16)
    colsum = wg.float().sum(1).contiguous()
    b2 = (wt @ beta + bias).contiguous()
    ln = nn.finalize_stats(stats, k, 1e-5)
    torch.cuda.synchronize()
    assert torch.allclose(ln, moments(xs), rtol=1e-4, atol=1e-5)
    out = nn.gemm(x, wg, bias=b2, act="gelu", ln_stats=ln, ln_colsum=colsum)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(xs, (k,), gamma, beta, 1e-5) @ wt.t() + bias)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err <= 3e-2 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.parametrize("nimg,hw,n,k", [(3, 49, 256, 64), (5, 169, 512, 128), (2, 2500, 64, 64), (7, 100, 2048, 512)])
def test_gemm_fused_global_average_pool(cuda_device, nimg, hw, n, k):
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(nimg * hw + n)
    a = (torch.randn(nimg * hw, k, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    b = (torch.randn(n, k, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    gap = nn.zero_(torch.empty((nimg, n), dtype=torch.int64, device=cuda_device))
    out = nn.gemm(a, b, gap=gap, gap_rows=hw)
    gap2 = nn.zero_(torch.empty((nimg, n), dtype=torch.int64, device=cuda_device))
    out2 = nn.gemm(a, b, gap=gap2, gap_rows=hw)
    torch.cuda.synchronize()
    assert torch.equal(gap, gap2) and torch.equal(out, out2)      # integer atomics: order-independent, bit-reproducible
    gap = gap.double() * FX
    ref = out.double().view(nimg, hw, n).sum(1)
    assert torch.allclose(gap, ref, rtol=1e-5, atol=1e-4 * hw ** 0.5), (gap - ref).abs().max().item()
    assert (out.float() - a.float() @ b.float().t()).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("nimg,hw,n,k", [(3, 49, 256, 64), (5, 169, 1024, 256), (2, 2500, 256, 64), (6, 49, 2048, 512)])
def test_gemm_se_tail_epilogue(cuda_device, nimg, hw, n, k):
    """SE bottleneck tail fused into the contraction: relu((acc + bias) * gate[image] + shortcut)
    (models/resnet_rs/resnet_rs_model.py:183,278-280)."""
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(nimg * hw + n + k)
    a = (torch.randn(nimg * hw, k, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    b = (torch.randn(n, k, generator=g) * 0.3).to(torch.bfloat16).to(cuda_device)
    bias = torch.randn(n, generator=g).to(cuda_device)
    gate = torch.rand(nimg, n, generator=g).to(cuda_device)
    res = torch.randn(nimg * hw, n, generator=g).to(torch.bfloat16).to(cuda_device)
    out = nn.gemm(a, b, bias=bias, act="relu", residual=res, row_gate=gate, gate_rows=hw)
    ref = torch.relu((a.float() @ b.float().t() + bias) * gate.repeat_interleave(hw, 0) + res.float())
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def test_dwconv_fused_squeeze(cuda_device):
    """DepthwiseConv2D + GELU with the SE squeeze (per-image channel sums) accumulated by the same kernel."""
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(11)
    x = (torch.randn(3, 14, 14, 64, generator=g)).to(torch.bfloat16).to(cuda_device)
    w = (torch.randn(3, 3, 64, generator=g) * 0.3).to(cuda_device)
    gap = nn.zero_(torch.empty((3, 64), dtype=torch.int64, device=cuda_device))
    y = nn.dwconv3x3(x, w, gelu=True, gap=gap)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.permute(2, 0, 1)[:, None], padding=1, groups=64)
    ref = torch.nn.functional.gelu(ref).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert (y.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    assert torch.allclose(gap.double() * FX, y.double().sum((1, 2)), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("m,c", [(1000, 64), (4099, 96), (777, 128), (513, 192), (300, 384), (65, 768)])
def test_layernorm_matches_torch(cuda_device, m, c):
    """vip_layernorm_bf16 (block.py:28,39; feature.py:100-101; gcvit.py:79): 4-lane rows for C <= 128, 32-lane rows
    kernel above, against fp32 torch on the same bf16 input; row_stats = (sum, sum of squares) of the ROUNDED output rows.
    Tolerance: one bf16 ulp of the output (2^-8 relative) + 1e-3."""
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(m + c)
    x = (torch.randn((m, c), generator=g) * 2.0 + 0.5).to(torch.bfloat16).to(cuda_device)
    gamma = (torch.rand((c,), generator=g) + 0.5).to(cuda_device)
    beta = (torch.randn((c,), generator=g) * 0.3).to(cuda_device)
    stats = torch.empty((m, 2), dtype=torch.float32, device=cuda_device)
    got = nn.layernorm(x, gamma, beta, eps=1e-5, ln_next=stats, next_eps=1e-5)
    ref = torch.nn.functional.layer_norm(x.float(), (c,), gamma, beta, 1e-5)
    torch.cuda.synchronize()
    err = (got.float() - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -8 + 1e-3).all(), err.max().item()
    assert torch.allclose(stats, moments(got), rtol=1e-4, atol=1e-5)      # of the ROUNDED output rows


@pytest.mark.parametrize("m,c,hidden", [(128, 96, 192), (1000, 96, 192), (40000, 96, 192), (777, 64, 192), (128 * 149 + 5, 64, 192)])
def test_mlp_fused_matches_two_gemms_and_torch(cuda_device, m, c, hidden):
    """vip_mlp_fused_bf16 (block.py:39-56,77-81: x + fc2(gelu(fc1(LN(x)))), hidden tile kept in TMEM) against (1) the two
    vip_gemm_bf16_ex calls it replaces on the same packed weights -- same arithmetic, so within one bf16 ulp of the output
    -- and (2) fp32 torch with exact-erf GELU (tolerance 2e-2: bf16 hidden + fitted GELU).  Ragged M, more tiles than SMs."""
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(m + c)
    x = (torch.randn((m, c), generator=g) * 1.5 + 0.2).to(torch.bfloat16).to(cuda_device)
    gamma, beta = torch.rand((c,), generator=g) + 0.5, torch.randn((c,), generator=g) * 0.2
    k1 = torch.randn((c, hidden), generator=g) / c ** 0.5
    b1 = torch.randn((hidden,), generator=g) * 0.1
    k2 = torch.randn((hidden, c), generator=g) / hidden ** 0.5
    b2 = torch.randn((c,), generator=g) * 0.1
    w1 = (k1 * gamma[:, None]).T.contiguous().to(torch.bfloat16)            # [hidden, c], gamma folded
    colsum1 = w1.float().sum(1).to(cuda_device)
    bias1 = (beta @ k1 + b1).to(cuda_device)
    w1 = w1.to(cuda_device)
    w2 = k2.T.contiguous().to(torch.bfloat16).to(cuda_device)               # [c, hidden]
    bias2 = b2.to(cuda_device)
    xf = x.float()
    stats = moments(x)
    x_lo_rows = (torch.randn((m, c), generator=g) * 2.0 ** -9).to(torch.bfloat16).to(cuda_device)   # low plane of the stream
    x_lo = nn.lo_plane_from_rows(x_lo_rows)
    ln_f = torch.empty((m, 2), dtype=torch.float32, device=cuda_device)
    got, got_lo = nn.mlp_fused(x, stats, w1, colsum1, bias1, w2, bias2, next_eps=1e-5, ln_next=ln_f, x_lo=x_lo, want_lo=True)
    hdn = nn.gemm(x, w1, bias=bias1, act="gelu", ln_stats=stats, ln_colsum=colsum1)
    rs_g = nn.row_stats_buffer(m, device=cuda_device)
    two_lo = nn.lo_plane(m, c, cuda_device)
    two = nn.gemm(hdn, w2, bias=bias2, residual=x, row_stats=rs_g, residual_lo=x_lo, out_lo=two_lo)
    torch.cuda.synchronize()
    got_lo, two_lo = nn.lo_plane_to_rows(got_lo, m), nn.lo_plane_to_rows(two_lo, m)
    d = (got.float() - two.float()).abs()
    assert (d <= two.float().abs() * 2.0 ** -7 + 1e-3).all(), d.max().item()
    d2 = ((got.float() + got_lo.float()) - (two.float() + two_lo.float())).abs()     # the two-plane sums agree much closer
    assert (d2 <= two.float().abs() * 2.0 ** -7 + 1e-3).all(), d2.max().item()
    ln_g = nn.finalize_stats(rs_g, c, 1e-5)
    torch.cuda.synchronize()
    assert torch.allclose(ln_f, ln_g, rtol=5e-3, atol=2e-3)
    assert torch.allclose(ln_f, moments(got.double() + got_lo.double()), rtol=1e-3, atol=1e-4)
    ln = torch.nn.functional.layer_norm(xf.cpu(), (c,), gamma, beta, 1e-5)
    ref = xf.cpu() + torch.nn.functional.gelu(ln @ k1 + b1) @ k2 + b2
    assert (got.float().cpu() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("groups,rows,n,k", [(3, 128, 96, 96), (5, 256, 64, 64), (2, 12544, 96, 96), (7, 384, 192, 192)])
def test_gemm_grouped_folds_a_per_image_input_scale(cuda_device, groups, rows, n, k):
    """vip_scale_weights_bf16 + vip_gemm_grouped_bf16 (SE gate of an MBConv block folded into per-image copies of the 1x1
    convolution weights, feature.py:144-150) against fp32 torch on (y * gate_image) @ W^T + residual."""
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(groups * 1000 + rows)
    m = groups * rows
    y = torch.randn((m, k), generator=g).to(torch.bfloat16).to(cuda_device)
    res = torch.randn((m, n), generator=g).to(torch.bfloat16).to(cuda_device)
    w = (torch.randn((n, k), generator=g) / k ** 0.5).to(torch.bfloat16).to(cuda_device)
    gate = torch.rand((groups, k), generator=g).to(cuda_device)
    wg = nn.scale_weights(w, gate)
    got = nn.gemm_grouped(y, wg, rows, residual=res)
    torch.cuda.synchronize()
    assert torch.equal(wg.float().view(groups, n, k), (w.float()[None] * gate[:, None, :]).to(torch.bfloat16).float())
    ref = torch.einsum("gmk,gnk->gmn", y.float().view(groups, rows, k), wg.float().view(groups, n, k)).reshape(m, n) + res.float()
    err = (got.float() - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -8 + 2e-3).all(), err.max().item()


@pytest.mark.parametrize("m,n,k", [(300, 64, 64), (4099, 96, 96), (1000, 384, 768), (777, 768, 1536), (128 * 150, 192, 192)])
def test_gemm_two_plane_residual_stream(cuda_device, m, n, k):
    """residual_lo / out_lo (vip_epilogue_t): the block residual stream of models/gcvit/layers/block.py:77-81 carried as
    hi + lo bf16 planes.  hi + lo must reproduce acc + bias + (res_hi + res_lo) to ~2^-16 relative (a single bf16 plane
    gives 2^-9), the row statistics must describe hi + lo, and two launches must agree bit for bit."""
    import torch

    from vipcup_b200 import nn

    g = torch.Generator(device="cpu").manual_seed(m + 3 * n + k)
    a = (torch.randn(m, k, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(torch.bfloat16).to(cuda_device)
    bias = torch.randn(n, generator=g).to(cuda_device)
    res = (torch.randn(m, n, generator=g) * 3.0 + 5.0)
    hi = res.to(torch.bfloat16)
    lo = (res - hi.float()).to(torch.bfloat16)
    hi, lo = hi.to(cuda_device), lo.to(cuda_device)
    lo_blk = nn.lo_plane_from_rows(lo)
    assert torch.equal(nn.lo_plane_to_rows(lo_blk, m), lo)
    outs = []
    for _ in range(2):
        st = nn.row_stats_buffer(m, device=cuda_device)
        out_lo = nn.lo_plane(m, n, cuda_device)
        out = nn.gemm(a, w, bias=bias, residual=hi, residual_lo=lo_blk, out_lo=out_lo, row_stats=st,
                      row_pivot=moments(hi.float() + lo.float()))
        outs.append((out, nn.lo_plane_to_rows(out_lo, m), st))
    torch.cuda.synchronize()
    for t0, t1 in zip(*outs):
        assert torch.equal(t0, t1)
    out, out_lo, st = outs[0]
    ref = a.double() @ w.double().t() + bias.double() + hi.double() + lo.double()
    two = out.double() + out_lo.double()
    err = (two - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -15 + 1e-5).all(), err.max().item()     # fp32 accumulation + 16-bit planes
    one = (out.double() - ref).abs()
    assert one.max() > 8 * err.max()                                          # the low plane really carries the remainder
    mean, var = decode_stats(st, n)
    assert torch.allclose(mean, ref.mean(1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(var, ref.var(1, unbiased=False), rtol=1e-4, atol=1e-5)
    # first block of a level: no incoming low plane
    out_lo2 = nn.lo_plane(m, n, cuda_device)
    out2 = nn.gemm(a, w, bias=bias, residual=hi, out_lo=out_lo2)
    ref2 = a.double() @ w.double().t() + bias.double() + hi.double()
    torch.cuda.synchronize()
    out_lo2 = nn.lo_plane_to_rows(out_lo2, m)
    assert ((out2.double() + out_lo2.double() - ref2).abs() <= ref2.abs() * 2.0 ** -15 + 1e-5).all()


@pytest.mark.parametrize("c,n", [(96, 288), (384, 768)])
def test_folded_layernorm_with_large_row_mean(cuda_device, c, n):
    """Rows with |mean| >> sigma (mean 50, sigma 0.1 -- outlier channels of trained ViT residual streams): the pivoted
    one-pass statistics (csrc/stats.cuh) must not lose the variance to cancellation.  Producer = a two-plane residual GEMM,
    consumer = a contraction with the LayerNorm folded in, reference = torch.nn.functional.layer_norm in float64."""
    import torch

    from vipcup_b200 import nn

    m = 1000
    g = torch.Generator(device="cpu").manual_seed(c)
    x0 = 50.0 + 0.1 * torch.randn(m, c, generator=g)
    hi = x0.to(torch.bfloat16)
    lo = (x0 - hi.float()).to(torch.bfloat16)
    hi, lo = hi.to(cuda_device), lo.to(cuda_device)
    a = (torch.randn(m, 64, generator=g) * 0.05).to(torch.bfloat16).to(cuda_device)
    w0 = (torch.randn(c, 64, generator=g) * 0.1).to(torch.bfloat16).to(cuda_device)
    st = nn.row_stats_buffer(m, device=cuda_device)
    x_lo = nn.lo_plane(m, c, cuda_device)
    x = nn.gemm(a, w0, residual=hi, residual_lo=nn.lo_plane_from_rows(lo), out_lo=x_lo, row_stats=st,
                row_pivot=moments(hi.float() + lo.float()))     # as in a block: the moments the previous LayerNorm used
    torch.cuda.synchronize()
    x_lo = nn.lo_plane_to_rows(x_lo, m)
    full = x.double() + x_lo.double()
    mean, var = decode_stats(st, c)
    assert torch.allclose(mean, full.mean(1), rtol=1e-6, atol=1e-5)
    assert torch.allclose(var, full.var(1, unbiased=False), rtol=2e-3, atol=1e-7), (var - full.var(1, unbiased=False)).abs().max()
    # the LayerNorm itself, through the folded contraction with identity-like weights: out = LN(x) @ W^T
    gamma = (1.0 + 0.1 * torch.randn(c, generator=g)).to(cuda_device)
    beta = (0.1 * torch.randn(c, generator=g)).to(cuda_device)
    wt = (torch.randn(n, c, generator=g) / c ** 0.5).to(cuda_device)
    wg = (wt * gamma[None, :]).to(torch.bfloat16)
    out = nn.gemm(x, wg, bias=(wt @ beta).contiguous(), ln_stats=nn.finalize_stats(st, c, 1e-5),
                  ln_colsum=wg.float().sum(1).contiguous(), out_dtype=torch.float32)
    # reference on the operand the kernel contracts (the hi plane) with the statistics of the full-precision stream
    mu, sd = full.mean(1, keepdim=True), (full.var(1, unbiased=False, keepdim=True) + 1e-5).sqrt()
    ref = ((x.double() - mu) / sd) @ wg.double().t() + (wt.double() @ beta.double())
    torch.cuda.synchronize()
    err = (out.double() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), err


def test_gemm_pair_mode_selected(cuda_device):
    """The two-CTA (tcgen05.mma.cta_group::2, 256-row tiles) configuration only engages for K >= 4096 by default; here the
    threshold is lowered in a child process (the library reads VIP_GEMM_PAIR_MIN_KB once) so that the deep shapes of the
    test above -- ragged M / N / K tails included -- run through it, and the kernel name is checked in the trace output."""
    import os
    import subprocess
    import sys

    if os.environ.get("VIP_GEMM_PAIR_MIN_KB") == "5":
        import torch

        from vipcup_b200 import nn

        for m, n, k in [(5000, 384, 768), (300, 256, 512), (33000, 1152, 384), (257, 128, 320), (1024, 1000, 520), (2048, 512, 4096)]:
            g = torch.Generator(device="cpu").manual_seed(m + n + k)
            a = (torch.randn(m, k, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
            b = (torch.randn(n, k, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
            bias = torch.randn(n, generator=g).to(cuda_device)
            res = torch.randn(m, n, generator=g).to(torch.bfloat16).to(cuda_device)
            ref = torch.relu(a.float() @ b.float().t() + bias) + res.float()
            out = nn.gemm(a, b, bias=bias, act="relu", residual=res).float()
            torch.cuda.synchronize()
            tol = 2e-2 * max(1.0, ref.abs().max().item())
            assert (out - ref).abs().max().item() < tol, (m, n, k)
        return
    env = dict(os.environ, VIP_GEMM_PAIR_MIN_KB="5", VIP_GEMM_TRACE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-s", __file__, "-k", "pair_mode_selected"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "pair=1" in r.stderr + r.stdout, "the two-CTA configuration was not selected"
