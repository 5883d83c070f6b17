"""GPU: the non-GEMM denoiser kernels (csrc/nn_kernels.cu) against plain PyTorch fp32 references of
the same ops -- GroupNorm(+SiLU) forward/backward on strided NHWC bf16 views (incl. the fused per-sample
column sums), bias-gradient column sums, nearest upsample and its adjoint, the attention core.
Tolerances: bf16 storage of inputs/outputs (rel 2^-8 per element) around fp32 arithmetic."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nhwc_view(N, H, W, C, c_total=None, c_off=0, seed=0, scale=1.0):
    """bf16 NHWC tensor, optionally a channel slice of a wider buffer (concat-slice addressing)"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    c_total = c_total or C
    buf = (torch.randn(N, H, W, c_total, device="cuda", generator=g) * scale).to(torch.bfloat16)
    return buf[..., c_off:c_off + C]


@pytest.mark.parametrize("cluster_bwd", ["0", "1"], ids=["bwd2pass", "bwdcluster"])
@pytest.mark.parametrize("N,H,C,c_total,c_off,silu", [
    (4, 32, 128, None, 0, True), (3, 16, 256, 384, 128, True), (2, 8, 384, None, 0, True), (5, 4, 512, 1024, 512, False),
    (2, 2, 768, None, 0, True), (7, 1, 1024, None, 0, True), (2, 64, 128, 256, 0, True),
    # cluster-per-sample single-pass kernels: 16-CTA clusters (fwd 384 ch: ragged 5-pixel rows; bwd 256 ch)
    (2, 32, 384, None, 0, True), (3, 32, 256, 512, 256, True), (150, 16, 128, None, 0, True),
    # training-batch small maps: a sample is split into channel slabs of whole groups (grid N x slabs), cpg 4 / 8 / 12 / 24
    (128, 16, 128, None, 0, True), (128, 8, 256, 512, 256, True), (96, 8, 384, None, 0, True), (128, 4, 768, None, 0, False),
    (128, 2, 512, 1024, 0, True)])
def test_groupnorm_silu_fwd_bwd(N, H, C, c_total, c_off, silu, cluster_bwd, monkeypatch):
    from mdm_b200 import denoiser_ops as ops
    monkeypatch.setenv("MDM_GN_CLUSTER_BWD", cluster_bwd)
    G, eps = 32, 1e-5
    x = _nhwc_view(N, H, H, C, c_total, c_off, seed=1)
    dy = _nhwc_view(N, H, H, C, seed=2)
    add = _nhwc_view(N, H, H, C, seed=3)
    gen = torch.Generator(device="cuda").manual_seed(4)
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=gen)
    beta = 0.2 * torch.randn(C, device="cuda", generator=gen)
    y = torch.empty(N, H, H, C, device="cuda", dtype=torch.bfloat16)
    stats = torch.empty(N, G, 2, device="cuda")
    ws = torch.empty(max(1, ops.gn_ws_floats(N, H * H, C)), device="cuda")
    ops.gn_silu_fwd(x, y, gamma, beta, stats, ws, N, H * H, C, G, eps, silu)
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.group_norm(xr, G, gr, br, eps)
    if silu:
        yr = F.silu(yr)
    want = yr.permute(0, 2, 3, 1)
    assert torch.allclose(y.float(), want, atol=2e-2, rtol=1e-2), (y.float() - want).abs().max()
    # backward
    yr.backward(dy.float().permute(0, 3, 1, 2))
    dx = torch.empty(N, H, H, C, device="cuda", dtype=torch.bfloat16)
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    colsum = torch.zeros(N, C + 64, device="cuda")
    dbias = torch.zeros(C, device="cuda")
    ops.gn_silu_bwd(x, dy, dx, gamma, beta, stats, dgamma, dbeta, ws, N, H * H, C, G, silu, add=add,
                    colsum=colsum[:, 32:], ld_colsum=colsum.shape[1], dbias=dbias)
    want_dx = xr.grad.permute(0, 2, 3, 1)
    scale = want_dx.abs().max().item()
    assert torch.allclose(dx.float(), want_dx + add.float(), atol=1e-2 * max(scale, 1.0) + 2e-2, rtol=2e-2)
    assert torch.allclose(dgamma, gr.grad, atol=2e-3 * gr.grad.abs().max().item() + 1e-3, rtol=1e-2)
    assert torch.allclose(dbeta, br.grad, atol=2e-3 * br.grad.abs().max().item() + 1e-3, rtol=1e-2)
    want_cs = xr.grad.sum(dim=(2, 3))                      # GroupNorm part only (without `add`)
    tol = 2e-3 * want_cs.abs().max().item() + 1e-3
    assert torch.allclose(colsum[:, 32:32 + C], want_cs, atol=tol, rtol=1e-2)
    assert torch.allclose(dbias, want_cs.sum(0), atol=N * tol, rtol=1e-2)
    assert colsum[:, :32].abs().max() == 0 and colsum[:, 32 + C:].abs().max() == 0


def test_groupnorm_bwd_inplace_accumulate():
    """dx may alias `add` (x.grad accumulated in place on the concat buffers)"""
    from mdm_b200 import denoiser_ops as ops
    N, H, C, G = 3, 16, 256, 32
    x, dy = _nhwc_view(N, H, H, C, seed=5), _nhwc_view(N, H, H, C, seed=6)
    acc = _nhwc_view(N, H, H, C, seed=7).contiguous()
    acc0 = acc.clone()
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    stats = torch.empty(N, G, 2, device="cuda")
    ws = torch.empty(max(1, ops.gn_ws_floats(N, H * H, C)), device="cuda")
    y = torch.empty_like(acc)
    ops.gn_silu_fwd(x, y, gamma, beta, stats, ws, N, H * H, C, G, 1e-5, True)
    ref = torch.empty_like(acc)
    ops.gn_silu_bwd(x, dy, ref, gamma, beta, stats, None, None, ws, N, H * H, C, G, True)
    ops.gn_silu_bwd(x, dy, acc, gamma, beta, stats, None, None, ws, N, H * H, C, G, True, add=acc)
    assert torch.allclose(acc.float(), ref.float() + acc0.float(), atol=3e-2, rtol=2e-2)


@pytest.mark.parametrize("rows,C,c_total", [(4096, 128, None), (1000, 512, 1024), (128, 9856, None), (77, 1536, None), (3, 384, None)])
def test_colsum(rows, C, c_total):
    from mdm_b200 import denoiser_ops as ops
    d = _nhwc_view(1, 1, rows, C, c_total, 0, seed=8)[0, 0]
    out = torch.zeros(C, device="cuda")
    out2 = torch.ones(C, device="cuda")
    ops.colsum(d, out, rows, C, out2=out2)
    want = d.float().sum(0)
    tol = 1e-3 * want.abs().max().item() + 1e-3
    assert torch.allclose(out, want, atol=tol) and torch.allclose(out2, want + 1, atol=tol)


def test_upsample_and_adjoint():
    from mdm_b200 import denoiser_ops as ops
    N, H, C = 3, 8, 256
    x = _nhwc_view(N, H, H, C, seed=9).contiguous()
    y = torch.empty(N, 2 * H, 2 * H, C, device="cuda", dtype=torch.bfloat16)
    ops.upsample2x_fwd(x, y, N, H, H, C)
    want = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.float(), want)
    dy = _nhwc_view(N, 2 * H, 2 * H, C, seed=10).contiguous()
    dx = torch.empty_like(x)
    ops.upsample2x_bwd(dy, dx, N, H, H, C)
    want_dx = dy.float().view(N, H, 2, H, 2, C).sum(dim=(2, 4))
    assert torch.allclose(dx.float(), want_dx, atol=3e-2, rtol=1e-2)


@pytest.mark.parametrize("N,L,C", [(4, 4, 512), (2, 16, 512), (3, 64, 512), (2, 256, 512), (2, 1, 512),
                                   (8, 256, 1024), (2, 1024, 64), (128, 4, 512), (3, 64, 24)])
def test_attention_core(N, L, C):
    from mdm_b200 import denoiser_ops as ops
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(N * L, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    dout = torch.randn(N * L, C, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(N * L, C, device="cuda", dtype=torch.bfloat16)
    ops.attention_fwd(qkv, out, N, L, C)
    heads = C // 8
    q, k, v = [t.float().view(N, L, heads, 8).transpose(1, 2).requires_grad_(True) for t in qkv.split(C, dim=1)]
    ref = F.scaled_dot_product_attention(q, k, v)
    want = ref.transpose(1, 2).reshape(N * L, C)
    assert torch.allclose(out.float(), want, atol=2e-2, rtol=2e-2)
    ref.backward(dout.float().view(N, L, heads, 8).transpose(1, 2))
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(qkv, dout, dqkv, N, L, C)
    want_d = torch.cat([t.grad.transpose(1, 2).reshape(N * L, C) for t in (q, k, v)], dim=1)
    assert torch.allclose(dqkv.float(), want_d, atol=3e-2 * max(1.0, want_d.abs().max().item()), rtol=3e-2)


@pytest.mark.parametrize("N,H,cin,cout,halo", [(3, 16, 128, 128, 0), (2, 32, 64, 256, 0), (2, 32, 128, 160, 1), (5, 16, 128, 384, 0)])
def test_conv_epilogue_groupnorm_statistics(N, H, cin, cout, halo, monkeypatch):
    """GroupNorm statistics fused into the conv store epilogue (quad sums) + the apply-only forward
    (mdm_gn_silu_fwd_q) == the stand-alone two-kernel forward on the conv's bf16 output"""
    from mdm_b200 import denoiser_ops as ops
    if halo:
        monkeypatch.setenv("MDM_IGEMM_HALO_FORCE", "1")
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(N, H, H, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, 9, cin, device="cuda", generator=g) / (9 * cin) ** 0.5).to(torch.bfloat16)
    b = torch.randn(cout, device="cuda", generator=g)
    res = torch.randn(N, H, H, cout, device="cuda", generator=g).to(torch.bfloat16)
    y = torch.empty(N, H, H, cout, device="cuda", dtype=torch.bfloat16)
    q = torch.zeros(N, cout // 4, 2, device="cuda")
    ops.conv_fprop(x, w, y, N, H, H, 3, 1, bias=b, resid=res, qsum=q)
    yf = y.float().view(N, H * H, cout // 4, 4)
    want_s, want_q = yf.sum(dim=(1, 3)), (yf * yf).sum(dim=(1, 3))
    # the epilogue sums the fp32 values BEFORE the bf16 rounding of the store: agreement to bf16 rounding noise
    assert torch.allclose(q[..., 0], want_s, atol=2e-2 * H, rtol=1e-2)
    assert torch.allclose(q[..., 1], want_q, atol=2e-2 * H, rtol=1e-2)
    if (cout // 32) % 4 == 0:
        gamma = 1 + 0.2 * torch.randn(cout, device="cuda", generator=g)
        beta = 0.2 * torch.randn(cout, device="cuda", generator=g)
        o1, o2 = torch.empty_like(y), torch.empty_like(y)
        st1, st2 = torch.empty(N, 32, 2, device="cuda"), torch.empty(N, 32, 2, device="cuda")
        ws = torch.empty(max(1, ops.gn_ws_floats(N, H * H, cout)), device="cuda")
        ops.gn_silu_fwd(y, o1, gamma, beta, st1, ws, N, H * H, cout, 32, 1e-5, True)
        half = (cout // 4) // 2
        qa, qb = q[:, :half].contiguous(), q[:, half:].contiguous()        # as if two producers had written a concatenation
        ops.gn_silu_fwd_q(y, o2, gamma, beta, st2, qa, qb, N, H * H, cout, 32, 1e-5, True)
        assert torch.allclose(st1, st2, atol=2e-3, rtol=2e-3)
        assert torch.allclose(o1.float(), o2.float(), atol=3e-2, rtol=2e-2)
