"""GPU: tcgen05 implicit-GEMM conv (fprop / dgrad / wgrad) against torch's fp32 conv2d on the
same bf16-rounded operands.  Tolerance: the kernel accumulates bf16 products in fp32 (exact
products, fp32 sums); the output is rounded to bf16 -> relative error <= 2^-8 per element plus
summation-order noise; stated as |err| <= 1e-2 * max|ref| (fprop/dgrad, bf16 output) and
<= 2e-3 * max|ref| for wgrad (fp32 output, split-K atomics)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def nhwc(t):   # NCHW fp32 -> NHWC bf16 contiguous
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


CASES = [
    # N, H(out), cin, cout, k, stride
    (2, 32, 128, 128, 3, 1),
    (3, 16, 128, 256, 3, 1),
    (2, 8, 256, 256, 3, 1),
    (5, 4, 512, 512, 3, 1),
    (16, 2, 512, 512, 3, 1),
    (7, 1, 1024, 512, 3, 1),
    (2, 16, 384, 256, 1, 1),
    (2, 16, 128, 128, 3, 2),
    (4, 4, 256, 256, 3, 2),
    (1, 128, 128, 128, 3, 1),
    (200, 1, 512, 512, 1, 1),        # a Linear: 200 rows
]


@pytest.mark.parametrize("N,H,cin,cout,k,stride", CASES)
def test_fprop(N, H, cin, cout, k, stride):
    from mdm_b200 import denoiser_ops as ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, cin, H * stride, H * stride, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    rv = torch.randn(N, cout, device="cuda", generator=g)
    res = torch.randn(N, cout, H, H, device="cuda", generator=g)
    xb, wb, resb = nhwc(x), ops.pack_conv_weight(w).to(torch.bfloat16), nhwc(res)
    y = torch.empty(N, H, H, cout, device="cuda", dtype=torch.bfloat16)
    ops.conv_fprop(xb, wb, y, N, H, H, k, stride, bias=b, rowvec=rv, resid=resb)
    ref = F.conv2d(nchw(xb), ops.unpack_conv_weight(wb.float(), k), b, stride=stride, padding=k // 2)
    ref = ref + rv[:, :, None, None] + nchw(resb)
    err = (nchw(y) - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item(), err


def test_fprop_fused_shortcut_and_slices():
    """conv3x3(a) + conv1x1(x) in one accumulator; inputs/outputs are channel slices of wider buffers."""
    from mdm_b200 import denoiser_ops as ops
    N, H, cin, cout, cx = 2, 16, 256, 256, 384
    g = torch.Generator(device="cuda").manual_seed(1)
    abuf = torch.randn(N, H, H, cin + 64, device="cuda", generator=g).to(torch.bfloat16)
    xbuf = torch.randn(N, H, H, cx + 128, device="cuda", generator=g).to(torch.bfloat16)
    a, x = abuf[..., 64:], xbuf[..., :cx]
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (9 * cin) ** 0.5)
    ws = (torch.randn(cout, cx, 1, 1, device="cuda", generator=g) / cx ** 0.5)
    wb, wsb = ops.pack_conv_weight(w).to(torch.bfloat16), ops.pack_conv_weight(ws).to(torch.bfloat16)
    ybuf = torch.zeros(N, H, H, cout + 256, device="cuda", dtype=torch.bfloat16)
    y = ybuf[..., 256:]
    yf = torch.empty(N * H * H, cout, device="cuda")
    ops.conv_fprop(a, wb, y, N, H, H, 3, 1, x2=x, w2=wsb, y_f32=yf)
    ref = F.conv2d(nchw(a), ops.unpack_conv_weight(wb.float(), 3), padding=1) + F.conv2d(nchw(x), ops.unpack_conv_weight(wsb.float(), 1))
    assert (nchw(y) - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert (yf.view(N, H, H, cout).permute(0, 3, 1, 2) - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert ybuf[..., :256].abs().max().item() == 0          # nothing written outside the slice


@pytest.mark.parametrize("N,H,cin,cout,k,stride", [c for c in CASES if c[5] == 1])
def test_dgrad(N, H, cin, cout, k, stride):
    from mdm_b200 import denoiser_ops as ops
    g = torch.Generator(device="cuda").manual_seed(2)
    dy = torch.randn(N, cout, H, H, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cout * k * k) ** 0.5
    dyb, wb = nhwc(dy), ops.pack_conv_weight(w).to(torch.bfloat16)
    dx = torch.empty(N, H, H, cin, device="cuda", dtype=torch.bfloat16)
    ops.conv_dgrad(dyb, wb, dx, N, H, H, k)
    ref = torch.nn.grad.conv2d_input((N, cin, H, H), ops.unpack_conv_weight(wb.float(), k), nchw(dyb), padding=k // 2)
    err = (nchw(dx) - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item(), err
    # accumulate into existing contents
    ops.conv_dgrad(dyb, wb, dx, N, H, H, k, accumulate=True)
    assert (nchw(dx) - 2 * ref).abs().max().item() <= 2.5e-2 * ref.abs().max().item()


@pytest.mark.parametrize("N,H,cin,cout,k,stride", CASES)
def test_wgrad(N, H, cin, cout, k, stride):
    from mdm_b200 import denoiser_ops as ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(N, cin, H * stride, H * stride, device="cuda", generator=g)
    dy = torch.randn(N, cout, H, H, device="cuda", generator=g)
    xb, dyb = nhwc(x), nhwc(dy)
    dw = torch.zeros(cout, k * k, cin, device="cuda")
    db = torch.full((cout,), 0.5, device="cuda")
    db2 = torch.zeros(cout, device="cuda")
    ops.conv_wgrad(xb, dyb, dw, N, H, H, k, stride, dbias=db, dbias2=db2)
    ref = torch.nn.grad.conv2d_weight(nchw(xb), (cout, cin, k, k), nchw(dyb), stride=stride, padding=k // 2)
    err = (ops.unpack_conv_weight(dw, k) - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-4, err
    # bias gradient fused into the same kernel (one N=16 MMA against a tile of ones): accumulates into dbias
    ref_b = dyb.float().sum(dim=(0, 1, 2))
    tol = 2e-3 * ref_b.abs().max().item() + 1e-3
    assert (db - 0.5 - ref_b).abs().max().item() <= tol and (db2 - ref_b).abs().max().item() <= tol


# ---- halo mode: the M tile is a pixel patch whose zero-padded input halo is TMA-loaded once per channel chunk;
# the nine taps are shifted UMMA views of it.  kind 2 = 16x16 patches (256-row work items, the default for big
# layers), kind 1 = 8x16 patches.  The small test shapes are forced into halo mode through the environment.
HALO_CASES = [
    # N, H, cin, cout
    (2, 32, 128, 128),
    (3, 16, 128, 256),
    (1, 64, 192, 160),      # narrow last cout tile, 3 channel chunks
    (150, 16, 64, 128),     # more work items than SMs: several items per CTA (ring / accumulator parities wrap)
]


@pytest.fixture(params=[2, 1], ids=["halo256", "halo128"])
def halo_env(request, monkeypatch):
    monkeypatch.setenv("MDM_IGEMM_HALO", "1" if request.param == 1 else "2")
    monkeypatch.setenv("MDM_IGEMM_HALO_FORCE", "1")
    return request.param


@pytest.mark.parametrize("N,H,cin,cout", HALO_CASES)
def test_halo_fprop_dgrad(halo_env, N, H, cin, cout):
    from mdm_b200 import denoiser_ops as ops
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, cin, H, H, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    rv = torch.randn(N, cout, device="cuda", generator=g)
    res = torch.randn(N, cout, H, H, device="cuda", generator=g)
    xb, wb, resb = nhwc(x), ops.pack_conv_weight(w).to(torch.bfloat16), nhwc(res)
    y = torch.empty(N, H, H, cout, device="cuda", dtype=torch.bfloat16)
    ops.conv_fprop(xb, wb, y, N, H, H, 3, 1, bias=b, rowvec=rv, resid=resb)
    ref = F.conv2d(nchw(xb), ops.unpack_conv_weight(wb.float(), 3), b, padding=1) + rv[:, :, None, None] + nchw(resb)
    assert (nchw(y) - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    if cin % 128 == 0 and cout % 64 == 0:
        dyb = nhwc(torch.randn(N, cout, H, H, device="cuda", generator=g))
        dx = torch.empty(N, H, H, cin, device="cuda", dtype=torch.bfloat16)
        ops.conv_dgrad(dyb, wb, dx, N, H, H, 3)
        refd = torch.nn.grad.conv2d_input((N, cin, H, H), ops.unpack_conv_weight(wb.float(), 3), nchw(dyb), padding=1)
        assert (nchw(dx) - refd).abs().max().item() <= 1e-2 * refd.abs().max().item()
        ops.conv_dgrad(dyb, wb, dx, N, H, H, 3, accumulate=True)
        assert (nchw(dx) - 2 * refd).abs().max().item() <= 2.5e-2 * refd.abs().max().item()


def test_halo_fused_shortcut_slices(halo_env):
    """conv3x3(a) + conv1x1(x) in one accumulator in halo mode; operands are channel slices of wider buffers."""
    from mdm_b200 import denoiser_ops as ops
    N, H, cin, cout, cx = 3, 16, 256, 256, 384
    g = torch.Generator(device="cuda").manual_seed(6)
    abuf = torch.randn(N, H, H, cin + 64, device="cuda", generator=g).to(torch.bfloat16)
    xbuf = torch.randn(N, H, H, cx + 128, device="cuda", generator=g).to(torch.bfloat16)
    a, x = abuf[..., 64:], xbuf[..., :cx]
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (9 * cin) ** 0.5)
    ws = (torch.randn(cout, cx, 1, 1, device="cuda", generator=g) / cx ** 0.5)
    wb, wsb = ops.pack_conv_weight(w).to(torch.bfloat16), ops.pack_conv_weight(ws).to(torch.bfloat16)
    ybuf = torch.zeros(N, H, H, cout + 256, device="cuda", dtype=torch.bfloat16)
    y = ybuf[..., 256:]
    ops.conv_fprop(a, wb, y, N, H, H, 3, 1, x2=x, w2=wsb)
    ref = F.conv2d(nchw(a), ops.unpack_conv_weight(wb.float(), 3), padding=1) + F.conv2d(nchw(x), ops.unpack_conv_weight(wsb.float(), 1))
    assert (nchw(y) - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert ybuf[..., :256].abs().max().item() == 0


def test_halo_tail_wave_single_items():
    """the last partial wave of a 16x16-patch halo layer is re-cut into single 128-row items: 150 patches x 2 cout tiles
    = 300 work items = 2 full waves of 148 + 4 pair items -> 8 single items; fprop with fused 1x1 shortcut and dgrad"""
    from mdm_b200 import denoiser_ops as ops
    N, H, cin, cout, cx = 150, 16, 128, 256, 64
    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(N, cin, H, H, device="cuda", generator=g)
    xs = torch.randn(N, cx, H, H, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5
    ws = torch.randn(cout, cx, 1, 1, device="cuda", generator=g) / cx ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    xb, xsb = nhwc(x), nhwc(xs)
    wb, wsb = ops.pack_conv_weight(w).to(torch.bfloat16), ops.pack_conv_weight(ws).to(torch.bfloat16)
    y = torch.empty(N, H, H, cout, device="cuda", dtype=torch.bfloat16)
    ops.conv_fprop(xb, wb, y, N, H, H, 3, 1, bias=b, x2=xsb, w2=wsb)
    ref = F.conv2d(nchw(xb), ops.unpack_conv_weight(wb.float(), 3), b, padding=1) + F.conv2d(nchw(xsb), ops.unpack_conv_weight(wsb.float(), 1))
    assert (nchw(y) - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    dyb = nhwc(torch.randn(N, cout, H, H, device="cuda", generator=g))
    res = nhwc(torch.randn(N, cin, H, H, device="cuda", generator=g))
    dx = torch.empty(N, H, H, cin, device="cuda", dtype=torch.bfloat16)
    ops.conv_dgrad(dyb, wb, dx, N, H, H, 3, resid=res)
    refd = torch.nn.grad.conv2d_input((N, cin, H, H), ops.unpack_conv_weight(wb.float(), 3), nchw(dyb), padding=1) + nchw(res)
    assert (nchw(dx) - refd).abs().max().item() <= 1e-2 * refd.abs().max().item()


@pytest.mark.parametrize("halo", ["0", "2"])
def test_dynamic_work_distribution_matches_static(monkeypatch, halo):
    """work items drawn from the atomic counter (every launch forced dynamic) vs the static lists: every output tile is
    computed by the same arithmetic whichever CTA gets it, so fprop (bias + residual + fused GroupNorm quad sums'
    output) and dgrad are BIT-identical; wgrad's fp32 reduce-adds only change order.  300+ items: several per CTA."""
    from mdm_b200 import denoiser_ops as ops
    monkeypatch.setenv("MDM_IGEMM_HALO", halo)
    N, H, cin, cout = 150, 16, 128, 256
    g = torch.Generator(device="cuda").manual_seed(33)
    xb = nhwc(torch.randn(N, cin, H, H, device="cuda", generator=g))
    dyb = nhwc(torch.randn(N, cout, H, H, device="cuda", generator=g))
    res = nhwc(torch.randn(N, cout, H, H, device="cuda", generator=g))
    wb = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5).to(torch.bfloat16)
    b = torch.randn(cout, device="cuda", generator=g)
    out = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("MDM_IGEMM_DYNAMIC", mode)
        y = torch.empty(N, H, H, cout, device="cuda", dtype=torch.bfloat16)
        q = torch.zeros(N, cout // 4, 2, device="cuda")
        ops.conv_fprop(xb, wb, y, N, H, H, 3, 1, bias=b, resid=res, qsum=q)
        dx = torch.empty(N, H, H, cin, device="cuda", dtype=torch.bfloat16)
        ops.conv_dgrad(dyb, wb, dx, N, H, H, 3)
        dw = torch.zeros(cout, 9, cin, device="cuda")
        db = torch.zeros(cout, device="cuda")
        ops.conv_wgrad(xb, dyb, dw, N, H, H, 3, 1, dbias=db)
        torch.cuda.synchronize()
        out[mode] = (y, dx, dw, db, q)
    assert torch.equal(out["0"][0], out["2"][0]) and torch.equal(out["0"][1], out["2"][1])
    for i in (2, 3, 4):
        a, c = out["0"][i], out["2"][i]
        assert (a - c).abs().max().item() <= 1e-4 * a.abs().max().item()
    ref = F.conv2d(nchw(xb), ops.unpack_conv_weight(wb.float(), 3), b, padding=1) + nchw(res)
    assert (nchw(out["2"][0]) - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("N,H,cin", [(2, 16, 128), (3, 32, 256), (1, 64, 128)])
def test_fused_upsample_conv(N, H, cin):
    """diffusers Upsample2D = conv3x3(interpolate(x, 2, 'nearest')): four parity launches over the low-resolution input
    with pre-summed 2x2 weights (4/9 of the FLOPs, no upsampled tensor), output incl. bias and the fused GroupNorm
    quad sums equals the explicit upsample + convolution"""
    from mdm_b200 import denoiser_ops as ops
    g = torch.Generator(device="cuda").manual_seed(5)
    cout = cin
    x = torch.randn(N, cin, H, H, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (9 * cin) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    xb = nhwc(x)
    wp = ops.pack_conv_weight(w).contiguous()                       # fp32 [cout, 9, cin]: the master-weight layout
    w4 = torch.empty(4, cout, 4, cin, device="cuda", dtype=torch.bfloat16)
    ops.up2x_weights(wp, w4, cout, cin)
    # the parity weights are exact sums of the fp32 taps, rounded once
    wr = wp.view(cout, 3, 3, cin)
    rows = {0: [wr[:, 0:1].sum(1), wr[:, 1:3].sum(1)], 1: [wr[:, 0:2].sum(1), wr[:, 2:3].sum(1)]}
    for a in range(2):
        for bb in range(2):
            for u in range(2):
                r = rows[a][u]                                      # [cout, 3 (s), cin]
                cols = [r[:, 0:1].sum(1), r[:, 1:3].sum(1)] if bb == 0 else [r[:, 0:2].sum(1), r[:, 2:3].sum(1)]
                for v in range(2):
                    torch.testing.assert_close(w4[2 * a + bb, :, 2 * u + v].float(), cols[v].to(torch.bfloat16).float(), rtol=1e-2, atol=1e-3)
    y = torch.zeros(N, 2 * H, 2 * H, cout, device="cuda", dtype=torch.bfloat16)
    q = torch.zeros(N, cout // 4, 2, device="cuda")
    ops.conv_up2x_fprop(xb, w4, y, N, H, H, bias=b, qsum=q)
    up = F.interpolate(nchw(xb), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w, b, padding=1)                             # fp32 weights: the fused path rounds the SUMS, not the taps
    err = (nchw(y) - ref).abs().max().item()
    assert err <= 1.5e-2 * ref.abs().max().item(), err
    yq = nchw(y).view(N, cout // 4, 4, -1)
    want = torch.stack([yq.sum(dim=(2, 3)), (yq * yq).sum(dim=(2, 3))], dim=-1)
    # the statistics are taken from the fp32 accumulators before the bf16 rounding of y
    assert (q - want).abs().max().item() <= 2e-2 * want.abs().max().item()


@pytest.mark.parametrize("N,H,cin,cout,shortcut,stats", [(2, 32, 128, 128, False, True), (3, 16, 256, 128, True, False),
                                                         (1, 64, 128, 128, False, False), (2, 32, 384, 256, True, True),
                                                         # several work items per CTA (the operand ring wraps across items,
                                                         # odd and even slot uses per item, two cout tiles, tail items)
                                                         (20, 64, 256, 128, True, True), (40, 32, 128, 256, True, False),
                                                         (160, 16, 256, 128, False, True)])
def test_fprop_with_folded_groupnorm(N, H, cin, cout, shortcut, stats, monkeypatch):
    """inference: GroupNorm + SiLU of the input applied inside the convolution (transform warps fill the halo tiles with
    silu(x * scale + shift), zero outside the map) == the stand-alone apply pass followed by the convolution"""
    from mdm_b200 import denoiser_ops as ops
    for dyn in ("1", "2"):
        monkeypatch.setenv("MDM_IGEMM_DYNAMIC", dyn)
        g = torch.Generator(device="cuda").manual_seed(11)
        x = torch.randn(N, cin, H, H, device="cuda", generator=g)
        w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (9 * cin) ** 0.5
        b = torch.randn(cout, device="cuda", generator=g)
        coef = torch.stack([1.0 + 0.5 * torch.randn(N, cin, device="cuda", generator=g), 0.5 * torch.randn(N, cin, device="cuda", generator=g)], dim=-1).contiguous()
        xb, wb = nhwc(x), ops.pack_conv_weight(w).to(torch.bfloat16)
        xs = torch.randn(N, 192, H, H, device="cuda", generator=g)
        ws_ = torch.randn(cout, 192, 1, 1, device="cuda", generator=g) / 192 ** 0.5
        res = torch.randn(N, cout, H, H, device="cuda", generator=g)
        y = torch.empty(N, H, H, cout, device="cuda", dtype=torch.bfloat16)
        q = torch.zeros(N, cout // 4, 2, device="cuda") if stats else None
        kw = dict(x2=nhwc(xs), w2=ops.pack_conv_weight(ws_).to(torch.bfloat16)) if shortcut else dict(resid=nhwc(res))
        ops.conv_fprop(xb, wb, y, N, H, H, 3, 1, bias=b, qsum=q, gn_coef=coef, **kw)
        a = nchw(xb) * coef[..., 0][:, :, None, None] + coef[..., 1][:, :, None, None]
        a = torch.nn.functional.silu(a).to(torch.bfloat16).float()          # the kernel rounds the normalised value to bf16
        ref = F.conv2d(a, ops.unpack_conv_weight(wb.float(), 3), b, padding=1)
        ref = ref + (F.conv2d(nchw(nhwc(xs)), ops.unpack_conv_weight(ops.pack_conv_weight(ws_).to(torch.bfloat16).float(), 1)) if shortcut else nchw(nhwc(res)))
        err = (nchw(y) - ref).abs().max().item()
        assert err <= 1.2e-2 * ref.abs().max().item(), (dyn, err)
        if stats:
            yq = ref.view(N, cout // 4, 4, -1)
            want = torch.stack([yq.sum(dim=(2, 3)), (yq * yq).sum(dim=(2, 3))], dim=-1)
            assert (q - want).abs().max().item() <= 2e-2 * want.abs().max().item()
