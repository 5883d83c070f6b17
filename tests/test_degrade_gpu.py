"""GPU: drop-in Scheduler (K1 kernels) against the reference goldens and the oracle.

Tolerances: masks, mask-derived tensors and generator state bit-exact; the fill value is a fp32
masked mean whose summation order differs from torch's -> |diff| <= 2e-6, and the composite
inherits exactly that difference at degraded pixels (kept pixels are bit-exact)."""
import numpy as np
import pytest
import torch

import scheduler
from oracle.mdm_oracle import OracleRNG, OracleScheduler, default_args
from tests.golden.make_golden import DEGRADE_CASES, SHIFT_TYPES, mk_args
from tests.helpers import torch_state_words, unpack_mask

pytestmark = pytest.mark.gpu
FILL_TOL = 2e-6


def make(a, seed):
    S = scheduler.Scheduler(a)
    S.update_ddpm_num_steps(a.ddpm_num_steps)
    torch.manual_seed(seed)
    S.adopt_torch_rng("cuda")
    return S


def state_equal(S, state_u8):
    key, pos = S.rng.export()
    gk, gp = torch_state_words(state_u8)
    return pos == gp and np.array_equal(key, gk)


@pytest.mark.parametrize("name", list(DEGRADE_CASES))
def test_degrade_against_reference_golden(golden, name):
    g = golden("degrade")
    a = mk_args(data_size=16, ddpm_num_steps=100, **DEGRADE_CASES[name])
    S = make(a, 3)
    x0 = torch.from_numpy(g[f"{name}/x0"]).to(a.weight_dtype).cuda()
    ts = torch.from_numpy(g[f"{name}/timesteps"]).cuda()
    n = S.get_black_area_num_pixels_time(ts)
    assert np.array_equal(n.cpu().numpy(), g[f"{name}/n"])
    d_img, masks, d_mask, mean_mask = S.degrade_training(n, x0, a.mean_option, a.mean_area)
    assert d_img.dtype == torch.float32 and masks.dtype == torch.float32      # SURVEY 3.2 dtype semantics
    gm = unpack_mask(g[f"{name}/masks"], g[f"{name}/masks_shape"])
    assert torch.equal(masks.cpu(), gm)                                       # bit-exact masks
    np.testing.assert_allclose(mean_mask[:, :, 0, 0].cpu().numpy(), g[f"{name}/fill"], atol=FILL_TOL, rtol=0)
    np.testing.assert_allclose(d_img.cpu().numpy(), g[f"{name}/degrade_img"], atol=FILL_TOL, rtol=0)
    np.testing.assert_allclose(d_mask.cpu().numpy(), g[f"{name}/degrade_mask"], atol=FILL_TOL, rtol=0)
    kept = gm.bool()
    assert torch.allclose(d_img.cpu()[kept], torch.from_numpy(g[f"{name}/degrade_img"])[kept], atol=0, rtol=0,
                          equal_nan=True)                                      # kept pixels exact (NaN rows: quirk q6)
    s_img, s_masks, s_mean = S.degrade_independent_base_sampling(n, x0.float(), a.mean_option, a.mean_area)
    np.testing.assert_allclose(s_img.cpu().numpy(), g[f"{name}/s_img"], atol=FILL_TOL, rtol=0)
    w = S.degrade_with_mask(x0.float(), s_masks, a.mean_option, a.mean_area)
    np.testing.assert_allclose(w.cpu().numpy(), g[f"{name}/w_img"], atol=FILL_TOL, rtol=0)
    assert state_equal(S, g[f"{name}/state_after"])


def test_dependent_two_threshold(golden):
    g = golden("degrade")
    a = mk_args(data_size=16, ddpm_num_steps=100, select_degrade_pixel="thresholding", ddpm_schedule="linear",
                degrade_channel="1-channel", mean_option="degraded_area", mean_area="image-wise")
    S = make(a, 5)
    ts = torch.from_numpy(g["dep2/timesteps"]).cuda()
    r = S.degrade_dependent_base_sampling(S.get_black_area_num_pixels_time(ts), S.get_black_area_num_pixels_time(ts - 1),
                                          torch.from_numpy(g["dep2/x0"]).cuda(), "degraded_area", "image-wise")
    np.testing.assert_allclose(r[0].cpu().numpy(), g["dep2/img_t"], atol=FILL_TOL, rtol=0)
    np.testing.assert_allclose(r[3].cpu().numpy(), g["dep2/img_n"], atol=FILL_TOL, rtol=0)
    assert torch.equal(r[1].cpu(), unpack_mask(g["dep2/mask_t"], r[1].shape))
    assert torch.equal(r[4].cpu(), unpack_mask(g["dep2/mask_n"], r[4].shape))


@pytest.mark.parametrize("st", SHIFT_TYPES)
@pytest.mark.parametrize("wd", ["fp32", "bf16"])
def test_shift_against_reference_golden(golden, st, wd):
    g = golden("shift")
    a = mk_args(data_size=16, ddpm_num_steps=100, select_degrade_pixel="thresholding", ddpm_schedule="linear",
                shift_type=st, noise_mean=0.25, weight_dtype=wd)
    S = make(a, 9)
    ts = torch.tensor([1., 7., 33., 64., 99., 100.]).cuda()
    sh = S.get_schedule_shift_time(ts, torch.zeros(6, 3, 16, 16, device="cuda"))
    assert sh.dtype == a.weight_dtype and sh.shape == (6, 3, 16, 16)
    tol = 2e-6 if wd == "fp32" else 1e-2      # normal_ differs across libm variants
    np.testing.assert_allclose(sh.float().cpu().numpy(), g[f"{st}/{wd}/shift"], atol=tol, rtol=0)
    assert state_equal(S, g[f"{st}/{wd}/state_after"])


def test_shift_quirk_q7(golden):
    g = golden("shift")
    a = mk_args(data_size=8, ddpm_num_steps=50, select_degrade_pixel="thresholding", ddpm_schedule="linear",
                shift_type="noise_with_perturbation")
    S = make(a, 9)
    sh = S.get_schedule_shift_time((torch.arange(1, 9).float() * 5).cuda(), torch.zeros(8, 3, 8, 8, device="cuda"))
    np.testing.assert_allclose(sh.cpu().numpy(), g["q7/shift"], atol=2e-6, rtol=0)


@pytest.mark.parametrize("cfg", [
    dict(select_degrade_pixel="thresholding", ddpm_schedule="linear", degrade_channel="1-channel", mean_option="degraded_area", mean_area="image-wise"),
    dict(select_degrade_pixel="indexing", ddpm_schedule="log", mean_option="degraded_area", mean_area="image-wise"),
    dict(select_degrade_pixel="thresholding", ddpm_schedule="linear", degrade_channel="3-channel", mean_option="non_degraded_area", mean_area="channel-wise"),
])
@pytest.mark.parametrize("shape", [(64, 1, 32), (16, 3, 64), (7, 3, 24)])
def test_degrade_against_oracle(cfg, shape):
    B, C, S_ = shape
    if cfg.get("degrade_channel") == "3-channel" and C != 3:
        pytest.skip("3-channel masks need C == 3")
    a = default_args(data_size=S_, in_channel=C, out_channel=C, ddpm_num_steps=200, **cfg)
    S = make(a, 12)
    O = OracleScheduler(a, OracleRNG(12))
    Tp = O.update_ddpm_num_steps()
    g = torch.Generator().manual_seed(4)
    x0 = torch.rand(B, C, S_, S_, generator=g) * 2 - 1
    ts = torch.randint(1, Tp + 1, (B,), generator=g)
    ts[0], ts[-1] = 1, Tp
    ref = O.degrade_training(O.get_black_area_num_pixels_time(ts), x0, a.mean_option, a.mean_area)
    out = S.degrade_training(S.get_black_area_num_pixels_time(ts.cuda()), x0.cuda(), a.mean_option, a.mean_area)
    assert torch.equal(out[1].cpu(), ref[1].contiguous())
    for o, r in zip((out[0], out[2], out[3]), (ref[0], ref[2], ref[3])):
        assert torch.allclose(o.cpu(), r, atol=FILL_TOL, rtol=1e-5, equal_nan=True)    # NaN pattern kept (quirk q6);
        # rtol: non_degraded_area divides a ~HW-term sum by a handful of pixels, |fill| can reach ~50
    k, p = S.rng.export()
    ok, op = O.rng.state_words()
    assert p == op and np.array_equal(k, ok)


def test_nan_when_nothing_degraded():
    """quirk q6: thresholding + degraded_area with zero degraded pixels -> NaN image, as the reference."""
    a = default_args(data_size=8, ddpm_num_steps=10, select_degrade_pixel="thresholding", ddpm_schedule="linear")
    S = make(a, 0)
    x0 = torch.rand(2, 3, 8, 8).cuda()
    ratio = torch.tensor([0.0, 1.0], dtype=torch.float64).cuda()      # nothing / everything degraded
    d, m, _, _ = S.degrade_training(ratio, x0, "degraded_area", "image-wise")
    torch.manual_seed(0)
    u = torch.FloatTensor(2, 64).uniform_(0, 1)
    if (u[0] > 0).all():
        assert torch.isnan(d[0]).all()
    assert torch.allclose(d[1], x0[1].mean().expand_as(d[1]), atol=1e-6)


def test_full_size_properties():
    """BASELINE c4 shape (256 x 3 x 128 x 128): size-independent properties."""
    B, C, S_ = 256, 3, 128
    a = default_args(data_size=S_, ddpm_num_steps=1000, select_degrade_pixel="thresholding", ddpm_schedule="linear")
    S = make(a, 1)
    x0 = (torch.rand(B, C, S_, S_, device="cuda") * 2 - 1)
    ts = torch.randint(1, 1001, (B,), device="cuda")
    n = S.get_black_area_num_pixels_time(ts)
    d, m, dm, mean = S.degrade_training(n, x0, "degraded_area", "image-wise")
    assert set(m.unique().tolist()) <= {0.0, 1.0}
    assert torch.equal(d[m.bool()], x0[m.bool()])                               # kept pixels untouched
    fill = mean[:, :1, :1, :1]
    assert torch.equal(d[~m.bool()], fill.expand_as(d)[~m.bool()])              # degraded pixels == fill
    ref_fill = (x0 * (1 - m)).sum(dim=(1, 2, 3)) / (1 - m).sum(dim=(1, 2, 3))
    assert torch.allclose(fill.flatten(), ref_fill, atol=FILL_TOL, rtol=0)
    frac = 1 - m[:, 0].mean(dim=(1, 2))
    assert torch.allclose(frac.double(), n, atol=0.02)                          # masked fraction ~ ratio
    # idempotence: degrading the result with the same mask and its own fill changes nothing
    again = S.degrade_with_mask(d, m, "degraded_area", "image-wise")
    assert torch.allclose(again, d, atol=FILL_TOL, rtol=0)
    # words consumed: B*HW
    assert S.rng.words_drawn == B * S_ * S_


def test_uint8_feed_equals_the_cpu_transforms():
    """raw uint8 batches are normalised inside K1's read exactly like the reference's CPU pipeline
    (torchvision ToTensor: u / 255, then Normalize(0.5, 0.5): (x - 0.5) / 0.5, utils/mydataset.py:81) -- x_t, masks and the
    fp32 image handed to the loss are bit-identical to feeding the CPU-normalised fp32 batch."""
    from tests.golden.make_golden import mk_args
    g = torch.Generator().manual_seed(2)
    for shape in ((5, 3, 16, 16), (3, 1, 32, 32), (2, 3, 8, 8)):
        u8 = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g)
        x_ref = u8.to(torch.float32).div(255).sub_(0.5).div_(0.5)           # ToTensor + Normalize(0.5, 0.5)
        for opt, area in (("degraded_area", "image-wise"), ("non_degraded_area", "channel-wise"), (0, "image-wise")):
            a = mk_args(data_size=shape[2], in_channel=shape[1], out_channel=shape[1], ddpm_num_steps=50,
                        select_degrade_pixel="indexing", ddpm_schedule="log", mean_option=opt, mean_area=area)
            outs = []
            for x in (x_ref, u8):
                S = scheduler.Scheduler(a)
                Tp = S.update_ddpm_num_steps(50)
                ts = torch.tensor([1, Tp // 2, Tp][: shape[0]] + [3] * max(0, shape[0] - 3), device="cuda")
                torch.manual_seed(7)
                S.adopt_torch_rng("cuda")
                outs.append((S.degrade_training(S.get_black_area_num_pixels_time(ts), x.cuda(), opt, area), S))
            (f32, _), (raw, S_u8) = outs
            assert torch.equal(raw[0], f32[0]) and torch.equal(raw[1], f32[1]) and torch.equal(raw[3], f32[3])
            assert raw[0].dtype == torch.float32 and torch.equal(S_u8.x0_normalised.cpu(), x_ref)


def test_uint8_batches_through_run_batch():
    """the trainer keeps a uint8 batch raw (4x fewer host -> device bytes), K1 normalises it and hands the fp32 image to the
    loss: same loss and gradients as the CPU-normalised batch"""
    import trainer_masked
    from tests.golden.make_golden import FakeAccelerator, TinyNet, mk_args
    a = mk_args(data_size=16, ddpm_num_steps=100, select_degrade_pixel="indexing", ddpm_schedule="log",
                mean_option="degraded_area", mean_area="image-wise", method="base")
    a.use_ema, a.timeindex_rng = False, "cpu_stream"
    u8 = torch.randint(0, 256, (6, 3, 16, 16), dtype=torch.uint8, generator=torch.Generator().manual_seed(4))
    res = []
    for x in (u8.to(torch.float32).div(255).sub_(0.5).div_(0.5), u8):
        net = TinyNet(3).cuda()
        opt = torch.optim.SGD(net.parameters(), lr=0.0)
        opt.zero_grad = lambda *args, **kw: None
        tr = trainer_masked.Trainer(a, None, None, net, None, opt, torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0), FakeAccelerator())
        tr.prepare_schedule()
        tr.timesteps_used_epoch = tr.Scheduler.get_timesteps_epoch(0, 1)
        torch.manual_seed(41)
        tr.Scheduler.adopt_torch_rng("cuda")
        loss = tr._run_batch(0, (x.cuda(),), 0, 1, 0, None, None)[0]
        res.append((loss, net.conv.weight.grad.clone()))
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])


def test_single_pass_cluster_kernels_equal_two_pass(monkeypatch):
    """MDM_DEGRADE_FUSED=1: K1 and K5 as one cluster kernel per sample (sample in registers, partial sums through DSMEM).
    Same masks, same composite given the fill; the fill itself may differ by the summation order (<= 2e-6).  Measured
    slower than the two-pass default on B200, kept opt-in."""
    import sampler
    from tests.golden.make_golden import ToyModel, mk_args
    outs = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("MDM_DEGRADE_FUSED", fused)
        res = []
        for C, S, opt, area in ((3, 64, "degraded_area", "image-wise"), (3, 128, "non_degraded_area", "channel-wise"),
                                (1, 32, "degraded_area", "image-wise"), (3, 16, 0, "image-wise")):
            a = mk_args(data_size=S, in_channel=C, out_channel=C, ddpm_num_steps=50, select_degrade_pixel="indexing",
                        ddpm_schedule="log", mean_option=opt, mean_area=area, sample_num=3,
                        shift_type="1-d_constant", sample_latent_shape="uniform")
            Sch = scheduler.Scheduler(a)
            Tp = Sch.update_ddpm_num_steps(50)
            g = torch.Generator().manual_seed(3)
            x0 = (torch.rand(3, C, S, S, generator=g) * 2 - 1).cuda()
            ts = torch.tensor([2, Tp // 2, Tp], device="cuda")
            torch.manual_seed(5)
            Sch.adopt_torch_rng("cuda")
            d = Sch.degrade_training(Sch.get_black_area_num_pixels_time(ts), x0, opt, area)
            Sch.release_rng_to_torch()
            torch.manual_seed(6)
            s0, _ = sampler.Sampler(None, a, Sch, None).sample(ToyModel(Tp, "cuda"), Sch.get_timesteps_epoch(0, 1)[-6:])
            res.append((d[0].clone(), d[1].clone(), d[2].clone(), d[3].clone(), s0.clone()))
        outs[fused] = res
    for two, one in zip(outs["0"], outs["1"]):
        assert torch.equal(two[1], one[1])                                      # masks
        for k in (0, 2, 3):
            torch.testing.assert_close(one[k], two[k], atol=2e-6, rtol=0)       # composite, degrade mask, fill
        torch.testing.assert_close(one[4], two[4], atol=2e-5, rtol=0)           # six restoration steps
