"""GPU: the fused optimiser tail (csrc/optim.cu: global-norm clip + Adam / AdamW / SGD + EMA + bf16 mirror in one
pass) against what the reference runs op by op -- `accelerator.clip_grad_norm_(params, 1.0)`, `optimizer.step()`
(torch.optim.Adam / AdamW / SGD, main_train_masked.py:134-141) and `ema_model.step(params)` (diffusers EMAModel,
SURVEY.md C.2; trainer_masked.py:144-153) -- three consecutive steps on random flat buffers.

Tolerances: fp32 state (parameters, moments, EMA) rtol 2e-6 / atol 1e-8 per step chain (torch's foreach kernels use
lerp / addcdiv, the fused kernel the textbook form: differences are single roundings); the bf16 mirror must be the
exactly rounded fp32 parameter."""
import numpy as np
import pytest
import torch

from mdm_b200 import optim_ops

pytestmark = pytest.mark.gpu

N = 4 * 25000 + 8          # not a multiple of the block tile


def _ema_decay(step, power=0.75, inv_gamma=1.0, max_decay=0.9999):
    s = max(0, step - 1)
    if s <= 0:
        return 0.0
    return max(min(1 - (1 + s / inv_gamma) ** -power, max_decay), 0.0)


@pytest.mark.parametrize("mode", ["adam", "adamw", "sgd"])
@pytest.mark.parametrize("clip", [0.0, 1.0])
@pytest.mark.parametrize("use_ema", [False, True])
@pytest.mark.parametrize("entry", ["host_scalars", "device_scalars"])
def test_fused_step_matches_torch(mode, clip, use_ema, entry):
    g = torch.Generator(device="cuda").manual_seed(7)
    p0 = torch.randn(N, device="cuda", generator=g)
    lr, b1, b2, eps = 3e-3, 0.9, 0.999, 1e-8
    wd = 0.01 if mode == "adamw" else 0.0
    # reference: torch
    pr = torch.nn.Parameter(p0.clone())
    opt = {"adam": lambda: torch.optim.Adam([pr], lr=lr), "adamw": lambda: torch.optim.AdamW([pr], lr=lr),
           "sgd": lambda: torch.optim.SGD([pr], lr=lr)}[mode]()
    ema_r = p0.clone()
    # fused
    p = p0.clone()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    ema = p0.clone() if use_ema else None
    p16 = torch.empty(N, dtype=torch.bfloat16, device="cuda")
    ws = torch.empty(1024, device="cuda")
    gsq = torch.zeros(1, device="cuda")
    hyper = torch.zeros(4, device="cuda")
    for step in range(1, 4):
        # gradients large enough that clipping at 1.0 is active (norm ~ 0.05 * sqrt(N) = 16)
        grad = torch.randn(N, device="cuda", generator=g) * 0.05
        decay = _ema_decay(step) if use_ema else 0.0
        pr.grad = grad.clone()
        if clip > 0:
            total = torch.nn.utils.clip_grad_norm_([pr], clip)
            assert total.item() > clip                      # the clip really scales the gradient
        opt.step()
        if use_ema:
            ema_r.sub_((1 - decay) * (ema_r - pr.detach()))
        if clip > 0:
            optim_ops.grad_sumsq(grad, ws, gsq)
            assert abs(gsq.sqrt().item() - total.item()) <= 1e-5 * total.item()
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        if entry == "host_scalars":
            optim_ops.adam_ema_step(p, grad, m, v, ema, p16, gsq if clip > 0 else None, lr, b1, b2, eps, wd, bc1, bc2,
                                    clip, decay, 1.0, optim_ops.MODE[mode])
        else:
            hyper.copy_(torch.tensor([lr, lr / bc1, bc2 ** 0.5, decay]))
            optim_ops.adam_ema_step_dev(p, grad, m, v, ema, p16, gsq if clip > 0 else None, hyper, b1, b2, eps, wd,
                                        clip, 1.0, optim_ops.MODE[mode])
        torch.testing.assert_close(p, pr.detach(), rtol=2e-6, atol=1e-7)
        assert torch.equal(p16, p.to(torch.bfloat16))       # mirror = exactly rounded master
        if use_ema:
            torch.testing.assert_close(ema, ema_r, rtol=2e-6, atol=1e-7)
    if mode != "sgd":
        st = opt.state[pr]
        # |m| ~ 5e-3, v ~ 2.5e-6 after three steps; elements where beta * m and (1 - beta) * g cancel differ by one rounding
        # of the LARGER term (torch: lerp, here: the textbook form) -> absolute floors at 4e-6 of the typical magnitude
        torch.testing.assert_close(m, st["exp_avg"], rtol=2e-6, atol=2e-8)
        torch.testing.assert_close(v, st["exp_avg_sq"], rtol=2e-6, atol=1e-11)


def test_grad_scale_is_the_data_parallel_mean():
    """`grad_scale = 1 / world` folds the all-reduce MEAN into the kernel: clip sees the norm of the averaged gradient"""
    g = torch.Generator(device="cuda").manual_seed(3)
    p0 = torch.randn(N, device="cuda", generator=g)
    grad_sum = torch.randn(N, device="cuda", generator=g) * 0.2      # what a SUM all-reduce over 4 ranks leaves
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([pr], lr=1e-3)
    pr.grad = grad_sum / 4
    torch.nn.utils.clip_grad_norm_([pr], 1.0)
    opt.step()
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    ws, gsq = torch.empty(1024, device="cuda"), torch.zeros(1, device="cuda")
    optim_ops.grad_sumsq(grad_sum, ws, gsq)
    optim_ops.adam_ema_step(p, grad_sum, m, v, None, None, gsq, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1 - 0.9, 1 - 0.999, 1.0, 0.0,
                            0.25, optim_ops.MODE["adamw"])
    torch.testing.assert_close(p, pr.detach(), rtol=2e-6, atol=1e-7)


def test_fused_optimizer_object_follows_reference_sequence():
    """FusedOptimizer + EMAModel (attach_ema) + LambdaSchedule driven the way `_run_batch` drives them, against
    torch.optim.AdamW + clip_grad_norm_ + the EMA rule + diffusers' cosine warm-up, on the real flat parameter buffer"""
    from mdm_b200.denoiser import UNet2DModelB200
    from mdm_b200.runtime import EMAModel, FusedOptimizer, get_scheduler
    cfg = dict(block_out_channels=[128, 128], down_block_types=["DownBlock2D", "AttnDownBlock2D"],
               up_block_types=["AttnUpBlock2D", "UpBlock2D"], sample_size=16)
    model = UNet2DModelB200(device="cuda", **cfg)
    model.reset_parameters(seed=1)
    ema = EMAModel(model.parameters(), decay=0.9999, use_ema_warmup=True, inv_gamma=1.0, power=0.75)
    opt = FusedOptimizer(model, "adamw", lr=1e-3)
    opt.attach_ema(ema)
    sched = get_scheduler("cosine", opt, num_warmup_steps=2, num_training_steps=10)
    pr = torch.nn.Parameter(model.flat_param.detach().clone())
    ropt = torch.optim.AdamW([pr], lr=1e-3)
    rsched = torch.optim.lr_scheduler.LambdaLR(ropt, lambda s: sched.fn(s, 2, 10, 0.5))
    ema_r = pr.detach().clone()
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(1, 5):
        grad = torch.randn(model.numel_flat, device="cuda", generator=g) * 1e-3
        model.flat_grad.copy_(grad)
        pr.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([pr], 1.0)
        ropt.step(); rsched.step(); ropt.zero_grad()
        ema_r.sub_((1 - _ema_decay(step)) * (ema_r - pr.detach()))
        opt.set_clip(1.0); opt.step(); sched.step(); opt.zero_grad(); ema.step(model.parameters())
        assert abs(sched.get_last_lr()[0] - rsched.get_last_lr()[0]) < 1e-12
        torch.testing.assert_close(model.flat_param, pr.detach(), rtol=2e-6, atol=1e-7)
        torch.testing.assert_close(ema.flat, ema_r, rtol=2e-6, atol=1e-7)
        assert torch.equal(model.flat_bf16, model.flat_param.to(torch.bfloat16)) and not model.params_dirty()
        assert float(model.flat_grad.abs().max()) == 0.0
    assert ema.optimization_step == 4 and opt.step_count == 4
