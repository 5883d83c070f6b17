"""CPU: host-side runtime pieces the reference takes from diffusers / accelerate, the
analytic FLOP count, the flag namespace, and that the product path refuses to run without a GPU."""
import math

import pytest
import torch

from mdm_b200.config import default_args, unet_forward_flops
from mdm_b200.denoiser import default_config
from mdm_b200.runtime import EMAModel, SCHEDULES, get_scheduler


def test_flops_match_survey_and_oracle():
    from oracle.unet_ref import fwd_flops_per_image, unet_config
    for C, S, want in ((1, 32, 7.76e9), (3, 32, 7.77e9), (3, 64, 31.03e9), (3, 128, 124.14e9)):
        got = unet_forward_flops(default_config(C, S), S)
        assert abs(got - want) / want < 2e-3
        assert got == fwd_flops_per_image(unet_config(C, S))                   # layer-by-layer hook count of the oracle
    assert abs(unet_forward_flops(default_config(3, 256, base=256), 256) - 1984.94e9) / 1984.94e9 < 2e-3   # c5


def test_lr_schedules_match_diffusers_formulas():
    class Opt:
        param_groups = [dict(lr=2.0)]
    w, total = 10, 110
    for name in SCHEDULES:
        s = get_scheduler(name, Opt(), num_warmup_steps=w, num_training_steps=total, num_cycles=0.5)
        assert s.get_last_lr()[0] == 0.0
        for step in range(1, 130):
            s.step()
            if step < w:
                want = step / w
            elif name == "constant":
                want = 1.0
            elif name == "linear":
                want = max(0.0, (total - step) / (total - w))
            elif name == "cosine":
                p = (step - w) / (total - w)
                want = max(0.0, 0.5 * (1 + math.cos(math.pi * 2 * 0.5 * p)))
            else:
                p = (step - w) / (total - w)
                want = 0.0 if p >= 1 else max(0.0, 0.5 * (1 + math.cos(math.pi * ((0.5 * p) % 1.0))))
            assert abs(s.get_last_lr()[0] - 2.0 * want) < 1e-12, (name, step)
    s = get_scheduler("linear", Opt(), 0, 100)
    s.steps_per_call = 4                                                       # accelerate: num_processes steps per call
    s.step()
    assert s.last_epoch == 4
    sd = s.state_dict()
    s2 = get_scheduler("linear", Opt(), 0, 100)
    s2.load_state_dict(sd)
    assert s2.last_epoch == 4 and s2.get_last_lr() == s.get_last_lr()


def test_ema_decay_and_update_rule():
    p = [torch.nn.Parameter(torch.ones(4))]
    ema = EMAModel(p, decay=0.9999, use_ema_warmup=True, inv_gamma=1.0, power=0.75)
    assert ema.get_decay(1) == 0.0
    assert abs(ema.get_decay(2) - (1 - (1 + 1) ** -0.75)) < 1e-12
    assert ema.get_decay(10 ** 9) == 0.9999
    ema2 = EMAModel(p, decay=0.5, use_ema_warmup=False)
    assert abs(ema2.get_decay(3) - min(0.5, 3 / 12)) < 1e-12
    with torch.no_grad():
        p[0].fill_(3.0)
    ema.step(p)                      # step 1: decay 0 -> shadow = param
    assert torch.equal(ema.shadow_params[0], torch.full((4,), 3.0))
    with torch.no_grad():
        p[0].fill_(5.0)
    ema.step(p)
    d = ema.get_decay(2)
    assert torch.allclose(ema.shadow_params[0], torch.full((4,), 3.0 - (1 - d) * (3.0 - 5.0)))
    ema.store(p)
    ema.copy_to(p)
    assert torch.equal(p[0].data, ema.shadow_params[0])
    ema.restore(p)
    assert torch.equal(p[0].data, torch.full((4,), 5.0))


def test_flag_namespace_matches_reference_parser():
    import main_train_masked as M
    ns = vars(M.build_parser().parse_args([]))
    d = vars(default_args())
    for k in ("ddpm_num_steps", "batch_size", "lr", "ema_power", "ema_max_decay", "shift_type", "momentum_adaptive",
              "sampling_mask_dependency", "sample_num", "scheduler_num_scale_timesteps", "gradient_accumulation_steps"):
        assert ns[k] == d[k], k                                                # main_train_masked.py:347-417 defaults
    assert ns["use_ema"] is True and ns["optim"] == "adamw" and ns["mean_area"] == "image-wise"


def test_no_cpu_fallback():
    from mdm_b200.denoiser import UNet2DModelB200
    with pytest.raises(RuntimeError):
        UNet2DModelB200(device="cpu")
    import scheduler
    a = default_args(data_size=8, ddpm_num_steps=10, ddpm_schedule="linear", select_degrade_pixel="thresholding",
                     degrade_channel="1-channel")
    S = scheduler.Scheduler(a)
    S.update_ddpm_num_steps(10)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            S.degrade_training(torch.tensor([0.5]), torch.zeros(1, 3, 8, 8), 0, "image-wise")


def test_complement_of_gradient_ranges():
    """host logic of the segmented all-reduce: the last segment reduces everything the earlier ranges did not cover"""
    from trainer_masked import _complement
    assert _complement([(10, 20), (5, 10)], 30) == [(0, 5), (20, 30)]
    assert _complement([(0, 30)], 30) == []
    assert _complement([], 7) == [(0, 7)]
    assert _complement([(3, 5), (8, 9)], 9) == [(0, 3), (5, 8)]
