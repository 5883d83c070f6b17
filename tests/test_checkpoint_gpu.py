"""GPU: checkpoint save -> fresh objects -> load -> resume (SURVEY.md sections 5.4, 8 f2).

The reference saves through accelerate's `save_state` with the pre-hooks of main_train_masked.py:195-225 (the
denoiser and its EMA go to `unet/` and `unet_ema/` in the diffusers layout, accelerate adds optimizer / scheduler /
RNG files) and resumes through `load_state` (:250-277).  The drop-in's `get_accelerator` registers the same hooks on
`mdm_b200.runtime.Accelerator`.  Checked here: the exact file set, diffusers key names and NCHW fp32 shapes, EMA fields
in `unet_ema/config.json`, bit-identical parameters / EMA / moments / step counters / LR schedule / CPU generator state
after loading into freshly built objects that had ALREADY captured their step graph with other weights (so a stale
bf16 mirror would show), and that the next training step after the resume reproduces the uninterrupted run."""
import json
import os

import numpy as np
import pytest
import torch

from mdm_b200.config import default_args

pytestmark = pytest.mark.gpu

SMALL = dict(block_out_channels=[128, 128, 256], layers_per_block=1, sample_size=16,
             down_block_types=["DownBlock2D", "DownBlock2D", "AttnDownBlock2D"],
             up_block_types=["AttnUpBlock2D", "UpBlock2D", "UpBlock2D"])


def _build(seed):
    import main_train_masked as M
    import trainer_masked
    from mdm_b200.denoiser import UNet2DModelB200
    from mdm_b200.runtime import get_scheduler
    a = default_args(data_size=16, ddpm_num_steps=50, select_degrade_pixel="indexing", ddpm_schedule="log",   # the reference's
                     mean_option="degraded_area", mean_area="image-wise", method="base", mixed_precision="bf16")  # flag namespace
    a.use_ema, a.cuda_graph, a.timeindex_rng = True, True, "cpu_stream"
    model = UNet2DModelB200(device="cuda", **SMALL)
    model.reset_parameters(seed=seed)
    ema = M.get_ema(a, model)
    acc = M.get_accelerator(a, ema)
    M.get_weight_type(a, acc)
    opt = M.get_optimizer(model, "adamw", 1e-3)
    sched = get_scheduler("cosine", opt, num_warmup_steps=2, num_training_steps=50)
    model, opt, sched = acc.prepare(model, opt, sched)
    tr = trainer_masked.Trainer(a, None, None, model, ema, opt, sched, acc)
    tr.prepare_schedule()
    tr.timesteps_used_epoch = tr.Scheduler.get_timesteps_epoch(0, 1)
    return a, model, ema, opt, sched, acc, tr


def _steps(tr, x0, n, first=0):
    tr.Scheduler.adopt_torch_rng("cuda")
    out = [tr._run_batch(first + i, (x0,), 0, 1, 0, None, None)[0] for i in range(n)]
    tr.Scheduler.release_rng_to_torch()
    return out


def test_save_load_resume_round_trip(tmp_path):
    g = torch.Generator().manual_seed(9)
    x0 = (torch.rand(8, 3, 16, 16, generator=g) * 2 - 1).cuda()
    xb = (torch.rand(8, 3, 16, 16, generator=g) * 2 - 1).cuda()
    # ---- run A: 4 steps (2 eager, capture, replay), checkpoint, one more step --------------------------------
    torch.manual_seed(123)
    a, model, ema, opt, sched, acc, tr = _build(seed=1)
    _steps(tr, x0, 4)
    ck = str(tmp_path / "checkpoint-epoch-3")
    acc.save_state(ck)
    snap = dict(p=model.flat_param.clone(), e=ema.flat.clone(), m=opt.m.clone(), v=opt.v.clone(), step=opt.step_count,
                ema_step=ema.optimization_step, last_epoch=sched.last_epoch, lr=sched.get_last_lr(),
                rng=torch.get_rng_state().clone())
    loss_next = _steps(tr, x0, 1, first=4)[0]
    p_next = model.flat_param.clone()
    # ---- the exact file set ---------------------------------------------------------------
    files = sorted(os.path.relpath(os.path.join(d, f), ck) for d, _, fs in os.walk(ck) for f in fs)
    assert files == sorted(["unet/config.json", "unet/diffusion_pytorch_model.safetensors", "unet_ema/config.json",
                            "unet_ema/diffusion_pytorch_model.safetensors", "optimizer.bin", "scheduler.bin",
                            "random_states_0.pkl"]), files
    from safetensors.torch import load_file
    sd = load_file(os.path.join(ck, "unet", "diffusion_pytorch_model.safetensors"))
    assert sd["conv_in.weight"].shape == (128, 3, 3, 3) and sd["conv_in.weight"].dtype == torch.float32
    assert sd["down_blocks.0.resnets.0.conv1.weight"].shape == (128, 128, 3, 3)          # NCHW, not the packed layout
    assert sd["down_blocks.2.attentions.0.to_q.weight"].shape == (256, 256)
    assert sd["up_blocks.0.resnets.0.conv_shortcut.weight"].shape == (256, 512, 1, 1)
    assert "mid_block.attentions.0.to_out.0.bias" in sd and "down_blocks.0.resnets.0.time_emb_proj.weight" in sd
    cfg = json.load(open(os.path.join(ck, "unet", "config.json")))
    assert cfg["_class_name"] == "UNet2DModel" and cfg["block_out_channels"] == [128, 128, 256]
    ecfg = json.load(open(os.path.join(ck, "unet_ema", "config.json")))
    for k in ("decay", "min_decay", "optimization_step", "update_after_step", "use_ema_warmup", "inv_gamma", "power"):
        assert k in ecfg, k
    assert ecfg["optimization_step"] == 4 and ecfg["use_ema_warmup"] is True
    rs = torch.load(os.path.join(ck, "random_states_0.pkl"), weights_only=False)     # torch.save, like accelerate
    assert {"random_state", "numpy_random_seed", "torch_manual_seed", "torch_cuda_manual_seed"} <= set(rs)
    # ---- run B: other weights, other data, graph already captured; then load ------------------------------
    torch.manual_seed(77)
    a2, model2, ema2, opt2, sched2, acc2, tr2 = _build(seed=2)
    _steps(tr2, xb, 4)
    assert not torch.equal(model2.flat_param, snap["p"])
    acc2.load_state(ck)
    assert torch.equal(model2.flat_param, snap["p"]) and torch.equal(ema2.flat, snap["e"])
    assert torch.equal(opt2.m, snap["m"]) and torch.equal(opt2.v, snap["v"]) and opt2.step_count == snap["step"]
    assert ema2.optimization_step == snap["ema_step"] and abs(ema2.decay - 0.9999) < 1e-12 and ema2.power == 0.75
    assert sched2.last_epoch == snap["last_epoch"] and sched2.get_last_lr() == snap["lr"]
    assert opt2.param_groups[0]["lr"] == snap["lr"][0]
    assert torch.equal(torch.get_rng_state(), snap["rng"])
    assert torch.equal(model2.flat_bf16, model2.flat_param.to(torch.bfloat16))          # mirror follows the load
    # ---- the resumed step reproduces the uninterrupted one (graph replay on the loaded weights) -----------------
    loss2 = _steps(tr2, x0, 1, first=4)[0]
    assert abs(loss2 - loss_next) <= 2e-3 * abs(loss_next), (loss2, loss_next)
    rel = ((model2.flat_param - p_next).norm() / (p_next - snap["p"]).norm()).item()
    assert rel < 5e-2, rel           # same update up to the fp32 atomics' summation order


def test_pretrained_round_trip_and_torch_layout_refused(tmp_path):
    from mdm_b200.denoiser import UNet2DModelB200
    from mdm_b200.runtime import EMAModel, FusedOptimizer
    model = UNet2DModelB200(device="cuda", **SMALL)
    model.reset_parameters(seed=4)
    model.save_pretrained(str(tmp_path / "unet"))
    back = UNet2DModelB200.from_pretrained(str(tmp_path), subfolder="unet")
    assert torch.equal(back.flat_param, model.flat_param) and back._cfg["block_out_channels"] == [128, 128, 256]
    ema = EMAModel(model.parameters(), decay=0.99, use_ema_warmup=True, power=0.75, model_cls=UNet2DModelB200,
                   model_config=model.config)
    ema.flat.mul_(0.5)
    ema.optimization_step = 17
    ema.save_pretrained(str(tmp_path / "unet_ema"))
    e2 = EMAModel.from_pretrained(str(tmp_path / "unet_ema"), UNet2DModelB200)
    assert torch.equal(e2.flat, ema.flat) and e2.optimization_step == 17 and e2.decay == 0.99
    opt = FusedOptimizer(model, "adamw")
    with pytest.raises(RuntimeError, match="torch.optim"):
        opt.load_state_dict({"state": {0: {"step": 1}}, "param_groups": [{}]})
