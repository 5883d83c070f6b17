"""Flag namespace of the reference (`main_train_masked.py:347-417`) without argparse, for programmatic
use (bench.py, tests), and the analytic FLOP count of the denoiser."""
from __future__ import annotations

from types import SimpleNamespace

import torch


def default_args(**kw):
    """Namespace with the reference's flag names and defaults; keyword arguments override."""
    d = dict(
        data_size=64, in_channel=3, out_channel=3, batch_size=128, num_attention=1, model="default", method="base",
        optim="adamw", lr=1e-4, lr_scheduler="linear", lr_warmup_steps=500, lr_cycle=0.5, num_epochs=1000,
        gradient_accumulation_steps=1, mixed_precision="no", use_ema=True, ema_inv_gamma=1.0, ema_power=0.75,
        ema_max_decay=0.9999, loss_weight_use=False, loss_weight_power_base=10.0, loss_space="x_0",
        ddpm_num_steps=1000, updated_ddpm_num_steps=1000, ddpm_schedule="linear", ddpm_schedule_base=10.0,
        scheduler_num_scale_timesteps=1, select_degrade_pixel="indexing", degrade_channel=None, mean_option=0,
        mean_area="image-wise", mean_value_accumulate=False, shift_type="noise_with_perturbation", noise_mean=0.0,
        sample_latent_shape="data", sampling="base", momentum_adaptive="base_momentum", adaptive_decay_rate=0.999,
        adaptive_momentum_rate=0.9, sampling_mask_dependency="independent", sample_num=100, sample_epoch_ratio=0.2,
        resume_from_checkpoint="False", num_workers=32, checkpointing_steps=500, save_images_epochs=10, output_dir=None,
        dir_dataset="/synthetic", data_name="synthetic", weight_dtype=torch.float32,
    )
    d.update(kw)
    return SimpleNamespace(**d)


def unet_forward_flops(cfg: dict, S: int) -> float:
    """forward FLOPs (2*MAC) per image of the UNet2DModel configuration `cfg` at S x S: every 3x3 / 1x1
    convolution, every Linear (time embedding, time_emb_proj, attention projections) and the attention
    matmuls.  Matches the layer-by-layer enumeration of SURVEY.md section 8d / Appendix A."""
    boc = list(cfg["block_out_channels"])
    n = cfg["layers_per_block"]
    temb = boc[0] * 4
    cin_img, cout_img = cfg["in_channels"], cfg["out_channels"]
    total = 0.0

    def conv(res, ci, co, k):
        nonlocal total
        total += 2.0 * res * res * co * ci * k * k

    def linear(rows, ci, co):
        nonlocal total
        total += 2.0 * rows * ci * co

    def resnet(res, ci, co):
        conv(res, ci, co, 3)
        conv(res, co, co, 3)
        linear(1, temb, co)
        if ci != co:
            conv(res, ci, co, 1)

    def attn(res, c):
        L = res * res
        linear(L, c, 3 * c)
        linear(L, c, c)
        nonlocal total
        total += 2.0 * 2.0 * L * L * c

    linear(1, boc[0], temb)
    linear(1, temb, temb)
    conv(S, cin_img, boc[0], 3)
    res, c = S, boc[0]
    skips = [c]
    for i, ty in enumerate(cfg["down_block_types"]):
        ci, c = c, boc[i]
        for j in range(n):
            resnet(res, ci if j == 0 else c, c)
            if ty.startswith("Attn"):
                attn(res, c)
            skips.append(c)
        if i != len(boc) - 1:
            res //= 2
            conv(res, c, c, 3)
            skips.append(c)
    resnet(res, c, c)
    attn(res, c)
    resnet(res, c, c)
    rev = list(reversed(boc))
    for i, ty in enumerate(cfg["up_block_types"]):
        cprev, c = c, rev[i]
        for j in range(n + 1):
            s = skips.pop()
            resnet(res, (cprev if j == 0 else c) + s, c)
            if ty.startswith("Attn"):
                attn(res, c)
        if i != len(boc) - 1:
            res *= 2
            conv(res, c, c, 3)
    conv(S, boc[0], cout_img, 3)
    return total
