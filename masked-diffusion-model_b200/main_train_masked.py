"""Drop-in for the reference's entry point `main_train_masked.py` (/root/reference/code/
main_train_masked.py:46-448): same flags (:347-417), same factory functions and the same
checkpoint layout (`checkpoint-epoch-E/{unet,unet_ema}/...` + optimizer/scheduler/RNG state), on the
B200 runtime (`mdm_b200.runtime`: Accelerator / EMAModel / LR schedules / fused optimiser -- the
reference takes these from `accelerate` and `diffusers`, which are third-party and absent here).

    torchrun --nproc-per-node 8 main_train_masked.py --data_name synthetic --data_size 32 --method base \
        --select_degrade_pixel indexing --ddpm_schedule log --mixed_precision bf16 ...

Differences: the import of the missing `trainer_masked_mean_shift_v2` (q1) is
dropped; `--resume_from_checkpoint latest` parses `checkpoint-epoch-<E>` correctly (q17);
`--data_name synthetic` provides uniform [-1, 1] images when no dataset directory exists (datasets,
wandb and the `test` method are out of scope)."""
from __future__ import annotations

import argparse
import json
import math
import os
import random
from datetime import timedelta

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

import utils.dirutils as dirutils
import utils.model as models
from mdm_b200.denoiser import UNet2DModelB200 as UNet2DModel
from mdm_b200.runtime import Accelerator, EMAModel, FusedOptimizer, InitProcessGroupKwargs, get_scheduler
from trainer_masked import Trainer as BaseTrainer
from trainer_masked_mean_shift import Trainer as MeanShiftTrainer


class SyntheticDataset(Dataset):
    """uniform [-1, 1] images with the batch contract of `utils/mydataset.py:278` -> (data, label, random)"""

    def __init__(self, num, channels, size, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.data = torch.rand(num, channels, size, size, generator=g) * 2 - 1

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        return self.data[i], 0, 0


def compute_mean_histogram(dataset, args):
    """main_train_masked.py:60-87: histogram of per-image (or per-channel) means -> initial latent"""
    if args.sample_latent_shape.lower() != 'data':
        return [None, None, None]
    means = []
    for batch in DataLoader(dataset, batch_size=100, drop_last=False, shuffle=False, num_workers=0):
        d = batch[0]
        if args.mean_area == 'channel-wise':
            means.append(torch.mean(d, dim=[2, 3]))
        else:
            means.append(torch.mean(d, dim=[1, 2, 3]).unsqueeze(-1))
    stats = torch.cat(means, 0)
    hist, edges = torch.histogramdd(stats, bins=args.sample_num, density=True)
    shape = hist.shape
    hist = torch.ravel(hist)
    hist = hist / torch.sum(hist)
    return [shape, edges, torch.cumsum(hist, dim=0)]


def get_dataset(data_path, data_name, data_set, data_height, data_width, data_subset, data_subset_num, method, args=None):
    if data_name != "synthetic":
        raise NotImplementedError("dataset loading is out of scope of the B200 hot path (SURVEY.md section 2, rows 11-12): "
                                  "use --data_name synthetic or pass your own Dataset to main()")
    n = data_subset_num if data_subset else 4096
    dataset = SyntheticDataset(n, args.in_channel, data_height)
    return dataset, compute_mean_histogram(dataset, args)


def get_dataloader(dataset, batch_size, num_workers):
    return DataLoader(dataset=dataset, batch_size=batch_size, drop_last=True, shuffle=True, pin_memory=True, num_workers=0)


def get_model(args):
    if args.model == 'default':
        return models.MyModel(dim_channel=args.in_channel, dim_height=args.data_size, dim_width=args.data_size,
                              num_attention=args.num_attention)
    return UNet2DModel.from_config(UNet2DModel.load_config(args.model))


def get_ema(args, model):
    if not args.use_ema:
        return None
    return EMAModel(model.parameters(), decay=args.ema_max_decay, use_ema_warmup=True, inv_gamma=args.ema_inv_gamma,
                    power=args.ema_power, model_cls=UNet2DModel, model_config=model.config)


def get_optimizer(model, optim_name, lr):
    name = optim_name.lower()
    if name not in ("sgd", "adam", "adamw"):
        raise ValueError(optim_name)
    return FusedOptimizer(model, name, lr=lr)


def get_lr_scheduler(scheduler_name, optimizer, dataloader, lr_warmup_steps, gradient_accumulation_steps, num_epochs, num_cycles):
    return get_scheduler(scheduler_name, optimizer, num_warmup_steps=lr_warmup_steps * gradient_accumulation_steps,
                         num_training_steps=len(dataloader) * num_epochs, num_cycles=num_cycles)


def get_accelerator(args, ema_model):
    accelerator = Accelerator(gradient_accumulation_steps=args.gradient_accumulation_steps, mixed_precision=args.mixed_precision,
                              kwargs_handlers=[InitProcessGroupKwargs(timeout=timedelta(seconds=7200))])

    def save_model_hook(models_, weights, output_dir):
        if accelerator.is_main_process:
            if args.use_ema:
                ema_model.save_pretrained(os.path.join(output_dir, "unet_ema"))
            for model in models_:
                model.save_pretrained(os.path.join(output_dir, "unet"))
                weights.pop()

    def load_model_hook(models_, input_dir):
        if args.use_ema:
            loaded = EMAModel.from_pretrained(os.path.join(input_dir, "unet_ema"), UNet2DModel)
            ema_model.load_state_dict(loaded.state_dict())
            ema_model.to(accelerator.device)
        for _ in range(len(models_)):
            model = models_.pop()
            loaded = UNet2DModel.from_pretrained(input_dir, subfolder="unet")
            model.load_state_dict(loaded.state_dict())

    accelerator.register_save_state_pre_hook(save_model_hook)
    accelerator.register_load_state_pre_hook(load_model_hook)
    return accelerator


def get_weight_type(args, accelerator):
    weight_dtype = torch.float32
    if accelerator.mixed_precision == "fp16":
        raise NotImplementedError("mixed_precision fp16 is not on the B200 path (bf16 activations, fp32 master weights)")
    if accelerator.mixed_precision == "bf16":
        weight_dtype = torch.bfloat16
        args.mixed_precision = accelerator.mixed_precision
    args.weight_dtype = weight_dtype


def resume_train(args, accelerator, num_update_steps_per_epoch, dirs):
    global_step, first_epoch, resume_step = 0, 0, 0
    if args.resume_from_checkpoint != "latest":
        path = args.resume_from_checkpoint
    else:
        root = args.output_dir or dirs.list_dir['checkpoint']
        cands = [d for d in os.listdir(root) if d.startswith("checkpoint")]
        cands = sorted(cands, key=lambda x: int(x.split("-")[-1]))           # q17: "checkpoint-epoch-<E>"
        path = os.path.join(root, cands[-1]) if cands else None
    if path is None or not os.path.isdir(path):
        accelerator.print(f"Checkpoint '{args.resume_from_checkpoint}' does not exist. Starting a new training run.")
        args.resume_from_checkpoint = None
        return global_step, first_epoch, resume_step
    accelerator.print(f"Resuming from checkpoint {path}")
    accelerator.load_state(path)
    global_step = int(os.path.basename(path).split("-")[-1])
    resume_global_step = global_step * args.gradient_accumulation_steps
    first_epoch = global_step // num_update_steps_per_epoch
    resume_step = resume_global_step % (num_update_steps_per_epoch * args.gradient_accumulation_steps)
    return global_step, first_epoch, resume_step


def main(dirs, args, dataset=None, dataset_hist=None):
    if dataset is None:
        dataset, dataset_hist = get_dataset(args.dir_dataset, args.data_name, args.data_set, args.data_size, args.data_size,
                                            args.data_subset, args.data_subset_num, args.method, args)
    dataloader = get_dataloader(dataset, args.batch_size, args.num_workers)
    accelerator_device_ready = torch.cuda.is_available()
    if not accelerator_device_ready:
        raise RuntimeError("main_train_masked: a CUDA device is required (no CPU fallback)")
    if "LOCAL_RANK" in os.environ:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    model = get_model(args)
    ema_model = get_ema(args, model)
    accelerator = get_accelerator(args, ema_model)
    get_weight_type(args, accelerator)
    optimizer = get_optimizer(model, args.optim, args.lr)
    lr_scheduler = get_lr_scheduler(args.lr_scheduler, optimizer, dataloader, args.lr_warmup_steps,
                                    args.gradient_accumulation_steps, args.num_epochs, args.lr_cycle)
    model, optimizer, dataloader, lr_scheduler = accelerator.prepare(model, optimizer, dataloader, lr_scheduler)
    if args.use_ema:
        ema_model.to(accelerator.device)
    num_update_steps_per_epoch = math.ceil(len(dataloader) / args.gradient_accumulation_steps)
    if args.resume_from_checkpoint != 'False':
        global_step, first_epoch, resume_step = resume_train(args, accelerator, num_update_steps_per_epoch, dirs)
    else:
        global_step, first_epoch, resume_step = 0, 0, 0
    method = args.method.lower()
    if method == 'base':
        trainer = BaseTrainer(args, dataloader, dataset, model, ema_model, optimizer, lr_scheduler, accelerator)
    elif method == 'mean_shift':
        trainer = MeanShiftTrainer(args, dataloader, dataset, dataset_hist, model, ema_model, optimizer, lr_scheduler, accelerator)
    else:
        raise NotImplementedError(f"method '{args.method}' is out of scope of the B200 hot path")
    trainer.train(first_epoch, args.num_epochs, resume_step, global_step, dirs, None)
    return trainer


def save_option(args, dir_save):
    with open(os.path.join(dir_save, 'option.ini'), 'w') as f:
        json.dump({k: (str(v) if isinstance(v, torch.dtype) else v) for k, v in args.__dict__.items()}, f, indent=2)


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--use_wandb', type=eval, default=False, choices=[True, False])
    parser.add_argument('--use_mlflow', type=eval, default=False, choices=[True, False])
    parser.add_argument('--task', type=str, choices=['train', 'sample', 'dataset'], default='train')
    parser.add_argument('--content', type=str, default='test_code')
    parser.add_argument('--dir_work', type=str, default='./')
    parser.add_argument('--dir_dataset', type=str, default='/nas2/dataset')
    parser.add_argument('--data_name', type=str, default='mnist')
    parser.add_argument('--data_set', type=str, default='train')
    parser.add_argument('--data_size', type=int, default=64)
    parser.add_argument('--data_subset', type=eval, default=False)
    parser.add_argument('--data_subset_num', type=int, default=1000)
    parser.add_argument('--date', type=str, default='')
    parser.add_argument('--time', type=str, default='')
    parser.add_argument('--wandb_name', type=str, default='diffusion')
    parser.add_argument('--method', type=str, default='base')
    parser.add_argument('--test_method', type=str, default='base')
    parser.add_argument('--title', type=str, default='')
    parser.add_argument('--model', type=str, default='default')
    parser.add_argument('--batch_size', type=int, default=128)
    parser.add_argument('--in_channel', type=int, default=3)
    parser.add_argument('--out_channel', type=int, default=3)
    parser.add_argument('--num_attention', type=int, default=1)
    parser.add_argument('--num_epochs', type=int, default=1000)
    parser.add_argument('--optim', type=str, choices=(['adam', 'adamw', 'sgd']), default='adamw')
    parser.add_argument('--lr', type=float, default=1e-4)
    parser.add_argument('--lr_scheduler', type=str, default='linear')
    parser.add_argument('--lr_warmup_steps', type=int, default=500)
    parser.add_argument('--lr_cycle', type=float, default=0.5)
    parser.add_argument('--gradient_accumulation_steps', type=int, default=1)
    parser.add_argument('--mixed_precision', type=str, default="no", choices=["no", "fp16", "bf16"])
    parser.add_argument('--use_ema', type=eval, default=True, choices=[True, False])
    parser.add_argument('--ema_inv_gamma', type=float, default=1.0)
    parser.add_argument('--ema_power', type=float, default=3 / 4)
    parser.add_argument('--ema_max_decay', type=float, default=0.9999)
    parser.add_argument('--loss_weight_use', type=eval, default=False)
    parser.add_argument('--loss_weight_power_base', type=float, default=10.0)
    parser.add_argument('--loss_space', type=str, default='x_0')
    parser.add_argument('--ddpm_num_steps', type=int, default=1000)
    parser.add_argument('--updated_ddpm_num_steps', type=int, default=1000)
    parser.add_argument("--ddpm_schedule", type=str, default="linear")
    parser.add_argument("--ddpm_schedule_base", type=float, default=10.0)
    parser.add_argument('--scheduler_num_scale_timesteps', type=int, default=1)
    parser.add_argument('--select_degrade_pixel', default='indexing')
    parser.add_argument('--degrade_channel', type=str)
    parser.add_argument('--mean_option', default=0)
    parser.add_argument('--mean_area', default='image-wise', choices=['channel-wise', 'image-wise'])
    parser.add_argument('--mean_value_accumulate', type=eval, default=False, choices=[True, False])
    parser.add_argument('--shift_type', type=str, default='noise_with_perturbation',
                        choices=['1-d_constant', '3-d_constant', 'noise_reduction', 'noise_std_reduction',
                                 'noise_with_perturbation', 'non_shift'])
    parser.add_argument('--noise_mean', type=float, default=0)
    parser.add_argument("--sample_latent_shape", type=str, default="data", choices=['data', 'zero', 'normal', 'uniform', 'grid'])
    parser.add_argument("--sampling", type=str, default="base")
    parser.add_argument("--momentum_adaptive", type=str, default="base_momentum",
                        choices=['base_momentum', 'base_sampling', 'momentum', 'boosting'])
    parser.add_argument('--adaptive_decay_rate', type=float, default=0.999)
    parser.add_argument('--adaptive_momentum_rate', type=float, default=0.9)
    parser.add_argument("--sampling_mask_dependency", type=str, default="independent",
                        choices=['dependent_prev', 'independent', 'dependent_t'])
    parser.add_argument('--sample_num', type=int, default=100)
    parser.add_argument('--sample_epoch_ratio', type=float, default=0.2)
    parser.add_argument('--resume_from_checkpoint', default="False")
    parser.add_argument('--num_workers', type=int, default=32)
    parser.add_argument("--checkpointing_steps", type=int, default=500)
    parser.add_argument("--save_images_epochs", type=int, default=10)
    parser.add_argument("--output_dir", type=str, default=None)
    parser.add_argument("--test_model_path", type=str, default=None)
    return parser


if __name__ == '__main__':
    args = build_parser().parse_args()
    dirs = dirutils.Dir(task=args.task, content=args.content, dir_work=args.dir_work, dir_dataset=args.dir_dataset,
                        data_name=args.data_name, data_set=args.data_set, data_size=args.data_size, date=args.date,
                        time=args.time, method=args.method, title=args.title)
    torch.manual_seed(0)
    torch.cuda.manual_seed(0)
    torch.cuda.manual_seed_all(0)
    np.random.seed(0)
    random.seed(0)
    save_option(args, dirs.list_dir['option'])
    main(dirs, args)
