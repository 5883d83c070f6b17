"""Drop-in for the reference's `trainer_masked_mean_shift.py` (class `Trainer`,
/root/reference/code/trainer_masked_mean_shift.py:28-467): the mean-shift variant.

Differences from the base trainer (trainer_masked_mean_shift.py:82-193), all kept:
  * constructor takes `dataset_hist` as the 4th positional argument (:29-39);
  * timesteps are cast to `weight_dtype` (:110);
  * `shift = get_schedule_shift_time(t, masks)` (:119), the denoiser sees `x_t + shift` (:120,140);
  * `inverse = (x_t + shift + net) - shift` (:142-145); the loss is fp32 MSE of `inverse` vs x0 (:153);
  * five `.mean()` statistics (:175-179); `_run_batch` returns only the loss (:193).
On B200 the shift noise comes from the device mt19937 stream (bit-aligned with the CPU generator),
`x_t + shift` and `(x_in + net) - shift`, the loss and dLoss/dnet are two fused kernels, and the
whole step replays as one CUDA graph (see trainer_masked.py)."""
from __future__ import annotations

import torch

from mdm_b200 import train_ops
import trainer_masked as _base


class Trainer(_base.Trainer):
    method = "mean_shift"

    def __init__(self, args, dataloader, dataset, dataset_hist, model, ema_model, optimizer, lr_scheduler, accelerator):
        super().__init__(args, dataloader, dataset, model, ema_model, optimizer, lr_scheduler, accelerator,
                         dataset_hist=dataset_hist)
        self.dataset_hist = dataset_hist
        self.train_visual_names = ['input', 'degraded_img', 'degrade_binary_masks', 'degradation_mask', 'mean_pixel',
                                   'shift', 'shifted_degrade_img', 'reconstructed_img',
                                   'inverse_shift_reconstructed_img', 'mask']
        self.loss_names = ['train_loss', 'inverse_reconstruct_train_mean', 'reconstruct_train_mean',
                           'shifted_degrade_img_mean', 'degraded_train_mean']
        self.mean_names = ['ema_sample_mean']

    def _forward_backward(self):
        a = self.args
        x0 = self.input
        B = x0.shape[0]
        wd = a.weight_dtype
        timeindex = self._draw_timeindex(B, x0.device)
        timesteps = torch.index_select(self._timesteps_table(x0.device), 0, timeindex)
        if wd == torch.bfloat16 and max(self.timesteps_used_epoch) > 256:
            # quirk q10: the reference casts timesteps to bf16 (:110) and integers above 256 round, up to an
            # out-of-range index at T' = 1000.  Timesteps stay integral here (documented divergence).
            pass
        else:
            timesteps = timesteps.to(wd)
        self.timeindex, self.timesteps = timeindex, timesteps
        black_area_num = self.Scheduler.get_black_area_num_pixels_time(timesteps)
        self.degraded_img, self.degrade_binary_masks, self.degradation_mask, self.mean_pixel = \
            self.Scheduler.degrade_training(black_area_num, x0, mean_option=a.mean_option, mean_area=a.mean_area,
                                            want_degrade_mask=bool(getattr(a, "materialize_visuals", False)))
        self.shift = self.Scheduler.get_schedule_shift_time(timesteps, self.degrade_binary_masks).to(wd)
        self.shifted_degrade_img = self.Scheduler.perturb_shift(self.degraded_img.to(wd), self.shift)
        with self.accelerator.accumulate(self.model):
            self.mask = self.model(self.shifted_degrade_img, timesteps).sample
            weight = None
            if a.loss_weight_use:
                weight = self.Scheduler.get_weight_timesteps(timeindex, a.loss_weight_power_base)
            # inverse = (x_in + net) - shift ; loss = mean(w (inverse - x0)^2) in fp32
            if x0.dtype == torch.uint8:
                x0 = self.Scheduler.x0_normalised         # written by K1 next to x_t
            self.reconstruct_loss, self.inverse_shift_reconstructed_img = train_ops.residual_mse(
                self.mask, self.shifted_degrade_img, x0, shift=self.shift, weight=weight)
            stats = self._publish(self._stats())      # forward-only statistics: readable before the backward has run
            self.accelerator.backward(self.reconstruct_loss)
        return stats

    def _stats(self):
        self.train_loss = self.reconstruct_loss.detach()
        self.inverse_reconstruct_train_mean = self.inverse_shift_reconstructed_img.mean()
        if bool(getattr(self.args, "materialize_visuals", False)):
            self.reconstructed_img = self.shifted_degrade_img.float() + self.mask.float()
            self.reconstruct_train_mean = self.reconstructed_img.mean()
        self.shifted_degrade_img_mean = self.shifted_degrade_img.float().mean()
        self.degraded_train_mean = self.degraded_img.mean()
        return torch.stack([self.train_loss.float(), self.inverse_reconstruct_train_mean.float(),
                            self.shifted_degrade_img_mean.float(), self.degraded_train_mean.float()])

    def _return_values(self, stats):
        vals = self._await_stats()
        return float(vals[0]) if vals is not None else stats[0].item()

    def _run_epoch(self, epoch: int, epoch_length: int, resume_step: int, dirs: dict, visualizer):
        loss_batch = []
        self.timesteps_used_epoch = self.Scheduler.get_timesteps_epoch(epoch, epoch_length)
        for i, input in enumerate(self.dataloader, 0):
            loss = self._run_batch(i, input, epoch, epoch_length, resume_step, dirs, visualizer)
            if self.accelerator.is_main_process:
                loss_batch.append(loss)
        return loss_batch

    def get_current_losses(self):
        out = {}
        for name in self.loss_names:
            if hasattr(self, name):
                out[name] = float(getattr(self, name))
        return out
