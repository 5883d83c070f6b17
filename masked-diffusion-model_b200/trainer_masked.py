"""Drop-in for the reference's `trainer_masked.py` (class `Trainer`, /root/reference/code/
trainer_masked.py:30-556): the base masked-diffusion trainer.

Same constructor and method names; `_run_batch` performs the same step --

    t ~ randint, D(x0, t) (K1), net = U-Net(D, t), recon = D + net, loss = mean((recon - x0)^2)
    backward, clip_grad_norm_(1.0), optimizer.step, lr_scheduler.step, zero_grad, EMA step

-- but on B200 every piece is a hand-written kernel, and when the model is the B200 denoiser
with the fused optimiser the whole device side of the step is ONE CUDA-graph replay: the host
only uploads four scalars (lr, bias corrections, EMA decay) and reads three floats back.

Deliberate differences: `Sampler` is built with 4 arguments (the reference's
3-argument call is a TypeError, q2); the sampling helpers unpack the `(sample, visual_list)`
2-tuple the sampler really returns (q3); plotting / wandb / image-grid code is out of scope."""
from __future__ import annotations

import os
import statistics
from timeit import default_timer as timer

import torch

from mdm_b200 import train_ops
from sampler import Sampler
from scheduler import Scheduler


def _complement(ranges, n):
    """the pieces of [0, n) not covered by `ranges`"""
    out, pos = [], 0
    for lo, hi in sorted(ranges):
        if lo > pos:
            out.append((pos, lo))
        pos = max(pos, hi)
    if pos < n:
        out.append((pos, n))
    return out


class Trainer:
    method = "base"

    def __init__(self, args, dataloader, dataset, model, ema_model, optimizer, lr_scheduler, accelerator,
                 dataset_hist=None):
        self.args = args
        self.dataloader = dataloader
        self.dataset = dataset
        self.model = model
        self.ema_model = ema_model
        self.optimizer = optimizer
        self.lr_scheduler = lr_scheduler
        self.lr_list = []
        self.accelerator = accelerator

        self.Scheduler = Scheduler(args)
        self.Sampler = Sampler(self.dataset, self.args, self.Scheduler, dataset_hist)

        self.global_step = 0
        self.train_visual_names = ['input', 'degraded_img', 'degradation_mask', 'mask', 'mean_pixel',
                                   'degrade_binary_masks', 'reconstructed_img']
        self.loss_names = ['reconstruct_loss', 'learning_rate', 'time_steps']
        self.mean_names = ['reconstruct_train_mean', 'degraded_train_mean', 'ema_sample_mean', 'ema_sample_t_mean',
                           'ema_sample_0_mean']
        self.timesteps_used_epoch = None
        self._graphs = {}
        self._ts_dev = {}
        self.degradation_mask = None
        self._pin, self._ticks_expected, self._stats_published = None, 0, False
        if self._fused() and getattr(args, "use_ema", False) and ema_model is not None:
            self.optimizer.attach_ema(ema_model)      # EMA update rides in the optimiser kernel

    # ------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------
    def _fused(self):
        """B200 denoiser + fused optimiser -> the step can be captured as a CUDA graph"""
        from mdm_b200.runtime import FusedOptimizer
        return hasattr(self.model, "flat_param") and isinstance(self.optimizer, FusedOptimizer)

    def _timesteps_table(self, device):
        key = (tuple(self.timesteps_used_epoch), str(device))
        t = self._ts_dev.get(key)
        if t is None:
            self._ts_dev.clear()
            t = torch.tensor(self.timesteps_used_epoch, device=device)
            self._ts_dev[key] = t
        return t

    def _draw_timeindex(self, B, device):
        """trainer_masked.py:114 -- `torch.randint(..., device=input.device)`: the CUDA generator when
        the batch is on the GPU.  `args.timeindex_rng = 'cpu_stream'` draws the indices from the
        CPU-generator stream instead (what the reference does when it runs on CPU; parity tests)."""
        n = len(self.timesteps_used_epoch)
        if getattr(self.args, "timeindex_rng", "torch") == "cpu_stream":
            return self.Scheduler._rng_for(device).randint(0, n, B)
        return torch.randint(low=0, high=n, size=(B,), device=device)

    def _extract_input(self, input):
        if 'huggingface' in getattr(self.args, "dir_dataset", ""):
            x = input["image"]
            self.label = input.get("label") if hasattr(input, "get") else None
        else:
            x = input[0]
        return x

    # ------------------------------------------------------------------------------------------
    # device side of one step (stream-ordered, no host sync -> capturable)
    # ------------------------------------------------------------------------------------------
    def _forward_backward(self):
        a = self.args
        x0 = self.input
        B = x0.shape[0]
        timeindex = self._draw_timeindex(B, x0.device)
        timesteps = torch.index_select(self._timesteps_table(x0.device), 0, timeindex)
        self.timeindex, self.timesteps = timeindex, timesteps
        black_area_num = self.Scheduler.get_black_area_num_pixels_time(timesteps)
        self.degraded_img, self.degrade_binary_masks, self.degradation_mask, self.mean_pixel = \
            self.Scheduler.degrade_training(black_area_num, x0, mean_option=a.mean_option, mean_area=a.mean_area,
                                            want_degrade_mask=bool(getattr(a, "materialize_visuals", False)))
        with self.accelerator.accumulate(self.model):
            self.mask = self.model(self.degraded_img, timesteps).sample
            weight = None
            if a.loss_weight_use:
                weight = self.Scheduler.get_weight_timesteps(timeindex, a.loss_weight_power_base)
            # recon = degraded + net ; loss = mean(w (recon - x0)^2)   (one fused kernel, fwd + grad)
            if x0.dtype == torch.uint8:
                x0 = self.Scheduler.x0_normalised         # written by K1 next to x_t
            self.reconstruct_loss, self.reconstructed_img = train_ops.residual_mse(self.mask, self.degraded_img, x0,
                                                                                   shift=None, weight=weight)
            stats = self._publish(self._stats())      # forward-only statistics: readable before the backward has run
            self.accelerator.backward(self.reconstruct_loss)
        return stats

    def _stats(self):
        self.reconstruct_train_mean = self.reconstructed_img.mean()
        self.degraded_train_mean = self.degraded_img.mean()
        return torch.stack([self.reconstruct_loss.detach().float(), self.reconstruct_train_mean.float(),
                            self.degraded_train_mean.float()])

    # ------------------------------------------------------------------------------------------
    # step statistics without a full device synchronisation
    # ------------------------------------------------------------------------------------------
    def _publish(self, stats):
        """fused path: a kernel stores the statistics and a tick into pinned host memory right after the forward part
        of the step; `_return_values` polls the tick instead of synchronising with the whole step, so the host enqueues
        step i + 1 while the GPU still runs the backward / optimiser of step i."""
        if not self._fused() or not bool(getattr(self.args, "async_stats", True)):
            return stats
        from mdm_b200 import optim_ops
        if getattr(self, "_pin", None) is None:
            self._pin = torch.zeros(32, dtype=torch.float32).pin_memory()
            self._pin_f = self._pin.numpy()
            self._pin_i = self._pin_f.view("int32")
            self._tick_dev = torch.zeros(1, dtype=torch.int32, device=stats.device)
        self._n_stats = stats.numel()
        optim_ops.publish_stats(stats.float().contiguous(), self._tick_dev, self._pin)
        return stats

    def _await_stats(self):
        """the statistics of the step just enqueued (None when nothing was published: generic path)"""
        if getattr(self, "_pin", None) is None or not self._stats_published:
            return None
        n, want = self._n_stats, self._ticks_expected
        t0 = timer()
        while int(self._pin_i[n]) != want:
            if timer() - t0 > 120.0:
                raise RuntimeError(f"step statistics never arrived (tick {int(self._pin_i[n])}, expected {want})")
        return [float(v) for v in self._pin_f[:n]]

    def _optimizer_tail_device(self):
        """clip + optimiser + EMA + zero_grad for the fused optimiser (device part)"""
        self.optimizer.set_clip(1.0)
        self.optimizer.launch()
        self.optimizer._clip = 0.0
        self.optimizer.zero_grad()

    def _step_fused(self):
        acc = self.accelerator
        single = acc.num_processes <= 1
        if acc.gradient_accumulation_steps > 1:      # micro-batches: eager, optimiser only on the last one
            stats = self._forward_backward()
            if acc.sync_gradients:
                self.optimizer.upload_hyper()
                self._optimizer_tail_device()
            return stats
        use_graph = bool(getattr(self.args, "cuda_graph", True))
        key = (tuple(self.input.shape), self.input.dtype, tuple(self.timesteps_used_epoch), single)
        g = self._graphs.get(key)
        if g is None:
            self._graphs.clear()
            p2p = getattr(acc, "p2p", None)
            if single or p2p is not None:
                # single GPU, or data parallel with the peer-memory all-reduce: the WHOLE step is one graph.  Under DP
                # the backward program calls `_dp_range_ready` at every cut (denoiser.py: dp_cut_prefixes): the range
                # that just became final is all-reduced by one kernel on a forked stream while the next segment
                # computes; the remaining ranges and the join come right before the optimiser.
                def body():
                    stats = self._forward_backward()
                    if not single:
                        self._dp_finish()
                    self._optimizer_tail_device()
                    return stats
                g = (train_ops.GraphedCallable(body, enabled=use_graph), None)
            else:
                # data parallel over NCCL (MDM_DP_ALLREDUCE=nccl): graph(fwd+bwd) -> NCCL all-reduce of the flat
                # gradient -> graph(optimiser tail).  (Capturing the NCCL all-reduce inside one graph hung on 2 GPUs
                # in round 1.)
                # With `dp_overlap` (default) the backward is cut into segments, each its own graph: the flat-gradient
                # range a segment finishes is all-reduced asynchronously (NCCL stream) under the next segment.
                segs = []
                m = self.model
                if (bool(getattr(self.args, "dp_overlap", True)) and os.environ.get("MDM_DP_OVERLAP", "1") != "0"
                        and hasattr(m, "grad_segment_ranges")):
                    ranges = m.grad_segment_ranges()
                    segs = [train_ops.GraphedCallable((lambda k=k: m.backward_segment(k)), enabled=use_graph)
                            for k in range(1, len(ranges) + 1)]
                g = (train_ops.GraphedCallable(self._forward_backward, enabled=use_graph),
                     train_ops.GraphedCallable(self._optimizer_tail_device, enabled=use_graph), segs)
            self._graphs[key] = g
        self.optimizer.upload_hyper()
        if self.model.params_dirty():
            # the captured step holds no fp32 -> bf16 cast (the optimiser kernel refreshes the mirror): anything that
            # rewrote the master weights since the last step (EMA copy_to / restore, load_state_dict) re-casts here
            self.model._refresh_bf16()
        if g[1] is None:
            dp = not single
            acc._defer_all_reduce = dp          # (DP: the ranges are reduced by the in-graph peer-memory kernels)
            if dp:
                self.model.grad_ready_hook = self._dp_range_ready
            try:
                stats = g[0]()
            finally:
                acc._defer_all_reduce = False
                if dp:
                    self.model.grad_ready_hook = None
        elif not g[2]:
            acc._defer_all_reduce = True        # the all-reduce runs between the two graphs
            try:
                stats = g[0]()
            finally:
                acc._defer_all_reduce = False
            acc.all_reduce_gradients()
            g[1]()
        else:
            m = self.model
            ranges = m.grad_segment_ranges()
            acc._defer_all_reduce = True
            m.bwd_segmented = True              # loss.backward() inside g[0] stops after backward segment 0
            try:
                stats = g[0]()
                works = [acc.all_reduce_async(m.flat_grad[ranges[0][0]:ranges[0][1]])]
                for k, seg in enumerate(g[2], start=1):
                    seg()
                    if k < len(ranges):
                        works.append(acc.all_reduce_async(m.flat_grad[ranges[k][0]:ranges[k][1]]))
            finally:
                acc._defer_all_reduce = False
                m.bwd_segmented = False
            for lo, hi in _complement(ranges, m.flat_grad.numel()):
                works.append(acc.all_reduce_async(m.flat_grad[lo:hi]))
            for w in works:
                w.wait()                        # stream-level wait: the optimiser graph follows the reductions
            g[1]()
        return stats

    # -- data parallel, peer-memory all-reduce inside the step graph ---------------------------------------
    def _comm(self):
        if getattr(self, "_comm_stream", None) is None:
            # high priority: the few small communication blocks should be placed as soon as their range is ready
            self._comm_stream = torch.cuda.Stream(device=self.accelerator.device, priority=-1)
            self._comm_used = False
        return self._comm_stream

    def _dp_range_ready(self, lo, hi, side_streams=()):
        """called by the backward program when the kernels that finish flat_grad[lo:hi] have been enqueued (dgrad chain on
        the current stream, weight gradients on `side_streams`): the communication stream waits for both and reduces
        the range there; the main stream does not wait"""
        cs = self._comm()
        cs.wait_stream(torch.cuda.current_stream(self.accelerator.device))
        for s_ in side_streams:
            cs.wait_stream(s_)
        with torch.cuda.stream(cs):
            self.accelerator.p2p.all_reduce(lo, hi)
        self._comm_used = True

    def _dp_finish(self):
        """after the backward: reduce whatever the cuts did not cover, then join the communication stream"""
        m, acc = self.model, self.accelerator
        main = torch.cuda.current_stream(acc.device)
        cs = self._comm()
        done = m.grad_segment_ranges() if self._comm_used else []
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            for lo, hi in _complement(done, m.flat_grad.numel()):
                acc.p2p.all_reduce(lo, hi, exposed=True)      # the backward is over: nothing left to disturb
        main.wait_stream(cs)
        self._comm_used = False

    def _step_generic(self):
        """any nn.Module + torch optimiser: the reference's sequence, op by op"""
        stats = self._forward_backward()
        if self.accelerator.sync_gradients:
            self.accelerator.clip_grad_norm_(self.model.parameters(), 1.0)
            # accelerate wraps optimiser and scheduler: both skip the micro-batches that only accumulate
            self.optimizer.step()
            self.lr_scheduler.step()
            self.optimizer.zero_grad()
        return stats

    # ------------------------------------------------------------------------------------------
    # trainer_masked.py:95-183
    # ------------------------------------------------------------------------------------------
    def _set_input(self, x):
        # uint8 batches stay raw: K1 normalises them on load (the GPU data-feeding path)
        if x.dtype != torch.uint8:
            x = x.to(self.args.weight_dtype)
        if not x.is_cuda:
            raise RuntimeError("Trainer: the batch must be on a CUDA device (no CPU fallback)")
        buf = getattr(self, "_input_buf", None)
        if buf is None or buf.shape != x.shape or buf.dtype != x.dtype or buf.device != x.device:
            self._input_buf = buf = torch.empty_like(x)
            self._graphs.clear()
        buf.copy_(x, non_blocking=True)
        self.input = buf

    def _run_batch(self, batch: int, input, epoch: int, epoch_length: int, resume_step: int, dirs: dict, visualizer):
        self._set_input(self._extract_input(input))
        self._stats_published = False
        if self._fused():
            if bool(getattr(self.args, "async_stats", True)):
                self._ticks_expected += 1            # exactly one publish kernel executes per step (eager, capture + replay, or replay)
                self._stats_published = True
            stats = self._step_fused()
            if self.accelerator.sync_gradients:      # accelerate's AcceleratedScheduler skips non-sync micro-batches
                self.lr_scheduler.step()
        else:
            stats = self._step_generic()
        if self.accelerator.sync_gradients:
            if self.args.use_ema:
                self.ema_model.step(self.model.parameters())
            self.global_step += 1
        self.learning_rate = self.lr_scheduler.get_last_lr()[0]
        self.lr_list.append(self.learning_rate)
        self.accelerator.wait_for_everyone()
        return self._return_values(stats)

    def _return_values(self, stats):
        vals = self._await_stats()
        if vals is None:
            vals = stats.tolist()                    # generic path: the single device->host sync of the step
        loss, rmean, dmean = vals
        return loss, rmean, dmean

    def _run_epoch(self, epoch: int, epoch_length: int, resume_step: int, dirs: dict, visualizer):
        loss_batch, reconstruct_train_mean_batch, degraded_train_mean_batch = [], [], []
        self.timesteps_used_epoch = self.Scheduler.get_timesteps_epoch(epoch, epoch_length)
        for i, input in enumerate(self.dataloader, 0):
            loss, rmean, dmean = self._run_batch(i, input, epoch, epoch_length, resume_step, dirs, visualizer)
            if self.accelerator.is_main_process:
                loss_batch.append(loss)
                reconstruct_train_mean_batch.append(rmean)
                degraded_train_mean_batch.append(dmean)
        return loss_batch, reconstruct_train_mean_batch, degraded_train_mean_batch

    def prepare_schedule(self):
        """trainer_masked.py:213-215"""
        updated = self.Scheduler.update_ddpm_num_steps(self.args.ddpm_num_steps)
        self.args.updated_ddpm_num_steps = updated
        self.time_steps = self.Scheduler.get_black_area_num_pixels_all()
        return updated

    def train(self, epoch_start: int, epoch_length: int, resume_step: int, global_step: int, dirs: dict, visualizer):
        self.prepare_schedule()
        self.global_step = global_step
        loss_mean_epoch, loss_std_epoch = [], []
        self.model.train()
        a = self.args
        for epoch in range(epoch_start, epoch_start + epoch_length):
            start = timer()
            if self.accelerator.is_main_process and visualizer is not None:
                visualizer.reset()
            # the masks continue the CPU generator's stream (device shadow; written back per epoch)
            self.Scheduler.adopt_torch_rng(self.accelerator.device)
            out = self._run_epoch(epoch, epoch_length, resume_step, dirs, visualizer)
            self.Scheduler.release_rng_to_torch()
            loss = out[0] if isinstance(out, tuple) else out
            self.elapsed_time = timer() - start
            if self.accelerator.is_main_process:
                loss_mean_epoch.append(statistics.mean(loss) if loss else float("nan"))
                save_now = (epoch > 0 and (epoch + 1) % a.save_images_epochs == 0) or epoch == (epoch_start + epoch_length - 1) \
                    or (epoch + 1) % (epoch_length / a.scheduler_num_scale_timesteps) == 0
                if save_now:
                    self._save_learning_curve(dirs, loss_mean_epoch, loss_std_epoch)
                    result = None
                    if a.use_ema:
                        if a.sampling == 'base':
                            result = self._save_ema_sample(dirs, epoch)
                        elif a.sampling == 'momentum':
                            result = self._save_ema_momentum_sample(dirs, epoch)
                    if visualizer is not None:
                        visualizer.display_current_results(epoch, result)
                        visualizer.plot_current_losses(epoch, self.get_current_mean(), 'value')
                    if dirs is not None:
                        save_path = os.path.join(dirs.list_dir['checkpoint'], f"checkpoint-epoch-{epoch}")
                        self.accelerator.save_state(save_path)
        self.loss_mean_epoch = loss_mean_epoch
        return loss_mean_epoch

    # ------------------------------------------------------------------------------------------
    # periodic EMA sampling (trainer_masked.py:379-470); grids / plots are out of scope
    # ------------------------------------------------------------------------------------------
    def _save_learning_curve(self, dirs, loss_mean, loss_std):
        if dirs is None:
            return
        try:
            path = os.path.join(dirs.list_dir['loss'], 'loss.csv')
        except (KeyError, AttributeError):
            return
        with open(path, 'w') as f:
            f.write(",".join(repr(v) for v in loss_mean) + "\n")

    def _ema_sample(self):
        self.ema_model.store(self.model.parameters())
        self.ema_model.copy_to(self.model.parameters())
        try:
            sample, visual = self.Sampler.sample(self.model.eval(), self.timesteps_used_epoch)
        finally:
            self.ema_model.restore(self.model.parameters())
            self.model.train()
        self.ema_sample = sample
        self.ema_sample_mean = sample.mean()
        return sample

    def _save_ema_sample(self, dirs, epoch):
        return self._ema_sample()

    def _save_ema_momentum_sample(self, dirs, epoch):
        return self._ema_sample()

    def get_current_mean(self):
        out = {}
        for name in self.mean_names:
            if hasattr(self, name):
                out[name] = float(getattr(self, name))
        return out

    def get_current_losses(self):
        return {"reconstruct_loss": float(self.reconstruct_loss), "learning_rate": self.learning_rate}
