#!/usr/bin/env python
"""bench.py -- masked-diffusion hot path on B200: training samples/s (BASELINE.json configs[1]:
masked U-Net, 3x32x32, bf16 activations, batch 128 per GPU, data parallel) with a secondary
masked-sampling line (configs[3] shape, 3x128x128), the roofline of the dominant kernel and the
reference's CPU path timed on the same box.

    python bench.py --gpus N --steps K --warmup W            # this framework (N>1: under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

One JSON line on stdout (rank 0).  A "step" = one `Trainer._run_batch` (trainer_masked.py:95-183):
timestep draw, K1 degradation, U-Net forward/backward, fused loss, clip + AdamW + EMA.
  value : samples/s with the batches already resident in HBM (device timed, max over ranks)
  e2e   : samples/s through the same public call with HOST (pinned) batches copied in every step and
          the step's loss read back -- the number to compare with the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "masked-diffusion-model_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch

METRIC = "train_samples_per_sec"
UNIT = "samples/s"


def workload_name(C, S, method):
    """the same string in both arms (the driver compares `config` of the two JSON lines)"""
    return (f"masked U-Net {C}x{S}x{S} training step ({method} trainer, trainer_masked.py:95-183), "
            f"113.67M-param UNet2DModel config")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch (reference default, main_train_masked.py:370)")
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--method", default="base", choices=["base", "mean_shift"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-sampling", action="store_true", help="skip the secondary masked-sampling measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--sample-size", type=int, default=128)
    ap.add_argument("--sample-batch", type=int, default=256)
    ap.add_argument("--sample-steps", type=int, default=1000,
                    help="denoising steps of the timed restoration loop (BASELINE configs[3]: the whole 1000-step loop, timed once)")
    ap.add_argument("--ref-batch", type=int, default=0,
                    help="samples per reference-arm step (0 = the workload's own per-GPU batch, so both arms run the same config)")
    ap.add_argument("--cpu-batch", type=int, default=8, help="samples per step of the bounded cpu_baseline leg of the b200 arm")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] / configs[4] sub-objects and the kernel micro-benchmarks")
    ap.add_argument("--c5-batch", type=int, default=8)
    return ap.parse_args()


def workload_args(a, mixed="bf16"):
    """the reference's flag set for this workload (main_train_masked.py:347-417)"""
    from mdm_b200.config import default_args
    ns = default_args(data_size=a.size, in_channel=a.channels, out_channel=a.channels, batch_size=a.batch,
                      ddpm_num_steps=1000, ddpm_schedule="log", select_degrade_pixel="indexing",
                      mean_option="degraded_area", mean_area="image-wise", method=a.method,
                      shift_type="noise_with_perturbation", sample_latent_shape="zero",
                      momentum_adaptive="base_momentum", sampling_mask_dependency="independent",
                      mixed_precision=mixed, use_ema=True, optim="adamw", lr=1e-4)
    return ns


# ---------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index=0, period=0.01):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.index, self.period = index, period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80,
                 "sw_thermal_slowdown": 0x20, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def build_trainer(a, wa, base=128):
    import main_train_masked as M
    torch.manual_seed(0)
    torch.cuda.manual_seed_all(0)
    np.random.seed(0)
    wa.model = "default"
    wa.num_attention = 1
    wa.gradient_accumulation_steps = 1
    wa.ema_max_decay, wa.ema_inv_gamma, wa.ema_power = 0.9999, 1.0, 0.75
    wa.cuda_graph = not a.no_graph
    if base == 128:
        model = M.get_model(wa)
    else:                                   # configs[4]: the large-channel U-Net (block_out_channels 256 .. 1024)
        from mdm_b200.denoiser import UNet2DModelB200, default_config
        model = UNet2DModelB200(device="cuda", **default_config(wa.in_channel, wa.data_size, base=base))
    ema = M.get_ema(wa, model)
    acc = M.get_accelerator(wa, ema)
    M.get_weight_type(wa, acc)
    opt = M.get_optimizer(model, wa.optim, wa.lr)
    sched = M.get_scheduler("cosine", opt, num_warmup_steps=500, num_training_steps=100000, num_cycles=0.5)
    model, opt, sched = acc.prepare(model, opt, sched)
    if wa.method == "base":
        tr = M.BaseTrainer(wa, None, None, model, ema, opt, sched, acc)
    else:
        tr = M.MeanShiftTrainer(wa, None, None, [None, None, None], model, ema, opt, sched, acc)
    tr.prepare_schedule()
    tr.timesteps_used_epoch = tr.Scheduler.get_timesteps_epoch(0, 1)
    model.train()
    return tr, model, acc


def conv_flops_per_step(model, B, S):
    """algorithmic FLOPs (2*MAC) of one training step = 3 x forward (fprop + dgrad + wgrad), SURVEY.md 8d"""
    from mdm_b200.config import unet_forward_flops
    return 3.0 * B * unet_forward_flops(model._cfg, S)


def timed_steps(run_step, steps, dev, world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        run_step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def roofline_pass(tr, model, batch_dev, peaks):
    """time every tcgen05 implicit-GEMM launch of ONE eager step with CUDA events on the launching
    stream; achieved = algorithmic conv/linear FLOPs of those launches / their summed duration"""
    from mdm_b200 import denoiser_ops as ops
    rec = []
    orig = {}

    def wrap(name, flops_fn):
        f = getattr(ops, name)
        orig[name] = f

        def g(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = f(*args, **kw)
            e1.record()
            rec.append((name, flops_fn(*args, **kw), e0, e1))
            return r
        setattr(ops, name, g)

    def fl_fprop(x, w, y, N, H, W, ksize=3, stride=1, **kw):
        cout = y.shape[-1] if y is not None else kw.get("cout")
        f = 2.0 * N * H * W * cout * ksize * ksize * x.shape[-1]
        if kw.get("x2") is not None:
            f += 2.0 * N * H * W * cout * kw["x2"].shape[-1]
        return f

    def fl_dgrad(dy, w, dx, N, H, W, ksize=3, **kw):
        cin = dx.shape[-1] if dx is not None else kw.get("cin")
        return 2.0 * N * H * W * cin * ksize * ksize * dy.shape[-1]

    def fl_wgrad(x, dy, dw, N, H, W, ksize=3, stride=1, **kw):
        return 2.0 * N * H * W * dy.shape[-1] * ksize * ksize * x.shape[-1]

    wrap("conv_fprop", fl_fprop)
    wrap("conv_dgrad", fl_dgrad)
    wrap("conv_wgrad", fl_wgrad)
    # the HBM-bound kernels of the step, with their ALGORITHMIC bytes (bf16 activations: read x (+dy, +add, +add2),
    # write y / dx); reported next to the dominant kernel as `roofline.hbm_kernels`
    hbm = []

    def wrap_bytes(name, bytes_fn):
        f = getattr(ops, name)
        orig[name] = f

        def g(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = f(*args, **kw)
            e1.record()
            hbm.append((name, bytes_fn(*args, **kw), e0, e1))
            return r
        setattr(ops, name, g)

    from mdm_b200 import optim_ops as _oo
    _adam = _oo.adam_ema_step_dev

    def adam_timed(p, g, m, v, ema, p16, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = _adam(p, g, m, v, ema, p16, *a, **k)
        e1.record()
        # read p, g, m, v (+ema), write p, m, v (+ema) fp32 + the bf16 mirror
        hbm.append(("adam_ema_step", float(p.numel()) * (4 * (7 + (2 if ema is not None else 0)) + 2), e0, e1))
        return r
    _oo.adam_ema_step_dev = adam_timed
    wrap_bytes("gn_silu_fwd", lambda x, y, gamma, beta, stats, ws, N, HW, C, *a, **k: 2.0 * 2 * N * HW * C)
    wrap_bytes("gn_silu_fwd_q", lambda x, y, gamma, beta, stats, qa, qb, N, HW, C, *a, **k: 2.0 * 2 * N * HW * C)
    wrap_bytes("gn_silu_bwd", lambda x, dy, dx, gamma, beta, stats, dgamma, dbeta, ws, N, HW, C, *a, **k:
               2.0 * N * HW * C * (3 + (k.get("add") is not None) + (k.get("add2") is not None)))
    saved = tr.args.cuda_graph
    tr.args.cuda_graph = False
    tr._graphs.clear()
    model.wgrad_side_stream = False     # serialise wgrad with the main chain: each event pair then times ONE kernel
    try:
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # a ~150 ms spin kernel first: the host enqueues the whole eager step behind it, so the stream never
        # starves and each event pair brackets its kernel only (not host launch latency)
        torch.cuda._sleep(int(8e8))
        s0.record()
        tr._run_batch(0, (batch_dev,), 0, 1, 0, None, None)
        s1.record()
        torch.cuda.synchronize()
    finally:
        for k, f in orig.items():
            setattr(ops, k, f)
        _oo.adam_ema_step_dev = _adam
        model.wgrad_side_stream = True
        tr.args.cuda_graph = saved
        tr._graphs.clear()
    hbm_peak = peaks.get("hbm_gbs", 6532.2)
    hbm_by = {}
    for name, nbytes, e0, e1 in hbm:
        d = hbm_by.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += nbytes
        d[2] += e0.elapsed_time(e1)
    hbm_kernels = {k: {"launches": v[0], "algorithmic_GB": round(v[1] / 1e9, 3), "ms": round(v[2], 3),
                       "GBps": round(v[1] / (v[2] * 1e-3) / 1e9, 1) if v[2] > 0 else None,
                       "frac_of_hbm_peak": round(v[1] / (v[2] * 1e-3) / 1e9 / hbm_peak, 3) if v[2] > 0 else None}
                   for k, v in hbm_by.items()}
    total_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in rec)
    total_fl = sum(f for _, f, _, _ in rec)
    by = {}
    for name, f, e0, e1 in rec:
        d = by.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += f
        d[2] += e0.elapsed_time(e1)
    achieved = total_fl / (total_ms * 1e-3) / 1e12 if total_ms > 0 else 0.0
    peak = peaks.get("bf16_tflops_sustained", 1373.8)
    return {
        "bound": "tensor", "kernel": "igemm_kernel (tcgen05 implicit-GEMM conv/linear: fprop+dgrad+wgrad)",
        "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
        **igemm_traffic(),
        "launches_per_step": len(rec), "avg_launch_us": round(1e3 * total_ms / max(1, len(rec)), 2),
        "kernel_ms_per_step": round(total_ms, 3),
        "by_kind": {k: {"launches": v[0], "tflops": round(v[1] / (v[2] * 1e-3) / 1e12, 2) if v[2] > 0 else None,
                        "ms": round(v[2], 3)} for k, v in by.items()},
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback (B200_PROFILING.md)",
        # GroupNorm(+SiLU) forward / backward call sites of the same step (1-2 kernels each, launch gaps of the eager
        # pass included), achieved GB/s from ALGORITHMIC bytes against the measured HBM copy peak
        "hbm_kernels": hbm_kernels,
    }


def _git_blob_sha1(path):
    import hashlib
    data = open(path, "rb").read()
    return hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()


def igemm_traffic():
    """`roofline.traffic`: dram__bytes_read.sum + dram__bytes_write.sum per igemm launch from the committed ncu launch
    list of this command (profiles/igemm_traffic.json, written by scripts/make_traffic.py from the CSV it names).  The
    file records the git blob hash of csrc/igemm.cu it was measured with; if the kernel source changed since, the
    figure is reported as stale (null) instead of being carried along."""
    path = os.path.join(ROOT, "profiles", "igemm_traffic.json")
    try:
        t = json.load(open(path))
        cur = _git_blob_sha1(os.path.join(PKG, "mdm_b200", "csrc", "igemm.cu"))
        if t.get("igemm_cu_blob") != cur:
            return {"traffic": None, "traffic_source": f"{t.get('source')}: STALE (measured with igemm.cu {str(t.get('igemm_cu_blob'))[:12]}, "
                                                       f"current {cur[:12]})"}
        return {"traffic": t["traffic_bytes_per_launch"], "traffic_source": f"{t['source']} (ncu, per launch, average over "
                f"{t['launches']} igemm launches; igemm.cu blob {cur[:12]} verified)"}
    except Exception as e:
        return {"traffic": None, "traffic_source": f"unavailable: {e!r}"[:200]}


def _max_over_ranks(ms, dev, world):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def sampling_line(a, model, dev, world, S=None, N=None, n=None, label="configs[3]"):
    """secondary metric: masked-sampling throughput at the configs[3] shape through `Sampler.sample`"""
    import sampler as sampler_mod
    import scheduler as scheduler_mod
    from mdm_b200.config import default_args
    from mdm_b200.denoiser import UNet2DModelB200, default_config
    rank = int(os.environ.get("RANK", 0))
    S, N, n = S or a.sample_size, N or a.sample_batch, n or a.sample_steps
    sa = default_args(data_size=S, in_channel=3, out_channel=3, ddpm_num_steps=1000, ddpm_schedule="linear",
                      select_degrade_pixel="thresholding", degrade_channel="1-channel", mean_option="0",
                      mean_area="image-wise", method="base", shift_type="noise_with_perturbation",
                      sample_latent_shape="zero", momentum_adaptive="base_momentum",
                      sampling_mask_dependency="independent", sample_num=N)
    sa.weight_dtype = torch.float32
    m = UNet2DModelB200(device=dev, **default_config(3, S))
    m.reset_parameters(seed=0)
    m.eval()
    Sch = scheduler_mod.Scheduler(sa)
    Tp = Sch.update_ddpm_num_steps(1000)
    ts = Sch.get_timesteps_epoch(0, 1)
    n = min(n, len(ts))
    smp = sampler_mod.Sampler(None, sa, Sch, [None, None, None])
    # batch shards with no communication: every rank is an independent reference process with its own CPU-generator
    # stream, seeded seed0 + rank -> different latents, masks and noise on every GPU
    torch.manual_seed(0 + rank)
    smp.sample(m, ts[-4:])                        # warm-up (plans, workspaces; the denoiser captures its forward graph on call 3)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, _ = smp.sample(m, ts[-n:])
    e1.record()
    torch.cuda.synchronize()
    total_ms = _max_over_ranks(e0.elapsed_time(e1), dev, world)      # the slowest shard bounds the job
    ms = total_ms / n
    from mdm_b200.config import unet_forward_flops
    fl = N * unet_forward_flops(m._cfg, S)
    finite = bool(torch.isfinite(out).all().item())
    chk = float(out.double().sum().item())
    del m
    torch.cuda.empty_cache()
    peak = 1373.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", peak)
    except Exception:
        pass
    full = n == len(ts)
    return {"metric": "masked_sampling_images_per_sec", "unit": "images/s (1000 denoising steps per image)",
            "value": round(world * N / (ms * 1e-3 * 1000), 4), "denoise_steps_per_s": round(1e3 / ms, 3),
            "ms_per_denoise_step": round(ms, 3), "timed_steps": n, "loop_ms": round(total_ms, 1),
            "whole_loop_timed": full, **({} if full else {"extrapolated_to_steps": 1000}),
            "timing": "CUDA events around ONE Sampler.sample() call, max over ranks",
            "denoiser_tflops": round(fl / (ms * 1e-3) / 1e12, 2),
            "frac_of_tensor_peak": round(fl / (ms * 1e-3) / 1e12 / peak, 4),
            "config": {"workload": f"{label}: sampler.py restoration loop, {N}x3x{S}x{S} per GPU, thresholding/linear, T'={len(ts)}",
                       "seed": f"{0}+rank", "finite": finite, "rank0_output_sum": round(chk, 3)}}


def train_section(a, dev, world, rank, C, S, B, method, base, steps, warmup, label):
    """one of the other BASELINE training configurations, measured exactly like the headline (device-timed steps of
    `Trainer._run_batch` on resident batches, max over ranks, data parallel when world > 1)"""
    import argparse as _ap
    from mdm_b200.config import unet_forward_flops
    a2 = _ap.Namespace(**vars(a))
    a2.size, a2.channels, a2.batch, a2.method = S, C, B, method
    wa = workload_args(a2)
    tr, model, acc = build_trainer(a2, wa, base=base)
    g = torch.Generator().manual_seed(2000 + rank)
    devb = [(torch.rand(B, C, S, S, generator=g) * 2 - 1).to(dev) for _ in range(4)]
    torch.manual_seed(0)
    tr.Scheduler.adopt_torch_rng(dev)

    def step(i):
        return tr._run_batch(i, (devb[i % 4],), 0, 1, 0, None, None)
    for i in range(max(3, warmup)):
        step(i)
    ms = timed_steps(step, steps, dev, world)
    loss = step(0)
    loss = loss[0] if isinstance(loss, tuple) else loss
    tr.Scheduler.release_rng_to_torch()
    fl = 3.0 * B * unet_forward_flops(model._cfg, S)
    peak = 1373.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", peak)
    except Exception:
        pass
    out = {"metric": METRIC, "unit": UNIT, "value": round(world * B * steps / (ms * 1e-3), 2), "ms_per_step": round(ms / steps, 3),
           "steps": steps, "n_gpus": world, "model_tflops_per_gpu": round(fl / (ms / steps * 1e-3) / 1e12, 2),
           "frac_of_tensor_peak": round(fl / (ms / steps * 1e-3) / 1e12 / peak, 4), "loss": round(float(loss), 5),
           "params_M": round(model.num_parameters() / 1e6, 2), "peak_mem_GiB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
           "config": {"workload": f"{label}: {workload_name(C, S, method)}" if base == 128 else
                      f"{label}: large-channel masked U-Net (ch={base}, {model.num_parameters() / 1e6:.2f}M parameters) {C}x{S}x{S} training step ({method} trainer)",
                      "per_gpu_batch": B, "global_batch": world * B, "parallelism": f"dp{world}", "cuda_graph": bool(wa.cuda_graph)}}
    acc.close()                    # (data parallel: unmap the peers' gradient buffers before this model is freed)
    del tr, model, acc, devb
    torch.cuda.empty_cache()
    return out


def micro_kernels(dev, peaks):
    """the HBM-bound kernels of the path that the instrumented training step does not isolate, each timed alone at the
    configs[3] shape as CUDA-graph replays with CUDA events (L2 flushed before every launch, flush time subtracted),
    achieved GB/s from ALGORITHMIC bytes against the measured HBM copy peak"""
    import scheduler as sched_mod
    from mdm_b200 import denoiser_ops as ops
    from mdm_b200._lib import check, lib, ptr, stream_ptr
    from mdm_b200.config import default_args
    hbm_peak = peaks.get("hbm_gbs", 6532.2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, iters=8):
        """device time of fn() alone: [flush L2, fn] x iters replayed as ONE CUDA graph minus the same graph without fn
        (eager timing of 20-100 us kernels measures the Python launch path, not the kernels)"""
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        times = []
        for with_fn in (True, False):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(iters):
                    flush.zero_()
                    if with_fn:
                        fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / iters)
            del g
        return max(times[0] - times[1], 1e-4)

    def entry(nbytes, ms, what):
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"launches": 1, "algorithmic_GB": round(nbytes / 1e9, 4), "ms": round(ms, 4), "GBps": round(gbs, 1),
                "frac_of_hbm_peak": round(gbs / hbm_peak, 3), "shape": what}
    out = {}
    B, C, S = 256, 3, 128
    hw = S * S
    a = default_args(data_size=S, in_channel=C, out_channel=C, ddpm_num_steps=1000, ddpm_schedule="linear",
                     select_degrade_pixel="thresholding", degrade_channel="1-channel", mean_option="degraded_area",
                     mean_area="image-wise", sample_num=B)
    Sc = sched_mod.Scheduler(a)
    Tp = Sc.update_ddpm_num_steps(1000)
    torch.manual_seed(0)
    Sc.adopt_torch_rng(dev)
    x0 = torch.rand(B, C, S, S, device=dev) * 2 - 1
    ts = torch.randint(1, Tp + 1, (B,), device=dev)
    n = Sc.get_black_area_num_pixels_time(ts)
    mb = Sc.make_mask_bytes(n, x0.device).clone()
    mb2 = Sc.make_mask_bytes(n, x0.device).clone()
    out["K1_degrade_composite"] = entry(B * (C * hw * 4 * 2 + hw * 4 + hw),
                                        timeit(lambda: Sc._composite(x0, mb, 1, a.mean_option, a.mean_area, want_mask=True, want_degrade_mask=False)),
                                        f"{B}x{C}x{S}x{S} fp32: read x0 + byte mask, write x_t + 1-channel fp32 mask")
    os.environ["MDM_DEGRADE_FUSED"] = "1"           # the opt-in single-pass cluster form of the same kernel
    out["K1_single_pass_cluster_variant"] = entry(B * (C * hw * 4 * 2 + hw * 4 + hw),
                                                  timeit(lambda: Sc._composite(x0, mb, 1, a.mean_option, a.mean_area, want_mask=True, want_degrade_mask=False)),
                                                  "same shape; one 6-CTA cluster per sample, sample in registers (not the default: slower)")
    os.environ["MDM_DEGRADE_FUSED"] = "0"
    t_mask = timeit(lambda: Sc.make_mask_bytes(n, x0.device))
    out["mask_generate_threshold"] = {"ms": round(t_mask, 4), "words_per_ns": round(B * hw / (t_mask * 1e6), 2),
                                      "shape": f"{B}x1x{S}x{S} threshold mask from the mt19937 stream (0 algorithmic HBM bytes in, {B * hw} B out)"}
    net, x_t, shift = torch.randn_like(x0), torch.randn_like(x0), torch.randn_like(x0)
    x_next, x_in_next = torch.empty_like(x0), torch.empty_like(x0)
    ws = torch.empty(max(1, 2 * lib().mdm_degrade_ws_floats(B, C, hw)), device=dev)

    def k5():
        check(lib().mdm_sampler_step(ptr(x_t), ptr(net), ptr(shift), C * hw, hw, 1, ptr(mb), ptr(mb2), 1, 1, 0.0, 0, 1, 1,
                                     ptr(shift), C * hw, hw, 1, ptr(x_next), ptr(x_in_next), None, ptr(ws), B, C, hw, stream_ptr(dev)))
    out["K5_sampler_step"] = entry(B * C * hw * 4 * 5 + 2 * B * hw, timeit(k5),
                                   f"{B}x{C}x{S}x{S} fp32: 5 image passes + two byte masks (SURVEY 8d)")
    os.environ["MDM_DEGRADE_FUSED"] = "1"
    out["K5_single_pass_cluster_variant"] = entry(B * C * hw * 4 * 5 + 2 * B * hw, timeit(k5), "same shape; opt-in cluster form (slower)")
    os.environ["MDM_DEGRADE_FUSED"] = "0"
    Sc.release_rng_to_torch()
    del x0, net, x_t, shift, x_next, x_in_next
    # K4 attention core at the configs[3] (64 tokens) and configs[4] (256 tokens) shapes: read qkv, write out (bf16)
    for (Bt, L, Ca, tag) in ((256, 64, 512, "c4"), (8, 256, 1024, "c5")):
        qkv = torch.randn(Bt * L, 3 * Ca, device=dev).to(torch.bfloat16)
        att = torch.empty(Bt * L, Ca, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.attention_fwd(qkv, att, Bt, L, Ca))
        e = entry(Bt * L * Ca * 2 * 4, ms, f"{Bt} x {L} tokens x {Ca} channels ({Ca // 8} heads of 8), bf16: read q,k,v, write out")
        e["tflops"] = round(4.0 * Bt * L * L * Ca / (ms * 1e-3) / 1e12, 2)
        out[f"K4_attention_fwd_{tag}"] = e
        # backward of the core: read q,k,v + dO, write dq,dk,dv; 72 FMAs per (query, key, head) = 18 B L^2 C flops
        dqkv = torch.empty_like(qkv)
        ms = timeit(lambda: ops.attention_bwd(qkv, att, dqkv, Bt, L, Ca))
        e = entry(Bt * L * Ca * 2 * 7, ms, f"{Bt} x {L} tokens x {Ca} channels, bf16: read q,k,v,dO, write dq,dk,dv")
        e["tflops"] = round(18.0 * Bt * L * L * Ca / (ms * 1e-3) / 1e12, 2)
        out[f"K4_attention_bwd_{tag}"] = e
        del qkv, att, dqkv
    # first / last convolution (3 image channels) at the configs[3] shape
    img = torch.rand(B, C, S, S, device=dev)
    g_in = torch.zeros(B, S, S, 64, device=dev, dtype=torch.bfloat16)
    out["conv_in_im2col"] = entry(B * hw * (C * 4 + 9 * C * 2), timeit(lambda: ops.im2col3x3(img, g_in, B, C, S, S)),
                                  f"{B}x{C}x{S}x{S}: read the fp32 image, write the 27 data columns of the [pixel][64] bf16 operand")
    z = torch.randn(B * hw, 32, device=dev)
    bias = torch.zeros(C, device=dev)
    o = torch.empty(B, C, S, S, device=dev)
    out["conv_out_tapsum"] = entry(B * hw * (27 * 4 + C * 4), timeit(lambda: ops.tapsum3x3(z, bias, o, B, C, S, S)),
                                   f"{B}x{C}x{S}x{S}: read 27 fp32 per-tap partials per pixel, write the fp32 image")
    del img, g_in, z, o, flush
    torch.cuda.empty_cache()
    return out


def cpu_reference_step(a, ref_batch, threads):
    """The reference algorithm on the host: scheduler/trainer algebra of the oracle (restating
    scheduler.py + trainer_masked.py:95-183) + the fp32 PyTorch restatement of diffusers.UNet2DModel,
    torch CPU, all host threads.  Returns a closure running ONE optimisation step on `ref_batch` samples."""
    from oracle.mdm_oracle import OracleRNG, OracleScheduler, default_args, oracle_train_step
    from oracle.unet_ref import UNet2DModelRef, unet_config
    torch.set_num_threads(threads)
    ra = default_args(data_size=a.size, in_channel=a.channels, out_channel=a.channels, ddpm_num_steps=1000,
                      ddpm_schedule="log", select_degrade_pixel="indexing", mean_option="degraded_area",
                      mean_area="image-wise", method=a.method, shift_type="noise_with_perturbation")
    ra.weight_dtype = torch.float32
    torch.manual_seed(0)
    net = UNet2DModelRef(**unet_config(a.channels, a.size))
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
    S = OracleScheduler(ra, OracleRNG(0))
    S.update_ddpm_num_steps()
    ts = S.get_timesteps_epoch(0, 1)
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(ref_batch, a.channels, a.size, a.size, generator=g) * 2 - 1

    def step():
        loss, _ = oracle_train_step(ra, S, net, x0, ts, a.method)
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        return float(loss)
    return step


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rb = a.ref_batch or a.batch            # the workload's own per-GPU batch: one step = the same work as a b200-arm step
    step = cpu_reference_step(a, rb, threads)
    budget = float(os.environ.get("MDM_REF_BUDGET_S", "270"))
    t_start = time.perf_counter()
    warm = 0
    for _ in range(max(1, a.warmup)):
        step()
        warm += 1
        if time.perf_counter() - t_start > 0.25 * budget:     # warm-up never eats more than a quarter of the budget
            break
    steps = max(1, a.steps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if time.perf_counter() - t_start > budget:
            break
    dt = time.perf_counter() - t0
    v = done * rb / dt
    sample = (f"{done} steps x {rb} samples of the {a.channels}x{a.size}x{a.size} training step (the per-GPU batch of the b200 arm), "
              f"fp32, torch CPU, {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": a.gpus, "steps": done,
            "warmup": warm, "ms_per_step": round(1e3 * dt / done, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.channels, a.size, a.method), "global_batch": a.gpus * a.batch,
                       "per_gpu_batch": a.batch, "parallelism": f"dp{a.gpus}", "reference_sample": sample,
                       "steps_requested": a.steps, "warmup_requested": a.warmup,
                       "note": "one host process whatever --gpus is: the CPU arm has no data-parallel dimension"},
            "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(a):
    import torch.distributed as dist
    from mdm_b200 import _lib
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this is the B200 path, there is no CPU fallback "
                         "(use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    wa = workload_args(a)
    tr, model, acc = build_trainer(a, wa)
    B, C, S = a.batch, a.channels, a.size
    # synthetic data: 8 distinct uniform [-1, 1] batches per rank (different shards per rank), pinned host copies
    g = torch.Generator().manual_seed(1000 + rank)
    host = [(torch.rand(B, C, S, S, generator=g) * 2 - 1).pin_memory() for _ in range(8)]
    devb = [h.to(dev) for h in host]
    torch.manual_seed(0)                       # every rank seeds 0 (main_train_masked.py:441-445): identical masks
    tr.Scheduler.adopt_torch_rng(dev)

    def step_resident(i):
        return tr._run_batch(i, (devb[i % 8],), 0, 1, 0, None, None)

    def step_host(i):
        x = host[i % 8].to(dev, non_blocking=True)          # H2D from pinned memory inside the timed region
        return tr._run_batch(i, (x,), 0, 1, 0, None, None)  # reads the loss back (D2H) every step

    for i in range(max(a.warmup, 3)):
        step_resident(i)
    torch.cuda.synchronize()
    l0 = _lib.lib().mdm_launch_count()
    if not a.no_graph:
        # launches of one step = what the captured graph holds; re-count them with one eager step
        tr.args.cuda_graph = False
        saved = tr._graphs
        tr._graphs = {}
        step_resident(0)
        torch.cuda.synchronize()
        per_step = _lib.lib().mdm_launch_count() - l0
        tr.args.cuda_graph = True
        tr._graphs = saved
    with ClockSampler(local, period=0.004) as clk:           # covers both timed regions (resident, then end to end)
        ms = timed_steps(step_resident, a.steps, dev, world)
        if a.no_graph:
            per_step = (_lib.lib().mdm_launch_count() - l0) // max(1, a.steps)
        for i in range(2):
            step_host(i)
        ms_e2e = timed_steps(step_host, a.steps, dev, world)
    loss_last = step_resident(0)
    loss_last = loss_last[0] if isinstance(loss_last, tuple) else loss_last
    value = world * B * a.steps / (ms * 1e-3)
    e2e = world * B * a.steps / (ms_e2e * 1e-3)
    fl_step = conv_flops_per_step(model, B, S)
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(C, S, a.method),
                   "precision": "bf16 activations / fp32 master weights + AdamW + EMA",
                   "global_batch": world * B, "per_gpu_batch": B, "parallelism": f"dp{world}",
                   "cuda_graph": not a.no_graph,
                   "igemm_work_distribution": "static" if os.environ.get("MDM_IGEMM_DYNAMIC", "1") == "0" else "dynamic",
                   "l2": "per-step working set (weights 0.68 GB + activations > 2 GB) exceeds the 126 MB L2; "
                         "8 distinct input batches rotate"},
        "e2e": {"value": round(e2e, 2), "unit": UNIT, "h2d_bytes_per_step": B * C * S * S * 4, "d2h_bytes_per_step": 12,
                "ms_per_step": round(ms_e2e / a.steps, 3)},
        "gpu_launches": int(per_step * a.steps),
        "gpu_launches_per_step": int(per_step),
        "model_tflops": round(fl_step * world / (ms / a.steps * 1e-3) / 1e12, 2),
        "frac_of_tensor_peak_whole_step": round(fl_step / (ms / a.steps * 1e-3) / 1e12 / peaks.get("bf16_tflops_sustained", 1373.8), 4),
        "loss": round(float(loss_last), 5),
        "clocks": clk.summary(),
    }
    if not a.no_roofline:
        # every rank runs the instrumented step (it contains the gradient all-reduce: a collective must not be
        # entered by rank 0 alone); rank 0's numbers are reported
        rl = roofline_pass(tr, model, devb[0], peaks)
        if rank == 0:
            line["roofline"] = rl
    tr.Scheduler.release_rng_to_torch()
    acc.close()                    # (data parallel: unmap the peers' gradient buffers before this model is freed)
    del tr, model, acc, devb
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()

    def section(key, fn, into=None):
        """secondary measurements must not take the headline down"""
        tgt = line if into is None else into
        try:
            tgt[key] = fn()
        except Exception as e:
            tgt[key] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    if not a.no_sampling:
        # BASELINE configs[3]: batch-sharded sampling 3x128x128, 1000 steps, 256 images per GPU -- the WHOLE loop timed once
        section("sampling", lambda: sampling_line(a, None, dev, world))
    if not a.no_extra:
        cfgs = line.setdefault("configs", {})
        # configs[2]: mean-shift trainer at 3x64x64 (batch 48: 64 would trigger quirk q7) + full-T' sampling of 100 images
        section("c3_train", lambda: train_section(a, dev, world, rank, 3, 64, 48, "mean_shift", 128, 20, 5, "configs[2]"), cfgs)
        if not a.no_sampling:
            section("c3_sampling", lambda: sampling_line(a, None, dev, world, S=64, N=100, n=1000, label="configs[2]"), cfgs)
        # configs[4]: ch=256 U-Net (454.46 M parameters, attention at 16x16) at 3x256x256, data parallel over the ranks
        section("c5_train", lambda: train_section(a, dev, world, rank, 3, 256, a.c5_batch, "base", 256, 6, 3, "configs[4]"), cfgs)
        if rank == 0 and "roofline" in line:
            section("micro", lambda: micro_kernels(dev, peaks))
            if isinstance(line.get("micro"), dict) and "error" not in line["micro"]:
                line["roofline"]["hbm_kernels"].update(line.pop("micro"))
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step = cpu_reference_step(a, a.cpu_batch, threads)
        step()
        t0 = time.perf_counter()
        n = 0
        while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 20):
            step()
            n += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": round(n * a.cpu_batch / dt, 4), "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n} steps x {a.cpu_batch} samples of the same training step, fp32 torch CPU "
                                          f"(oracle restatement of trainer_masked.py + UNet2DModel)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
