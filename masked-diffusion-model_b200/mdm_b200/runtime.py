"""Host-side runtime the reference gets from `accelerate` / `diffusers` (both absent from the image
and third-party): a data-parallel `Accelerator`, `EMAModel`, the warm-up
LR schedules and a fused optimiser.  Same call surface the reference uses
(main_train_masked.py:116-227,299-307; trainer_masked.py:124-166,268).

One process per GPU (torchrun / RANK, LOCAL_RANK, WORLD_SIZE); gradients are averaged with NCCL
all-reduce on the flat fp32 gradient buffer (gloo on CPU for the logic tests)."""
from __future__ import annotations

import contextlib
import json
import math
import os
import pickle
import random
from datetime import timedelta

import numpy as np
import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------
# LR schedules (diffusers.optimization get_*_schedule_with_warmup semantics)
# ---------------------------------------------------------------------------------------------
def _cosine(step, w, total, cycles):
    if step < w:
        return step / max(1, w)
    p = (step - w) / max(1, total - w)
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * cycles * 2.0 * p)))


def _hard_cosine(step, w, total, cycles):
    if step < w:
        return step / max(1, w)
    p = (step - w) / max(1, total - w)
    if p >= 1.0:
        return 0.0
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * ((cycles * p) % 1.0))))


def _constant(step, w, total, cycles):
    return step / max(1.0, w) if step < w else 1.0


def _linear(step, w, total, cycles):
    if step < w:
        return step / max(1, w)
    return max(0.0, (total - step) / max(1, total - w))


SCHEDULES = {"cosine": _cosine, "hard_cosine": _hard_cosine, "constant": _constant, "linear": _linear}


class LambdaSchedule:
    """torch LambdaLR-shaped schedule for any optimiser exposing `param_groups`."""

    def __init__(self, optimizer, name, num_warmup_steps, num_training_steps=0, num_cycles=0.5):
        self.optimizer = optimizer
        self.fn = SCHEDULES[name]
        self.w, self.total, self.cycles = num_warmup_steps, num_training_steps, num_cycles
        self.base_lrs = [g.setdefault("initial_lr", g["lr"]) for g in optimizer.param_groups]
        self.last_epoch = 0
        self.steps_per_call = 1           # accelerate: num_processes scheduler steps per optimiser step
        self._apply()

    def _apply(self):
        lam = self.fn(self.last_epoch, self.w, self.total, self.cycles)
        self._last = [b * lam for b in self.base_lrs]
        for g, lr in zip(self.optimizer.param_groups, self._last):
            g["lr"] = lr

    def step(self):
        self.last_epoch += self.steps_per_call
        self._apply()

    def get_last_lr(self):
        return list(self._last)

    def state_dict(self):
        return {"last_epoch": self.last_epoch, "base_lrs": self.base_lrs}

    def load_state_dict(self, sd):
        self.last_epoch = sd["last_epoch"]
        self.base_lrs = sd["base_lrs"]
        self._apply()


def get_scheduler(name, optimizer, num_warmup_steps, num_training_steps=0, num_cycles=0.5):
    return LambdaSchedule(optimizer, name, num_warmup_steps, num_training_steps, num_cycles)


# ---------------------------------------------------------------------------------------------
# EMA (diffusers.training_utils.EMAModel semantics)
# ---------------------------------------------------------------------------------------------
class EMAModel:
    def __init__(self, parameters, decay=0.9999, min_decay=0.0, update_after_step=0, use_ema_warmup=False,
                 inv_gamma=1.0, power=2.0 / 3.0, model_cls=None, model_config=None, **_):
        parameters = list(parameters)
        self._model = getattr(parameters[0], "_mdm_model", None) if parameters else None
        if self._model is not None:
            self.flat = self._model.flat_param.detach().clone()
            self.shadow_params = [self.flat[s.offset:s.offset + s.numel].view(s.shape) for s in self._model._specs]
        else:
            self.flat = None
            self.shadow_params = [p.clone().detach() for p in parameters]
        self.temp_stored_params = None
        self.decay, self.min_decay, self.update_after_step = decay, min_decay, update_after_step
        self.use_ema_warmup, self.inv_gamma, self.power = use_ema_warmup, inv_gamma, power
        self.optimization_step = 0
        self.cur_decay_value = None
        self.model_cls, self.model_config = model_cls, model_config
        self._fused_done = False

    def get_decay(self, optimization_step: int) -> float:
        step = max(0, optimization_step - self.update_after_step - 1)
        if step <= 0:
            return 0.0
        cur = 1 - (1 + step / self.inv_gamma) ** -self.power if self.use_ema_warmup else (1 + step) / (10 + step)
        return max(min(cur, self.decay), self.min_decay)

    @torch.no_grad()
    def step(self, parameters):
        self.optimization_step += 1
        decay = self.get_decay(self.optimization_step)
        self.cur_decay_value = decay
        if self._fused_done:              # the fused optimiser kernel already applied this update
            self._fused_done = False
            return
        one_minus = 1 - decay
        if self.flat is not None:
            self.flat.sub_(one_minus * (self.flat - self._model.flat_param))
            return
        for s, p in zip(self.shadow_params, list(parameters)):
            if p.requires_grad:
                s.sub_(one_minus * (s - p.to(s.dtype)))
            else:
                s.copy_(p)

    def copy_to(self, parameters):
        if self.flat is not None:
            self._model.flat_param.data.copy_(self.flat)
            self._model.mark_params_updated()
            return
        for s, p in zip(self.shadow_params, list(parameters)):
            p.data.copy_(s.to(p.device).data)

    def store(self, parameters):
        if self.flat is not None:
            self.temp_stored_params = self._model.flat_param.detach().clone()
        else:
            self.temp_stored_params = [p.detach().cpu().clone() for p in parameters]

    def restore(self, parameters):
        if self.temp_stored_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        if self.flat is not None:
            self._model.flat_param.data.copy_(self.temp_stored_params)
            self._model.mark_params_updated()
        else:
            for c, p in zip(self.temp_stored_params, parameters):
                p.data.copy_(c.data)
        self.temp_stored_params = None

    def to(self, device=None, dtype=None):
        if self.flat is None:
            self.shadow_params = [p.to(device=device, dtype=dtype) if p.is_floating_point() else p.to(device=device)
                                  for p in self.shadow_params]
        return self

    def state_dict(self):
        return {"decay": self.decay, "min_decay": self.min_decay, "optimization_step": self.optimization_step,
                "update_after_step": self.update_after_step, "use_ema_warmup": self.use_ema_warmup,
                "inv_gamma": self.inv_gamma, "power": self.power, "shadow_params": [p.clone() for p in self.shadow_params]}

    def load_state_dict(self, sd):
        for k in ("decay", "min_decay", "optimization_step", "update_after_step", "use_ema_warmup", "inv_gamma", "power"):
            if k in sd:
                setattr(self, k, sd[k])
        sp = sd.get("shadow_params")
        if sp is not None:
            for s, n in zip(self.shadow_params, sp):
                s.copy_(n.to(s.device))

    def save_pretrained(self, path):
        if self.model_cls is None:
            raise ValueError("`save_pretrained` can only be used if `model_cls` was defined at __init__.")
        model = self.model_cls.from_config(self.model_config if isinstance(self.model_config, dict) else vars(self.model_config))
        sd = self.state_dict()
        sd.pop("shadow_params")
        model.register_to_config(**sd)
        self.copy_to(model.parameters()) if self.flat is None else model.flat_param.data.copy_(self.flat)
        model.mark_params_updated() if hasattr(model, "mark_params_updated") else None
        model.save_pretrained(path)

    @classmethod
    def from_pretrained(cls, path, model_cls):
        cfg = model_cls.load_config(path)
        model = model_cls.from_pretrained(path)
        ema = cls(model.parameters(), model_cls=model_cls, model_config=model.config)
        ema.load_state_dict({k: cfg[k] for k in ("decay", "min_decay", "optimization_step", "update_after_step",
                                                 "use_ema_warmup", "inv_gamma", "power") if k in cfg})
        return ema


# ---------------------------------------------------------------------------------------------
# fused optimiser
# ---------------------------------------------------------------------------------------------
class FusedOptimizer:
    """torch.optim-shaped optimiser over the denoiser's flat parameter buffer: one kernel does
    clip + Adam/AdamW/SGD (+ EMA, + bf16 weight mirror).  csrc/optim.cu."""

    def __init__(self, model, name="adamw", lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=None):
        from . import optim_ops
        self._ops = optim_ops
        self.model = model
        self.name = name.lower()
        if weight_decay is None:
            weight_decay = 0.01 if self.name == "adamw" else 0.0
        self.param_groups = [dict(params=model.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]
        self.m = torch.zeros_like(model.flat_param) if self.name != "sgd" else None
        self.v = torch.zeros_like(model.flat_param) if self.name != "sgd" else None
        self.step_count = 0
        self._ws = torch.empty(1024, dtype=torch.float32, device=model.device)
        self._gnorm_sq = torch.zeros(1, dtype=torch.float32, device=model.device)
        self._hyper = torch.zeros(4, dtype=torch.float32, device=model.device)
        self._clip = 0.0
        self.ema = None
        self.grad_scale = 1.0

    def attach_ema(self, ema: EMAModel):
        if ema is not None and ema.flat is not None:
            self.ema = ema

    def set_clip(self, max_norm: float):
        """global-norm clipping fused into the next step(); returns the (lazy, device) total norm"""
        self._ops.grad_sumsq(self.model.flat_grad, self._ws, self._gnorm_sq)
        self._clip = float(max_norm)
        return self._gnorm_sq.sqrt() * self.grad_scale

    def upload_hyper(self):
        """host part of a step: advance the step count and send (lr, bias corrections, EMA decay) to the
        device.  Kept apart from `launch()` so that `launch()` can live inside a replayed CUDA graph."""
        g = self.param_groups[0]
        self.step_count += 1
        b1, b2 = g["betas"]
        ema_decay = 0.0
        if self.ema is not None:
            ema_decay = self.ema.get_decay(self.ema.optimization_step + 1)
            self.ema._fused_done = True
        # the scalars torch.optim.Adam derives per step, in double, rounded to fp32 once: step_size, sqrt(bias_correction2)
        bc1, bc2 = 1.0 - b1 ** self.step_count, 1.0 - b2 ** self.step_count
        self._hyper.copy_(torch.tensor([g["lr"], g["lr"] / bc1, math.sqrt(bc2), ema_decay], dtype=torch.float32),
                          non_blocking=True)

    def launch(self, clip=None):
        """device part of a step (one kernel): clip + Adam/AdamW/SGD + EMA + bf16 mirror"""
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        clip = self._clip if clip is None else clip
        self._ops.adam_ema_step_dev(self.model.flat_param, self.model.flat_grad, self.m, self.v,
                                    self.ema.flat if self.ema is not None else None, self.model.flat_bf16,
                                    self._gnorm_sq if clip > 0 else None, self._hyper, b1, b2, g["eps"],
                                    g["weight_decay"], clip, self.grad_scale, self._ops.MODE[self.name])
        # the kernel refreshed the bf16 mirror itself
        self.model._bf16_stale = False
        self.model._bf16_version = self.model.flat_param._version

    def step(self):
        self.upload_hyper()
        self.launch()
        self._clip = 0.0

    def zero_grad(self, set_to_none=False):
        self.model.flat_grad.zero_()

    def state_dict(self):
        return {"step": self.step_count, "m": self.m, "v": self.v,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        if "step" not in sd or ("state" in sd and "m" not in sd):
            # torch.optim layout ({"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups"}) as written by the
            # reference's accelerate run: the per-parameter moments follow diffusers' module order, which this
            # framework cannot verify offline -- refuse loudly instead of loading moments into the wrong slots.
            raise RuntimeError("FusedOptimizer.load_state_dict: expected this framework's flat layout "
                               "{'step', 'm', 'v', 'param_groups'}; a torch.optim state dict (reference-produced "
                               "optimizer.bin) is not interchangeable -- only unet/ and unet_ema/ are (INTEGRATION.md)")
        self.step_count = sd["step"]
        if self.m is not None:
            self.m.copy_(sd["m"]); self.v.copy_(sd["v"])
        for g, n in zip(self.param_groups, sd["param_groups"]):
            g.update(n)


# ---------------------------------------------------------------------------------------------
# gradient all-reduce over NVLink peer memory (csrc/allreduce.cu)
# ---------------------------------------------------------------------------------------------
class P2PAllReduce:
    """Maps every rank's flat gradient buffer into every process (CUDA IPC; handles exchanged through the process
    group) and reduces ranges of it with ONE hand-written kernel per range -- capturable inside the training-step
    graph, small blocks that co-reside with the persistent GEMM CTAs (NCCL's channels wait for free SMs: the
    round-1 all-reduce stayed exposed).  SUM semantics, like the NCCL path: the fused optimiser folds in 1/world."""

    def __init__(self, model, rank, world, device):
        from . import comm_ops
        self._ops = comm_ops
        if world > comm_ops.MAX_RANKS:
            raise RuntimeError(f"P2PAllReduce: at most {comm_ops.MAX_RANKS} ranks (one NVSwitch domain)")
        self.rank, self.world, self.device = rank, world, device
        self.blocks = int(os.environ.get("MDM_P2P_BLOCKS", "48"))
        # "sm": one kernel per range moves everything with peer loads / stores; "ce": the copy engines move the bytes,
        # the SMs only synchronise and reduce locally (csrc/allreduce.cu)
        self.mode = os.environ.get("MDM_P2P_MODE", "ce").lower()
        # copy nodes of one range run back to back on ONE stream by default: measured on 8 GPUs 94.7 % weak scaling with
        # 1 stream, 94.1 % with 3, 88.3 % with 7 (concurrent peer copies compete with the backward for HBM / the fabric)
        self.n_sub = max(1, int(os.environ.get("MDM_P2P_COPY_STREAMS", "1")))
        self.ce_min_bytes = int(os.environ.get("MDM_P2P_CE_MIN_BYTES", str(32 << 20)))
        self._subs = None
        self.staging = None
        n = model.numel_flat
        # a dedicated allocation for the gradients (the IPC handle covers a whole cudaMalloc segment)
        self.buf = torch.zeros(n, dtype=torch.float32, device=device)
        self.flags = torch.zeros(comm_ops.flag_words(), dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        mine = (comm_ops.ipc_export(self.buf), comm_ops.ipc_export(self.flags))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        self.comm = comm_ops.P2PCommStruct()
        self.comm.rank, self.comm.world = rank, world
        self._opened = []
        for p in range(world):
            if p == rank:
                self.comm.buf[p], self.comm.flag[p] = self.buf.data_ptr(), self.flags.data_ptr()
            else:
                (hb, ob), (hf, of) = everyone[p]
                self.comm.buf[p] = comm_ops.ipc_open(hb, ob)
                self._opened.append((self.comm.buf[p], ob))
                self.comm.flag[p] = comm_ops.ipc_open(hf, of)
                self._opened.append((self.comm.flag[p], of))
        model.rehome_grad(self.buf)
        dist.barrier()

    def close(self):
        """collective: unmap the peers' buffers (before any rank frees or re-exports its own)"""
        if self._opened is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier()
        for ptr_, off in self._opened:
            self._ops.ipc_close(ptr_, off)
        self._opened = None
        dist.barrier()

    def all_reduce(self, lo, hi, exposed=False):
        """enqueue the SUM all-reduce of flat_grad[lo:hi] on the current stream.  `exposed`: nothing else runs
        meanwhile (the last range of a step): the one-kernel variant then uses every block it may"""
        # big ranges: copy engines (no SM interference with the co-running backward); small ranges: the one-kernel
        # variant (two in-kernel flag exchanges instead of 2 (world - 1) copy nodes: lower latency, and the final
        # range of a step is exposed)
        if self.mode == "ce" and 4 * (hi - lo) >= self.ce_min_bytes:
            return self._all_reduce_ce(lo, hi)
        self._ops.p2p_allreduce(self.comm, lo, hi - lo, 128 if exposed else self.blocks, self.device)

    def _all_reduce_ce(self, lo, hi):
        ops, W, r, dev = self._ops, self.world, self.rank, self.device
        if self.staging is None:
            self._stride = ((self.buf.numel() + W - 1) // W + 3) // 4 * 4           # floats per peer copy (largest slice)
            self.staging = torch.empty((W - 1) * self._stride, dtype=torch.float32, device=dev)
            self._subs = [torch.cuda.Stream(device=dev) for _ in range(min(self.n_sub, W - 1))]
        n4 = (hi - lo) // 4
        slice_ = (n4 + W - 1) // W * 4                                              # floats per rank
        s0 = lo + r * slice_
        my_n = max(0, min(s0 + slice_, hi) - s0)
        cs = torch.cuda.current_stream(dev)
        ops.p2p_barrier(self.comm, 0, dev)                                          # every rank's gradients of the range are final
        if my_n > 0:
            mine = self.buf.data_ptr() + 4 * s0

            def copies(make):
                for sub in self._subs:
                    sub.wait_stream(cs)
                for k in range(W - 1):
                    with torch.cuda.stream(self._subs[k % len(self._subs)]):
                        make(k, (r + 1 + k) % W)
                for sub in self._subs:
                    cs.wait_stream(sub)
            # pull the peers' copies of this rank's slice (copy engines), reduce locally in rank order, push the result
            copies(lambda k, p: ops.memcpy_async(self.staging.data_ptr() + 4 * k * self._stride, self.comm.buf[p] + 4 * s0, 4 * my_n, dev))
            ops.reduce_slices(mine, self.staging.data_ptr(), self._stride, my_n, r, W, dev)
            copies(lambda k, p: ops.memcpy_async(self.comm.buf[p] + 4 * s0, mine, 4 * my_n, dev))
        ops.p2p_barrier(self.comm, 1, dev)                                          # all pushes landed; nobody reads this buffer any more


# ---------------------------------------------------------------------------------------------
# Accelerator
# ---------------------------------------------------------------------------------------------
class InitProcessGroupKwargs:
    def __init__(self, timeout=timedelta(seconds=1800), **_):
        self.timeout = timeout


class _ShardedLoader:
    """every `world`-th batch of the wrapped loader, starting at `rank` (accelerate's BatchSamplerShard)"""

    def __init__(self, loader, rank, world, device):
        self.loader, self.rank, self.world, self.device = loader, rank, world, device

    def __len__(self):
        return len(self.loader) // self.world

    def __iter__(self):
        n = len(self) * self.world
        for i, batch in enumerate(self.loader):
            if i >= n:
                break
            if i % self.world == self.rank:
                yield _to_device(batch, self.device)


def _to_device(batch, device):
    if torch.is_tensor(batch):
        return batch.to(device, non_blocking=True)
    if isinstance(batch, (list, tuple)):
        return type(batch)(_to_device(b, device) for b in batch)
    if isinstance(batch, dict):
        return {k: _to_device(v, device) for k, v in batch.items()}
    return batch


class Accelerator:
    def __init__(self, gradient_accumulation_steps=1, mixed_precision="no", kwargs_handlers=None, device=None, **_):
        self.gradient_accumulation_steps = int(gradient_accumulation_steps)
        self.mixed_precision = mixed_precision or "no"
        self.rank = int(os.environ.get("RANK", 0))
        self.num_processes = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        use_cuda = torch.cuda.is_available() and device != "cpu"
        if use_cuda:
            torch.cuda.set_device(self.local_rank)
            self.device = torch.device("cuda", self.local_rank)
        else:
            self.device = torch.device("cpu")
        timeout = timedelta(seconds=1800)
        for h in kwargs_handlers or []:
            timeout = getattr(h, "timeout", timeout)
        if self.num_processes > 1 and not dist.is_initialized():
            dist.init_process_group(backend="nccl" if use_cuda else "gloo", timeout=timeout)
        self.sync_gradients = True
        self._accum = 0
        self._models, self._optimizers, self._schedulers = [], [], []
        self._save_hooks, self._load_hooks = [], []
        self.p2p = None                   # P2PAllReduce of the (single) B200 denoiser, data parallel on CUDA

    # -- properties ----------------------------------------------------------------------------
    @property
    def is_main_process(self):
        return self.rank == 0

    @property
    def is_local_main_process(self):
        return self.local_rank == 0

    def print(self, *a, **k):
        if self.is_local_main_process:
            print(*a, **k)

    def init_trackers(self, *a, **k):
        pass

    def wait_for_everyone(self):
        # The reference barriers every batch (trainer_masked.py:166); the gradient all-reduce already
        # orders the ranks, so this is a no-op kept for API parity.
        pass

    def barrier(self):
        if self.num_processes > 1:
            dist.barrier()

    # -- prepare ---------------------------------------------------------------------------------
    def prepare(self, *objs):
        out = []
        for o in objs:
            if hasattr(o, "flat_param") or isinstance(o, torch.nn.Module):
                self._models.append(o)
                if isinstance(o, torch.nn.Module):
                    o.to(self.device)
                self._broadcast_params(o)
                self._setup_p2p(o)
                out.append(o)
            elif isinstance(o, (FusedOptimizer, torch.optim.Optimizer)):
                self._optimizers.append(o)
                if isinstance(o, FusedOptimizer):
                    o.grad_scale = 1.0 / self.num_processes if self.num_processes > 1 else 1.0
                out.append(o)
            elif hasattr(o, "get_last_lr"):
                if hasattr(o, "steps_per_call"):
                    o.steps_per_call = self.num_processes
                self._schedulers.append(o)
                out.append(o)
            elif hasattr(o, "__iter__") and hasattr(o, "__len__"):
                out.append(_ShardedLoader(o, self.rank, self.num_processes, self.device) if self.num_processes > 1 or self.device.type == "cuda" else o)
            else:
                out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def _setup_p2p(self, model):
        """data parallel on CUDA with the B200 denoiser: gradients are reduced by the peer-memory kernel
        (MDM_DP_ALLREDUCE=nccl keeps the NCCL all-reduce; a failed IPC set-up falls back to it on every rank)"""
        if (self.num_processes <= 1 or self.device.type != "cuda" or not hasattr(model, "rehome_grad") or self.p2p is not None
                or os.environ.get("MDM_DP_ALLREDUCE", "p2p").lower() != "p2p"):
            return
        ok = torch.ones(1, device=self.device)
        p2p = None
        try:
            p2p = P2PAllReduce(model, self.rank, self.num_processes, self.device)
        except Exception as e:                                   # e.g. IPC not permitted in this container
            print(f"[mdm_b200] rank {self.rank}: peer-memory all-reduce unavailable ({e!r}); using NCCL", flush=True)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)                # all ranks or none
        self.p2p = p2p if ok.item() > 0 else None

    def close(self):
        """collective tear-down of the peer mappings (call before dropping the model of a data-parallel run)"""
        if self.p2p is not None:
            self.p2p.close()
            self.p2p = None

    def _broadcast_params(self, model):
        if self.num_processes <= 1:
            return
        if hasattr(model, "flat_param"):
            dist.broadcast(model.flat_param.data, src=0)
            model.mark_params_updated()
        else:
            for p in model.parameters():
                dist.broadcast(p.data, src=0)

    # -- training step ---------------------------------------------------------------------------
    @contextlib.contextmanager
    def accumulate(self, model=None):
        self._accum += 1
        self.sync_gradients = self._accum % self.gradient_accumulation_steps == 0
        yield

    def backward(self, loss):
        if self.gradient_accumulation_steps > 1:
            loss = loss / self.gradient_accumulation_steps
        loss.backward()
        if self.sync_gradients and not getattr(self, "_defer_all_reduce", False):
            self.all_reduce_gradients()

    def all_reduce_gradients(self):
        if self.num_processes <= 1:
            return
        fused = any(isinstance(o, FusedOptimizer) for o in self._optimizers)
        for m in self._models:
            if hasattr(m, "flat_grad"):
                if self.p2p is not None and m.flat_grad.data_ptr() == self.p2p.buf.data_ptr():
                    self.p2p.all_reduce(0, m.flat_grad.numel())
                else:
                    dist.all_reduce(m.flat_grad, op=dist.ReduceOp.SUM)
                if not fused:                       # the fused optimiser folds 1/world into its kernel
                    m.flat_grad.mul_(1.0 / self.num_processes)
            else:
                for p in m.parameters():
                    if p.grad is not None:
                        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
                        p.grad.mul_(1.0 / self.num_processes)

    def all_reduce_async(self, t):
        """SUM all-reduce of a slice of the flat gradient on NCCL's stream; `.wait()` orders the current stream after it"""
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)

    # -- overlapped all-reduce of the flat gradient (B200 denoiser): the range finished first by the backward pass
    #    is reduced asynchronously while the rest of the backward still runs
    def start_late_all_reduce(self, model, lo, hi):
        self._late = (model, lo, hi, dist.all_reduce(model.flat_grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def finish_all_reduce(self, model):
        late = getattr(self, "_late", None)
        if late is None or late[0] is not model:
            dist.all_reduce(model.flat_grad, op=dist.ReduceOp.SUM)
            return
        _, lo, hi, work = late
        self._late = None
        if lo > 0:
            dist.all_reduce(model.flat_grad[:lo], op=dist.ReduceOp.SUM)
        if hi < model.flat_grad.numel():
            dist.all_reduce(model.flat_grad[hi:], op=dist.ReduceOp.SUM)
        work.wait()

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        for o in self._optimizers:
            if isinstance(o, FusedOptimizer):
                return o.set_clip(max_norm)
        return torch.nn.utils.clip_grad_norm_(list(parameters), max_norm, norm_type=norm_type)

    # -- checkpoints (accelerate layout) ---------------------------------------
    def register_save_state_pre_hook(self, hook):
        self._save_hooks.append(hook)

    def register_load_state_pre_hook(self, hook):
        self._load_hooks.append(hook)

    def save_state(self, output_dir):
        os.makedirs(output_dir, exist_ok=True)
        models = list(self._models)
        weights = [m.state_dict() for m in models]
        for h in self._save_hooks:
            h(models, weights, output_dir)
        if self.is_main_process:
            for i, w in enumerate(weights):      # whatever the hooks did not pop
                torch.save(w, os.path.join(output_dir, "pytorch_model.bin" if i == 0 else f"pytorch_model_{i}.bin"))
            for i, o in enumerate(self._optimizers):
                torch.save(o.state_dict(), os.path.join(output_dir, "optimizer.bin" if i == 0 else f"optimizer_{i}.bin"))
            for i, s in enumerate(self._schedulers):
                torch.save(s.state_dict(), os.path.join(output_dir, "scheduler.bin" if i == 0 else f"scheduler_{i}.bin"))
        states = {"random_state": random.getstate(), "numpy_random_seed": np.random.get_state(),
                  "torch_manual_seed": torch.get_rng_state()}
        if torch.cuda.is_available():
            states["torch_cuda_manual_seed"] = torch.cuda.get_rng_state_all()
        # accelerate writes this file with torch.save (checkpointing.py: save_accelerator_state)
        torch.save(states, os.path.join(output_dir, f"random_states_{self.rank}.pkl"))
        return output_dir

    def load_state(self, input_dir):
        models = list(self._models)
        for h in self._load_hooks:
            h(models, input_dir)
        for i, m in enumerate(models):           # models the hooks did not pop
            p = os.path.join(input_dir, "pytorch_model.bin" if i == 0 else f"pytorch_model_{i}.bin")
            if os.path.exists(p):
                m.load_state_dict(torch.load(p, map_location="cpu"))
        for i, o in enumerate(self._optimizers):
            p = os.path.join(input_dir, "optimizer.bin" if i == 0 else f"optimizer_{i}.bin")
            if os.path.exists(p):
                o.load_state_dict(torch.load(p, map_location=self.device, weights_only=False))
        for i, s in enumerate(self._schedulers):
            p = os.path.join(input_dir, "scheduler.bin" if i == 0 else f"scheduler_{i}.bin")
            if os.path.exists(p):
                s.load_state_dict(torch.load(p, weights_only=False))
        p = os.path.join(input_dir, f"random_states_{self.rank}.pkl")
        if os.path.exists(p):
            try:
                st = torch.load(p, weights_only=False)
            except Exception:                        # round-1 checkpoints of this framework: plain pickle
                with open(p, "rb") as f:
                    st = pickle.load(f)
            random.setstate(st["random_state"])
            np.random.set_state(st["numpy_random_seed"])
            torch.set_rng_state(st["torch_manual_seed"])
            if "torch_cuda_manual_seed" in st and torch.cuda.is_available():
                torch.cuda.set_rng_state_all(st["torch_cuda_manual_seed"])

    def save_model(self, model, save_directory, **_):
        os.makedirs(save_directory, exist_ok=True)
        torch.save(model.state_dict(), os.path.join(save_directory, "pytorch_model.bin"))
