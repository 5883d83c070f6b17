"""B200-native denoiser: the U-Net the reference builds at utils/model.py:3-33
(`diffusers.UNet2DModel`), executed entirely by the sm_100a kernels behind
include/mdm.h -- tcgen05 implicit-GEMM convolutions/linears, fused GroupNorm+SiLU, attention core.

Surface kept from diffusers (what the reference calls): `model(x, t).sample`, `.device`,
`.config`, `.parameters()`, `.train()/.eval()`, `.state_dict()/.load_state_dict()` with the
diffusers key names and NCHW fp32 tensors, `.save_pretrained()/.from_pretrained()`.

Data layout in HBM
  * parameters: ONE flat fp32 master buffer (and one flat fp32 gradient buffer of the same
    layout) + a flat bf16 mirror read by the GEMMs.  3x3/1x1 conv weights are stored packed
    [cout][kh*kw][cin] (K-major for fprop, directly the layout wgrad produces); q/k/v weights
    are adjacent so one [3C][C] GEMM does the three projections; every resnet's
    `time_emb_proj` is adjacent so ONE GEMM projects the time embedding for all 32 resnets.
    `state_dict()` converts to/from the checkpoint layout.
  * activations: NHWC bf16.  Skip connections are never copied: the producer of a skip tensor
    writes straight into its slice of the up-path concat buffer (channel stride = C_h + C_skip)
    and every consumer addresses the slice through the stride.
  * gradients of activations mirror that aliasing, so the concat backward is free as well.

The forward/backward "programs" are built once per (batch, size) and replayed; every launch is
stream-ordered with no host sync, so a whole step can be captured in a CUDA graph.
"""
from __future__ import annotations

import json
import math
import os
from types import SimpleNamespace

import torch

from . import denoiser_ops as ops

G = 32  # norm_num_groups


def default_config(dim_channel=3, dim_height=32, num_attention=1, base=128, layers_per_block=2):
    """utils/model.py:3-33 (`MyModel`) as a config dict with the diffusers field names."""
    attn_down = {1: (4,), 2: (3, 4), 3: (2, 3, 4), 4: (1, 2, 3, 4), 5: (1, 2, 3, 4, 5)}
    attn_up = {1: (1,), 2: (1, 2), 3: (1, 2, 3), 4: (1, 2, 3, 4), 5: (0, 1, 2, 3, 4)}
    if num_attention not in attn_down:
        raise NotImplementedError("not implemented")
    b = base
    return dict(
        _class_name="UNet2DModel", sample_size=dim_height, in_channels=dim_channel, out_channels=dim_channel,
        layers_per_block=layers_per_block, block_out_channels=[b, b, 2 * b, 2 * b, 4 * b, 4 * b],
        down_block_types=["AttnDownBlock2D" if i in attn_down[num_attention] else "DownBlock2D" for i in range(6)],
        up_block_types=["AttnUpBlock2D" if i in attn_up[num_attention] else "UpBlock2D" for i in range(6)],
        act_fn="silu", attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5,
        time_embedding_type="positional", flip_sin_to_cos=True, freq_shift=0,
        downsample_padding=1, add_attention=True, center_input_sample=False, dropout=0.0,
        mid_block_scale_factor=1, downsample_type="conv", upsample_type="conv",
        resnet_time_scale_shift="default", attn_norm_num_groups=None,
        class_embed_type=None, num_class_embeds=None, num_train_timesteps=None,
    )


# ---------------------------------------------------------------------------------------------
# parameter book-keeping
# ---------------------------------------------------------------------------------------------
class ParamSpec:
    __slots__ = ("name", "shape", "kind", "offset", "numel")

    def __init__(self, name, shape, kind):
        self.name, self.shape, self.kind = name, tuple(shape), kind   # kind: conv | plain
        self.numel = math.prod(shape)
        self.offset = None


class Act:
    """symbolic NHWC activation; storage assigned after the graph is known"""
    __slots__ = ("N", "H", "W", "C", "alias", "val", "grad", "has_upstream_grad", "name", "q", "no_q", "parts")

    def __init__(self, N, H, W, C, name=""):
        self.N, self.H, self.W, self.C = N, H, W, C
        self.alias = None          # (parent Act, channel offset)
        self.val = None
        self.grad = None
        self.has_upstream_grad = False   # a later consumer (the up path) already wrote into .grad
        self.name = name
        self.q = None              # _QBuf: GroupNorm quad sums of this tensor, written by its producer's conv epilogue
        self.no_q = False          # producer cannot emit them (attention output projection)
        self.parts = None          # channel concatenation: (first Act, second Act)


class _QBuf:
    """[N][C/4][2] fp32 slice of the plan's statistics arena (allocated after every emitter has asked)"""
    __slots__ = ("N", "C", "t")

    def __init__(self, N, C):
        self.N, self.C, self.t = N, C, None


class _Win:
    """batch window [b0, b0 + nb) of an activation: what one LANE of a low-resolution region works on"""
    __slots__ = ("a", "b0", "nb")

    def __init__(self, a, b0, nb):
        self.a, self.b0, self.nb = a, b0, nb

    @property
    def val(self):
        return self.a.val[self.b0:self.b0 + self.nb]

    @property
    def grad(self):
        return self.a.grad[self.b0:self.b0 + self.nb]

    def __getattr__(self, k):
        return getattr(self.a, k)


class UNet2DModelB200:
    def __init__(self, device="cuda", **config):
        cfg = default_config()
        cfg.update(config)
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        # the kernels implement the configuration `utils/model.py:3-33` builds; other values of these diffusers fields
        # would silently compute a different network, so they are refused (the fp32 oracle honours them: that is how
        # it is pinned to the reference's in-repo unet6, tests/test_oracle_unet6.py)
        fixed = dict(attention_head_dim=8, norm_num_groups=32, flip_sin_to_cos=True, freq_shift=0, downsample_padding=1,
                     act_fn="silu", time_embedding_type="positional", resnet_time_scale_shift="default",
                     downsample_type="conv", upsample_type="conv", add_attention=True, center_input_sample=False)
        for k, v in fixed.items():
            if cfg.get(k, v) != v:
                raise NotImplementedError(f"UNet2DModelB200: config field {k}={cfg[k]!r} is not implemented (only {v!r})")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("UNet2DModelB200 runs on a CUDA device only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.training = True
        self.dtype = torch.float32
        self._specs = []           # ordered ParamSpec list (flat layout order)
        self._by_name = {}
        self._build_layers()
        self._alloc_params()
        self._plans = {}
        self.reset_parameters()

    # -- structure ------------------------------------------------------------------------------
    def _p(self, name, shape, kind="plain"):
        s = ParamSpec(name, shape, kind)
        self._specs.append(s)
        self._by_name[name] = s
        return name

    def _build_layers(self):
        cfg = self._cfg
        boc = cfg["block_out_channels"]
        n = cfg["layers_per_block"]
        temb = boc[0] * 4
        self.temb_dim = temb
        L = SimpleNamespace()
        L.resnets = []             # in execution order, for the batched time-embedding projection
        self._p("conv_in.weight", (boc[0], cfg["in_channels"], 3, 3))
        self._p("conv_in.bias", (boc[0],))
        self._p("time_embedding.linear_1.weight", (temb, boc[0]))
        self._p("time_embedding.linear_1.bias", (temb,))
        self._p("time_embedding.linear_2.weight", (temb, temb))
        self._p("time_embedding.linear_2.bias", (temb,))

        def resnet(prefix, cin, cout):
            r = SimpleNamespace(prefix=prefix, cin=cin, cout=cout, shortcut=cin != cout)
            self._p(f"{prefix}.norm1.weight", (cin,)); self._p(f"{prefix}.norm1.bias", (cin,))
            self._p(f"{prefix}.conv1.weight", (cout, 9, cin), "conv"); self._p(f"{prefix}.conv1.bias", (cout,))
            self._p(f"{prefix}.norm2.weight", (cout,)); self._p(f"{prefix}.norm2.bias", (cout,))
            self._p(f"{prefix}.conv2.weight", (cout, 9, cout), "conv"); self._p(f"{prefix}.conv2.bias", (cout,))
            if r.shortcut:
                self._p(f"{prefix}.conv_shortcut.weight", (cout, 1, cin), "conv"); self._p(f"{prefix}.conv_shortcut.bias", (cout,))
            L.resnets.append(r)
            return r

        def attention(prefix, c):
            a = SimpleNamespace(prefix=prefix, c=c)
            self._p(f"{prefix}.group_norm.weight", (c,)); self._p(f"{prefix}.group_norm.bias", (c,))
            for nm in ("to_q", "to_k", "to_v"):          # adjacent -> one [3C][C] matrix
                self._p(f"{prefix}.{nm}.weight", (c, c))
            for nm in ("to_q", "to_k", "to_v"):
                self._p(f"{prefix}.{nm}.bias", (c,))
            self._p(f"{prefix}.to_out.0.weight", (c, c)); self._p(f"{prefix}.to_out.0.bias", (c,))
            return a

        L.down = []
        c = boc[0]
        for i, ty in enumerate(cfg["down_block_types"]):
            cin, c = c, boc[i]
            blk = SimpleNamespace(resnets=[], attns=[], down=None)
            for j in range(n):
                blk.resnets.append(resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c))
            if ty.startswith("Attn"):
                blk.attns = [attention(f"down_blocks.{i}.attentions.{j}", c) for j in range(n)]
            if i != len(boc) - 1:
                blk.down = f"down_blocks.{i}.downsamplers.0.conv"
                self._p(blk.down + ".weight", (c, 9, c), "conv"); self._p(blk.down + ".bias", (c,))
            L.down.append(blk)
        L.mid = SimpleNamespace()
        L.mid.res0 = resnet("mid_block.resnets.0", boc[-1], boc[-1])
        L.mid.attn = attention("mid_block.attentions.0", boc[-1])
        L.mid.res1 = resnet("mid_block.resnets.1", boc[-1], boc[-1])
        L.up = []
        rev = list(reversed(boc))
        c = rev[0]
        for i, ty in enumerate(cfg["up_block_types"]):
            cprev, c = c, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            blk = SimpleNamespace(resnets=[], attns=[], up=None, skip_c=[])
            for j in range(n + 1):
                skip = cin if j == n else c
                rin = cprev if j == 0 else c
                blk.resnets.append(resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, c))
                blk.skip_c.append(skip)
            if ty.startswith("Attn"):
                blk.attns = [attention(f"up_blocks.{i}.attentions.{j}", c) for j in range(n + 1)]
            if i != len(boc) - 1:
                blk.up = f"up_blocks.{i}.upsamplers.0.conv"
                self._p(blk.up + ".weight", (c, 9, c), "conv"); self._p(blk.up + ".bias", (c,))
            L.up.append(blk)
        self._p("conv_norm_out.weight", (boc[0],)); self._p("conv_norm_out.bias", (boc[0],))
        self._p("conv_out.weight", (cfg["out_channels"], boc[0], 3, 3)); self._p("conv_out.bias", (cfg["out_channels"],))
        # time_emb_proj of every resnet, adjacent (ONE GEMM for all of them)
        off = 0
        for r in L.resnets:
            self._p(f"{r.prefix}.time_emb_proj.weight", (r.cout, temb))
            r.tproj_off = off
            off += r.cout
        for r in L.resnets:
            self._p(f"{r.prefix}.time_emb_proj.bias", (r.cout,))
        L.tproj_total = off
        # flat layout: the projections sit right behind the time-embedding MLP, i.e. in FRONT of the down path.  Their
        # gradient is the last one the backward finishes (it sums over every resnet), like that of conv_in and of
        # down_blocks.0: with all of them at the front the data-parallel all-reduce has ONE final range.
        tp = [s for s in self._specs if ".time_emb_proj." in s.name]
        rest = [s for s in self._specs if ".time_emb_proj." not in s.name]
        at = next(i for i, s in enumerate(rest) if s.name == "time_embedding.linear_2.bias") + 1
        self._specs = rest[:at] + tp + rest[at:]
        self.layers = L

    def _alloc_params(self):
        off = 0
        for s in self._specs:
            off = (off + 63) // 64 * 64          # 256-byte alignment of every tensor (TMA needs 16)
            s.offset = off
            off += s.numel
        self.numel_flat = (off + 63) // 64 * 64
        dev = self.device
        self.flat_param = torch.zeros(self.numel_flat, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(self.numel_flat, dtype=torch.float32, device=dev)
        self.flat_bf16 = torch.zeros(self.numel_flat, dtype=torch.bfloat16, device=dev)
        self._params = {}
        for s in self._specs:
            p = torch.nn.Parameter(self.flat_param[s.offset:s.offset + s.numel].view(s.shape), requires_grad=True)
            p.grad = self.flat_grad[s.offset:s.offset + s.numel].view(s.shape)
            p._mdm_model = self
            self._params[s.name] = p
        self._bf16_stale = True

    def rehome_grad(self, buf):
        """move the flat gradient into `buf` (a dedicated, IPC-exportable allocation: data-parallel peer-memory all-reduce);
        every p.grad becomes a view of it, plans built against the old buffer are dropped"""
        assert buf.numel() == self.numel_flat and buf.dtype == torch.float32 and buf.device == self.flat_grad.device
        buf.copy_(self.flat_grad)
        self.flat_grad = buf
        for s in self._specs:
            self._params[s.name].grad = buf[s.offset:s.offset + s.numel].view(s.shape)
        self._plans = {}

    def num_parameters(self):
        return sum(s.numel for s in self._specs)

    def late_grad_range(self):
        """[lo, hi) of the flat buffers holding mid_block, up_blocks, conv_norm_out and conv_out (contiguous in the
        layout order): the parameters whose gradients the backward pass finishes FIRST"""
        lo = self._by_name["mid_block.resnets.0.norm1.weight"].offset
        last = self._by_name["conv_out.bias"]
        return lo, (last.offset + last.numel + 63) // 64 * 64

    # views ------------------------------------------------------------------------------------
    def w32(self, name):
        s = self._by_name[name]
        return self.flat_param[s.offset:s.offset + s.numel].view(s.shape)

    def g32(self, name):
        s = self._by_name[name]
        return self.flat_grad[s.offset:s.offset + s.numel].view(s.shape)

    def w16(self, name):
        s = self._by_name[name]
        return self.flat_bf16[s.offset:s.offset + s.numel].view(s.shape)

    def _span(self, first, last, shape, which):
        a, b = self._by_name[first], self._by_name[last]
        buf = {"w16": self.flat_bf16, "w32": self.flat_param, "g32": self.flat_grad}[which]
        n = math.prod(shape)
        assert b.offset + b.numel - a.offset == n, "adjacent parameters are not contiguous"
        return buf[a.offset:a.offset + n].view(shape)

    # -- nn.Module-like surface ----------------------------------------------------------------------
    def parameters(self):
        return list(self._params.values())

    def named_parameters(self):
        return list(self._params.items())

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def to(self, *a, **k):
        return self

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()

    def mark_params_updated(self):
        """call after the fp32 master was changed outside the fused optimiser kernel (load_state_dict, EMA
        copy_to / restore, broadcast, a kernel torch cannot see).  The bf16 mirror the GEMMs read is re-cast
        IMMEDIATELY: a captured step graph contains no fp32 -> bf16 cast (the optimiser kernel writes the mirror),
        so a lazily refreshed mirror would be read stale by the next replay.  In-place torch ops on the parameters
        are detected through the tensor version counter (`params_dirty`)."""
        self._bf16_stale = True
        if not torch.cuda.is_current_stream_capturing():
            self._refresh_bf16()

    def params_dirty(self):
        """the bf16 mirror does not reflect the fp32 master (checked by the trainer before every graph replay)"""
        return self._bf16_stale or self.flat_param._version != getattr(self, "_bf16_version", None)

    def _refresh_bf16(self):
        v = self.flat_param._version
        if self._bf16_stale or v != getattr(self, "_bf16_version", None):
            ops.cast_f32_bf16(self.flat_param, self.flat_bf16)
            self._bf16_stale = False
            self._bf16_version = v

    def reset_parameters(self, seed=None):
        """PyTorch default init of the corresponding nn.Conv2d / nn.Linear / nn.GroupNorm."""
        g = torch.Generator(device="cpu")
        if seed is not None:
            g.manual_seed(seed)
        else:
            g.manual_seed(torch.initial_seed() & 0x7FFFFFFF)
        with torch.no_grad():
            for s in self._specs:
                w = self.w32(s.name)
                if s.name.endswith("norm1.weight") or s.name.endswith("norm2.weight") or s.name.endswith("group_norm.weight") or s.name == "conv_norm_out.weight":
                    w.fill_(1.0)
                elif "norm" in s.name and s.name.endswith(".bias"):
                    w.zero_()
                elif s.name.endswith(".weight"):
                    fan_in = math.prod(s.shape[1:])
                    bound = 1.0 / math.sqrt(fan_in)
                    w.copy_(((torch.rand(s.shape, generator=g) * 2 - 1) * bound).to(w.device))
                else:  # bias of conv / linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                    wname = s.name[:-4] + "weight"
                    fan_in = math.prod(self._by_name[wname].shape[1:])
                    bound = 1.0 / math.sqrt(fan_in)
                    w.copy_(((torch.rand(s.shape, generator=g) * 2 - 1) * bound).to(w.device))
        self.mark_params_updated()

    # checkpoint layout: diffusers key names, NCHW fp32
    def state_dict(self):
        out = {}
        order = self._checkpoint_order()
        for name in order:
            s = self._by_name[name]
            w = self.w32(name).detach()
            if s.kind == "conv":
                k = int(round(math.sqrt(s.shape[1])))
                w = ops.unpack_conv_weight(w, k)
            out[name] = w.clone()
        return out

    def state_dict_grads(self):
        """gradients in the checkpoint layout (tests / DDP debugging)"""
        out = {}
        for s in self._specs:
            g = self.g32(s.name).detach()
            if s.kind == "conv":
                g = ops.unpack_conv_weight(g, int(round(math.sqrt(s.shape[1]))))
            out[s.name] = g.clone()
        return out

    def _checkpoint_order(self):
        return [s.name for s in self._specs]

    def load_state_dict(self, sd, strict=True):
        missing = [n for n in self._by_name if n not in sd]
        unexpected = [n for n in sd if n not in self._by_name]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]}..., unexpected {unexpected[:5]}...")
        with torch.no_grad():
            for name, t in sd.items():
                if name not in self._by_name:
                    continue
                s = self._by_name[name]
                t = t.to(device=self.device, dtype=torch.float32)
                if s.kind == "conv":
                    t = ops.pack_conv_weight(t)
                if tuple(t.shape) != s.shape:
                    raise RuntimeError(f"size mismatch for {name}: {tuple(t.shape)} vs {s.shape}")
                self.w32(name).copy_(t)
        self.mark_params_updated()
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def register_to_config(self, **kw):
        for k, v in kw.items():
            setattr(self.config, k, v)
            self._cfg[k] = v

    def save_pretrained(self, save_directory, **_):
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        cfg = dict(self._cfg)
        cfg["_diffusers_version"] = "0.26.0"
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2)
        save_file({k: v.cpu().contiguous() for k, v in self.state_dict().items()},
                  os.path.join(save_directory, "diffusion_pytorch_model.safetensors"))

    @classmethod
    def load_config(cls, path, subfolder=None, **_):
        d = os.path.join(path, subfolder) if subfolder else path
        if os.path.isdir(d):
            d = os.path.join(d, "config.json")
        with open(d) as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config, device="cuda", **_):
        cfg = {k: v for k, v in dict(config).items() if not k.startswith("_") or k == "_class_name"}
        return cls(device=device, **cfg)

    @classmethod
    def from_pretrained(cls, path, subfolder=None, device="cuda", **_):
        from safetensors.torch import load_file
        d = os.path.join(path, subfolder) if subfolder else path
        model = cls.from_config(cls.load_config(d), device=device)
        model.load_state_dict(load_file(os.path.join(d, "diffusion_pytorch_model.safetensors")))
        return model

    # ------------------------------------------------------------------------------------------
    # plan: symbolic graph -> buffers -> forward / backward programs
    # ------------------------------------------------------------------------------------------
    def _plan(self, B, S, need_grad, tag=0):
        key = (B, S, need_grad, tag)
        if key in self._plans:
            return self._plans[key]
        P = _Plan(self, B, S, need_grad)
        self._plans[key] = P
        return P

    def _plans_for(self, B, S, need_grad):
        """one plan for the whole batch, or two half-batch plans that run concurrently on two streams: the
        low-resolution layers of this U-Net launch far fewer CTAs than there are SMs, so two independent
        half-batches overlap there (GroupNorm statistics are per sample: splitting the batch changes nothing)"""
        split = getattr(self, "batch_split", 1)     # measured (3x32x32, batch 128): 9.87 ms split vs 9.54 ms whole -> off
        if split == 2 and B % 2 == 0 and B >= 2 * getattr(self, "batch_split_min", 16) and B * S * S <= getattr(self, "batch_split_max_pixels", 1 << 19):
            if getattr(self, "_split_streams", None) is None:
                self._split_streams = [torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)]
            plans = [self._plan(B // 2, S, need_grad, tag=1), self._plan(B // 2, S, need_grad, tag=2)]
            for p in plans:
                p.static_output = True          # the halves are concatenated into a fresh tensor anyway
            return plans
        return [self._plan(B, S, need_grad)]

    def _run_plans(self, plans, fn):
        """fn(plan, index) on the current stream (one plan) or on the two split streams (fork / join)"""
        if len(plans) == 1:
            return [fn(plans[0], 0)]
        main = torch.cuda.current_stream(self.device)
        out = []
        for i, (p, s) in enumerate(zip(plans, self._split_streams)):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                out.append(fn(p, i))
        for s in self._split_streams:
            main.wait_stream(s)
        return out

    def __call__(self, sample, timestep, return_dict=True):
        return self.forward(sample, timestep)

    def forward(self, sample: torch.Tensor, timestep):
        if not sample.is_cuda:
            raise RuntimeError("denoiser input must be a CUDA tensor (no CPU fallback)")
        B, C, H, W = sample.shape
        if H != W or (H & (H - 1)) != 0:
            raise RuntimeError(f"denoiser: square power-of-two images only (got {H}x{W})")
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([float(t)], device=sample.device)
        t = t.to(device=sample.device, dtype=torch.float32)
        if t.dim() == 0:
            t = t[None]
        t = t.expand(B).contiguous()
        need_grad = self.training and torch.is_grad_enabled()
        plans = self._plans_for(B, H, need_grad)
        self._refresh_bf16()
        x = sample.float().contiguous()
        self._last_plans = plans
        if need_grad:
            # a fresh leaf created on the CURRENT stream hooks the kernels into autograd (a long-lived leaf
            # would pin its AccumulateGrad node to the construction-time stream and break graph capture)
            anchor = torch.zeros(1, device=self.device, requires_grad=True)
            out = _DenoiserFn.apply(anchor, self, plans, x, t)
        else:
            out = self._forward_plans(plans, x, t)
        return SimpleNamespace(sample=out)

    def _forward_plans(self, plans, x, t):
        n = len(plans)
        h = x.shape[0] // n
        outs = self._run_plans(plans, lambda p, i: p.run_forward(x[i * h:(i + 1) * h], t[i * h:(i + 1) * h]))
        return outs[0] if n == 1 else torch.cat(outs, 0)

    def _backward_plans(self, plans, d_out):
        n = len(plans)
        h = d_out.shape[0] // n
        seg = 0 if (getattr(self, "bwd_segmented", False) and n == 1) else None
        self._run_plans(plans, lambda p, i: p.run_backward(d_out[i * h:(i + 1) * h], seg))

    def backward_segment(self, k):
        """segment k >= 1 of the last backward (after `bwd_segmented = True` made loss.backward() stop at segment 0)"""
        self._last_plans[0].run_backward(None, k)

    def dp_cut_prefixes(self):
        """resnet prefixes at which the backward program is cut for data-parallel overlap, in backward order (the first
        is always the mid block: everything from there to the head lies contiguously at the end of the flat buffer).
        MDM_DP_CUTS overrides (comma-separated `down_blocks.<i>` indices, descending), e.g. "4,2,1"."""
        n = len(self._cfg["block_out_channels"])
        spec = os.environ.get("MDM_DP_CUTS", getattr(self, "dp_cuts", "4,2,1"))
        cuts = [int(c) for c in str(spec).split(",") if c.strip() != ""]
        cuts = [c for c in sorted(set(cuts), reverse=True) if 0 < c < n]
        return ["mid_block.resnets.0"] + [f"down_blocks.{c}.resnets.0" for c in cuts]

    def grad_segment_ranges(self):
        """[(lo, hi)] of the flat gradient finished by backward segment 0, 1, ...; the last segment finishes the
        complement.  Matches `_Plan.bwd_marks`."""
        lo0, hi0 = self.late_grad_range()
        out = [(lo0, hi0)]
        prev = lo0
        for pre in self.dp_cut_prefixes()[1:]:
            off = self._by_name[f"{pre}.norm1.weight"].offset
            out.append((off, prev))
            prev = off
        return out

    def backward(self, d_out: torch.Tensor):
        """Backpropagate d(loss)/d(sample output) (NCHW fp32) through the last training forward;
        accumulates into `flat_grad` / every `p.grad`."""
        plans = self._last_plans
        if not plans[0].need_grad:
            raise RuntimeError("backward() needs a forward in train mode with grad enabled")
        self._backward_plans(plans, d_out.float().contiguous())


class _DenoiserFn(torch.autograd.Function):
    """Makes `loss.backward()` (what the reference trainers call through accelerate) run the
    hand-written backward program; parameter gradients are accumulated into `flat_grad`."""

    @staticmethod
    def forward(ctx, anchor, model, plans, x, t):
        ctx.model, ctx.plans = model, plans
        return model._forward_plans(plans, x, t)

    @staticmethod
    def backward(ctx, d_out):
        ctx.model._backward_plans(ctx.plans, d_out.float().contiguous())
        return None, None, None, None, None


class _Plan:
    """Buffers and the two launch sequences for one (batch, size)."""

    def __init__(self, m: UNet2DModelB200, B, S, need_grad):
        self.m, self.B, self.S, self.need_grad = m, B, S, need_grad
        self.dev = m.device
        self.fwd, self.bwd = [], []
        self.acts = []
        self._scratch = {}
        # weight-gradient GEMMs only feed the optimiser: they run on a side stream, concurrently with the
        # dgrad / GroupNorm chain of the main stream (small layers leave most SMs idle otherwise)
        self.side = torch.cuda.Stream(device=self.dev) if need_grad else None
        # optional further wgrad side streams (round robin; MDM_WGRAD_STREAMS).  Measured on B200 (3x32x32, batch 128):
        # 1 stream 7.87 ms/step, 2: 7.92, 3: 7.99, 4: 7.96 -- the weight gradients are not on the critical path
        # (the forward / dgrad / GroupNorm chain is), so one side stream is the default.
        n_side = int(os.environ.get("MDM_WGRAD_STREAMS", getattr(m, "wgrad_streams", 1)))
        self.sides = ([self.side] + [torch.cuda.Stream(device=self.dev) for _ in range(max(1, n_side) - 1)]) if need_grad else []
        self._side_rr = 0
        self._sides_used = set()
        self._side_readers = {}        # scratch tag -> event recorded after its last side-stream reader
        # LANES (opt-in, MDM_LANES=2|4 or model.lanes): the low-resolution part of the network can run as independent
        # batch windows on separate streams (GroupNorm statistics are per sample: exact), each with its own wgrad
        # side stream.  Measured on B200 (3x32x32, batch 128): 8.37 ms whole batch, 8.56 ms with 2 lanes, 8.71 ms
        # with 4 -- those layers are bound by streaming their WEIGHTS (split-K already spreads every layer over all
        # SMs), and every lane streams them again.  Kept for small-batch serving shapes; default off.
        self.lanes = int(os.environ.get("MDM_LANES", getattr(m, "lanes", 1)))
        if self.lanes < 1 or B % max(1, self.lanes) != 0 or B // max(1, self.lanes) < 8:
            self.lanes = 1
        self.lane_rows = int(os.environ.get("MDM_LANE_ROWS", getattr(m, "lane_rows", 4096)))   # split ops with B*H*W <= this
        self.lane_streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.lanes)] if self.lanes > 1 else []
        self.lane_sides = [torch.cuda.Stream(device=self.dev) for _ in range(self.lanes)] if (self.lanes > 1 and need_grad) else []
        self._cur_lane = None
        self._lane_side_used = [False] * max(1, self.lanes)
        self._qbufs = []
        self._fwd_graphs, self._fwd_calls = {}, 0       # inference graphs by SM reservation (grids are baked in)
        self._infer_graph_ok = bool(int(os.environ.get("MDM_INFER_GRAPH", "1"))) and bool(getattr(m, "graph_inference", True))
        self.fused_stats = bool(int(os.environ.get("MDM_GN_FUSED_STATS", "1")))
        self._build()

    # -- storage helpers ---------------------------------------------------------------------------
    def act(self, H, C, name=""):
        a = Act(self.B, H, H, C, name)
        self.acts.append(a)
        return a

    def new(self, shape, dtype=torch.bfloat16, zero=False):
        return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.dev)

    def scratch(self, tag, shape, dtype=torch.bfloat16):
        """backward temporaries, shared by all layers (sized for the largest user)"""
        n = math.prod(shape)
        key = (tag, dtype)
        cur = self._scratch.get(key)
        if cur is None or cur.numel() < n:
            self._scratch[key] = cur = torch.empty(n, dtype=dtype, device=self.dev)
        return (tag, dtype, tuple(shape))

    def sget(self, ref):
        tag, dtype, shape = ref
        return self._scratch[(tag, dtype)][:math.prod(shape)].view(shape)

    # -- side stream (wgrad) ---------------------------------------------------------------------------
    def on_side(self, fn, reads=()):
        """run `fn` (weight-gradient launches) on the side stream after everything enqueued so far on the
        main stream; `reads` = scratch tags the launches read (guarded against the next main-stream writer)"""
        if not getattr(self.m, "wgrad_side_stream", True):
            fn()
            return
        main = torch.cuda.current_stream(self.dev)
        if self._cur_lane is None:
            side = self.sides[self._side_rr % len(self.sides)]
            self._side_rr += 1
            self._sides_used.add(side)
        else:
            side = self.lane_sides[self._cur_lane]
            self._lane_side_used[self._cur_lane] = True
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            fn()
            if reads:
                done = torch.cuda.Event()
                done.record(side)
                for tag in reads:
                    self._side_readers[tag] = done

    def before_write(self, *tags):
        """main stream is about to overwrite these scratch buffers: wait for their side-stream readers"""
        main = torch.cuda.current_stream(self.dev)
        for tag in tags:
            ev = self._side_readers.pop(tag, None)
            if ev is not None:
                main.wait_event(ev)

    def join_side(self):
        if self.side is not None:
            cur = torch.cuda.current_stream(self.dev)
            for s in self.sides:
                if s is self.side or s in self._sides_used:     # (never wait on a stream that did not join this capture)
                    cur.wait_stream(s)
            self._sides_used.clear()
            self._side_rr = 0
            self._side_readers.clear()

    def _run_lanes(self, progs):
        """progs[l] = launches of lane l: fork the lane streams off the current one, enqueue, join (their wgrad side
        streams join too, so nothing of the region is in flight afterwards)"""
        main = torch.cuda.current_stream(self.dev)
        for l, prog in enumerate(progs):
            s = self.lane_streams[l]
            s.wait_stream(main)
            self._cur_lane = l
            self._lane_side_used[l] = False
            try:
                with torch.cuda.stream(s):
                    for op in prog:
                        op()
                    if self._lane_side_used[l]:      # (never wait on a stream that did not join this capture)
                        s.wait_stream(self.lane_sides[l])
            finally:
                self._cur_lane = None
        for l in range(len(progs)):
            main.wait_stream(self.lane_streams[l])

    def _materialise(self):
        for a in self.acts:
            if a.alias is None and a.val is None:
                a.val = self.new((a.N, a.H, a.W, a.C))
                if self.need_grad:
                    a.grad = self.new((a.N, a.H, a.W, a.C))
        for a in self.acts:
            if a.alias is not None:
                parent, c0 = a.alias
                a.val = parent.val[..., c0:c0 + a.C]
                if self.need_grad:
                    a.grad = parent.grad[..., c0:c0 + a.C]

    # -- graph -------------------------------------------------------------------------------------
    def _build(self):
        m, B, S = self.m, self.B, self.S
        L = m.layers
        cfg = m._cfg
        boc = cfg["block_out_channels"]
        eps = cfg["norm_eps"]
        self.eps = eps
        ng = self.need_grad
        dev = self.dev
        temb = m.temb_dim
        # ---- time embedding ------------------------------------------------------------------------
        self.t_in = self.new((B,), torch.float32)
        self.x_in = self.new((B, cfg["in_channels"], S, S), torch.float32)
        self.out = self.new((B, cfg["out_channels"], S, S), torch.float32)
        te0 = self.new((B, boc[0]))
        e1_f32 = self.new((B, temb), torch.float32)
        e1_act = self.new((B, temb))
        e2_f32 = self.new((B, temb), torch.float32)
        e2_act = self.new((B, temb))
        tproj = self.new((B, L.tproj_total), torch.float32)
        self.tproj = tproj
        first, last = L.resnets[0].prefix, L.resnets[-1].prefix
        w_tp = m._span(f"{first}.time_emb_proj.weight", f"{last}.time_emb_proj.weight", (L.tproj_total, temb), "w16")
        b_tp = m._span(f"{first}.time_emb_proj.bias", f"{last}.time_emb_proj.bias", (L.tproj_total,), "w32")
        f = self.fwd
        f.append(lambda: ops.timestep_embedding(self.t_in, te0, B, boc[0]))
        f.append(lambda: ops.conv_fprop(te0, m.w16("time_embedding.linear_1.weight"), None, B, 1, 1, 1, 1,
                                        bias=m.w32("time_embedding.linear_1.bias"), y_f32=e1_f32, cout=temb))
        f.append(lambda: ops.silu_fwd(e1_f32, e1_act))
        f.append(lambda: ops.conv_fprop(e1_act, m.w16("time_embedding.linear_2.weight"), None, B, 1, 1, 1, 1,
                                        bias=m.w32("time_embedding.linear_2.bias"), y_f32=e2_f32, cout=temb))
        f.append(lambda: ops.silu_fwd(e2_f32, e2_act))
        f.append(lambda: ops.conv_fprop(e2_act, w_tp, None, B, 1, 1, 1, 1, bias=b_tp, y_f32=tproj, cout=L.tproj_total))
        if ng:
            self.d_tproj = self.new((B, L.tproj_total), torch.float32)
            d_tproj16 = self.new((B, L.tproj_total))
            d_e2act = self.new((B, temb), torch.float32)
            d_e2 = self.new((B, temb))
            d_e1act = self.new((B, temb), torch.float32)
            d_e1 = self.new((B, temb))
            g_tp = m._span(f"{first}.time_emb_proj.weight", f"{last}.time_emb_proj.weight", (L.tproj_total, 1, temb), "g32")
            gb_tp = m._span(f"{first}.time_emb_proj.bias", f"{last}.time_emb_proj.bias", (L.tproj_total,), "g32")

            def temb_bwd():
                ops.cast_f32_bf16(self.d_tproj, d_tproj16)
                self.on_side(lambda: ops.conv_wgrad(e2_act, d_tproj16, g_tp, B, 1, 1, 1, 1, dbias=gb_tp))
                ops.conv_dgrad(d_tproj16, w_tp, None, B, 1, 1, 1, dx_f32=d_e2act, cin=temb)
                ops.silu_bwd(e2_f32, d_e2act, d_e2)
                self.on_side(lambda: ops.conv_wgrad(e1_act, d_e2, m.g32("time_embedding.linear_2.weight").view(temb, 1, temb),
                                                    B, 1, 1, 1, 1, dbias=m.g32("time_embedding.linear_2.bias")))
                ops.conv_dgrad(d_e2, m.w16("time_embedding.linear_2.weight"), None, B, 1, 1, 1, dx_f32=d_e1act, cin=temb)
                ops.silu_bwd(e1_f32, d_e1act, d_e1)
                self.on_side(lambda: ops.conv_wgrad(te0, d_e1, m.g32("time_embedding.linear_1.weight").view(temb, 1, boc[0]),
                                                    B, 1, 1, 1, 1, dbias=m.g32("time_embedding.linear_1.bias")))
            self._temb_bwd = temb_bwd

        # ---- symbolic pass -------------------------------------------------------------------------
        self._ops = []        # (kind, dict) in forward order
        h = self.act(S, boc[0], "conv_in")
        self._ops.append(("conv_in", dict(out=h)))
        skips = [h]
        res = S
        for i, blk in enumerate(L.down):
            for j, r in enumerate(blk.resnets):
                h = self._sym_resnet(r, h, res)
                if blk.attns:
                    h = self._sym_attn(blk.attns[j], h, res)
                skips.append(h)
            if blk.down:
                o = self.act(res // 2, h.C, blk.down)
                self._ops.append(("down", dict(name=blk.down, x=h, out=o, H=res // 2)))
                res //= 2
                h = o
                skips.append(h)
        h = self._sym_resnet(L.mid.res0, h, res)
        h = self._sym_attn(L.mid.attn, h, res)
        h = self._sym_resnet(L.mid.res1, h, res)
        for i, blk in enumerate(L.up):
            for j, r in enumerate(blk.resnets):
                s = skips.pop()
                assert s.C == blk.skip_c[j] and s.H == res, (s.C, blk.skip_c[j], s.H, res)
                cat = self.act(res, h.C + s.C, f"cat.{r.prefix}")
                assert h.alias is None and s.alias is None
                h.alias = (cat, 0)
                s.alias = (cat, h.C)
                cat.parts = (h, s)
                s.has_upstream_grad = True
                h = self._sym_resnet(r, cat, res)
                if blk.attns:
                    h = self._sym_attn(blk.attns[j], h, res)
            if blk.up:
                o = self.act(res * 2, h.C, blk.up)
                self._ops.append(("up", dict(name=blk.up, x=h, out=o, H=res)))
                res *= 2
                h = o
        self._ops.append(("head", dict(x=h)))
        self._materialise()
        # GroupNorm sites fed by convolution epilogues (decided before emission: producers come first)
        for kind, d in self._ops:
            if kind == "resnet":
                HW = d["H"] * d["H"]
                d["q1"] = self._fused_stats_ok(HW, d["r"].cin) and self._want_q(d["x"])
                d["q2"] = self._fused_stats_ok(HW, d["r"].cout)
            elif kind == "head":
                d["q"] = self._fused_stats_ok(S * S, d["x"].C) and self._want_q(d["x"])
        # ---- emit programs ---------------------------------------------------------------------------
        emit = dict(conv_in=self._emit_conv_in, resnet=self._emit_resnet, attn=self._emit_attn, down=self._emit_down,
                    up=self._emit_up, head=self._emit_head)
        entries = []
        cut_at = {}                     # op index -> position in dp_cut_prefixes() (segment boundary BEFORE this op in forward order)
        cut_names = m.dp_cut_prefixes()
        for i, (kind, d) in enumerate(self._ops):
            if kind == "resnet" and d["r"].prefix in cut_names:
                cut_at[i] = cut_names.index(d["r"].prefix)
            rows = None
            if self.lanes > 1 and kind in ("resnet", "attn", "down", "up"):
                hh = 2 * d["H"] if kind == "up" else d["H"]          # resolution the op's GEMMs run at
                rows = B * hh * hh
            if rows is not None and rows <= self.lane_rows:
                nb = B // self.lanes
                per_lane = [emit[kind](**d, win=(l * nb, nb, l)) for l in range(self.lanes)]
                entries.append(("lanes", [pl[0] for pl in per_lane], [pl[1] for pl in per_lane]))
            else:
                fw, bw = emit[kind](**d)
                entries.append(("plain", fw, bw))
        # forward program: consecutive lane-split ops form ONE region (fork once, join once)
        region = None
        for kind_, fw, _ in entries:
            if kind_ == "lanes":
                if region is None:
                    region = [[] for _ in range(self.lanes)]
                    self.fwd.append(lambda progs=region: self._run_lanes(progs))
                for l in range(self.lanes):
                    region[l] += fw[l]
            else:
                region = None
                self.fwd += fw
        bw_chunks = [(k, bw) for k, _, bw in entries]
        self._alloc_q()
        # backward segments for data-parallel overlap: the gradients of a contiguous range of the flat buffer are
        # final at the end of each segment (UNet2DModelB200.grad_segment_ranges), so their all-reduce runs while
        # the next segment computes.  segment 0: head, up path, mid block; then one segment per cut of dp_cut_prefixes()
        # (default: down blocks 5..4, 3..2, 1); the last one: the rest.
        self.bwd_marks = []
        if ng:
            region = None
            for i in range(len(bw_chunks) - 1, -1, -1):
                kind_, bw = bw_chunks[i]
                if kind_ == "lanes":
                    if region is None:
                        region = [[] for _ in range(self.lanes)]
                        self.bwd.append(lambda progs=region: self._run_lanes(progs))
                    for l in range(self.lanes):
                        region[l] += bw[l]
                else:
                    region = None
                    self.bwd += bw
                if i in cut_at:
                    region = None            # a segment boundary closes the region: its lanes have joined
                    self.bwd.append(lambda k=cut_at[i]: self._grads_ready(k))
                    self.bwd_marks.append(len(self.bwd))
            self.bwd.append(self._temb_bwd)

    def _grads_ready(self, k):
        """end of backward segment k: the gradients of grad_segment_ranges()[k] are final.  `model.grad_ready_hook(lo, hi)`
        (the trainer's in-graph peer-memory all-reduce) is called with the main stream ordered after the side streams."""
        hook = getattr(self.m, "grad_ready_hook", None)
        if hook is not None:
            # the weight gradients of that range ran on the side stream(s): the hook orders ITS stream after them (and
            # after the main stream) -- the main stream itself does not wait, the dgrad / GroupNorm chain goes on
            lo, hi = self.m.grad_segment_ranges()[k]
            hook(lo, hi, [s for s in self.sides if s in self._sides_used])

    def _sym_resnet(self, r, x, res):
        out = self.act(res, r.cout, r.prefix)
        self._ops.append(("resnet", dict(r=r, x=x, out=out, H=res)))
        return out

    def _sym_attn(self, a, x, res):
        out = self.act(res, a.c, a.prefix)
        out.no_q = True
        self._ops.append(("attn", dict(a=a, x=x, out=out, H=res)))
        return out

    # -- emitters: return (forward launches, backward launches) ---------------------------------------
    def _gn_ws(self, HW, C, B=None, lane=None):
        n = max(1, ops.gn_ws_floats(B or self.B, HW, C))
        store = self.__dict__.setdefault("_gnws_by_lane", {})
        cur = store.get(lane)
        if cur is None or cur.numel() < n:
            store[lane] = torch.empty(n, dtype=torch.float32, device=self.dev)
        return lambda: store[lane]

    def _lane(self, win):
        """(batch, first sample, lane index, scratch-tag suffix) of an emitter's batch window (None: the whole batch)"""
        if win is None:
            return self.B, 0, None, ""
        b0, nb, l = win
        return nb, b0, l, f"@{l}"

    # -- GroupNorm statistics fused into the producing convolution's epilogue ------------------------------
    def qbuf(self, N, C):
        q = _QBuf(N, C)
        self._qbufs.append(q)
        return q

    def _fused_stats_ok(self, HW, C):
        """a GroupNorm site whose forward would need the statistics pass + apply pass (big maps): its producers'
        store epilogues accumulate the statistics instead and the activation is read once"""
        if not self.fused_stats or HW % 128 != 0 or (C // G) % 4 != 0:
            return False
        kind = ops.gn_fwd_kind(self.B, HW, C, G)
        if kind == 2:
            return True
        # inference: a medium map (one-pass cluster kernel) whose consumer convolution can fold the normalisation into its
        # operand path (_emit_resnet) takes its statistics from the producers as well -- the GroupNorm pass disappears
        H = math.isqrt(HW)
        if (not self.need_grad and int(os.environ.get("MDM_GN_FOLD", "1")) > 0 and kind == 1 and H * H == HW and H % 16 == 0
                and C % 64 == 0 and C >= 128):
            return True
        # training, medium maps: statistics from the producers' epilogues + the apply pass instead of the one-pass cluster
        # kernel.  Opt-in (MDM_GN_FUSED_STATS_MEDIUM=1): on one GPU c2 7.95 -> 7.83 ms, c3 10.43 -> 10.23 ms per step, but the
        # one 8-GPU run taken with it was SLOWER (8.86 vs 8.43 ms per step) and the round's GPU budget ended before that
        # could be repeated -- the default stays the configuration whose scaling was measured twice.
        return kind == 1 and int(os.environ.get("MDM_GN_FUSED_STATS_MEDIUM", "0")) > 0

    def _want_q(self, x):
        """mark the producers of x (both halves of a concatenation) to emit quad sums; False if one of them cannot"""
        parts = x.parts or (x,)
        if any(p.no_q or p.parts is not None for p in parts):
            return False
        for p in parts:
            if p.q is None:
                p.q = self.qbuf(self.B, p.C)
        return True

    def _q_of(self, x):
        parts = x.parts or (x,)
        return parts[0].q, (parts[1].q if len(parts) > 1 else None)

    def _alloc_q(self):
        total = sum(q.N * (q.C // 4) * 2 for q in self._qbufs)
        self.q_arena = torch.zeros(max(1, total), dtype=torch.float32, device=self.dev)
        off = 0
        for q in self._qbufs:
            n = q.N * (q.C // 4) * 2
            q.t = self.q_arena[off:off + n].view(q.N, q.C // 4, 2)
            off += n
        if self._qbufs:
            self.fwd.insert(0, lambda: self.q_arena.zero_())     # one memset per forward for every site

    def _emit_conv_in(self, out):
        """first conv as a tensor-core GEMM: gather the 3x3 neighbourhood of the C-channel image into a
        [pixel][64] bf16 matrix (im2col3x3), then a 1x1 GEMM against W'[cout][tap*C + c]"""
        m, B, S = self.m, self.B, self.S
        C, co = m._cfg["in_channels"], out.C
        g_in = self.new((B, S, S, 64), zero=True)      # im2col3x3 only rewrites the data columns
        wq = self.new((co, 1, 64), zero=True)

        def prep_w():
            wq[:, 0, :9 * C].copy_(m.w32("conv_in.weight").view(co, C, 9).permute(0, 2, 1).reshape(co, 9 * C))
        fw = [prep_w,
              lambda: ops.im2col3x3(self.x_in, g_in, B, C, S, S),
              lambda: ops.conv_fprop(g_in, wq, out.val, B, S, S, 1, 1, bias=m.w32("conv_in.bias"),
                                     qsum=out.q.t if out.q is not None else None)]
        bw = []
        if self.need_grad:
            dwq = self.new((co, 1, 64), torch.float32)

            def backward():
                dwq.zero_()
                ops.conv_wgrad(g_in, out.grad, dwq, B, S, S, 1, 1, dbias=m.g32("conv_in.bias"))
                m.g32("conv_in.weight").add_(dwq[:, 0, :9 * C].view(co, 9, C).permute(0, 2, 1).reshape(co, C, 3, 3))
            bw = [backward]
        return fw, bw

    def _emit_resnet(self, r, x, out, H, win=None, q1=False, q2=False):
        m = self.m
        B, b0, lane, sfx = self._lane(win)
        q_in = self._q_of(x) if (q1 and win is None) else None            # statistics of x from its producers
        q_h1 = self.qbuf(B, r.cout) if (q2 and win is None) else None      # statistics of h1 from conv1's epilogue
        q_out = out.q if win is None else None
        if win is not None:
            x, out = _Win(x, b0, B), _Win(out, b0, B)
        HW = H * H
        p = r.prefix
        ng = self.need_grad
        a1 = self.new((B, H, H, r.cin)) if ng else None
        h1 = self.new((B, H, H, r.cout)) if ng else None
        a2 = self.new((B, H, H, r.cout)) if ng else None
        st1 = self.new((B, G, 2), torch.float32)
        st2 = self.new((B, G, 2), torch.float32)
        if not ng:   # inference: temporaries shared by all layers
            ra1 = self.scratch("a1" + sfx, (B, H, H, r.cin))
            rh1 = self.scratch("h1" + sfx, (B, H, H, r.cout))
            ra2 = self.scratch("a2" + sfx, (B, H, H, r.cout))
        ws1, ws2 = self._gn_ws(HW, r.cin, B, lane), self._gn_ws(HW, r.cout, B, lane)
        tp = self.tproj[b0:b0 + B, r.tproj_off:r.tproj_off + r.cout]
        ld_tp = self.tproj.shape[1]
        eps = self.eps

        def A1():
            return a1 if ng else self.sget(ra1)

        def H1():
            return h1 if ng else self.sget(rh1)

        def A2():
            return a2 if ng else self.sget(ra2)

        def qt(q):
            return q.t if q is not None else None

        def norm1():
            if q_in is not None:
                ops.gn_silu_fwd_q(x.val, A1(), m.w32(f"{p}.norm1.weight"), m.w32(f"{p}.norm1.bias"), st1, q_in[0].t, qt(q_in[1]),
                                  B, HW, r.cin, G, eps, True)
            else:
                ops.gn_silu_fwd(x.val, A1(), m.w32(f"{p}.norm1.weight"), m.w32(f"{p}.norm1.bias"), st1, ws1(), B, HW, r.cin, G, eps, True)

        def norm2():
            if q_h1 is not None:
                ops.gn_silu_fwd_q(H1(), A2(), m.w32(f"{p}.norm2.weight"), m.w32(f"{p}.norm2.bias"), st2, q_h1.t, None,
                                  B, HW, r.cout, G, eps, True)
            else:
                ops.gn_silu_fwd(H1(), A2(), m.w32(f"{p}.norm2.weight"), m.w32(f"{p}.norm2.bias"), st2, ws2(), B, HW, r.cout, G, eps, True)
        # inference, big maps: GroupNorm + SiLU folded into the consumer convolution's operand path (igemm.cu: kNorm).  The
        # statistics come from the producers' epilogues (q sums); a small kernel turns them into a per-sample (scale, shift)
        # table and the convolution normalises the raw activation while it fills its halo tiles -- the normalised tensors
        # a1 / a2 are never written or read (2 x 2 B per element per site).  MDM_GN_FOLD=0 off, 2 = also on small grids.
        fold_mode = int(os.environ.get("MDM_GN_FOLD", "1"))
        items = B * (H // 16) * (H // 16) * ((r.cout + 127) // 128) if H % 16 == 0 else 0
        fold_ok = (not ng) and win is None and fold_mode > 0 and H % 16 == 0 and (items >= 96 or fold_mode >= 2)
        fold1 = fold_ok and q_in is not None and r.cin % 64 == 0 and r.cin >= 128
        fold2 = fold_ok and q_h1 is not None and r.cout % 64 == 0 and r.cout >= 128
        coef1 = self.new((B, r.cin, 2), torch.float32) if fold1 else None
        coef2 = self.new((B, r.cout, 2), torch.float32) if fold2 else None
        fw = []
        if fold1:
            fw.append(lambda: ops.gn_coef_q(q_in[0].t, qt(q_in[1]), m.w32(f"{p}.norm1.weight"), m.w32(f"{p}.norm1.bias"), coef1,
                                            B, HW, r.cin, G, eps, stats=st1))
            fw.append(lambda: ops.conv_fprop(x.val, m.w16(f"{p}.conv1.weight"), H1(), B, H, H, 3, 1, bias=m.w32(f"{p}.conv1.bias"),
                                             rowvec=tp, ld_rowvec=ld_tp, qsum=qt(q_h1), gn_coef=coef1))
        else:
            fw.append(norm1)
            fw.append(lambda: ops.conv_fprop(A1(), m.w16(f"{p}.conv1.weight"), H1(), B, H, H, 3, 1, bias=m.w32(f"{p}.conv1.bias"),
                                             rowvec=tp, ld_rowvec=ld_tp, qsum=qt(q_h1)))
        if fold2:
            fw.append(lambda: ops.gn_coef_q(q_h1.t, None, m.w32(f"{p}.norm2.weight"), m.w32(f"{p}.norm2.bias"), coef2,
                                            B, HW, r.cout, G, eps, stats=st2))
        else:
            fw.append(norm2)
        a2_in = (lambda: H1()) if fold2 else A2
        if r.shortcut:
            fw.append(lambda: ops.conv_fprop(a2_in(), m.w16(f"{p}.conv2.weight"), out.val, B, H, H, 3, 1, bias=m.w32(f"{p}.conv2.bias"),
                                             bias2=m.w32(f"{p}.conv_shortcut.bias"), x2=x.val, w2=m.w16(f"{p}.conv_shortcut.weight"),
                                             qsum=qt(q_out), gn_coef=coef2))
        else:
            fw.append(lambda: ops.conv_fprop(a2_in(), m.w16(f"{p}.conv2.weight"), out.val, B, H, H, 3, 1, bias=m.w32(f"{p}.conv2.bias"),
                                             resid=x.val, qsum=qt(q_out), gn_coef=coef2))
        bw = []
        if ng:
            rd_a2 = self.scratch("d_a2" + sfx, (B, H, H, r.cout))
            rd_h1 = self.scratch("d_h1" + sfx, (B, H, H, r.cout))
            rd_a1 = self.scratch("d_a1" + sfx, (B, H, H, r.cin))
            has_up = x.has_upstream_grad

            def wgrad_out():
                d_out = out.grad
                # weight gradients; the bias gradients (column sums of d_out) ride in the same kernel
                ops.conv_wgrad(a2, d_out, m.g32(f"{p}.conv2.weight"), B, H, H, 3, 1, dbias=m.g32(f"{p}.conv2.bias"),
                               dbias2=m.g32(f"{p}.conv_shortcut.bias") if r.shortcut else None)
                if r.shortcut:
                    ops.conv_wgrad(x.val, d_out, m.g32(f"{p}.conv_shortcut.weight"), B, H, H, 1, 1)

            def backward():
                d_out = out.grad
                d_a2, d_h1, d_a1 = self.sget(rd_a2), self.sget(rd_h1), self.sget(rd_a1)
                self.on_side(wgrad_out)
                ops.conv_dgrad(d_out, m.w16(f"{p}.conv2.weight"), d_a2, B, H, H, 3)
                self.before_write("d_h1" + sfx)
                # d_h1 plus, in the same pass, its per-sample column sums = d(time_emb_proj output) and conv1.bias grad
                ops.gn_silu_bwd(h1, d_a2, d_h1, m.w32(f"{p}.norm2.weight"), m.w32(f"{p}.norm2.bias"), st2,
                                m.g32(f"{p}.norm2.weight"), m.g32(f"{p}.norm2.bias"), ws2(), B, HW, r.cout, G, True,
                                colsum=self.d_tproj[b0:b0 + B, r.tproj_off:], ld_colsum=self.d_tproj.shape[1],
                                dbias=m.g32(f"{p}.conv1.bias"))
                self.on_side(lambda: ops.conv_wgrad(a1, d_h1, m.g32(f"{p}.conv1.weight"), B, H, H, 3, 1), reads=("d_h1" + sfx,))
                ops.conv_dgrad(d_h1, m.w16(f"{p}.conv1.weight"), d_a1, B, H, H, 3)
                add = None
                if r.shortcut:
                    ops.conv_dgrad(d_out, m.w16(f"{p}.conv_shortcut.weight"), x.grad, B, H, H, 1, accumulate=has_up)
                    add = x.grad
                elif has_up:
                    add = x.grad
                ops.gn_silu_bwd(x.val, d_a1, x.grad, m.w32(f"{p}.norm1.weight"), m.w32(f"{p}.norm1.bias"), st1,
                                m.g32(f"{p}.norm1.weight"), m.g32(f"{p}.norm1.bias"), ws1(), B, HW, r.cin, G, True,
                                add=add, add2=None if r.shortcut else d_out)
            bw = [backward]
        return fw, bw

    def _emit_attn(self, a, x, out, H, win=None):
        m = self.m
        B, b0, lane, sfx = self._lane(win)
        if win is not None:
            x, out = _Win(x, b0, B), _Win(out, b0, B)
        L_, C = H * H, a.c
        p = a.prefix
        ng = self.need_grad
        y = self.new((B * L_, C))
        qkv = self.new((B * L_, 3 * C))
        att = self.new((B * L_, C))
        st = self.new((B, G, 2), torch.float32)
        ws = self._gn_ws(L_, C, B, lane)
        w_qkv = m._span(f"{p}.to_q.weight", f"{p}.to_v.weight", (3 * C, C), "w16")
        b_qkv = m._span(f"{p}.to_q.bias", f"{p}.to_v.bias", (3 * C,), "w32")
        eps = self.eps
        fw = [
            lambda: ops.gn_silu_fwd(x.val, y, m.w32(f"{p}.group_norm.weight"), m.w32(f"{p}.group_norm.bias"), st, ws(), B, L_, C, G, eps, False),
            lambda: ops.conv_fprop(y, w_qkv, qkv, B * L_, 1, 1, 1, 1, bias=b_qkv),
            lambda: ops.attention_fwd(qkv, att, B, L_, C),
            lambda: ops.conv_fprop(att, m.w16(f"{p}.to_out.0.weight"), out.val, B * L_, 1, 1, 1, 1,
                                   bias=m.w32(f"{p}.to_out.0.bias"), resid=x.val),
        ]
        bw = []
        if ng:
            rd_att = self.scratch("d_att" + sfx, (B * L_, C))
            rd_qkv = self.scratch("d_qkv" + sfx, (B * L_, 3 * C))
            rd_y = self.scratch("d_y" + sfx, (B * L_, C))
            g_qkv = m._span(f"{p}.to_q.weight", f"{p}.to_v.weight", (3 * C, 1, C), "g32")
            gb_qkv = m._span(f"{p}.to_q.bias", f"{p}.to_v.bias", (3 * C,), "g32")
            has_up = x.has_upstream_grad

            def backward():
                d_out = out.grad
                d_att, d_qkv, d_y = self.sget(rd_att), self.sget(rd_qkv), self.sget(rd_y)
                self.on_side(lambda: ops.conv_wgrad(att, d_out, m.g32(f"{p}.to_out.0.weight").view(C, 1, C), B * L_, 1, 1, 1, 1,
                                                    dbias=m.g32(f"{p}.to_out.0.bias")))
                ops.conv_dgrad(d_out, m.w16(f"{p}.to_out.0.weight"), d_att, B * L_, 1, 1, 1)
                self.before_write("d_qkv" + sfx)
                ops.attention_bwd(qkv, d_att, d_qkv, B, L_, C)
                self.on_side(lambda: ops.conv_wgrad(y, d_qkv, g_qkv, B * L_, 1, 1, 1, 1, dbias=gb_qkv), reads=("d_qkv" + sfx,))
                ops.conv_dgrad(d_qkv, w_qkv, d_y, B * L_, 1, 1, 1)
                ops.gn_silu_bwd(x.val, d_y, x.grad, m.w32(f"{p}.group_norm.weight"), m.w32(f"{p}.group_norm.bias"), st,
                                m.g32(f"{p}.group_norm.weight"), m.g32(f"{p}.group_norm.bias"), ws(), B, L_, C, G, False,
                                add=x.grad if has_up else None, add2=d_out)
            bw = [backward]
        return fw, bw

    def _emit_down(self, name, x, out, H, win=None):
        m = self.m
        B, b0, lane, sfx = self._lane(win)
        if win is not None:
            x, out = _Win(x, b0, B), _Win(out, b0, B)
        C = x.C
        q_out = out.q if win is None else None
        fw = [lambda: ops.conv_fprop(x.val, m.w16(name + ".weight"), out.val, B, H, H, 3, 2, bias=m.w32(name + ".bias"),
                                     qsum=q_out.t if q_out is not None else None)]
        bw = []
        if self.need_grad:
            rz = self.scratch("zins" + sfx, (B, 2 * H, 2 * H, C))
            has_up = x.has_upstream_grad

            def backward():
                z = self.sget(rz)
                ops.zero_insert2x(out.grad, z, B, H, H, C)
                self.on_side(lambda: ops.conv_wgrad(x.val, out.grad, m.g32(name + ".weight"), B, H, H, 3, 2, dbias=m.g32(name + ".bias")))
                ops.conv_dgrad(z, m.w16(name + ".weight"), x.grad, B, 2 * H, 2 * H, 3, accumulate=has_up)
            bw = [backward]
        return fw, bw

    def _emit_up(self, name, x, out, H, win=None):
        m = self.m
        B, b0, lane, sfx = self._lane(win)
        if win is not None:
            x, out = _Win(x, b0, B), _Win(out, b0, B)
        C = x.C
        ng = self.need_grad
        fused = (not ng) and win is None and H % 16 == 0 and bool(int(os.environ.get("MDM_UP2X_FUSED", "1")))
        u = self.new((B, 2 * H, 2 * H, C)) if ng else None
        ru = None if (ng or fused) else self.scratch("ups" + sfx, (B, 2 * H, 2 * H, C))

        def U():
            return u if ng else self.sget(ru)
        q_out = out.q if win is None else None
        # inference on maps that are multiples of 16 x 16: the nearest-2x upsample is fused into the convolution (four
        # parity launches over the LOW-resolution input with pre-summed 2x2 weights: 4/9 of the FLOPs, no upsampled
        # tensor).  The parity weights are rebuilt from the fp32 master by one small kernel per forward (inside the
        # inference graph, so they always follow the weights).  Training keeps the explicit upsample (its backward
        # reads it).  MDM_UP2X_FUSED=0 switches back.
        if fused:
            w4 = self.new((4, C, 4, C))
            fw = [lambda: ops.up2x_weights(m.w32(name + ".weight"), w4, C, C),
                  lambda: ops.conv_up2x_fprop(x.val, w4, out.val, B, H, H, bias=m.w32(name + ".bias"),
                                              qsum=q_out.t if q_out is not None else None)]
        else:
            fw = [lambda: ops.upsample2x_fwd(x.val, U(), B, H, H, C),
                  lambda: ops.conv_fprop(U(), m.w16(name + ".weight"), out.val, B, 2 * H, 2 * H, 3, 1, bias=m.w32(name + ".bias"),
                                         qsum=q_out.t if q_out is not None else None)]
        bw = []
        if ng:
            rdu = self.scratch("d_ups" + sfx, (B, 2 * H, 2 * H, C))

            def backward():
                du = self.sget(rdu)
                self.on_side(lambda: ops.conv_wgrad(u, out.grad, m.g32(name + ".weight"), B, 2 * H, 2 * H, 3, 1, dbias=m.g32(name + ".bias")))
                ops.conv_dgrad(out.grad, m.w16(name + ".weight"), du, B, 2 * H, 2 * H, 3)
                ops.upsample2x_bwd(du, x.grad, B, H, H, C)
            bw = [backward]
        return fw, bw

    def _emit_head(self, x, q=False):
        """GroupNorm + SiLU + last conv.  The conv is a 1x1 GEMM producing the 9*Co per-tap partial outputs
        z[pixel][tap*Co + c] (fp32) followed by the 9-tap scatter-sum; its backward is the flipped gather of
        d_out feeding a GEMM (dgrad) and a wgrad GEMM."""
        m, B, S = self.m, self.B, self.S
        C, Co = x.C, m._cfg["out_channels"]
        a = self.new((B, S, S, C))
        st = self.new((B, G, 2), torch.float32)
        ws = self._gn_ws(S * S, C)
        eps = self.eps
        w2 = self.new((32, 1, C), zero=True)
        z = self.new((B * S * S, 32), torch.float32)

        def prep_w2():
            w2[:9 * Co, 0, :].copy_(m.w32("conv_out.weight").view(Co, C, 9).permute(2, 0, 1).reshape(9 * Co, C))
        q_in = self._q_of(x) if q else None

        def norm():
            if q_in is not None:
                ops.gn_silu_fwd_q(x.val, a, m.w32("conv_norm_out.weight"), m.w32("conv_norm_out.bias"), st, q_in[0].t,
                                  q_in[1].t if q_in[1] is not None else None, B, S * S, C, G, eps, True)
            else:
                ops.gn_silu_fwd(x.val, a, m.w32("conv_norm_out.weight"), m.w32("conv_norm_out.bias"), st, ws(), B, S * S, C, G, eps, True)
        fw = [norm,
              prep_w2,
              lambda: ops.conv_fprop(a, w2, None, B, S, S, 1, 1, y_f32=z, cout=32),
              lambda: ops.tapsum3x3(z, m.w32("conv_out.bias"), self.out, B, Co, S, S)]
        bw = []
        if self.need_grad:
            rda = self.scratch("d_head", (B, S, S, C))
            self.d_out = self.new((B, Co, S, S), torch.float32)
            g_out = self.new((B, S, S, 64), zero=True)
            w3 = self.new((C, 1, 64), zero=True)
            dw2 = self.new((64, 1, C), torch.float32)

            def backward():
                da = self.sget(rda)
                ops.im2col3x3(self.d_out, g_out, B, Co, S, S, flip=True)
                w3[:, 0, :9 * Co].copy_(m.w32("conv_out.weight").view(Co, C, 9).permute(1, 2, 0).reshape(C, 9 * Co))
                ops.conv_fprop(g_out, w3, da, B, S, S, 1, 1)
                dw2.zero_()
                ops.conv_wgrad(a, g_out, dw2, B, S, S, 1, 1)
                m.g32("conv_out.weight").add_(dw2[:9 * Co, 0, :].view(9, Co, C).permute(1, 2, 0).reshape(Co, C, 3, 3))
                m.g32("conv_out.bias").add_(self.d_out.sum(dim=(0, 2, 3)))
                ops.gn_silu_bwd(x.val, da, x.grad, m.w32("conv_norm_out.weight"), m.w32("conv_norm_out.bias"), st,
                                m.g32("conv_norm_out.weight"), m.g32("conv_norm_out.bias"), ws(), B, S * S, C, G, True)
            bw = [backward]
        return fw, bw

    # -- execution ---------------------------------------------------------------------------------
    def run_forward(self, x, t):
        self.x_in.copy_(x)
        self.t_in.copy_(t)
        # inference plans replay their forward program as ONE CUDA graph (after two eager warm-up calls): the sampler's
        # small shapes are otherwise bound by the ~250 Python launches per denoising step (64x3x32x32: GPU busy 2.2 ms
        # of a 7.9 ms step).  Training forwards are captured by the trainer together with the backward.
        if (not self.need_grad and self._infer_graph_ok and not torch.cuda.is_current_stream_capturing()):
            key = ops.reserve_sms(-1)
            g = self._fwd_graphs.get(key)
            if g is None and self._fwd_calls >= 2:
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for op in self.fwd:
                        op()
                self._fwd_graphs[key] = g
            if g is not None:
                g.replay()
                return self.out if getattr(self, "static_output", False) else self.out.clone()
            self._fwd_calls += 1
        if not getattr(self, "static_output", False):
            self.out = torch.empty_like(self.out)      # callers may keep the previous result
        for op in self.fwd:
            op()
        return self.out

    def run_backward(self, d_out, seg=None):
        """the whole backward program (seg None) or one of its segments (see `bwd_marks`)"""
        if seg is None or seg == 0:
            self.d_out.copy_(d_out)
            self.d_tproj.zero_()      # accumulated by the fused column sums of the norm2 backward
        prog = self.bwd
        if seg is not None:
            bounds = [0] + self.bwd_marks + [len(self.bwd)]
            prog = self.bwd[bounds[seg]:bounds[seg + 1]]
        for op in prog:
            op()
        self.join_side()              # every weight gradient has landed before the caller (optimiser / all-reduce) runs
