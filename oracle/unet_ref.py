"""TEST INFRASTRUCTURE ONLY -- fp32 PyTorch restatement of the denoiser the reference builds at
`utils/model.py:3-33` (`diffusers.UNet2DModel`, block_out_channels (128,128,256,256,512,512),
2 layers per block, attention in the blocks `num_attention` selects and in the mid block).

`diffusers` is a third-party dependency that is neither vendored under /root/reference nor installed (no
requirements file pins a version; run dates suggest ~0.26).  This module restates its published architecture
 with the diffusers state-dict key names so that a
real checkpoint loads.

**Pinned as a whole network to reference-held code**: the reference ships a runnable U-Net of the same topology,
`models/unet/unet6.py:365-506` instantiated as `UNet(C, 128, C, [1,1,2,2,4,4], 2, [F,F,F,F,T,F])`
(`models_Unet.py:153-159`).  It differs from the diffusers network in four places, each of which IS a diffusers
config field that this restatement honours the way diffusers does -- `unet6_compat()` below: GroupNorm eps 1e-6
(`norm_eps`), timestep embedding `[sin, cos]` with `half - 1` in the exponent (`flip_sin_to_cos=False`,
`freq_shift=1`), one attention head scaled 1/sqrt(C) (`attention_head_dim = C`), stride-2 convolutions padded
right / bottom only (`downsample_padding=0`).  `tests/golden/make_golden_unet6.py` runs the REFERENCE network with
seeded weights (mapped key by key from the diffusers names, `unet6_key_map`) and `tests/test_oracle_unet6.py` checks
this module under `unet6_compat()` against its output to 1e-5 relative L2: skip wiring, channel plan, up / down
ordering, where the time embedding and the attention enter are pinned.  What stays UNPINNED is only the default
value of those four fields (diffusers' own defaults: 1e-5, `[cos, sin]` / `half`, head_dim 8, symmetric padding).
The block algebra (ResnetBlock2D, Attention, the multi-head core) IS pinned against the reference's in-repo
`unet6.ResidualBlock` / `unet6.AttentionBlock` / `unet4.QKVAttentionLegacy` (tests/golden/unet_blocks.npz,
tests/test_oracle_blocks.py).  Further checkable anchors: parameter count 113,673,219 (C=3) / 113,668,609 (C=1) / 454,461,443
(ch=256), equal to the in-repo `models/unet/unet6.py` configuration `[1,1,2,2,4,4]`.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


def unet_config(dim_channel=3, dim_height=32, num_attention=1, base=128, layers_per_block=2):
    """Mirror of utils/model.py:3-33 (`MyModel`) as a plain config dict."""
    down = {
        1: ("DownBlock2D", "DownBlock2D", "DownBlock2D", "DownBlock2D", "AttnDownBlock2D", "DownBlock2D"),
        2: ("DownBlock2D", "DownBlock2D", "DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "DownBlock2D"),
        3: ("DownBlock2D", "DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "DownBlock2D"),
        4: ("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "DownBlock2D"),
        5: ("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
    }
    up = {
        1: ("UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"),
        2: ("UpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"),
        3: ("UpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D"),
        4: ("UpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
        5: ("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
    }
    if num_attention not in down:
        raise NotImplementedError("not implemented")
    b = base
    return dict(
        _class_name="UNet2DModel", sample_size=dim_height, in_channels=dim_channel, out_channels=dim_channel,
        layers_per_block=layers_per_block, block_out_channels=[b, b, 2 * b, 2 * b, 4 * b, 4 * b],
        down_block_types=list(down[num_attention]), up_block_types=list(up[num_attention]),
        act_fn="silu", attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5,
        time_embedding_type="positional", flip_sin_to_cos=True, freq_shift=0,
        downsample_padding=1, add_attention=True, center_input_sample=False, dropout=0.0,
        mid_block_scale_factor=1, downsample_type="conv", upsample_type="conv",
        resnet_time_scale_shift="default", attn_norm_num_groups=None,
        class_embed_type=None, num_class_embeds=None, num_train_timesteps=None,
    )


def timestep_embedding(t: torch.Tensor, dim: int = 128, flip_sin_to_cos: bool = True, freq_shift: float = 0) -> torch.Tensor:
    """diffusers `Timesteps(dim, flip_sin_to_cos, downscale_freq_shift)` [upstream]: exponent denominator
    `half - freq_shift`, `[sin, cos]` order unless flipped.  (flip=False, shift=1) is the embedding of the
    reference's in-repo models/unet/unet6.py:18-35."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift)
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    if flip_sin_to_cos:
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)


# the four places where the reference's in-repo unet6 differs from the diffusers defaults, as diffusers config fields
def unet6_compat(channels_at_attention: int = 512) -> dict:
    return dict(norm_eps=1e-6, flip_sin_to_cos=False, freq_shift=1, attention_head_dim=channels_at_attention,
                downsample_padding=0)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, c, head_dim, groups, eps):
        super().__init__()
        self.heads = c // head_dim
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        y = self.group_norm(x.view(B, C, H * W)).transpose(1, 2)          # (B, L, C)
        q, k, v = self.to_q(y), self.to_k(y), self.to_v(y)
        sp = lambda z: z.view(B, -1, self.heads, C // self.heads).transpose(1, 2)
        q, k, v = sp(q), sp(k), sp(v)
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]), dim=-1) @ v
        att = att.transpose(1, 2).reshape(B, -1, C)
        out = self.to_out[0](att).transpose(1, 2).reshape(B, C, H, W)
        return out + x


class Downsample2D(nn.Module):
    """diffusers Downsample2D(use_conv=True, padding=downsample_padding): padding 0 pads right / bottom by one
    instead (`F.pad(x, (0, 1, 0, 1))` [upstream]) -- the `SamePad2d(3, 2)` of unet6.py:258-275 on even maps"""

    def __init__(self, c, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0.0)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, n, attn, down, hd, groups, eps, down_pad=1):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, temb, groups, eps) for j in range(n)])
        if attn:
            self.attentions = nn.ModuleList([Attention(cout, hd, groups, eps) for _ in range(n)])
        self.has_attn = attn
        if down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout, down_pad)])
        self.has_down = down

    def forward(self, h, temb):
        outs = []
        for j, r in enumerate(self.resnets):
            h = r(h, temb)
            if self.has_attn:
                h = self.attentions[j](h)
            outs.append(h)
        if self.has_down:
            h = self.downsamplers[0](h)
            outs.append(h)
        return h, outs


class UpBlock(nn.Module):
    def __init__(self, cin, cout, cprev, temb, n, attn, up, hd, groups, eps):
        super().__init__()
        res = []
        for j in range(n):
            skip = cin if j == n - 1 else cout
            rin = cprev if j == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(res)
        if attn:
            self.attentions = nn.ModuleList([Attention(cout, hd, groups, eps) for _ in range(n)])
        self.has_attn = attn
        if up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.has_up = up

    def forward(self, h, skips, temb):
        for j, r in enumerate(self.resnets):
            h = r(torch.cat([h, skips.pop()], dim=1), temb)
            if self.has_attn:
                h = self.attentions[j](h)
        if self.has_up:
            h = self.upsamplers[0](h)
        return h


class MidBlock(nn.Module):
    def __init__(self, c, temb, hd, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, groups, eps), ResnetBlock2D(c, c, temb, groups, eps)])
        self.attentions = nn.ModuleList([Attention(c, hd, groups, eps)])

    def forward(self, h, temb):
        return self.resnets[1](self.attentions[0](self.resnets[0](h, temb)), temb)


class UNet2DModelRef(nn.Module):
    """`model(x, t).sample`, `.device`, `.config` -- the surface sampler.py:111,145 and
    trainer_masked.py:125 use."""

    def __init__(self, **config):
        super().__init__()
        cfg = unet_config()
        cfg.update(config)
        self.config = SimpleNamespace(**cfg)
        boc = cfg["block_out_channels"]
        g, eps, hd = cfg["norm_num_groups"], cfg["norm_eps"], cfg["attention_head_dim"]
        n = cfg["layers_per_block"]
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.down_blocks = nn.ModuleList()
        c = boc[0]
        for i, ty in enumerate(cfg["down_block_types"]):
            cin, c = c, boc[i]
            self.down_blocks.append(DownBlock(cin, c, temb, n, ty.startswith("Attn"), i != len(boc) - 1, hd, g, eps,
                                              cfg["downsample_padding"]))
        self.mid_block = MidBlock(boc[-1], temb, hd, g, eps)
        rev = list(reversed(boc))
        self.up_blocks = nn.ModuleList()
        c = rev[0]
        for i, ty in enumerate(cfg["up_block_types"]):
            cprev, c = c, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(cin, c, cprev, temb, n + 1, ty.startswith("Attn"), i != len(boc) - 1, hd, g, eps))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, sample, timestep):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], device=sample.device)
        if t.dim() == 0:
            t = t[None]
        t = t.to(sample.device).expand(sample.shape[0])
        temb = self.time_embedding(timestep_embedding(t, self.config.block_out_channels[0], self.config.flip_sin_to_cos,
                                                      self.config.freq_shift).to(sample.dtype))
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk(h, temb)
            skips += outs
        h = self.mid_block(h, temb)
        for blk in self.up_blocks:
            h = blk(h, skips, temb)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return SimpleNamespace(sample=h)


def seeded_state_dict(model: nn.Module, seed: int) -> dict:
    """deterministic weights for every parameter of `model` in state-dict order (CPU mt19937 generator): N(0, 1)
    scaled by 1/sqrt(fan_in) for matrices / filters, 1 + 0.1 N(0,1) for norm scales, 0.1 N(0,1) for biases -- nothing
    is left at the zero initialisation the reference gives its last convolutions, so every path contributes.
    The golden generator feeds the SAME tensors to the reference network (447 MB of weights never enter the repo)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in model.state_dict().items():
        r = torch.randn(v.shape, generator=g)
        if v.dim() >= 2:
            out[k] = r / math.sqrt(v[0].numel())
        elif "norm" in k and k.endswith("weight"):
            out[k] = 1.0 + 0.1 * r
        else:
            out[k] = 0.1 * r
    return out


def unet6_key_map(cfg: dict) -> dict:
    """diffusers state-dict key (without .weight/.bias) -> module path inside the reference's `unet6.UNet`
    (models/unet/unet6.py:365-506).  Attention q/k/v map to slices of `project_in` (handled by the caller)."""
    n = cfg["layers_per_block"]
    L = len(cfg["block_out_channels"])
    m = {"conv_in": "in_conv", "time_embedding.linear_1": "embed.0", "time_embedding.linear_2": "embed.2",
         "conv_norm_out": "out_conv.0", "conv_out": "out_conv.2"}
    res = {"norm1": "norm1", "conv1": "conv1", "time_emb_proj": "fc", "norm2": "norm2", "conv2": "conv2", "conv_shortcut": "skip"}
    att = {"group_norm": "norm", "to_out.0": "project_out"}

    def block(dprefix, uprefix, attn):
        for a, b in res.items():
            m[f"{dprefix[0]}.{a}"] = f"{uprefix}.0.{b}" if attn else f"{uprefix}.{b}"
        if attn:
            for a, b in att.items():
                m[f"{dprefix[1]}.{a}"] = f"{uprefix}.1.{b}"
            m[f"{dprefix[1]}.qkv"] = f"{uprefix}.1.project_in"

    for i, ty in enumerate(cfg["down_block_types"]):
        for j in range(n):
            block((f"down_blocks.{i}.resnets.{j}", f"down_blocks.{i}.attentions.{j}"), f"downsamples.level_{i}.{j}", ty.startswith("Attn"))
        if i != L - 1:
            m[f"down_blocks.{i}.downsamplers.0.conv"] = f"downsamples.level_{i}.{n}.1"
    block(("mid_block.resnets.0", None), "middle.0", False)
    for a, b in att.items():
        m[f"mid_block.attentions.0.{a}"] = f"middle.1.{b}"
    m["mid_block.attentions.0.qkv"] = "middle.1.project_in"
    block(("mid_block.resnets.1", None), "middle.2", False)
    for i, ty in enumerate(cfg["up_block_types"]):
        lvl = L - 1 - i
        for j in range(n + 1):
            block((f"up_blocks.{i}.resnets.{j}", f"up_blocks.{i}.attentions.{j}"), f"upsamples.level_{lvl}.{j}", ty.startswith("Attn"))
        if i != L - 1:
            m[f"up_blocks.{i}.upsamplers.0.conv"] = f"upsamples.level_{lvl}.{n + 1}.1"
    return m


def to_unet6_state_dict(sd: dict, cfg: dict) -> dict:
    """re-key a diffusers-named state dict for `unet6.UNet`: Linear q/k/v -> one 1x1 `project_in` filter (q, k, v
    stacked: unet6.py:327 chunks in that order), Linear `to_out.0` -> a 1x1 filter"""
    km = unet6_key_map(cfg)
    out = {}
    for k, v in sd.items():
        stem, leaf = k.rsplit(".", 1)
        if stem.rsplit(".", 1)[-1] in ("to_q", "to_k", "to_v"):
            continue
        tgt = km[stem]
        if tgt.endswith("project_out"):
            v = v[:, :, None, None] if v.dim() == 2 else v
        out[f"{tgt}.{leaf}"] = v
    for stem, tgt in km.items():
        if stem.endswith(".qkv"):
            base = stem[:-4]
            out[f"{tgt}.weight"] = torch.cat([sd[f"{base}.to_{c}.weight"] for c in "qkv"], 0)[:, :, None, None]
            out[f"{tgt}.bias"] = torch.cat([sd[f"{base}.to_{c}.bias"] for c in "qkv"], 0)
    return out


def fwd_flops_per_image(cfg: dict) -> float:
    """Forward FLOPs (2*MAC) per image, convs + linears + attention matmuls."""
    total = 0.0
    S = cfg["sample_size"]
    model = UNet2DModelRef(**cfg).to("meta")
    hooks = []

    def conv_hook(m, i, o):
        nonlocal total
        total += 2.0 * o.shape[1] * o.shape[2] * o.shape[3] * m.in_channels * m.kernel_size[0] * m.kernel_size[1]

    def lin_hook(m, i, o):
        nonlocal total
        total += 2.0 * o.numel() / o.shape[0] * m.in_features

    def attn_hook(m, i, o):
        nonlocal total
        L = o.shape[2] * o.shape[3]
        total += 2.0 * 2.0 * L * L * o.shape[1]

    for mod in model.modules():
        if isinstance(mod, nn.Conv2d):
            hooks.append(mod.register_forward_hook(conv_hook))
        elif isinstance(mod, nn.Linear):
            hooks.append(mod.register_forward_hook(lin_hook))
        elif isinstance(mod, Attention):
            hooks.append(mod.register_forward_hook(attn_hook))
    x = torch.zeros(1, cfg["in_channels"], S, S, device="meta")
    model(x, torch.zeros(1, device="meta"))
    for h in hooks:
        h.remove()
    return total
