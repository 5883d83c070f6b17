"""TEST INFRASTRUCTURE ONLY -- fp32 PyTorch restatement of the denoiser the reference builds at
`utils/model.py:3-33` (`diffusers.UNet2DModel`, block_out_channels (128,128,256,256,512,512),
2 layers per block, attention in the blocks `num_attention` selects and in the mid block).

**Parity unpinned**: `diffusers` is a third-party dependency that is neither vendored under
/root/reference nor installed (no requirements file pins a version; run dates suggest ~0.26).
This module restates its published architecture with
the diffusers state-dict key names so that a real checkpoint loads.
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


def timestep_embedding(t: torch.Tensor, dim: int = 128) -> torch.Tensor:
    """diffusers `Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)` [upstream]."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


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
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, n, attn, down, hd, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, temb, groups, eps) for j in range(n)])
        if attn:
            self.attentions = nn.ModuleList([Attention(cout, hd, groups, eps) for _ in range(n)])
        self.has_attn = attn
        if down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
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
            self.down_blocks.append(DownBlock(cin, c, temb, n, ty.startswith("Attn"), i != len(boc) - 1, hd, g, eps))
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
        temb = self.time_embedding(timestep_embedding(t, self.config.block_out_channels[0]).to(sample.dtype))
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
