"""Golden fixtures that pin the denoiser oracle (oracle/unet_ref.py) to the reference's OWN building blocks.

`diffusers.UNet2DModel` itself is absent (parity unpinned), but the reference tree holds in-repo
U-Nets with the same block algebra:
  * `models/unet/unet6.py:336-362` ResidualBlock   = GN -> SiLU -> conv3x3 -> + Linear(SiLU(t_emb)) -> GN -> SiLU ->
    conv3x3 -> + (1x1 skip | identity)             <-> oracle ResnetBlock2D (eps mapped: 1e-6 there)
  * `models/unet/unet6.py:296-333`  AttentionBlock  = GN -> 1x1 qkv -> softmax(q k^T / sqrt(C)) v -> 1x1 out -> + x
    (one head)                                     <-> oracle Attention with head_dim = C
  * `models/unet/unet4.py:694-719`  QKVAttentionLegacy (multi-head core, scale 1/sqrt(ch) split over q and k)
                                                   <-> the oracle's multi-head softmax core (head_dim 8)
This script RUNS those reference classes on CPU under a fixed seed and stores inputs, weights and outputs in
tests/golden/unet_blocks.npz.  Run in the build container only:  python tests/golden/make_golden_blocks.py"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("MDM_REFERENCE_CODE", "/root/reference/code")
OUT = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    from oracle.ref_shims import install_stubs
    install_stubs()
    sys.path.insert(0, os.path.join(REF, "models", "unet"))
    sys.path.insert(0, os.path.join(REF, "models"))
    sys.path.insert(0, REF)
    u6 = load(os.path.join(REF, "models", "unet", "unet6.py"), "ref_unet6")
    u4 = load(os.path.join(REF, "models", "unet", "unet4.py"), "ref_unet4")
    out = {}
    g = torch.Generator().manual_seed(1234)

    def rnd(*shape, scale=1.0):
        return torch.randn(*shape, generator=g) * scale

    # ---- ResidualBlock: identity skip (64 -> 64) and 1x1 skip (32 -> 64) ---------------------------------
    for tag, cin, cout in (("res_same", 64, 64), ("res_proj", 32, 64)):
        temb = 48
        blk = u6.ResidualBlock(cin, cout, temb, drop_rate=0.0).eval()
        with torch.no_grad():
            for p in blk.parameters():                       # conv2 is zero-initialised in the reference: perturb everything
                p.copy_(rnd(*p.shape, scale=0.2))
            x, t = rnd(2, cin, 8, 8), rnd(2, temb)
            y = blk(x.clone(), t)
        out[f"{tag}.x"], out[f"{tag}.t"], out[f"{tag}.y"] = x.numpy(), t.numpy(), y.numpy()
        for k, v in blk.state_dict().items():
            out[f"{tag}.w.{k}"] = v.numpy()
    # ---- AttentionBlock (single head) ---------------------------------------------------------------------
    blk = u6.AttentionBlock(64).eval()
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(rnd(*p.shape, scale=0.2))
        x = rnd(2, 64, 4, 4)
        y = blk(x)
    out["attn1.x"], out["attn1.y"] = x.numpy(), y.numpy()
    for k, v in blk.state_dict().items():
        out[f"attn1.w.{k}"] = v.numpy()
    # ---- QKVAttentionLegacy: 4 heads of dim 8, 16 tokens ----------------------------------------------------
    core = u4.QKVAttentionLegacy(4)
    qkv = rnd(3, 4 * 3 * 8, 16)
    with torch.no_grad():
        a = core(qkv)
    out["qkv_legacy.qkv"], out["qkv_legacy.out"] = qkv.numpy(), a.numpy()
    np.savez_compressed(os.path.join(OUT, "unet_blocks.npz"), **out)
    print("wrote unet_blocks.npz:", {k: v.shape for k, v in out.items() if not ".w." in k})


if __name__ == "__main__":
    main()
