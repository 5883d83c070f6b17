"""Robustness check on the other BASELINE shapes: forward+backward finite, a few optimiser steps reduce the loss.
    python scripts/smoke_shapes.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
from mdm_b200.denoiser import UNet2DModelB200, default_config
from mdm_b200.runtime import FusedOptimizer
for name, C, S, base, B in (("c1 1x32x32", 1, 32, 128, 64), ("c3 3x64x64", 3, 64, 128, 48), ("c4 3x128x128", 3, 128, 128, 8), ("c5 3x256x256 ch=256", 3, 256, 256, 2)):
    torch.manual_seed(0)
    m = UNet2DModelB200(device="cuda", **default_config(C, S, base=base)); m.reset_parameters(seed=0); m.train()
    opt = FusedOptimizer(m, "adamw", lr=1e-4)
    x0 = torch.rand(B, C, S, S, device="cuda") * 2 - 1
    x = x0 * (torch.rand(B, 1, S, S, device="cuda") > 0.5)
    t = torch.randint(1, 1000, (B,), device="cuda").float()
    losses = []
    for it in range(4):
        m.zero_grad()
        out = m(x, t).sample
        loss = torch.nn.functional.mse_loss(x + out, x0)
        loss.backward()
        gn = opt.set_clip(1.0)
        opt.step()
        losses.append(loss.item())
    torch.cuda.synchronize()
    ok = all(torch.isfinite(torch.tensor(losses))) and torch.isfinite(m.flat_param).all().item()
    print(f"{name}: params {m.num_parameters()/1e6:.1f}M batch {B} losses {[round(l,4) for l in losses]} finite={ok} mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
    del m, opt; torch.cuda.empty_cache()
