"""BASELINE configs[4]: large-channel U-Net (ch=256) at 3x256x256 -- eager training-step throughput on one GPU
(forward + backward + fused clip/AdamW; synthetic data).  python scripts/bench_c5.py [batch] [steps]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
from mdm_b200.denoiser import UNet2DModelB200, default_config
from mdm_b200.runtime import FusedOptimizer
from mdm_b200.config import unet_forward_flops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
C, S, base = 3, 256, 256
torch.manual_seed(0)
m = UNet2DModelB200(device="cuda", **default_config(C, S, base=base)); m.reset_parameters(seed=0); m.train()
opt = FusedOptimizer(m, "adamw", lr=1e-4)
x0 = torch.rand(B, C, S, S, device="cuda") * 2 - 1
x = x0 * (torch.rand(B, 1, S, S, device="cuda") > 0.5)
t = torch.randint(1, 1000, (B,), device="cuda").float()
def step():
    m.zero_grad()
    out = m(x, t).sample
    loss = torch.nn.functional.mse_loss(x + out, x0)
    loss.backward()
    opt.set_clip(1.0)
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
fl = 3.0 * B * unet_forward_flops(m._cfg, S)
print(f"c5 ch=256 3x256x256 batch {B}: {ms:.1f} ms/step = {B/ms*1e3:.1f} samples/s, {fl/ms/1e9:.0f} TFLOP/s of conv/linear math, "
      f"params {m.num_parameters()/1e6:.1f}M, loss {loss.item():.4f}, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
