"""one forward + backward GroupNorm call per shape, for `ncu --set full -k regex:gn_` (no warm-up: every launch is profiled)
    python scripts/ncu_gn.py [N H C ...]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
from mdm_b200 import denoiser_ops as ops
a = [int(v) for v in sys.argv[1:]] or [128, 32, 128, 64, 128, 128]
for N, H, C in zip(a[0::3], a[1::3], a[2::3]):
    x = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)
    y = torch.empty_like(x); dx = torch.empty_like(x)
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    stats = torch.empty(N, 32, 2, device="cuda")
    ws = torch.empty(max(1, ops.gn_ws_floats(N, H * H, C)), device="cuda")
    ops.gn_silu_fwd(x, y, gamma, beta, stats, ws, N, H * H, C, 32, 1e-5, True)
    ops.gn_silu_bwd(x, dy, dx, gamma, beta, stats, dg, db, ws, N, H * H, C, 32, True, add2=y)
    torch.cuda.synchronize()
