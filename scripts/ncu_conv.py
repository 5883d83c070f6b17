"""one fprop + one dgrad of a big 3x3 layer for `ncu --set full -k regex:igemm_kernel`:  python scripts/ncu_conv.py [N H cin cout] [--stats] [--fold]
(--fold: the fprop twice -- plain, then with GroupNorm + SiLU of the input folded into its operand path -- and no dgrad;
warm-up launches first: capture with `--launch-skip 2 -c 2`)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
from mdm_b200 import denoiser_ops as ops
stats = "--stats" in sys.argv        # fprop with the GroupNorm quad sums of its output fused into the store epilogue (kStats)
fold = "--fold" in sys.argv
argv = [v for v in sys.argv[1:] if v not in ("--stats", "--fold")]
N, H, ci, co = [int(v) for v in argv[:4]] if len(argv) >= 4 else (64, 128, 128, 128)
x = torch.randn(N, H, H, ci, device="cuda").to(torch.bfloat16)
dy = torch.randn(N, H, H, co, device="cuda").to(torch.bfloat16)
w = (torch.randn(co, 9, ci, device="cuda") * 0.05).to(torch.bfloat16)
y = torch.empty(N, H, H, co, device="cuda", dtype=torch.bfloat16)
dx = torch.empty(N, H, H, ci, device="cuda", dtype=torch.bfloat16)
b = torch.randn(co, device="cuda")
q = torch.zeros(N, co // 4, 2, device="cuda") if stats else None
if fold:
    coef = torch.stack([1.0 + 0.1 * torch.randn(N, ci, device="cuda"), 0.1 * torch.randn(N, ci, device="cuda")], dim=-1).contiguous()
    for _ in range(2):
        ops.conv_fprop(x, w, y, N, H, H, 3, 1, bias=b, qsum=q)
        ops.conv_fprop(x, w, y, N, H, H, 3, 1, bias=b, qsum=q, gn_coef=coef)
else:
    for _ in range(2):
        ops.conv_fprop(x, w, y, N, H, H, 3, 1, bias=b, resid=dy if ci == co else None, qsum=q)
        ops.conv_dgrad(dy, w, dx, N, H, H, 3)
torch.cuda.synchronize()
