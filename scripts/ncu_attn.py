"""one attention forward at the configs[3] / configs[4] shapes and backward for `ncu --set full -k regex:attention_`:  python scripts/ncu_attn.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
from mdm_b200 import denoiser_ops as ops
for (B, L, C) in ((256, 64, 512), (8, 256, 1024)):
    qkv = torch.randn(B * L, 3 * C, device="cuda").to(torch.bfloat16)
    out = torch.empty(B * L, C, device="cuda", dtype=torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    for _ in range(2):
        ops.attention_fwd(qkv, out, B, L, C)
        ops.attention_bwd(qkv, out, dqkv, B, L, C)
torch.cuda.synchronize()
