"""Plain vs folded (GroupNorm + SiLU in the operand path) convolution per layer shape of a c4 forward, CUDA events:
    python scripts/bench_fold_conv.py      (FT_SHORT=1: three shapes)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
import torch
from mdm_b200 import denoiser_ops as ops
def run(N, H, cin, cout, sc, st):
    g = torch.Generator(device="cuda").manual_seed(11)
    xb = torch.randn(N, H, H, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (9 * cin) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    coef = torch.stack([1.0 + 0.5 * torch.randn(N, cin, device="cuda", generator=g), 0.5 * torch.randn(N, cin, device="cuda", generator=g)], dim=-1).contiguous()
    wb = ops.pack_conv_weight(w).to(torch.bfloat16)
    xs = torch.randn(N, H, H, 192, device="cuda", generator=g).to(torch.bfloat16)
    ws_ = torch.randn(cout, 192, 1, 1, device="cuda", generator=g) / 192 ** 0.5
    y = torch.empty(N, H, H, cout, device="cuda", dtype=torch.bfloat16)
    q = torch.zeros(N, cout // 4, 2, device="cuda") if st else None
    kw = dict(x2=xs, w2=ops.pack_conv_weight(ws_).to(torch.bfloat16)) if sc else {}
    res = []
    for coefv in (None, coef):
        for _ in range(3): ops.conv_fprop(xb, wb, y, N, H, H, 3, 1, bias=b, qsum=q, gn_coef=coefv, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): ops.conv_fprop(xb, wb, y, N, H, H, 3, 1, bias=b, qsum=q, gn_coef=coefv, **kw)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10 * 1e3)
    fl = 2 * N * H * H * cout * (9 * cin + (192 if sc else 0))
    print(f"N={N} H={H} {cin}->{cout} sc={sc} stats={st}: plain {res[0]:.0f} us ({fl/res[0]/1e6:.0f} TF/s)  folded {res[1]:.0f} us ({fl/res[1]/1e6:.0f} TF/s)  delta {res[1]-res[0]:+.0f} us; apply pass would move {2*xb.numel()*2/1e6:.0f} MB", flush=True)
cases = [(256, 128, 128, 128, 0, 1), (256, 128, 128, 128, 0, 0), (256, 128, 256, 128, 1, 1), (256, 64, 128, 128, 0, 1), (256, 32, 256, 256, 0, 1), (256, 32, 512, 256, 1, 1), (256, 16, 768, 256, 1, 0)]
if os.environ.get("FT_SHORT"): cases = [cases[1], cases[2], cases[6]]
for c in cases:
    run(*c)
