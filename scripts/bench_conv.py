"""micro-benchmark of the igemm kernel on the heavy conv shapes: python scripts/bench_conv.py [shape-set]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
from mdm_b200 import denoiser_ops as ops
SHAPES = {  # N, H, cin, cout, k
    "c2": [(128, 32, 128, 128, 3), (128, 32, 256, 128, 3), (128, 16, 128, 128, 3), (128, 8, 256, 256, 3), (128, 4, 256, 256, 3),
           (128, 2, 512, 512, 3), (128, 1, 512, 512, 3), (128, 32, 256, 128, 1)],
    "c4": [(64, 128, 128, 128, 3), (64, 128, 256, 128, 3), (64, 64, 128, 128, 3), (64, 32, 256, 256, 3), (64, 16, 256, 256, 3),
           (64, 8, 512, 512, 3), (64, 4, 512, 512, 3)],
}
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for N, H, ci, co, k in SHAPES[which]:
    x = torch.randn(N, H, H, ci, device="cuda").to(torch.bfloat16)
    dy = torch.randn(N, H, H, co, device="cuda").to(torch.bfloat16)
    w = (torch.randn(co, k * k, ci, device="cuda") * 0.05).to(torch.bfloat16)
    y = torch.empty(N, H, H, co, device="cuda", dtype=torch.bfloat16)
    dx = torch.empty(N, H, H, ci, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(co, k * k, ci, device="cuda")
    b = torch.randn(co, device="cuda")
    fl = 2.0 * N * H * H * co * ci * k * k
    tf = timeit(lambda: ops.conv_fprop(x, w, y, N, H, H, k, 1, bias=b, resid=dy))
    td = timeit(lambda: ops.conv_dgrad(dy, w, dx, N, H, H, k)) if co % 64 == 0 and ci % 128 == 0 else float("nan")
    tw = timeit(lambda: ops.conv_wgrad(x, dy, dw, N, H, H, k, 1))
    print(f"N={N} H={H} ci={ci} co={co} k={k}: fprop {tf:7.1f} us {fl/tf/1e6:6.0f} TF/s | dgrad {td:7.1f} us {fl/td/1e6:6.0f} TF/s | wgrad {tw:7.1f} us {fl/tw/1e6:6.0f} TF/s", flush=True)
