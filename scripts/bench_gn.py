"""micro-benchmark of the GroupNorm / colsum kernels: python scripts/bench_gn.py [iters]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
from mdm_b200 import denoiser_ops as ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
SHAPES = [(128, 32, 128), (128, 32, 256), (128, 16, 128), (128, 8, 256), (128, 2, 512), (64, 128, 128), (64, 64, 256), (256, 128, 128)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                      # evict L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3
for N, H, C in SHAPES:
    x = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)
    y = torch.empty_like(x); dx = torch.empty_like(x)
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    stats = torch.empty(N, 32, 2, device="cuda")
    ws = torch.empty(max(1, ops.gn_ws_floats(N, H * H, C)), device="cuda")
    nb = x.numel() * 2
    tf = timeit(lambda: ops.gn_silu_fwd(x, y, gamma, beta, stats, ws, N, H * H, C, 32, 1e-5, True))
    tb = timeit(lambda: ops.gn_silu_bwd(x, dy, dx, gamma, beta, stats, dg, db, ws, N, H * H, C, 32, True, add2=y))
    tb0 = timeit(lambda: ops.gn_silu_bwd(x, dy, dx, gamma, beta, stats, dg, db, ws, N, H * H, C, 32, True))     # norm2 sites: no add / add2
    tc = timeit(lambda: ops.colsum(dy, dg, N * H * H, C))
    # apply pass alone (statistics handed over as quad sums, the path behind a convolution): read x + write y
    qa = torch.zeros(N, C // 4, 2, device="cuda"); qa[:, :, 1] = 4.0 * H * H
    ta1 = timeit(lambda: ops.gn_silu_fwd_q(x, y, gamma, beta, stats, qa, None, N, H * H, C, 32, 1e-5, True))
    ta0 = timeit(lambda: ops.gn_silu_fwd_q(x, y, gamma, beta, stats, qa, None, N, H * H, C, 32, 1e-5, False))
    print(f"   apply-only: silu {ta1:6.1f} us = {2*nb/ta1/1e6:5.2f} TB/s | no silu {ta0:6.1f} us = {2*nb/ta0/1e6:5.2f} TB/s", flush=True)
    # algorithmic bytes: fwd = read x + write y; bwd = read x, dy, add2 + write dx
    print(f"N={N} H={H} C={C} ({nb/1e6:.1f} MB): fwd {tf:6.1f} us = {2*nb/tf/1e6:5.2f} TB/s (2 passes) | bwd {tb:6.1f} us = {4*nb/tb/1e6:5.2f} TB/s (4 passes) | bwd(no add) {tb0:6.1f} us = {3*nb/tb0/1e6:5.2f} TB/s (3 passes) | colsum {tc:6.1f} us = {nb/tc/1e6:5.2f} TB/s", flush=True)
