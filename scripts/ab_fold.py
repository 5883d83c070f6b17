"""A/B of MDM_GN_FOLD (GroupNorm + SiLU folded into the consumer convolution, inference) on one graph-replayed
denoiser forward: python scripts/ab_fold.py [B] [S] -- prints ms per forward for fold = 0 / 1 and the output difference."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
if len(sys.argv) > 3:                       # child: one setting
    import torch
    from mdm_b200.denoiser import UNet2DModelB200, default_config
    m = UNet2DModelB200(device="cuda", **default_config(3, S)); m.reset_parameters(seed=0); m.eval()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(B, 3, S, S, device="cuda", generator=g) * 2 - 1
    t = torch.full((B,), 500.0, device="cuda")
    with torch.no_grad():
        for _ in range(3): y = m(x, t)
        y = y.sample if hasattr(y, "sample") else y
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): m(x, t)
        e1.record(); torch.cuda.synchronize()
    torch.save(y.float().cpu(), sys.argv[3])
    print(f"MDM_GN_FOLD={os.environ.get('MDM_GN_FOLD')}: {e0.elapsed_time(e1) / 10:.3f} ms per forward, finite {bool(torch.isfinite(y).all())}")
else:
    import torch
    outs = []
    for f in ("0", "1"):
        p = f"/tmp/ab_fold_{f}.pt"
        subprocess.run([sys.executable, __file__, str(B), str(S), p], env=dict(os.environ, MDM_GN_FOLD=f), check=True, timeout=600)
        outs.append(torch.load(p))
    d = (outs[0] - outs[1]).norm() / outs[0].norm()
    print(f"rel L2 (fold 1 vs fold 0) = {d:.3e}")
