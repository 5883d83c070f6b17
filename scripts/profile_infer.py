"""Per-op device-time breakdown of one denoiser forward (inference plan): python scripts/profile_infer.py [B] [S]"""
import os, sys, collections
os.environ["MDM_INFER_GRAPH"] = "0"      # per-op events need the eager program
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
import torch
from mdm_b200 import denoiser_ops as ops
from mdm_b200.denoiser import UNet2DModelB200, default_config
from mdm_b200.config import unet_forward_flops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
m = UNet2DModelB200(device="cuda", **default_config(3, S)); m.reset_parameters(seed=0); m.eval()
x = torch.rand(B, 3, S, S, device="cuda") * 2 - 1
t = torch.full((B,), 500.0, device="cuda")
with torch.no_grad():
    for _ in range(2): m(x, t)
rec = []
orig = {}
def desc(args, kw):
    out = []
    for v in list(args) + list(kw.values()):
        if torch.is_tensor(v): out.append("x".join(map(str, v.shape)))
        elif isinstance(v, (int, bool)): out.append(str(int(v)))
    return ",".join(out[:8])
for n in dir(ops):
    f = getattr(ops, n)
    if not callable(f) or n.startswith("_") or not hasattr(f, "__code__") or n in ("pix_ld", "pack_conv_weight", "unpack_conv_weight", "gn_ws_floats"): continue
    def mk(n, f):
        def g(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = f(*a, **k); e1.record(); rec.append((n, desc(a, k), e0, e1)); return r
        return g
    setattr(ops, n, mk(n, f))
torch.cuda.synchronize(); torch.cuda._sleep(int(3e8))
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    s0.record(); m(x, t); s1.record()
torch.cuda.synchronize()
rows = [(n, d, e0.elapsed_time(e1) * 1e3) for n, d, e0, e1 in rec]
tot = sum(r[2] for r in rows)
print(f"forward B={B} S={S}: ops {tot/1e3:.2f} ms, wall {s0.elapsed_time(s1):.2f} ms; {B*unet_forward_flops(m._cfg,S)/(s0.elapsed_time(s1)*1e-3)/1e12:.0f} TF/s overall")
agg = collections.defaultdict(lambda: [0, 0.0])
for n, d, us in rows: agg[n][0] += 1; agg[n][1] += us
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{us:9.0f} us {100*us/tot:5.1f}%  n={c:4d}  {n}")
agg2 = collections.defaultdict(lambda: [0, 0.0])
for n, d, us in rows: agg2[(n, d)][0] += 1; agg2[(n, d)][1] += us
for (n, d), (c, us) in sorted(agg2.items(), key=lambda kv: -kv[1][1])[:28]: print(f"{us:9.0f} us {100*us/tot:5.1f}%  n={c:3d} avg={us/c:8.1f}  {n}  {d}")
