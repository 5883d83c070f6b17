"""Per-op device-time breakdown of one eager training step (CUDA events around every denoiser_ops call).
    python scripts/profile_ops.py [--batch 128 --size 32] -> gpurun_out/ops_profile.json + table on stdout"""
import argparse, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
import torch
import bench
from mdm_b200 import denoiser_ops as ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128); ap.add_argument("--size", type=int, default=32)
ap.add_argument("--channels", type=int, default=3); ap.add_argument("--method", default="base")
ap.add_argument("--out", default="gpurun_out/ops_profile.json")
ap.add_argument("--base", type=int, default=128, help="block_out_channels base (256 = BASELINE configs[4])")
pa = ap.parse_args()
a = argparse.Namespace(batch=pa.batch, size=pa.size, channels=pa.channels, method=pa.method, no_graph=True)
tr, model, acc = bench.build_trainer(a, bench.workload_args(a), base=pa.base)
model.wgrad_side_stream = False      # serialise the weight-gradient GEMMs with the main chain: one kernel at a time under every event pair
dev = torch.device("cuda", 0)
x = (torch.rand(pa.batch, pa.channels, pa.size, pa.size) * 2 - 1).to(dev)
torch.manual_seed(0); tr.Scheduler.adopt_torch_rng(dev)
for i in range(3):
    tr._run_batch(i, (x,), 0, 1, 0, None, None)
rec = []
names = [n for n in dir(ops) if callable(getattr(ops, n)) and not n.startswith("_") and n not in ("pix_ld", "pack_conv_weight", "unpack_conv_weight", "gn_ws_floats", "check", "lib", "ptr", "stream_ptr", "ConvArgs", "POINTER", "Structure")]
orig = {}
def desc(name, args, kw):
    out = []
    for v in list(args) + list(kw.values()):
        if torch.is_tensor(v): out.append("x".join(map(str, v.shape)))
        elif isinstance(v, (int, bool)): out.append(str(int(v)))
    return ",".join(out[:9])
for n in names:
    f = getattr(ops, n)
    if not hasattr(f, "__code__"): continue
    orig[n] = f
    def mk(n, f):
        def g(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = f(*args, **kw); e1.record()
            rec.append((n, desc(n, args, kw), e0, e1))
            return r
        return g
    setattr(ops, n, mk(n, f))
torch.cuda.synchronize()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(3e8))     # keep the GPU busy while the host enqueues the step: events then bracket kernels only
s0.record(); tr._run_batch(0, (x,), 0, 1, 0, None, None); s1.record(); torch.cuda.synchronize()
rows = [(n, d, e0.elapsed_time(e1) * 1e3) for n, d, e0, e1 in rec]
tot = sum(r[2] for r in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, d, us in rows:
    agg[n][0] += 1; agg[n][1] += us
print(f"ops total {tot:.0f} us over {len(rows)} calls; step wall {s0.elapsed_time(s1)*1e3:.0f} us")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:9.0f} us {100*us/tot:5.1f}%  n={c:4d}  avg={us/c:7.1f}  {n}")
agg2 = collections.defaultdict(lambda: [0, 0.0])
for n, d, us in rows:
    agg2[(n, d)][0] += 1; agg2[(n, d)][1] += us
print("--- by shape (top 60)")
for (n, d), (c, us) in sorted(agg2.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOP", 60))]:
    print(f"{us:9.0f} us {100*us/tot:5.1f}%  n={c:3d}  avg={us/c:7.1f}  {n}  {d}")
os.makedirs(os.path.dirname(pa.out), exist_ok=True)
json.dump(rows, open(pa.out, "w"))
