"""Kernel timeline of graph-replayed training steps (torch.profiler / CUPTI): per-kernel busy time, idle gaps,
concurrency.  python scripts/timeline.py [--batch 128 --size 32 --steps 3] -> gpurun_out/timeline.json + table"""
import argparse, json, os, sys, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128); ap.add_argument("--size", type=int, default=32)
ap.add_argument("--steps", type=int, default=3); ap.add_argument("--method", default="base")
ap.add_argument("--sample", action="store_true", help="profile the sampler.py restoration loop (inference) instead of training steps")
pa = ap.parse_args()
dev = torch.device("cuda", 0)
if pa.sample:
    import sampler as sampler_mod, scheduler as scheduler_mod
    from mdm_b200.config import default_args
    from mdm_b200.denoiser import UNet2DModelB200, default_config
    S, N = pa.size, pa.batch
    sa = default_args(data_size=S, in_channel=3, out_channel=3, ddpm_num_steps=1000, ddpm_schedule="linear",
                      select_degrade_pixel="thresholding", degrade_channel="1-channel", mean_option="0", mean_area="image-wise",
                      method="base", shift_type="noise_with_perturbation", sample_latent_shape="zero",
                      momentum_adaptive="base_momentum", sampling_mask_dependency="independent", sample_num=N)
    sa.weight_dtype = torch.float32
    m = UNet2DModelB200(device=dev, **default_config(3, S)); m.reset_parameters(seed=0); m.eval()
    Sch = scheduler_mod.Scheduler(sa); Sch.update_ddpm_num_steps(1000); ts = Sch.get_timesteps_epoch(0, 1)
    smp = sampler_mod.Sampler(None, sa, Sch, [None, None, None])
    torch.manual_seed(0)
    smp.sample(m, ts[-4:]); torch.cuda.synchronize()      # the denoiser captures its forward graph on call 3
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        smp.sample(m, ts[-pa.steps:])
        torch.cuda.synchronize()
else:
    a = argparse.Namespace(batch=pa.batch, size=pa.size, channels=3, method=pa.method, no_graph=False)
    tr, model, acc = bench.build_trainer(a, bench.workload_args(a))
    x = (torch.rand(pa.batch, 3, pa.size, pa.size) * 2 - 1).to(dev)
    torch.manual_seed(0); tr.Scheduler.adopt_torch_rng(dev)
    for i in range(6):
        tr._run_batch(i, (x,), 0, 1, 0, None, None)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(pa.steps):
            tr._run_batch(i, (x,), 0, 1, 0, None, None)
        torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = "gpurun_out/timeline_trace.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
if not ks:
    print("no kernel events captured"); sys.exit(0)
t0, t1 = ks[0]["ts"], max(e["ts"] + e["dur"] for e in ks)
def short(n):
    n = re.sub(r"^void ", "", n); n = re.sub(r"\(.*", "", n); return n[:60]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ks:
    agg[short(e["name"])][0] += 1; agg[short(e["name"])][1] += e["dur"]
# union of busy intervals, and time covered by >= 2 kernels
pts = []
for e in ks:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
busy = multi = 0.0; depth = 0; last = pts[0][0]
for t, d in pts:
    if depth >= 1: busy += t - last
    if depth >= 2: multi += t - last
    depth += d; last = t
wall = t1 - t0
n = pa.steps
print(f"{n} steps: wall {wall/n:.0f} us/step, busy (>=1 kernel) {busy/n:.0f} us, idle {(wall-busy)/n:.0f} us, >=2 kernels {multi/n:.0f} us, kernels/step {len(ks)/n:.0f}, sum of durations {sum(e['dur'] for e in ks)/n:.0f} us")
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{us/n:9.0f} us/step  n={c/n:6.1f}  avg={us/c:7.1f}  {name}")
# duration histogram of the small activation GEMMs and the finalize kernel
for pat in ("igemm_kernel<0, false, 1>", "igemm_kernel<1, false, 2>", "splitk_finalize", "gn_small_fwd_kernel<4>"):
    d = sorted(e["dur"] for e in ks if pat in e["name"])
    if d:
        q = lambda f: d[min(len(d) - 1, int(f * len(d)))]
        print(f"  {pat}: n/step={len(d)/n:.0f} min {d[0]:.1f} p25 {q(.25):.1f} p50 {q(.5):.1f} p75 {q(.75):.1f} p90 {q(.9):.1f} max {d[-1]:.1f} us")
# gaps between consecutive kernels on the busiest stream
by_stream = collections.defaultdict(list)
for e in ks: by_stream[e.get("tid")].append(e)
main = max(by_stream.values(), key=len)
gaps = [main[i + 1]["ts"] - (main[i]["ts"] + main[i]["dur"]) for i in range(len(main) - 1)]
gaps = [g for g in gaps if g < 1000]
print(f"main stream: {len(main)/n:.0f} kernels/step, kernel time {sum(e['dur'] for e in main)/n:.0f} us, gaps total {sum(g for g in gaps if g > 0)/n:.0f} us, median gap {sorted(gaps)[len(gaps)//2]:.2f} us")
# per-launch list of the LAST step (start relative to its first kernel, duration, stream, grid, block, name)
if n >= 1:
    per = len(ks) // n
    last_step = ks[-per:]
    base = last_step[0]["ts"]
    with open("gpurun_out/timeline_launches.csv", "w") as f:
        f.write("start_us,dur_us,stream,grid,block,kernel\n")
        for e in last_step:
            a = e.get("args", {})
            f.write(f"{e['ts'] - base:.1f},{e['dur']:.1f},{a.get('stream', '')},{'x'.join(map(str, a.get('grid', [])))},{'x'.join(map(str, a.get('block', [])))},\"{short(e['name'])}\"\n")
json.dump({"wall_us_per_step": wall / n, "busy_us": busy / n, "by_kernel": {k: v for k, v in agg.items()}}, open("gpurun_out/timeline.json", "w"))
os.remove(path)
