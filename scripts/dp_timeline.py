"""Kernel timeline of data-parallel training steps on rank 0 (torch.profiler / CUPTI), under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_timeline.py
-> gpurun_out/dp_timeline_N.txt: per-kernel busy time, when each all-reduce kernel ran and how long, step wall time."""
import argparse, collections, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
import torch
import torch.distributed as dist
import bench
from torch.profiler import profile, ProfilerActivity

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
a = argparse.Namespace(batch=128, size=32, channels=3, method="base", no_graph=False)
tr, model, acc = bench.build_trainer(a, bench.workload_args(a))
g = torch.Generator().manual_seed(1000 + rank)
x = (torch.rand(128, 3, 32, 32, generator=g) * 2 - 1).to(dev)
torch.manual_seed(0); tr.Scheduler.adopt_torch_rng(dev)
for i in range(8):
    tr._run_batch(i, (x,), 0, 1, 0, None, None)
torch.cuda.synchronize(); dist.barrier() if world > 1 else None
steps = 3
if rank == 0:
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(steps):
            tr._run_batch(i, (x,), 0, 1, 0, None, None)
        torch.cuda.synchronize()
else:
    for i in range(steps):
        tr._run_batch(i, (x,), 0, 1, 0, None, None)
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    path = f"gpurun_out/dp_trace_{world}.json"
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    ks = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e], key=lambda e: e["ts"])
    os.remove(path)
    t0, t1 = ks[0]["ts"], max(e["ts"] + e["dur"] for e in ks)
    short = lambda n: re.sub(r"\(.*", "", re.sub(r"^void ", "", n))[:56]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ks:
        agg[short(e["name"])][0] += 1; agg[short(e["name"])][1] += e["dur"]
    out = [f"world {world}: {steps} steps, wall {(t1 - t0) / steps:.0f} us/step, {len(ks) / steps:.0f} kernels/step, mode {os.environ.get('MDM_DP_ALLREDUCE', 'p2p')}"]
    for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        out.append(f"  {d / steps:9.0f} us/step n={n / steps:6.1f} avg={d / n:8.1f}  {k}")
    comm = [e for e in ks if "allreduce" in e["name"].lower() or "nccl" in e["name"].lower()]
    out.append("communication kernels (start relative to the first kernel of the trace, duration):")
    for e in comm:
        out.append(f"  t={e['ts'] - t0:9.0f} us  dur={e['dur']:8.0f} us  {short(e['name'])}")
    adam = [e for e in ks if "adam" in e["name"]]
    out.append("optimiser kernels at: " + ", ".join(f"{e['ts'] - t0:.0f}" for e in adam))
    open(f"gpurun_out/dp_timeline_{world}_{os.environ.get('MDM_DP_ALLREDUCE', 'p2p')}.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
