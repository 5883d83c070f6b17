"""Data-parallel consistency check (run under torchrun, >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py
Trains a few steps twice from the same seed -- once with the gradient all-reduce between the graphs (3-piece step) and
once with the segmented backward whose all-reduces overlap the following segments -- and compares losses and
parameters (tolerances 1e-4 / 5e-5: the fp32 reduce-add order of the weight gradients differs run to run); also
checks that every rank holds bit-identical parameters afterwards."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
import torch
import torch.distributed as dist
import bench

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6


def run(overlap):
    a = argparse.Namespace(batch=32, size=32, channels=3, method="base", no_graph=False)
    wa = bench.workload_args(a)
    wa.dp_overlap = overlap
    tr, model, acc = bench.build_trainer(a, wa)
    g = torch.Generator().manual_seed(1000 + rank)
    xs = [(torch.rand(32, 3, 32, 32, generator=g) * 2 - 1).to(dev) for _ in range(4)]
    torch.manual_seed(0)
    tr.Scheduler.adopt_torch_rng(dev)
    losses = []
    for i in range(steps):
        out = tr._run_batch(i, (xs[i % 4],), 0, 1, 0, None, None)
        losses.append(float(out[0] if isinstance(out, tuple) else out))
    torch.cuda.synchronize()
    p = model.flat_param.clone()
    return losses, p


l0, p0 = run(False)
l1, p1 = run(True)
ref = p1.clone()
dist.broadcast(ref, src=0)
same_ranks = bool((ref == p1).all().item())
diff = (p0 - p1).abs().max().item()
if rank == 0:
    print("losses 3-piece :", [round(v, 6) for v in l0])
    print("losses overlap :", [round(v, 6) for v in l1])
    print(f"max |param diff| between the two schedules: {diff:.3e}; identical parameters on every rank: {same_ranks}")
ok = torch.tensor([1.0 if (same_ranks and diff <= 5e-5 and all(abs(a - b) <= 1e-4 for a, b in zip(l0, l1))) else 0.0], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DP_CHECK", "OK" if ok.item() == 1.0 else "FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok.item() == 1.0 else 1)
