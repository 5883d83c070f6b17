"""Data-parallel step time under different overlap / NCCL settings, all inside ONE gpurun call (A/B on the same box).
    python scripts/dp_sweep.py N [config ...]     -> gpurun_out/dp_sweep_N.json
Every configuration runs `bench.py --gpus N` (headline only) under torchrun; N = 1 is measured first as the yardstick."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
NCCL = {"MDM_DP_ALLREDUCE": "nccl"}
CONFIGS = {
    "ce":             {"MDM_P2P_MODE": "ce"},                        # default: copy-engine all-reduce inside the step graph
    "ce_1stream":     {"MDM_P2P_MODE": "ce", "MDM_P2P_COPY_STREAMS": "1"},
    "ce_7streams":    {"MDM_P2P_MODE": "ce", "MDM_P2P_COPY_STREAMS": "7"},
    "sm48":           {"MDM_P2P_MODE": "sm", "MDM_P2P_BLOCKS": "48"},
    "sm32":           {"MDM_P2P_MODE": "sm", "MDM_P2P_BLOCKS": "32"},
    "sm24":           {"MDM_P2P_MODE": "sm", "MDM_P2P_BLOCKS": "24"},
    "sm16":           {"MDM_P2P_MODE": "sm", "MDM_P2P_BLOCKS": "16"},
    "p2p":            {"MDM_P2P_MODE": "sm"},
    "p2p_cuts_2":     {"MDM_DP_CUTS": "2"},
    "p2p_cuts_54321": {"MDM_DP_CUTS": "5,4,3,2,1"},
    "p2p_blocks48":   {"MDM_P2P_BLOCKS": "48"},
    "p2p_blocks128":  {"MDM_P2P_BLOCKS": "128"},
    "p2p_static":     {"MDM_IGEMM_DYNAMIC": "0"},
    "nccl":           dict(NCCL),                                    # round-1 path: segmented graphs + async NCCL
    "nccl_cuts_2":    dict(NCCL, MDM_DP_CUTS="2"),
    "nccl_static":    dict(NCCL, MDM_IGEMM_DYNAMIC="0"),
    "nccl_nch8":      dict(NCCL, NCCL_MAX_NCHANNELS="8"),
    "nccl_nthreads256": dict(NCCL, NCCL_NTHREADS="256", NCCL_MAX_NCHANNELS="16"),
    "nccl_reserve16": dict(NCCL, MDM_IGEMM_MAX_CTAS="132"),
    "nccl_no_overlap": dict(NCCL, MDM_DP_OVERLAP="0"),
}
want = sys.argv[2:] or list(CONFIGS)
flags = ["--steps", "30", "--warmup", "8", "--no-sampling", "--no-extra", "--no-cpu-baseline", "--no-roofline"]


def run(n, env_over):
    env = dict(os.environ)
    env.update(env_over)
    if n == 1:
        cmd = [sys.executable, "bench.py", "--gpus", "1"] + flags
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
               "--master-port", "29533", "bench.py", "--gpus", str(n)] + flags
    t0 = time.time()
    try:
        p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    except subprocess.TimeoutExpired:
        return {"error": "timeout"}
    line = next((l for l in p.stdout.splitlines() if l.startswith("{")), None)
    if line is None:
        return {"error": (p.stderr or "")[-400:], "rc": p.returncode}
    d = json.loads(line)
    return {"ms_per_step": d["ms_per_step"], "value": d["value"], "e2e": d["e2e"]["value"], "wall_s": round(time.time() - t0, 1),
            "clocks": d.get("clocks", {}).get("sm_mhz")}


out = {"n": N, "single": run(1, {})}
print("single", out["single"], flush=True)
for name in want:
    out[name] = run(N, CONFIGS[name])
    r = out[name]
    if "ms_per_step" in r and "ms_per_step" in out["single"]:
        r["efficiency"] = round(out["single"]["ms_per_step"] / r["ms_per_step"], 4)
    print(name, CONFIGS[name], r, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"dp_sweep_{N}.json"), "w"), indent=1)
