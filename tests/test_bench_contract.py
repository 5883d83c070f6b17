"""CPU: the bench.py contract.  (1) The argument parser keeps the driver's flags (--gpus/--steps/--warmup/--impl).
(2) The line the last GPU run printed (committed as profiles/bench_r02_final*.json) carries every key the contract
names -- metric/value/unit, e2e with its byte counts, roofline with the live kernel figures and the stamped ncu
traffic, cpu_baseline, clocks, gpu_launches -- and its numbers are self-consistent (value = batch x steps / time).
(3) profiles/igemm_traffic.json is stamped with the blob hash of the csrc/igemm.cu that is checked in, so
`roofline.traffic` will not be refused as stale."""
import glob
import hashlib
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lines():
    out = []
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r02_final*.json"))):
        with open(p) as f:
            rows = [json.loads(l) for l in f if l.startswith("{")]
        assert rows, p
        out.append((os.path.basename(p), rows[-1]))
    return out


def test_parser_keeps_the_driver_flags():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    src = open(os.path.join(ROOT, "bench.py")).read()
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert f'"{flag}"' in src, flag
    assert spec is not None


@pytest.mark.parametrize("name,d", _lines())
def test_committed_bench_line_has_the_contract_keys(name, d):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, (name, k)
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"].startswith("synthetic") and "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.8 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value is whole-job throughput: per-GPU batch x GPUs x steps / time
    b = d["config"]["per_gpu_batch"]
    want = d["n_gpus"] * b / (d["ms_per_step"] * 1e-3)
    assert abs(d["value"] - want) <= 2e-3 * want, (d["value"], want)
    r = d["roofline"]
    if r:      # the multi-GPU convenience runs skip it with --no-roofline
        assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
        assert abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-3
        assert r["kernel_ms_per_step"] <= d["ms_per_step"] * 1.5


def test_single_gpu_line_has_cpu_baseline_sampling_and_configs():
    d = dict(_lines())["bench_r02_final.json"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    s = d["sampling"]
    assert s["whole_loop_timed"] and s["timed_steps"] == 1000 and s["config"]["finite"]
    assert abs(s["value"] * s["ms_per_denoise_step"] - 256.0) <= 0.5        # images/s x ms per step = batch 256 over 1000 steps
    for k in ("c3_train", "c3_sampling", "c5_train"):
        assert d["configs"][k]["value"] > 0
    assert d["roofline"]["traffic"] and "verified" in d["roofline"]["traffic_source"]


def test_traffic_stamp_matches_the_checked_in_kernel():
    with open(os.path.join(ROOT, "profiles", "igemm_traffic.json")) as f:
        t = json.load(f)
    src = open(os.path.join(ROOT, "masked-diffusion-model_b200", "mdm_b200", "csrc", "igemm.cu"), "rb").read()
    blob = hashlib.sha1(b"blob %d\0" % len(src) + src).hexdigest()
    assert t["igemm_cu_blob"] == blob, "csrc/igemm.cu changed since profiles/launches_r02.csv was taken: rerun scripts/make_traffic.py"
    assert t["launches"] > 0 and t["traffic_bytes_per_launch"] > 0
