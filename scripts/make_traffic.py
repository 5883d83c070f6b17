"""From a raw ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv --log-file RAW.csv python bench.py ...`, long format: one row per metric per launch):

  * profiles/<name>.csv            -- the trimmed wide list (ID, kernel, grid, us, DRAM read MB, DRAM write MB)
  * profiles/<name>_summary.txt    -- per-kernel totals and shares of the serialised kernel time
  * profiles/igemm_traffic.json    -- average dram bytes (read + write) per igemm launch, stamped with the git blob
                                      hash of csrc/igemm.cu so that bench.py refuses a figure from another kernel version

    python scripts/make_traffic.py gpurun_out/launches_raw.csv launches_r02"""
import csv, hashlib, json, os, sys
from collections import OrderedDict, defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, name = sys.argv[1], sys.argv[2]
lines = [l for l in open(src, errors="replace") if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ik, im, iv, iu, iid = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
ig = hdr.index("Grid Size") if "Grid Size" in hdr else None
scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
L = OrderedDict()
for r in rows[1:]:
    d = L.setdefault(r[iid], {"kernel": r[ik], "grid": r[ig] if ig is not None else "", "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r[iv].replace(",", "")) * scale.get(r[iu].lower(), 1.0)
    if r[im] == "gpu__time_duration.sum":
        d["us"] = v
    elif r[im] == "dram__bytes_read.sum":
        d["rd"] = v
    elif r[im] == "dram__bytes_write.sum":
        d["wr"] = v
with open(os.path.join(ROOT, "profiles", name + ".csv"), "w") as f:
    f.write("ID,Kernel Name,Grid Size,gpu__time_duration.sum [us],dram__bytes_read.sum [MB],dram__bytes_write.sum [MB]\n")
    for k, d in L.items():
        f.write(f'{k},"{d["kernel"][:70]}","{d["grid"]}",{d["us"]:.2f},{d["rd"] / 1e6:.3f},{d["wr"] / 1e6:.3f}\n')
tot = sum(d["us"] for d in L.values())
by = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in L.values():
    key = d["kernel"].split("(")[0][:60]
    b = by[key]
    b[0] += 1; b[1] += d["us"]; b[2] += d["rd"]; b[3] += d["wr"]
with open(os.path.join(ROOT, "profiles", name + "_summary.txt"), "w") as f:
    f.write(f"{len(L)} launches, {tot / 1e3:.2f} ms of serialised cold-cache kernel time ({src})\n")
    for key, b in sorted(by.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{b[1]:10.1f} us {100 * b[1] / tot:5.1f}%  n={b[0]:4d} avg={b[1] / b[0]:8.1f} us  rd={b[2] / 1e6:9.1f} MB wr={b[3] / 1e6:9.1f} MB  {key}\n")
ig_l = [d for d in L.values() if "igemm_kernel" in d["kernel"]]
data = open(os.path.join(ROOT, "masked-diffusion-model_b200", "mdm_b200", "csrc", "igemm.cu"), "rb").read()
blob = hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()
out = {"traffic_bytes_per_launch": round(sum(d["rd"] + d["wr"] for d in ig_l) / max(1, len(ig_l)), 1), "launches": len(ig_l),
       "total_GB": round(sum(d["rd"] + d["wr"] for d in ig_l) / 1e9, 3),
       "igemm_share_of_kernel_time": round(sum(d["us"] for d in ig_l) / max(tot, 1e-9), 4),
       "source": f"profiles/{name}.csv", "igemm_cu_blob": blob}
json.dump(out, open(os.path.join(ROOT, "profiles", "igemm_traffic.json"), "w"), indent=1)
print(out)
