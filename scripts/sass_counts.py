"""SASS evidence that the GEMM path is Blackwell-native: per-kernel counts of the tcgen05 / TMEM / TMA mnemonics
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = cp.async.bulk.tensor load / store /
reduce, UBLKCP = cp.async.bulk) and of the legacy warp-level HMMA (must be 0) in libmdm_sm100.so.
    python scripts/sass_counts.py > profiles/sass_counts_r02.txt      (runs here: cuobjdump needs no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "masked-diffusion-model_b200", "mdm_b200", "libmdm_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "IMMA", "ELECT"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        per[cur]["_total"] += 1
        if op in MN:
            per[cur][op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"cuobjdump -sass {os.path.relpath(lib, ROOT)}: {len(per)} kernels")
print(f"{'kernel':84s} {'instrs':>7s} " + " ".join(f"{m:>8s}" for m in MN))
for (k, c), name in zip(per.items(), demangle):
    tot.update(c)
    if any(c[m] for m in MN if m not in ("ELECT", "SYNCS")) or "igemm" in name:
        short = re.sub(r"\(.*", "", name.replace("mdm::", ""))[:84]
        print(f"{short:84s} {c['_total']:7d} " + " ".join(f"{c[m]:8d}" for m in MN))
print(f"{'TOTAL (all kernels)':84s} {tot['_total']:7d} " + " ".join(f"{tot[m]:8d}" for m in MN))
print("legacy HMMA / IMMA (warp-level mma.sync) in the library:", tot["HMMA"] + tot["IMMA"])
