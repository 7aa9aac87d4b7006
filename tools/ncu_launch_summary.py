"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of `python bench.py ...` next to the bench JSON of the
same command run plainly: per-kernel launches / ms per timed step and the GEMM share by ncu vs by CUDA events.
usage: python tools/ncu_launch_summary.py <launches.csv> <bench.json> <steps+warmup+e2e passes over the pipeline> > summary.txt"""
import csv, json, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", "")); u = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        mk = re.search(r"\bk_\w+(<[^>(]*>)?", r["Kernel Name"])
        rows.append((mk.group(0) if mk else re.sub(r"\(.*", "", r["Kernel Name"]).strip(), ms))
bench = json.loads([ln for ln in open(sys.argv[2]) if ln.startswith("{")][-1])
passes = float(sys.argv[3])
agg = defaultdict(lambda: [0, 0.0])
for k, ms in rows:
    agg[k][0] += 1; agg[k][1] += ms
mine = {k: v for k, v in agg.items() if k.startswith("k_") or "rsvdb" in k}
tot = sum(v[1] for v in mine.values())
print(f"ncu launch list: {len(rows)} launches, {len(mine)} distinct kernels of this library; {passes:g} passes over the rSVD pipeline in the command")
print(f"total of this library's kernels: {tot:.2f} ms  ({tot / passes:.2f} ms per pass; per-launch times are cold-cache and serialised: compare SHARES)\n")
for k, (n, ms) in sorted(mine.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:<44} launches/pass={n / passes:6.1f}  ms/pass={ms / passes:9.3f}  share={100 * ms / tot:5.1f}%")
g = sum(v[1] for k, v in mine.items() if k.startswith("k_gemm_a") or k.startswith("k_reduce_splits"))
r = bench["roofline"]
# the same two kernels also run the skinny products of the CholeskyQR2 orthonormalisations (Gram matrices, Y * R^-1); the passes over A
# are the launches of the headline shape, told apart by their duration (> 25 % of the longest launch)
big = max(ms for k, ms in rows if k.startswith("k_gemm_a"))
ga = sum(ms for k, ms in rows if k.startswith("k_gemm_a") and ms > 0.25 * big)
na = sum(1 for k, ms in rows if k.startswith("k_gemm_a") and ms > 0.25 * big)
print(f"\nall launches of k_gemm_an + k_gemm_at + k_reduce_splits: ncu {100 * g / tot:.1f}% of the library's kernel time")
print(f"passes over A only ({na / passes:g} launches per pass, {ga / passes:.2f} ms per pass): ncu {100 * ga / tot:.1f}%   vs   CUDA events in bench.py "
      f"(phases gemm_an + gemm_at) {100 * r['gemm_ms_per_step'] / bench['ms_per_step']:.1f}%")
