"""Per-kernel SASS evidence for librsvdb.so: counts of the mnemonics that show what a kernel runs on
(DMMA = FP64 tensor core, UTMALDG / UBLKCP = TMA, SYNCS / mbarrier traffic, LDS / STS, SHFL, cluster barriers), plus registers
and shared memory from the ptxas log of a forced rebuild.  Run in the dev container (no GPU needed):

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rsvd_kamaneh_raganato_terrana_b200 import build as B  # noqa: E402

MNEMONICS = ["DMMA", "UTMALDG", "UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "UCGABAR", "ATOM", "RED", "LDGSTS"]


def main():
    lib = B.LIB
    out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict(); cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); kernels[cur] = Counter(); continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    kernels[cur][mn] += 1
    dem = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, dem if len(dem) == len(kernels) else list(kernels)))
    print(f"SASS summary of {lib.name} (cuobjdump -sass; sm_100a).  Columns: instructions, then counts of the listed mnemonics.")
    print("DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64 lowers to DMMA.8x8x4); UTMALDG = TMA tensor load (cp.async.bulk.tensor);")
    print("SYNCS = mbarrier arrive/try_wait; UCGABAR = barrier.cluster.  There is no tcgen05 / TMEM form for FP64 operands.\n")
    hdr = f"{'kernel':<58}{'inst':>7}" + "".join(f"{m:>8}" for m in MNEMONICS)
    print(hdr)
    for k, c in kernels.items():
        mk = re.search(r"\bk_\w+(<[^>]*>)?", names[k])
        nm = mk.group(0).replace("(int)", "") if mk else names[k][:57]
        print(f"{nm[:57]:<58}{c['_total']:>7}" + "".join(f"{c[m]:>8}" for m in MNEMONICS))


if __name__ == "__main__":
    main()
