"""(CPU) SASS census of libmmmot_sm100a.so: per kernel, how many tcgen05 / TMA / TMEM instructions the binary holds.
    python tools/sass_census.py [lib.so] > profiles/rNN_sass_census.txt
Mnemonics (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG,
cp.async.bulk -> UBLKCP, mbarrier -> SYNCS; HMMA would be the legacy mma.sync path (must be absent)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mmmot_b200", "libmmmot_sm100a.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "SYNCS", "HMMA", "LDG.E", "STG.E", "LDS", "STS", "FFMA", "F2FP", "total"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        per[cur]["total"] += 1
        for k in KEYS[:-1]:
            if op.startswith(k):
                per[cur][k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass); columns = instruction counts in the binary")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in KEYS))
tot = collections.Counter()
for (name, c), dm in zip(per.items(), demangle):
    short = re.sub(r"\(.*", "", dm.replace("(anonymous namespace)::", ""))[:70]
    print(f"{short:70s} " + " ".join(f"{c[k]:8d}" for k in KEYS))
    tot.update(c)
print(f"{'ALL KERNELS':70s} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
