"""Per-kernel census of the tensor-core / TMA / TMEM instructions in libldm_b200.so (cuobjdump -sass), the evidence that
the convolution path is tcgen05 + TMEM + TMA and which kernels are still on the legacy mma.sync (HMMA) pipe.

    python tools/sass_census.py [lib] > profiles/r02_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "latent-diffusion-models_b200", "libldm_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "SYNCS", "LDGSTS", "FFMA", "MUFU", "BAR.SYNC"]
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = {}
cur = None
counts = collections.OrderedDict()
total = collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_insts"] += 1
    for k in KEYS:
        if op.startswith(k):
            counts[cur][k] += 1
            total[k] += 1
names = list(counts)
try:
    out = subprocess.run(["c++filt"] + names, stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
    demangle = dict(zip(names, out))
except Exception:
    demangle = {n: n for n in names}
print(f"# SASS census of {os.path.basename(lib)} (sm_100a): instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load, LDTM/STTM = tcgen05.ld/st (TMEM),")
print("# HMMA = legacy mma.sync, SYNCS = mbarrier, LDGSTS = cp.async")
print(f"{'insts':>7} " + " ".join(f"{k:>8}" for k in KEYS) + "  kernel")
for n in sorted(names, key=lambda n: -(counts[n]['UTCHMMA'] * 1000 + counts[n]['HMMA'])):
    c = counts[n]
    if not any(c[k] for k in ("UTCHMMA", "HMMA", "UTMALDG", "LDTM")) and c["_insts"] < 400:
        continue
    short = re.sub(r"\(anonymous namespace\)::", "", demangle.get(n, n))
    short = re.sub(r"\(.*", "", short)
    print(f"{c['_insts']:7d} " + " ".join(f"{c[k]:8d}" for k in KEYS) + f"  {short[:90]}")
print("total   " + " ".join(f"{total[k]:8d}" for k in KEYS))
