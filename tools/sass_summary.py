"""Per-kernel SASS evidence for profiles/: which kernels of libedrl_b200.so carry tcgen05 / TMEM / TMA instructions
(UTC*MMA, LDTM / STTM, UTMALDG), packed-fp32 math, and that no legacy HMMA is present.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "robust-multimodal-learning-for-ophthalmic-disease-grading-via-disentangled-representation_b200",
                   "lib", "libedrl_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "FMUL2",
       "FADD2", "ATOMS", "REDUX", "CREDUX", "FMNMX3", "SHFL", "LDG.E.128", "STG.E.128"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void edrl::", "")
        counts[name] = collections.Counter()
        total[name] = 0
        continue
    if name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        total[name] += 1
        body = line.split("*/", 1)[1]
        for p in PAT:
            if re.search(r"\b" + re.escape(p), body):
                counts[name][p] += 1
print("kernel | instructions | " + " | ".join(PAT))
agg = collections.Counter()
for k, c in counts.items():
    agg.update(c)
    if total[k]:
        print(f"{k} | {total[k]} | " + " | ".join(str(c.get(p, 0)) for p in PAT))
print("TOTAL | " + str(sum(total.values())) + " | " + " | ".join(str(agg.get(p, 0)) for p in PAT))
print("legacy tensor path (HMMA) present:", agg.get("HMMA", 0) > 0)
