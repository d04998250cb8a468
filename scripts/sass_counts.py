"""Count the SASS mnemonics that prove which hardware paths libnaf_b200.so uses (per kernel and total):
tcgen05 MMA (UTCHMMA / UTCQMMA ...), TMEM load/store (LDTM / STTM), tcgen05 commit (UTCBAR), bulk copies
(UBLKCP), tensor-map TMA loads / stores (UTMALDG / UTMASTG), packed fp32 FMA (FFMA2).
    python scripts/sass_counts.py [path/to/lib.so] > profiles/rN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "naf_b200", "csrc", "libnaf_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKPF",
       "FFMA2", "SYNCS", "MUFU.EX2", "MUFU.TANH", "HMMA", "LDG.E.256", "STG.E.256"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for p in PAT:
        if re.search(r"\b" + re.escape(p), ln):
            per[cur][p] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  ({os.path.getsize(lib) / 1e6:.1f} MB, {len(per)} kernels)")
print("TOTAL " + "  ".join(f"{p}={tot[p]}" for p in PAT if tot[p]))
for k, c in per.items():
    if any(c[p] for p in PAT[:10]):
        print(f"{k}: " + "  ".join(f"{p}={c[p]}" for p in PAT if c[p]))
