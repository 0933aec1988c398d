"""profiles/sass_digest.txt: per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce), HMMA (legacy mma.sync).

    python scripts/make_sass_digest.py > profiles/sass_digest.txt        (CPU only: cuobjdump on jimm_b200/libjimm_b200.so)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "jimm_b200", "libjimm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UTCxMMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
                                ("UTMAREDG", r"\bUTMAREDG"), ("HMMA", r"\bHMMA"), ("MUFU", r"\bMUFU"), ("SYNCS", r"\bSYNCS")])
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k, p in pats.items():
        if re.search(p, line):
            counts[cur][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS digest of {os.path.relpath(lib, ROOT)} ({os.path.getsize(lib)} bytes); columns: " + " ".join(pats))
tot = collections.Counter()
rows = []
for (mang, c), name in zip(counts.items(), dem):
    name = re.sub(r"\(CUtensorMap_st.*", "", name).replace("void jimm::", "").replace("(anonymous namespace)::", "")
    rows.append((name, c))
    tot.update(c)
for name, c in sorted(rows):
    print(f"{name[:110]:110s} " + " ".join(f"{c[k]:6d}" for k in pats))
print(f"{'TOTAL':110s} " + " ".join(f"{tot[k]:6d}" for k in pats))
hm = [n for n, c in rows if c["HMMA"] and not c["UTCxMMA"]]
print("\n# kernels that still use the legacy mma.sync tensor path (HMMA without UTC*MMA): " + (", ".join(sorted(set(n.split('<')[0] for n in hm))) or "none"))
