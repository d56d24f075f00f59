"""Blackwell evidence from the shipped library, no GPU needed: per kernel, how many tcgen05 / TMEM / TMA / bulk-copy / packed-fp32x2
instructions its SASS holds (mnemonics of B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG /
UTMASTG = cp.async.bulk.tensor load / store, UBLKCP = cp.async.bulk 1-D, UTCBAR = tcgen05.commit, SYNCS = mbarrier).
  python tools/sass_evidence.py [lib.so] > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "anyv2v_b200", "lib", "libanyv2v_b200.so")
KEYS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "UTMAPF", "HMMA")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
cur, cnt, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        if op.startswith(KEYS):
            cnt[cur][re.sub(r"\.$", "", op)] += 1
print(f"# {os.path.relpath(lib, ROOT)}: cuobjdump -sass, instruction counts per kernel (static)")
for f in sorted(cnt, key=lambda f: subprocess.run(["c++filt", f], capture_output=True, text=True).stdout):
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((CUtensorMap_st|av2v::|__half|const __half).*$", "", name)
    fam = collections.Counter()
    for k, v in cnt[f].items():
        fam[next(key for key in KEYS if k.startswith(key))] += v
    print(f"{name}   [{total[f]} instructions]")
    print("    " + "  ".join(f"{k}={v}" for k, v in sorted(fam.items())))
    detail = {k: v for k, v in cnt[f].items() if k.startswith(("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UBLKCP"))}
    if detail:
        print("    " + "  ".join(f"{k}={v}" for k, v in sorted(detail.items())))
