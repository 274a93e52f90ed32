#!/usr/bin/env python3
"""cuobjdump -sass libvoitta_b200.so | python tools/sass_summary.py > profiles/rNN_sass_summary.txt
Counts, per kernel, the SASS mnemonics that prove which hardware path the code uses (B200_PROFILING.md):
UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit,
SYNCS = mbarrier operations."""
import collections
import re
import subprocess
import sys

KEYS = ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "LDG", "STG", "ATOMG", "ATOMS", "RED", "DFMA", "DADD", "DMUL",
        "FFMA", "BAR", "SHFL", "VOTE", "REDUX", "LDS", "STS")
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
cur, counts = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = pat.match(line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for key in KEYS:
            if op.startswith(key):
                counts[cur][key] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass voitta-rag_b200/libvoitta_b200.so — instruction mnemonics per kernel (sm_100a)")
print("# UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops")
for (k, c), name in zip(counts.items(), names):
    if "cub::" in name or "CUB_" in k:
        continue
    name = re.sub(r"\(.*", "", name)
    items = " ".join(f"{a}={b}" for a, b in c.items() if a != "_total" and b)
    print(f"{name[:64]:64s} total={c['_total']:6d}  {items}")
