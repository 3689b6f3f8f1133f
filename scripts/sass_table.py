"""SASS evidence table: per kernel of libb2llava.so, the count of the Blackwell-specific mnemonics (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA tensor), UBLKCP (TMA bulk), UTCBAR (tcgen05.commit),
HMMA (mma.sync), plus registers from the cubin. Runs in the build container (cuobjdump, no GPU):
    python scripts/sass_table.py > profiles/r2_sass_instruction_table.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "llava-plus-codebase_b200", "lib", "libb2llava.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "SYNCS", "ACQBULK", "UTCCP"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["total"] += 1
        for p in PAT:
            if op.startswith(p):
                counts[cur][p] += 1
print("# cuobjdump -sass llava-plus-codebase_b200/lib/libb2llava.so (sm_100a): occurrences of Blackwell-specific mnemonics per kernel")
print("# UTCHMMA/UTCQMMA = tcgen05.mma (bf16 / e4m3), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store,")
print("# UBLKCP = cp.async.bulk (1-D TMA), UTCBAR = tcgen05.commit, HMMA = mma.sync, SYNCS = mbarrier ops")
print(f"{'kernel':72s} {'instrs':>7s} " + " ".join(f"{p:>8s}" for p in PAT))
for k, c in counts.items():
    name = demangle(k).replace("(anonymous namespace)::", "").replace("void ", "").replace("b2::", "")
    name = re.sub(r"\(.*", "", name)
    if not any(c[p] for p in PAT) and c["total"] < 400:
        continue
    print(f"{name[:72]:72s} {c['total']:7d} " + " ".join(f"{c[p]:8d}" for p in PAT))
