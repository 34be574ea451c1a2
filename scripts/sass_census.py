"""Per-kernel SASS mnemonic census of the shipped library:  cuobjdump -sass <lib> | python scripts/sass_census.py"""
import collections
import re
import subprocess
import sys

keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDGMC", "REDG", "ATOMG", "BAR.SYNC",
        "ACQBULK", "LDG", "STG", "LDS", "STS", "FFMA", "HFMA2"]
cur, acc = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        acc[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        acc[cur]["_n"] += 1
        for k in keys:
            if m.group(1).startswith(k):
                acc[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(acc), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic census of protopformer_b200/lib/libprotohead_b200.so (cuobjdump -sass)")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = TMA load, SYNCS = mbarrier, "
      "LDGSTS = cp.async, LDGMC = multimem.ld_reduce (the multimem.st is an STG on the multicast address)")
for (f, c), d in zip(acc.items(), names):
    print(f"{d[:130]}\n    instructions {c['_n']:6d}  " + "  ".join(f"{k}={c[k]}" for k in keys if c[k]))
