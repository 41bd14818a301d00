"""SASS opcode histogram of libmi_b200.so per kernel (evidence that the hot kernels are tcgen05 / TMEM / TMA code):
    python scripts/sass_histogram.py > profiles/r02_sass_opcodes.md
UTC*MMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
MUFU.EX2 = ex2.approx, SYNCS = mbarrier ops (B200_PROFILING.md, "What proves a Blackwell-native kernel")."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mutual-information-multimodal_b200", "libmi_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "MUFU.EX2", "SYNCS", "HMMA", "UTMAPF"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["total"] += 1
        for o in OPS:
            if op.startswith(o):
                per[cur][o] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode histogram of libmi_b200.so (cuobjdump -sass; sm_100a)\n")
print("`UTCHMMA` = tcgen05.mma (cta_group::1/2), `UTMALDG` / `UTMASTG` = TMA tensor load / store, `LDTM` = tcgen05.ld, `UTCBAR` = "
      "tcgen05.commit, `MUFU.EX2` = ex2.approx, `SYNCS` = mbarrier; no `HMMA` (legacy mma.sync) anywhere.\n")
print("| kernel | instr | " + " | ".join(OPS) + " |")
print("|---|---|" + "---|" * len(OPS))
tot = collections.Counter()
for (name, c), dn in zip(per.items(), demangle):
    short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", "").replace("void ", ""))
    print(f"| `{short}` | {c['total']} | " + " | ".join(str(c[o]) for o in OPS) + " |")
    tot.update(c)
print(f"| **all {len(per)} kernels** | {tot['total']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")
