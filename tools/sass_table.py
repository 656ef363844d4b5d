"""`cuobjdump -sass` mnemonic counts per kernel of the built library (evidence for which hardware
paths each kernel uses: DMMA = FP64 tensor pipe, UTMALDG / UBLKCP = TMA tensor / bulk copies,
UTCIMMA / LDTM = tcgen05 INT8 MMA and TMEM loads, SYNCS = mbarrier, LDGSTS = cp.async).

    python tools/sass_table.py > profiles/r02_sass_counts.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "tnrkit.jl_b200", "lib", "libtnrcuda.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
PAT = ["DMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTCIMMA", "LDTM", "SYNCS", "LDGSTS", "LDG.E.128",
       "STG.E.128", "LDS.128", "LDG.E.64", "STG.E.64", "BAR.SYNC", "REDUX", "ATOMG"]
counts = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"tnr::\(anonymous namespace\)::|tnr::_GLOBAL__N__\w+::|tnr::", "", name)
        name = re.sub(r"\((?!anonymous).*$", "", name)
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    for p in PAT:
        if re.search(r"\b" + re.escape(p) + r"\b", ln) or (p.endswith("128") and p in ln) or \
                (p.endswith("64") and p in ln):
            cur[p] += 1
print("# SASS mnemonic counts per kernel of libtnrcuda.so (sm_100a), `tools/sass_table.py`\n")
print("| kernel | " + " | ".join(PAT) + " |")
print("|---|" + "---:|" * len(PAT))
for name, c in counts.items():
    if not any(c.values()):
        continue
    print(f"| `{name}` | " + " | ".join(str(c[p]) if c[p] else "" for p in PAT) + " |")
