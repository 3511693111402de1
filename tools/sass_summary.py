"""profiles/sass_summary.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell-native path
(B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier; HMMA / HGMMA would be the legacy tensor paths).  Runs here (no GPU): cuobjdump -sass on the built library."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200", "rvc_b200", "librvcb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT", "HMMA", "HGMMA", "LDGSTS", "REDUX", "ATOMS", "MEMBAR"]
per = collections.OrderedDict()
cur = None
arch = set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("rvcb::", "")
        per.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if cur:
        for mn in MN:
            if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line):
                per[cur][mn] += 1
with open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w") as f:
    f.write(f"# cuobjdump -sass librvcb200.so  (arch: {', '.join(sorted(arch))}); instruction counts per kernel, tools/sass_summary.py\n")
    f.write(f"# {'kernel':70s} " + " ".join(f"{m:>9s}" for m in MN) + "\n")
    tot = collections.Counter()
    for k, c in per.items():
        if not any(c.values()):
            continue
        tot.update(c)
        f.write(f"{k[:72]:72s} " + " ".join(f"{c[m]:9d}" for m in MN) + "\n")
    f.write(f"{'TOTAL':72s} " + " ".join(f"{tot[m]:9d}" for m in MN) + "\n")
print(open(os.path.join(ROOT, "profiles", "sass_summary.txt")).read()[:3000])
