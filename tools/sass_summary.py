"""Development tool: per-kernel counts of the Blackwell tensor-core / TMA / TMEM / mbarrier instructions in the built library
(cuobjdump -sass), written to profiles/<tag>_conv_sass.txt.   usage: python tools/sass_summary.py r2"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
so = os.path.join(ROOT, "lgd_b200", "liblgd_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEEP = re.compile(r"^(UTC|UTMA|LDTM|STTM|SYNCS|UBLKCP|UTMAPF|UTMACCTL)")
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = per.setdefault(demangle(m.group(1)), collections.Counter())
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and cur is not None and KEEP.match(m.group(1)):
        cur[m.group(1)] += 1
out = ["# %s: cuobjdump -sass lgd_b200/liblgd_b200.so -- Blackwell tensor-core / TMA / TMEM instruction counts per kernel (sm_100a)" % tag,
       "# UTCHMMA = tcgen05.mma (kind::f16), UTCMMA/UTCQMMA... = other kinds, UTMALDG = TMA tensor load (IM2COL = im2col mode),",
       "# LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, UTMAPF = descriptor prefetch",
       ""]
tot = collections.Counter()
for name, cnt in per.items():
    if not any(k.startswith(("UTC", "UTMA", "LDTM")) for k in cnt):
        continue
    out.append("== " + name)
    for k in sorted(cnt):
        out.append("   %-40s %d" % (k, cnt[k]))
        tot[k] += cnt[k]
out += ["", "== totals over the library"] + ["   %-40s %d" % (k, tot[k]) for k in sorted(tot)]
path = os.path.join(ROOT, "profiles", "%s_conv_sass.txt" % tag)
open(path, "w").write("\n".join(out) + "\n")
print(path, len(per), "kernels")
