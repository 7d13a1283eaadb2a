"""Per-kernel census of the tensor-core / TMEM / TMA SASS instructions in libb200ode.so (cuobjdump -sass):
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load), UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops.  usage: python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "differential_equations_resnet_b200", "libb200ode.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA"]
cur, counts, arch = None, collections.OrderedDict(), set()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    for p in pats:
        if re.search(r"\b%s\b|\b%s\." % (p, p), line):
            counts[cur][p] += 1
print("libb200ode.so  arch %s  (%d kernels)" % (sorted(arch), len(counts)))
print("%-78s %s" % ("kernel", " ".join("%8s" % p for p in pats)))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    if any(c[p] for p in pats[:10]):
        print("%-78s %s" % (k[:78], " ".join("%8d" % c[p] for p in pats)))
print("%-78s %s" % ("TOTAL (all kernels)", " ".join("%8d" % tot[p] for p in pats)))
