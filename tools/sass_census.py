"""SASS census of libfunasr_b200.so: per kernel, the counts of the instructions that prove a Blackwell-native path
(UTCHMMA = tcgen05.mma, UTMALDG = TMA bulk tensor load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier)
and of the legacy tensor instructions (HMMA = mma.sync).  Usage: python tools/sass_census.py > profiles/r2_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "funasr_b200", "libfunasr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
pat = {"UTCHMMA": r"\bUTCHMMA", "UTCHMMA.2CTA": r"UTCHMMA\.2CTA", "UTMALDG": r"\bUTMALDG", "UTMALDG.MULTICAST": r"UTMALDG[^;]*MULTICAST", "LDTM": r"\bLDTM",
       "STTM": r"\bSTTM", "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS", "HMMA": r"\bHMMA", "SHFL": r"\bSHFL", "MUFU.EX2": r"MUFU\.EX2"}
kern = None
counts = collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    counts[kern]["instructions"] += 1 if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln) else 0
    for k, p in pat.items():
        if re.search(p, ln):
            counts[kern][k] += 1
cols = ["instructions"] + list(pat)
print("# cuobjdump -sass census of funasr_b200/libfunasr_b200.so (sm_100a); one row per kernel with tensor / TMA / TMEM instructions")
print("%-78s " % "kernel" + " ".join("%9s" % c[:9] for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    if any(c[x] for x in pat if x not in ("SHFL", "MUFU.EX2", "SYNCS")):
        print("%-78s " % k[:78] + " ".join("%9d" % c[x] for x in cols))
print("%-78s " % ("TOTAL over %d kernels" % len(counts)) + " ".join("%9d" % tot[x] for x in cols))
