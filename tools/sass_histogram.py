"""Instruction histogram of every kernel in libctx_b200.so (cuobjdump -sass), the tcgen05 / TMA mnemonics first:
    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "imitation_from_observation_b200", "libctx_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "RED", "ATOM",
       "LDG", "STG", "LDS", "STS", "FFMA", "HFMA2", "F2FP", "BAR"]
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("ctx::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
        if m.group(1) in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UBLKCP"):
            hist[kern][m.group(1) + m.group(2)] += 1
print("# cuobjdump -sass %s: instructions per kernel (static counts)" % os.path.relpath(so, ROOT))
print("# UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy, LDTM / STTM = tcgen05.ld / st,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops")
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    tot = sum(v for n, v in h.items() if "." not in n)
    tc = [n for n in h if n.split(".")[0] in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UBLKCP") and "." in n]
    print("%-58s %6d instr | %s" % (k[:58], tot, " ".join("%s=%d" % (n, h[n]) for n in KEY if h[n])))
    if tc:
        print("%-58s        | %s" % ("", " ".join("%s=%d" % (n, h[n]) for n in sorted(tc))))
