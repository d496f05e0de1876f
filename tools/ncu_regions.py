"""Stall samples of tc_igemm_kernel per warp role (regions delimited by marker instructions in the SASS):
    python tools/ncu_regions.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], rows[hi + 1:]
ci = {h: i for i, h in enumerate(hdr)}
samp, src = ci["# Samples"], ci["Source"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
def first(pat, start=0):
    for k in range(start, len(data)):
        if pat in data[k][src]:
            return k
    return len(data)
tma = first("UTMALDG"); mma = first("UTCHMMA"); ldtm = first("LDTM")
# splitter region = between last UTMALDG and first UTCHMMA
last_tma = max(k for k in range(len(data)) if "UTMALDG" in data[k][src])
regions = [("setup+producer", 0, last_tma + 40), ("splitter", last_tma + 40, mma - 60), ("mma issuer", mma - 60, ldtm - 150),
           ("epilogue", ldtm - 150, len(data))]
tot = sum(int(r[samp] or 0) for r in data)
print("total samples", tot, " markers: first TMA %d, last TMA %d, first MMA %d, first LDTM %d" % (tma, last_tma, mma, ldtm))
for name, a, b in regions:
    n = sum(int(r[samp] or 0) for r in data[a:b])
    st = {}
    for r in data[a:b]:
        for i in stall_cols:
            st[hdr[i][6:]] = st.get(hdr[i][6:], 0) + int(r[i] or 0)
    top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
    print("%-16s lines %5d-%5d  samples %6d (%4.1f%%)  %s" % (name, a, b, n, 100.0 * n / max(tot, 1),
                                                            " ".join("%s:%d" % kv for kv in top)))
