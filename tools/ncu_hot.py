"""Top stall sites of an .ncu-rep's source page (SASS view):  python tools/ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; data = rows[hdr_i + 1:]
ci = {h: i for i, h in enumerate(hdr)}
samp = ci["# Samples"]; src = ci["Source"]; ex = ci["Instructions Executed"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[samp] or 0) for r in data)
print("total samples", tot)
ranked = sorted(range(len(data)), key=lambda k: -int(data[k][samp] or 0))[:top]
for k in sorted(ranked):
    r = data[k]
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print("%5d %5.1f%% ex=%-9s %-70s %s" % (k, 100.0 * int(r[samp] or 0) / max(tot, 1), r[ex], r[src][:70],
                                          " ".join("%s:%d" % (n, c) for c, n in st if c)))
