"""Summarise an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass) per code
segment between BAR.SYNCs: share of warp-stall samples, FFMA instructions, top stall reasons."""
import collections, csv, sys
path = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
r = csv.reader(open(path))
rows, hdr = [], None
for row in r:
    if row and row[0] == "Address":
        if hdr is not None and rows:
            break               # first kernel only
        hdr = row
        continue
    if hdr is not None and len(row) == len(hdr):
        rows.append(row)
ix = {h: i for i, h in enumerate(hdr)}
def I(x, k):
    v = x[ix[k]]
    try: return int(v)
    except ValueError: return 0
tot = sum(I(x, "# Samples") for x in rows)
print("instructions", len(rows), "total samples", tot)
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seg, cur = [], dict(n=0, samples=0, ffma=0, inst=0, start=0, stall=collections.Counter())
for i, x in enumerate(rows):
    cur["n"] += 1; cur["samples"] += I(x, "# Samples"); cur["inst"] += I(x, "Instructions Executed")
    if "FFMA" in x[ix["Source"]]: cur["ffma"] += I(x, "Instructions Executed")
    for c in stallcols: cur["stall"][c] += I(x, c)
    if "BAR.SYNC" in x[ix["Source"]]:
        seg.append(cur); cur = dict(n=0, samples=0, ffma=0, inst=0, start=i + 1, stall=collections.Counter())
seg.append(cur)
allst = collections.Counter()
for s in seg: allst.update(s["stall"])
print("overall:", ", ".join(f"{k[6:]}:{100*v/tot:.1f}%" for k, v in allst.most_common(10)))
for s in seg:
    if s["samples"] > tot * thr:
        top = ", ".join(f"{k[6:]}:{100*v/max(1,s['samples']):.0f}%" for k, v in s["stall"].most_common(5))
        print(f"sass[{s['start']:6d}+{s['n']:5d}] samples {100*s['samples']/tot:5.1f}%  ffma/inst {s['ffma']/max(1,s['inst']):.2f} inst {s['inst']:>11d} | {top}")
