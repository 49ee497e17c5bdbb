"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line: samples and instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
fname = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; ns = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= ns:
        continue
    if r[0].isdigit():
        cur = (fname, int(r[0]), r[1].strip()[:100])
        agg.setdefault(cur, [0, 0])
    if r[ns].isdigit() and r[2].startswith("0x"):
        agg[cur][0] += int(r[ns]); agg[cur][1] += int(r[ie])
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print("total samples", tot, "instructions", toti)
byfile = collections.Counter()
for (f, l, s), v in agg.items():
    byfile[f] += v[0]
print(dict(byfile))
for (f, l, s), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot:5.2f}% smp {100*v[1]/max(toti,1):5.2f}% ins  {f}:{l}  {s}")
