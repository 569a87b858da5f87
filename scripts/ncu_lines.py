"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[2] == "-":   # a source-line summary row
        try:
            key = (int(r[0]), r[1].strip()[:120])
            data[key] = (data.get(key, (0, 0))[0] + int(r[si]), data.get(key, (0, 0))[1] + int(r[ie]))
        except ValueError:
            pass
tot = sum(v[0] for v in data.values()) or 1
toti = sum(v[1] for v in data.values()) or 1
print("total samples", tot, "instructions", toti)
for (ln, txt), (s, i) in sorted(data.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% ins  L{ln:<5} {txt}")
