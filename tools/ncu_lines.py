"""Summarise an `ncu --page source --csv --print-source cuda` export: hottest CUDA lines by
executed instructions and by stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci, cs, src = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    try:
        st = {hdr[i]: int(r[i] or 0) for i in stall_cols}
        data.append((int(r[ci] or 0), int(r[cs] or 0), r[0], r[src], st))
    except (ValueError, IndexError):
        pass
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print("total warp-instructions", ti, "stall samples", ts)
agg = {}
for d in data:
    for k, v in d[4].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join(f"{k[6:]} {100*v/max(ts,1):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 22
print("--- by stall samples")
for d in sorted(data, key=lambda x: -x[1])[:n]:
    top = max(d[4].items(), key=lambda x: x[1])[0][6:] if d[1] else ""
    print(f"{100*d[1]/max(ts,1):5.1f}% smp {100*d[0]/ti:5.1f}% inst [{top:9s}] L{d[2]}: {d[3].strip()[:105]}")
print("--- by instructions")
for d in sorted(data, key=lambda x: -x[0])[:n]:
    print(f"{100*d[0]/ti:5.1f}% inst {100*d[1]/max(ts,1):5.1f}% smp  L{d[2]}: {d[3].strip()[:110]}")
