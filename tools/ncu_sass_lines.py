"""Join an `ncu --page source --csv` (SASS view) export with `nvdisasm -g -c` line markers:
hottest CUDA source lines by stall samples and by executed warp instructions.
usage: ncu_sass_lines.py sass.csv nvdisasm.txt kernel_substring [N]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci, cs, src = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
inst = []
for r in rows[2:]:
    try:
        inst.append((int(r[ci] or 0), int(r[cs] or 0), r[src], {hdr[i][6:]: int(r[i] or 0) for i in stall_cols}))
    except (ValueError, IndexError):
        pass

# nvdisasm: instruction lines carry /*addr*/ ; line markers: //## File "...", line N
lines, cur, infunc = [], None, False
for l in open(sys.argv[2]):
    if l.startswith('\t.section') or l.startswith('//-----'):
        infunc = sys.argv[3] in l if '.text.' in l else infunc
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if infunc and re.match(r'\s+/\*[0-9a-f]{4}\*/', l):
        lines.append(cur)
print("ncu instructions", len(inst), "nvdisasm instructions", len(lines))
n = min(len(inst), len(lines))
agg = {}
for i in range(n):
    k = lines[i]
    a = agg.setdefault(k, [0, 0, {}])
    a[0] += inst[i][0]; a[1] += inst[i][1]
    for s, v in inst[i][3].items():
        a[2][s] = a[2].get(s, 0) + v
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
mix = {}
for a in agg.values():
    for s, v in a[2].items():
        mix[s] = mix.get(s, 0) + v
print("total warp-instructions", ti, "stall samples", ts)
print("stall mix:", ", ".join(f"{k} {100*v/max(ts,1):.1f}%" for k, v in sorted(mix.items(), key=lambda x: -x[1])[:8]))
srcs = {}
def text(k):
    if k is None: return "?"
    f, ln = k
    if f not in srcs:
        import glob
        c = glob.glob(f"/root/repo/komb_b200/csrc/{f}") + glob.glob(f"/usr/local/cuda/include/**/{f}", recursive=True)
        srcs[f] = open(c[0]).read().split('\n') if c else []
    return srcs[f][ln-1].strip()[:95] if ln-1 < len(srcs[f]) else ""
N = int(sys.argv[4]) if len(sys.argv) > 4 else 20
print("--- by stall samples")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:N]:
    top = ",".join(f"{s}:{100*v//max(a[1],1)}" for s, v in sorted(a[2].items(), key=lambda x: -x[1])[:2])
    print(f"{100*a[1]/max(ts,1):5.1f}% smp {100*a[0]/ti:5.1f}% inst [{top:28s}] {k[0] if k else '?'}:{k[1] if k else 0}: {text(k)}")
print("--- by instructions")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:N]:
    print(f"{100*a[0]/ti:5.1f}% inst {100*a[1]/max(ts,1):5.1f}% smp  {k[0] if k else '?'}:{k[1] if k else 0}: {text(k)}")
