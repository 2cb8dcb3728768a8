"""Summaries of ncu exports for profiles/: launch-list shares and key raw metrics per kernel.
usage: ncu_summary.py launches.csv | ncu_summary.py --raw kernel_raw.csv"""
import collections, csv, re, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    tot, cnt, seq = collections.OrderedDict(), collections.Counter(), []
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('kg::<unnamed>::', '').replace('void ', '')
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
        tot[name] = tot.get(name, 0) + v; cnt[name] += 1; seq.append((name, v))
    # the timed step = launches after the last L2-flush fill kernel
    idx = max((i for i, (n, _) in enumerate(seq) if 'FillFunctor' in n), default=-1)
    step = seq[idx + 1:]
    t2, c2 = collections.OrderedDict(), collections.Counter()
    for n, v in step:
        t2[n] = t2.get(n, 0) + v; c2[n] += 1
    T = sum(t2.values())
    print(f"timed step: {len(step)} launches, {T:.1f} us summed kernel time (ncu: cold cache, serialised; compare SHARES)")
    for k, v in sorted(t2.items(), key=lambda x: -x[1])[:18]:
        print(f"{v:9.1f} us {100 * v / T:5.1f}%  x{c2[k]:3d}  {k[:100]}")


def raw(path):
    r = list(csv.reader(open(path)))
    hdr, units, vals = r[0], r[1], r[-1]
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else path
    print(re.sub(r'\(.*', '', name))
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"  {h:70s} {vals[i]:>16s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "--raw":
        for p in sys.argv[2:]:
            raw(p)
    else:
        launches(sys.argv[1])
