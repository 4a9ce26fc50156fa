"""Turn ncu exports into the small text summaries committed under profiles/.

  python tools/profile_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
  python tools/profile_summary.py kernel   gpurun_out/prof.ncu-rep [name]     > profiles/rNN_<kernel>.txt
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'sm__cycles_active.min', 'sm__cycles_active.max', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[0] == 'ID':
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        v = v / 1e3 if d['Metric Unit'] == 'ns' else (v * 1e3 if d['Metric Unit'] == 'ms' else v)
        k = d['Kernel Name'].split('(')[0][-60:]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(t for _, t in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)")
    print("%-62s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-62s %8d %12.1f %10.2f %6.1f%%" % (k, n, t, t / n, 100 * t / tot))


def kernel(rep, name=None):
    raw = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv'], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    sel = [r for r in rows[2:] if name is None or name in r[ki]]
    print("# ncu --set full --clock-control none; %d launches of %s" % (len(sel), sel[0][ki][:90]))
    out = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals = [r[i] for r in sel]
            print("%-70s %-12s %s" % (k, units[i], ' '.join(vals)))
            out[k] = (units[i], vals)
    src = subprocess.check_output(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'] +
                                  (['-k', 'regex:' + name] if name else []), stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name':
            break
        if len(r) >= len(hdr) and r[0] != 'Address':
            data.append(r)
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
    tot = float(sum(agg.values())) or 1.0
    print("# warp stall sampling (first launch), share of samples")
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        print("  %-28s %5.1f%%" % (k, 100 * v / tot))
    print("# hottest SASS instructions (samples, executions, avg active threads)")
    for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:12]:
        print("  %6s %10s %5s  %s" % (r[idx['# Samples']], r[idx['Instructions Executed']], r[idx['Avg. Threads Executed']], r[1].strip()[:70]))
    return out


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
