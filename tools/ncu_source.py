"""Per-instruction view of one kernel of an ncu report (needs -lineinfo / --import-source on).

  python tools/ncu_source.py gpurun_out/prof.ncu-rep [kernel-substring] [launch-index]

Prints address offset, executions, average active threads, stall samples and the dominant stall reason per SASS
instruction, plus totals per region between backward branches (loops)."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    name = sys.argv[2] if len(sys.argv) > 2 else None
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.check_output(['ncu', '-i', rep, '--page', 'source', '--csv'], stderr=subprocess.DEVNULL).decode()
    blocks, cur = [], None
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': [], 'hdr': None}
            blocks.append(cur)
        elif cur is not None and r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] and len(r) == len(cur['hdr']):
            cur['rows'].append(r)
    sel = [b for b in blocks if name is None or name in b['name']]
    b = sel[which]
    h = b['hdr']
    ia, isrc, ismp, iex, ith = (h.index(k) for k in ('Address', 'Source', '# Samples', 'Instructions Executed', 'Avg. Threads Executed'))
    stall_cols = [i for i, k in enumerate(h) if k.startswith('stall_') and 'Not Issued' not in k]
    base = int(b['rows'][0][ia], 16)
    tot_ex = sum(int(r[iex]) for r in b['rows'])
    tot_smp = sum(int(r[ismp]) for r in b['rows'])
    print("# %s: %d instructions, %d warp-instructions executed, %d samples" % (b['name'], len(b['rows']), tot_ex, tot_smp))
    for r in b['rows']:
        ex, smp = int(r[iex]), int(r[ismp])
        if ex == 0 and smp == 0:
            continue
        st = max(stall_cols, key=lambda i: int(r[i]))
        print("%04x %9d %3s %6d %-22s %s" % (int(r[ia], 16) - base, ex, r[ith], smp, h[st] if int(r[st]) else '', r[isrc].strip()))


if __name__ == '__main__':
    main()
