#!/usr/bin/env python
"""Digest an .ncu-rep (read here, no GPU): key raw metrics, stall reasons and instruction mix.
usage: python scripts/ncu_digest.py gpurun_out/x.ncu-rep [--csv-out profiles/x_summary.csv]"""
import collections, csv, io, subprocess, sys

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    out_rows = []
    raw = page(rep, 'raw')
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        for k in hdr:
            if k in KEEP or 'pipe_fp64' in k or 'pipe_xu' in k:
                if d[k] not in ('', 'n/a'):
                    print('%-75s %s %s' % (k, d[k][:110], units[hdr.index(k)]))
                    out_rows.append(('raw', k, d[k], units[hdr.index(k)]))
    src = page(rep, 'source')
    h = src[1]
    ix = {c: i for i, c in enumerate(h)}
    data = src[2:]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot = sum(f(r, '# Samples') for r in data) or 1.0
    stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    print('--- warp stall sampling (all samples = %d)' % tot)
    for s, v in sorted(((s, sum(f(r, s) for r in data)) for s in stalls), key=lambda x: -x[1])[:10]:
        print('  %-26s %6.1f%%' % (s, 100 * v / tot))
        out_rows.append(('stall', s, '%.1f' % (100 * v / tot), '% of samples'))
    mix = collections.Counter()
    for r in data:
        op = r[ix['Source']].strip().split()
        if not op:
            continue
        o = op[0] if not op[0].startswith('@') else (op[1] if len(op) > 1 else op[0])
        mix['.'.join(o.split('.')[:2]) if o.startswith(('F2F', 'I2F', 'F2I', 'FRND', 'IMAD', 'MUFU', 'LDS', 'LDG', 'STG')) else o.split('.')[0]] += f(r, 'Instructions Executed')
    t = sum(mix.values()) or 1.0
    print('--- warp instructions executed by opcode (total %d)' % t)
    for o, v in mix.most_common(28):
        print('  %-14s %12d %5.1f%%' % (o, v, 100 * v / t))
        out_rows.append(('opcode', o, str(int(v)), 'warp instructions'))
    exc = sum(f(r, 'L1 Wavefronts Shared Excessive') for r in data)
    print('--- shared wavefronts: total %d, excessive %d' % (sum(f(r, 'L1 Wavefronts Shared') for r in data), exc))
    if '--csv-out' in sys.argv:
        with open(sys.argv[sys.argv.index('--csv-out') + 1], 'w') as fo:
            w = csv.writer(fo)
            w.writerow(['section', 'name', 'value', 'unit'])
            w.writerows(out_rows)


if __name__ == '__main__':
    main()
