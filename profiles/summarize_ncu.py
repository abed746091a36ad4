#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the small markdown summary kept under profiles/.
usage: python profiles/summarize_ncu.py REPORT.ncu-rep OUT.md "title" [kernel-substring]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def main():
    rep, out, title = sys.argv[1:4]
    raw = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv']).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index('Kernel Name')
    with open(out, 'w') as f:
        f.write('# %s\n\nSource report: `%s` (ncu --set full --clock-control none; per-launch values, cold cache, serialised)\n\n' % (title, rep))
        f.write('Kernels: %s\n\n' % ', '.join(sorted(set(r[kn][:80] for r in data))))
        f.write('| metric | unit | ' + ' | '.join('launch %d' % (i + 1) for i in range(len(data))) + ' |\n|---|---|' + '---|' * len(data) + '\n')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write('| %s | %s | %s |\n' % (k, units[i], ' | '.join(r[i] for r in data)))
        f.write('\nWarp stall reasons (cycles per issued instruction, launch 1):\n\n')
        st = []
        for h in hdr:
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                st.append((float(data[0][hdr.index(h)]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        for v, n in sorted(st, reverse=True)[:8]:
            f.write('- %s: %.2f\n' % (n, v))
        # hottest source lines
        try:
            src = subprocess.check_output(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass',
                                           '--launch-skip', '0', '--launch-count', '1'], stderr=subprocess.DEVNULL).decode()
            agg = collections.OrderedDict()
            cur, h2, func = None, None, ''
            for r in csv.reader(io.StringIO(src)):
                if not r:
                    continue
                if r[0] == 'File Path':
                    cur = r[1].split('/')[-1]
                elif r[0] == 'Function Name':
                    func = r[1]
                elif r[0] == 'Line No':
                    h2 = r
                elif h2 and r[0]:
                    try:
                        ins = int(r[h2.index('Instructions Executed')])
                        smp = int(r[h2.index('# Samples')]) if r[h2.index('# Samples')] not in ('-', '') else 0
                    except Exception:
                        continue
                    key = (cur, int(r[0]))
                    a = agg.get(key, [0, 0, r[1][:100]])
                    a[0] += ins
                    a[1] += smp
                    agg[key] = a
            ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
            f.write('\nHottest source lines (launch 1; share of warp instructions / of stall samples):\n\n| file:line | instr % | samples % | source |\n|---|---|---|---|\n')
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
                f.write('| %s:%d | %.1f | %.1f | `%s` |\n' % (k[0], k[1], 100 * v[0] / max(ti, 1), 100 * v[1] / max(ts, 1), v[2].replace('|', '\\|')))
            f.write('\nLines with the most stall samples (launch 1):\n\n| file:line | samples % | instr % | source |\n|---|---|---|---|\n')
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
                f.write('| %s:%d | %.1f | %.1f | `%s` |\n' % (k[0], k[1], 100 * v[1] / max(ts, 1), 100 * v[0] / max(ti, 1), v[2].replace('|', '\\|')))
        except Exception as e:  # pragma: no cover
            f.write('\n(source page unavailable: %s)\n' % e)
    print(open(out).read())


if __name__ == '__main__':
    main()
