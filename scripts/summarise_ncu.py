#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep + launches.csv into the small text summaries committed under profiles/.

    python scripts/summarise_ncu.py gpurun_out/prof_k_safe.ncu-rep gpurun_out/launches.csv profiles/r01
"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def main():
    rep, launches, out = sys.argv[1:4]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    with open(out + '_ncu_full_summary.txt', 'w') as f:
        f.write('# ncu --set full --clock-control none --import-source on (one capture per kernel; cold-cache, serialised)\n')
        for row in rows[2:]:
            f.write('\n== %s\n' % row[hdr.index('Kernel Name')][:160])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write('%-72s %s %s\n' % (w, row[i], units[i]))
            st = sorted(((float(row[hdr.index(h)]), h) for h in stall), reverse=True)[:8]
            for v, h in st:
                f.write('  stall/issue %-36s %.3f\n' % (h.replace('smsp__average_warps_issue_stalled_', '').replace(
                    '_per_issue_active.ratio', ''), v))
    d = collections.OrderedDict()
    with open(launches) as fh:
        rr = [r for r in csv.reader(l for l in fh if not l.startswith('=='))]
    h = rr[0]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    for r in rr[1:]:
        if len(r) > vi:
            d.setdefault(r[ki][:110], []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in d.values())
    with open(out + '_ncu_launches_summary.txt', 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none, same command as the bench line; per-launch '
                'times are cold-cache and serialised: compare SHARES\n')
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write('%-112s n=%4d total=%10.1f us mean=%9.1f us share=%.3f\n' % (k, len(v), sum(v) / 1e3,
                                                                              sum(v) / len(v) / 1e3, sum(v) / tot))


if __name__ == '__main__':
    main()
