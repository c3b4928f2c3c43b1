"""Key metrics + top stall reasons of every kernel in an ncu report:  python scripts/ncu_brief.py REPORT.ncu-rep"""
import csv, io, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarise_ncu import WANT

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    print('==', row[hdr.index('Kernel Name')][:140])
    for w in WANT + ['launch__occupancy_limit_warps', 'sm__maximum_warps_per_active_cycle_pct', 'launch__waves_per_multiprocessor',
                     'smsp__inst_executed_pipe_fp64.sum', 'l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
                     'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_bytes.sum', 'smsp__cycles_active.avg']:
        if w in hdr:
            i = hdr.index(w)
            print('%-72s %s %s' % (w, row[i], units[i]))
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    for v, h in sorted(((float(row[hdr.index(h)]), h) for h in stall), reverse=True)[:9]:
        print('  stall/issue %-32s %.3f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
