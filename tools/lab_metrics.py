#!/usr/bin/env python
"""tools/lab_metrics.py <ncu --csv log> — one row per profiled kernel instantiation (its last profiled launch)."""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 12]
hdr = rows[0]
kn, mn, mv, idc = (hdr.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID'))
per = OrderedDict()
for r in rows[1:]:
    per.setdefault(r[idc], {'name': r[kn]})[r[mn]] = r[mv]
short = OrderedDict([('gpu__time_duration.sum', 'ns'), ('smsp__inst_executed.sum', 'Minst'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'), ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%'),
    ('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'fmaH%'), ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu%'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsuwf%'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'longsb'), ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'math'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'notsel'), ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'wait'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'bar'), ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'shortsb'),
    ('smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'disp'), ('launch__registers_per_thread', 'regs'),
    ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB')])
last = OrderedDict()
for k, d in per.items():
    last[d['name']] = d
print('%-44s' % 'kernel', ' '.join('%7s' % v for v in short.values()))
for name, d in last.items():
    m = re.search(r'(\w+)<(.*)>', name)
    label = (m.group(1)[-18:] + '<' + m.group(2).replace(' ', '') + '>') if m else name[:44]
    out = []
    for mk, s in short.items():
        try:
            v = float(d.get(mk, 'nan').replace(',', ''))
            if s in ('Minst', 'rdMB', 'wrMB'):
                v /= 1e6
            out.append('%7.2f' % v if v < 1e5 else '%7.0f' % v)
        except Exception:
            out.append('%7s' % '-')
    print('%-44s' % label[:44], ' '.join(out))
