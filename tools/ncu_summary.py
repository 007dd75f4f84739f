"""Summarise an .ncu-rep: headline metrics per kernel + stall mix + hottest source lines (needs -lineinfo)."""
import csv
import io
import subprocess
import sys
from collections import Counter, defaultdict

rep = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('##', r[idx['Kernel Name']].split('(')[0])
    for w in WANT:
        if w in idx:
            print(f'  {w}: {r[idx[w]]} {units[idx[w]]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for i, r in enumerate(rows):
    if r and r[0] == 'Address':
        h = i
        break
if h is None:
    sys.exit(0)
hdr = rows[h]
idx = {x: i for i, x in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) >= len(hdr) - 2]
def f(r, k):
    try:
        return float(r[idx[k]])
    except Exception:
        return 0.0
tot = sum(f(r, '# Samples') for r in data) or 1
toti = sum(f(r, 'Instructions Executed') for r in data) or 1
print(f'\nSASS lines {len(data)}, samples {tot:.0f}, warp instructions {toti:.0f}')
stalls = [x for x in hdr if x.startswith('stall_') and 'Not Issued' not in x]
agg = {x: sum(f(r, x) for r in data) for x in stalls}
print('stall mix: ' + ', '.join(f'{k[6:]} {100 * v / tot:.1f}%' for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
ops = Counter()
for r in data:
    t = r[idx['Source']].strip().split()
    if not t:
        continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    ops[op.split('.')[0]] += f(r, 'Instructions Executed')
print('opcode mix: ' + ', '.join(f'{k} {100 * v / toti:.1f}%' for k, v in ops.most_common(14)))
