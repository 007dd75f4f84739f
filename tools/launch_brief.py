"""Per-launch lines of an `ncu --csv --metrics ...` log: kernel, grid and the collected metrics (reading gpurun_out/ quickly)."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
d = defaultdict(dict)
for r in rows:
    if not r[0].isdigit(): continue
    i = int(r[0]); d[i]['k'] = r[4].split('(')[0][-34:]; d[i]['grid'] = r[8]; d[i][r[-3]] = r[-1]
for i in sorted(d):
    x = d[i]
    print(i, x['k'], x['grid'], ' '.join(f"{k.split('.')[0].replace('smsp__','').replace('gpu__','')}={v}" for k, v in x.items() if k not in ('k', 'grid')))
