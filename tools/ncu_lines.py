"""Per-source-line view of an ncu SASS profile: joins `ncu --page source` (SASS rows, in order) with
`nvdisasm -g` line markers of the same cubin. usage: ncu_lines.py rep.ncu-rep cubin mangled_kernel_substr [top]"""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
lines = []; cur = None; inside = False
for l in dis:
    if l.startswith('.text.') and l.endswith(':'):
        inside = kern in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', l): lines.append(cur)
import os
kf = os.environ.get('NCU_KERNEL')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + (['--kernel-name', 'regex:' + kf, '--launch-count', '1'] if kf else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[h]; idx = {x: i for i, x in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) >= len(hdr) - 2]
print(f'sass rows {len(data)} vs disasm {len(lines)}')
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
stalls = [x for x in hdr if x.startswith('stall_') and 'Not Issued' not in x]
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else None
    a = agg[key]; a[0] += f(r, '# Samples'); a[1] += f(r, 'Instructions Executed')
    for s in stalls: a[2][s[6:]] += f(r, s)
tot = sum(a[0] for a in agg.values()) or 1; toti = sum(a[1] for a in agg.values()) or 1
srcs = {}
def text(key):
    if not key: return ''
    fn, ln = key
    if fn not in srcs:
        import glob
        c = glob.glob('/root/repo/**/' + fn, recursive=True)
        srcs[fn] = open(c[0]).read().split('\n') if c else []
    return srcs[fn][ln - 1].strip()[:90] if 0 < ln <= len(srcs[fn]) else ''
print(f'{"samples%":>8} {"inst%":>6}  line')
for key, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = ', '.join(f'{k} {100 * v / max(a[0], 1):.0f}%' for k, v in sorted(a[2].items(), key=lambda x: -x[1])[:3])
    print(f'{100 * a[0] / tot:8.1f} {100 * a[1] / toti:6.1f}  {key}  {text(key)}   [{st}]')
