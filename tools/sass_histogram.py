"""SASS evidence for profiles/: per kernel of libcbcg.so the instruction count and the opcodes that prove what the design
claims (bulk-copy TMA: UBLKCP; mbarrier: SYNCS; warp reductions: REDUX; FP64 reciprocal of the coder: MUFU.RCP64H / DFMA; no
tensor-core or TMEM instruction anywhere: nothing on this path is a contraction). Runs here: cuobjdump needs no GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cbc_b200", "_build", "libcbcg.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kern, cur = collections.OrderedDict(), None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur:
        kern[cur][m.group(1)] += 1
WATCH = ["UBLKCP", "SYNCS", "REDUX", "SHFL", "VOTE", "MUFU.RCP64H", "DFMA", "DMUL", "I2F.F64", "F2I", "ATOM", "RED", "LDS", "STS", "LDG", "STG", "BSSY", "BRX", "CALL", "NANOSLEEP"]
TENSOR = ("HMMA", "IMMA", "DMMA", "UTCMMA", "UTCHMMA", "TCGEN", "LDTM", "STTM", "UTCBAR")
print(f"# SASS of {os.path.relpath(so, ROOT)} ({', '.join(arch)})\n")
print("| kernel | instructions | " + " | ".join(WATCH) + " | tensor / TMEM |")
print("|---|---|" + "---|" * (len(WATCH) + 1))
for k, c in kern.items():
    tot = sum(c.values())
    def cnt(p):
        return sum(v for o, v in c.items() if o == p or o.startswith(p + "."))
    tens = sum(v for o, v in c.items() if o.startswith(TENSOR))
    print(f"| `{k[:60]}` | {tot} | " + " | ".join(str(cnt(w)) for w in WATCH) + f" | {tens} |")
