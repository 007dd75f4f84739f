"""Writes a named configuration as SAM / FASTA text and runs the cbc command line on it, printing the program's own lines
(run under gpurun): where the wall time of `cbc -c` / `cbc -d` goes."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cbc_b200 import synth                      # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
extra = sys.argv[3:]
cfg = synth.SynthConfig.named(name, scale=scale)
g = synth.make_genome(cfg)
b = synth.make_reads(cfg, g)
cbc = os.path.join(ROOT, "cbc_b200", "_build", "cbc")
if os.environ.get("HOLD_CONTEXT"):
    import torch
    _keep = torch.zeros(1, device="cuda")          # a live context in another process: the driver stays initialised
with tempfile.TemporaryDirectory() as d:
    fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
    synth.write_fasta(fa, g)
    synth.write_sam(sam, b, g)
    print("sam bytes", os.path.getsize(sam), flush=True)
    for it in range(3):
        for cmd in ([cbc, "-c"] + extra + [sam, os.path.join(d, "o.cbc"), fa], [cbc, "-d"] + [x for x in extra if x.startswith("-g") or x[0].isdigit() and "," in x] + [os.path.join(d, "o.cbc"), os.path.join(d, "o.txt"), fa]):
            t0 = time.perf_counter()
            p = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, CBCH_TRACE="1", CBC_TRACE="1"))
            print(f"--- run {it} wall {time.perf_counter() - t0:.3f} s rc={p.returncode}: {' '.join(cmd[1:3])}")
            print(p.stdout.strip()); print(p.stderr.strip()[-600:], flush=True)
    with open(os.path.join(d, "o.txt"), "rb") as f:
        print("decoded ok:", f.read() == b.seq_lines())
