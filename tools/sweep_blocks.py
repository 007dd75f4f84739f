"""Block-size sweep on one GPU: stage times and container size per block_reads (run under gpurun)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cbc_b200 import synth                      # noqa: E402
from cbc_b200.codec import Codec, pin_batch     # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
sizes = [int(x) if x != "auto" else 0xffffffff for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [256, 512, 1024, 2048, 4096, 16384, 65536]
gen_mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = synth.SynthConfig.named(os.environ.get("CBC_CONFIG", "config2"), scale=scale)
L_HDR = {"config1": 100, "config5": 250}.get(os.environ.get("CBC_CONFIG", "config2"), 150)
g = synth.make_genome(cfg)
b = synth.make_reads(cfg, g)
c = Codec(0)
c.set_reference(g)
c.upload(pin_batch(b))
ref = b.seq_lines()
for R in sizes:
    rows = []
    for it in range(4):
        c.encode_resident(L_HDR, R, gen_mode)
        se = c.stats()
        c.decode_resident()
        sd = c.stats()
        rows.append((se["ms_k1"], se["ms_extract"], se["ms_plan"], se["ms_code"], se["ms_gather"], se["ms_total"],
                     sd["ms_code"], sd["ms_k3"], sd["ms_reconstruct"], sd["ms_total"]))
    ok = c.fetch_decoded().tobytes() == ref
    m = np.median(np.array(rows[1:]), axis=0)
    print(json.dumps({"gen_mode": gen_mode, "block_reads": R, "blocks": se["n_blocks"], "ok": ok, "container_bytes": se["container_bytes"],
                      "bits_per_base": 8.0 * se["container_bytes"] / b.total_bases(), "n_symbols": se["n_symbols"],
                      "n_edits": se["n_edits"],
                      "ms": dict(zip(["k1", "extract", "plan", "k2e", "gather", "enc_total", "k2d", "k3", "recon", "dec_total"],
                                     [round(float(x), 4) for x in m]))}), flush=True)
