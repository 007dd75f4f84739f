"""Time of each chain of the model kernel on its own (run under ncu on the GPU box): one full encode, then encodes with
CBCG_K2M_ROLES = 1, 2, 4, 8 (POS, FLAG, counts, edits); the masked encodes leave stale intervals and are never decoded."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cbc_b200 import synth                      # noqa: E402
from cbc_b200.codec import Codec, pin_batch     # noqa: E402

name = os.environ.get("CBC_CONFIG", "config2")
cfg = synth.SynthConfig.named(name, scale=float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
L_HDR = {"config1": 100, "config5": 250}.get(name, 150)
g = synth.make_genome(cfg)
b = synth.make_reads(cfg, g)
c = Codec(0)
c.set_reference(g)
c.upload(pin_batch(b))
os.environ["CBCG_NO_OVERLAP"] = "1"
for mask in (15, 15, 1, 2, 4, 8, 15):
    os.environ["CBCG_K2M_ROLES"] = str(mask)
    try:
        c.encode_resident(L_HDR, 0xffffffff, 1)
        print(mask, c.stats()["ms_code"], flush=True)
    except Exception as e:      # stale intervals may not code
        print(mask, "error", e, flush=True)
