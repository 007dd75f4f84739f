import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from cbc_b200 import synth
from cbc_b200.codec import Codec, CbcgError
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
cfg = synth.SynthConfig.named("config2", scale=scale); g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
c = Codec(0); c.set_reference(g)
orecs, oedits = O.extract(b, g)
for R, G in ((512, 0), (512, 1)):
    ocont = O.encode_blocked(b, g, 150, R, G)
    for it in range(3):
        cont = c.compress(b, 150, R, G)
        same = cont == ocont
        try:
            recs, chr_, edits = c.decode_edits(cont)
        except CbcgError as e:
            print(f"R={R} G={G} it={it} enc_ok={same} decode error: {e}", flush=True); continue
        okr = np.array_equal(recs, orecs); oke = np.array_equal(edits, oedits)
        msg = ""
        if not okr:
            bad = np.nonzero(recs != orecs)[0]; msg = f"first bad rec {bad[0]} (block {bad[0] // R}, ord {bad[0] % R}) n_bad {len(bad)}: {recs[bad[0]]} vs {orecs[bad[0]]}"
        elif not oke:
            bad = np.nonzero(edits != oedits)[0]; msg = f"edit diffs {len(bad)} first {bad[:4]}"
        print(f"R={R} G={G} it={it} enc_ok={same} recs={okr} edits={oke} {msg}", flush=True)
ref = b.seq_lines()
for R, G in ((512, 0), (512, 1)):
    for it in range(3):
        c.upload(b); c.encode_resident(150, R, G)
        try:
            c.decode_resident()
            ok = c.fetch_decoded().tobytes() == ref
            print(f"resident R={R} G={G} it={it} text_ok={ok}", flush=True)
        except CbcgError as e:
            print(f"resident R={R} G={G} it={it} error {e}", flush=True)
