import glob, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from cbc_b200 import synth
from cbc_b200.codec import Codec
meta = json.load(open(os.path.join(ROOT, "tests/golden/indels_100.json")))
cfg = synth.SynthConfig(**meta["synth"]); g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
c = Codec(0); c.set_reference(g)
orecs, oedits = O.extract(b, g)
for R, G in ((1, 0), (1, 1), (7, 0), (64, 1)):
    for it in range(4):
        cont = c.compress(b, 100, R, G)
        recs, chr_, edits = c.decode_edits(cont)
        text, n = c.decompress(cont)
        okr = np.array_equal(recs, orecs); oke = np.array_equal(edits, oedits); okt = text == b.seq_lines()
        msg = ""
        if not oke and len(edits) == len(oedits):
            bad = np.nonzero(edits != oedits)[0]; msg = f"edit diffs at {bad[:5]} got {edits[bad[:5]]} want {oedits[bad[:5]]}"
        if not okr:
            bad = np.nonzero(recs != orecs)[0]; msg += f" rec diffs at {bad[:5]}: {recs[bad[:3]]} vs {orecs[bad[:3]]}"
        print(f"R={R} G={G} it={it} recs={okr} edits={oke} text={okt} {msg}", flush=True)
