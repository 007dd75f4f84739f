import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from cbc_b200 import synth
from cbc_b200.codec import Codec, CbcgError
cfg = synth.SynthConfig.named("config2", scale=0.3); g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
c = Codec(0); c.set_reference(g)
orecs, oedits = O.extract(b, g)
off = orecs["edit_off"].astype(np.int64)
for R in (64, 1000):
    cont = c.compress(b, 150, R, 0)
    assert cont == O.encode_blocked(b, g, 150, R, 0)
    for it in range(3):
        try:
            recs, chr_, edits = c.decode_edits(cont)
        except CbcgError as e:
            print(f"R={R} it={it} error {e}", flush=True); continue
        if np.array_equal(edits, oedits) and np.array_equal(recs, orecs):
            print(f"R={R} it={it} ok", flush=True); continue
        bad_e = np.nonzero(edits != oedits)[0] if len(edits) == len(oedits) else np.array([-1])
        bad_r = np.nonzero(recs != orecs)[0]
        desc = []
        for e in bad_e[:8]:
            r = int(np.searchsorted(off, e, side="right") - 1)
            last_entry = e == off[r] + int(orecs[r]["n_dels"]) + int(orecs[r]["n_snps"]) + int(orecs[r]["n_ins"]) - 1
            # is r the last read with edits in its block?
            blk_end = (r // R + 1) * R
            later = np.nonzero(orecs["match"][r + 1:min(blk_end, len(orecs))] == 0)[0]
            desc.append(f"e{e}:r{r}(ord {r % R}/{R}, last_entry={bool(last_entry)}, later_edited_reads_in_block={len(later)}) got {hex(edits[e])} want {hex(oedits[e])}")
        print(f"R={R} it={it} bad edits {len(bad_e)} bad recs {len(bad_r)} first bad rec {bad_r[:3]}: " + "; ".join(desc), flush=True)
