"""CPU experiment (oracle only): container size of generation-primed blocks vs the single reference stream."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O                      # noqa: E402
from cbc_b200 import synth                  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
nreads = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
cfg = synth.SynthConfig.named(name, scale=1.0)
scale = nreads / cfg.n_reads
cfg = synth.SynthConfig.named(name, scale=scale)
g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
L = cfg.len_max
t = time.time(); single, _ = O.encode_legacy(b, g, L); print("single", len(single), f"{8*len(single)/b.total_bases():.4f} b/base", f"{time.time()-t:.1f}s", flush=True)


def sched(R, counts, reads):
    gg = O._G(g); cb = b.c_struct(); out = O._Buf()
    n = len(counts)
    ca = (C.c_uint32 * max(n, 1))(*counts); ra = (C.c_uint32 * max(n, 1))(*reads)
    t = time.time()
    rc = O.lib().cbco_encode_scheduled(C.byref(cb), C.byref(gg.s), C.c_uint32(L), C.c_uint32(R), C.c_uint32(n), ca, ra, C.byref(out))
    assert rc == 0, rc
    data = O._take(out)
    return data, time.time() - t


import struct
def split(data):
    nb, nchr = struct.unpack_from("<II", data, 24)
    o = 40
    for _ in range(nchr):
        nl, = struct.unpack_from("<I", data, o); o += 4 + nl + ((4 - (nl & 3)) & 3)
    ixb, = struct.unpack_from("<I", data, o); return nb, o + 4 + ixb, len(data) - o - 4 - ixb

SCHEDS = [([16, 112, 384, 1536], [16, 32, 64, 128]), ([128, 384, 1536], [16, 64, 128]), ([64, 448, 1536], [16, 48, 128]),
          ([256, 768, 1536], [8, 32, 128]), ([32, 224, 1536], [16, 64, 128])]
for R in (1179,):
    for counts, reads in SCHEDS:
        data, dt = sched(R, counts, reads)
        nb, head, pay = split(data)
        print(f"serial={sum(reads)} early_reads={sum(c*r for c,r in zip(counts,reads))} R={R} sched={list(zip(counts, reads))} blocks={nb} payload={pay} ({100*(pay-len(single))/len(single):+.2f}%) head={head} ({100*head/len(single):.2f}%) flush~{100*nb*4/len(single):.2f}% {dt:.1f}s", flush=True)
