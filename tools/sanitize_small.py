"""Small end-to-end pass for compute-sanitizer: K1, K2 (cold, primed, legacy), K3 (copy path, slow path, closed-form and
look-back offsets), and -- with CBCG_PIPE_MIN_READS lowered -- the pipelined host-buffer calls."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cbc_b200 import synth                      # noqa: E402
from cbc_b200.codec import Codec                # noqa: E402

AUTO = 0xffffffff
c = Codec(0)
for kw, L in ((dict(seed=5, genome_len=200_000, n_reads=20_000, len_min=150, len_max=150, p_sub=0.005, p_indel=0.002, p_clip=0.05), 150),
              (dict(seed=6, genome_len=150_000, n_chr=2, n_reads=8_000, len_min=50, len_max=250, p_sub=0.01, p_indel=0.02, p_clip=0.3), 250)):
    cfg = synth.SynthConfig(**kw)
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    c.set_reference(g)
    for R, G in ((0, 0), (512, 0), (512, 1), (AUTO, 1)):
        if R == 0 and kw["len_min"] != kw["len_max"]:
            continue
        cont = c.compress(b, L, block_reads=R, gen_mode=G)
        text, n = c.decompress(cont, legacy=(R == 0))
        assert n == b.n_reads and text == b.seq_lines(), (R, G)
    print("ok", kw["seed"], flush=True)
# many distinct FLAG values: rule F1 in the blocks, rule F2 in the merges, the single-block mode's table in the workspace;
# CIGAR recovery (class kernel, emit kernel) beside a blocked container and beside the reference's own stream
import numpy as np                              # noqa: E402
from cbc_b200.batch import Batch                # noqa: E402
cfg = synth.SynthConfig(seed=8, genome_len=150_000, n_reads=12_000, len_min=100, len_max=100, p_sub=0.01, p_indel=0.01, p_clip=0.2)
g = synth.make_genome(cfg); b0 = synth.make_reads(cfg, g)
rng = np.random.default_rng(3)
values = rng.choice(4096, size=900, replace=False).astype(np.uint16)
b = Batch(b0.pos, np.ascontiguousarray(values[rng.integers(0, 900, size=b0.n_reads)]), b0.seq_len, b0.chr, b0.seq_off, b0.seq,
          b0.cigar_off, b0.cigar, b0.md_off, b0.md)
c.set_reference(g)
cig = b"".join(b.cigar[int(b.cigar_off[r]):int(b.cigar_off[r + 1])].tobytes() + b"\n" for r in range(b.n_reads))
sec = c.cigar_pack(b)
for R, G, S in ((0, 0, 1), (2000, 0, 1), (1000, 1, 1), (AUTO, 1, 0), (1500, 1, 4)):
    cont = c.compress(b, 100, R, G, None, S)
    text, n = c.decompress(cont, legacy=(R == 0))
    assert n == b.n_reads and text == b.seq_lines(), (R, G, S)
    ctext, n = c.cigar_unpack(cont, sec, legacy=(R == 0))
    assert n == b.n_reads and ctext == cig, (R, G, S)
print("ok flags + cigar", flush=True)
if len(sys.argv) > 1:
    os.environ["CBCG_PIPE_MIN_READS"] = "1000"
    cfg = synth.SynthConfig(seed=7, genome_len=1_200_000, n_reads=int(sys.argv[1]), len_min=100, len_max=100, p_sub=0.005, p_indel=0.0, p_clip=0.0)
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    c.set_reference(g)
    cont = c.compress(b, 100, block_reads=AUTO, gen_mode=1)
    text, n = c.decompress(cont)
    assert n == b.n_reads and text == b.seq_lines()
    print("ok pipelined", flush=True)
c.close()
