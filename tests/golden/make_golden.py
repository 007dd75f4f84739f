"""Generates tests/golden/*.json + *.cbc from the UNMODIFIED reference (oracle/_ref/cbc_ref,
cbc_trace, built from /root/reference by oracle/Makefile). Run in the build container only:

    python tests/golden/make_golden.py

Each fixture records a synthetic-generator configuration (cbc_b200/csrc/host/synth.c is
deterministic, so the input is reproducible anywhere), the byte stream the reference encoder
wrote for it, and SHA-256 digests of the reference's symbol trace and of the reference
decoder's output. The reference ships no golden vectors of its own."""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cbc_b200 import synth          # noqa: E402
import oracle_lib as O             # noqa: E402

SHAPES = {
    # name: (SynthConfig kwargs, header read length)
    "subs_100":        (dict(seed=1, genome_len=60_000, n_reads=3000, len_min=100, len_max=100, p_sub=0.005), 100),
    "indels_100":      (dict(seed=2, genome_len=60_000, n_reads=3000, len_min=100, len_max=100, p_sub=0.005, p_indel=0.01), 100),
    "clips_100":       (dict(seed=3, genome_len=60_000, n_reads=3000, len_min=100, len_max=100, p_sub=0.005, p_indel=0.004, p_clip=0.3), 100),
    "fwd_only_100":    (dict(seed=4, genome_len=60_000, n_reads=2000, len_min=100, len_max=100, p_sub=0.01, p_indel=0.01, p_rev=0.0), 100),
    "rev_only_100":    (dict(seed=5, genome_len=60_000, n_reads=2000, len_min=100, len_max=100, p_sub=0.01, p_indel=0.01, p_rev=1.0), 100),
    "subs_150":        (dict(seed=6, genome_len=40_000, n_reads=8000, len_min=150, len_max=150, p_sub=0.005), 150),
    "indels_250":      (dict(seed=7, genome_len=60_000, n_reads=2000, len_min=250, len_max=250, p_sub=0.005, p_indel=0.02), 250),
    "paired_flags_n":  (dict(seed=8, genome_len=60_000, n_reads=3000, len_min=100, len_max=100, p_sub=0.005, p_indel=0.002, p_n=0.002, flag_mode=1), 100),
    "two_chr":         (dict(seed=11, genome_len=800_000, n_chr=2, n_reads=3000, len_min=100, len_max=100, p_sub=0.005, p_indel=0.002), 100),
    "sparse_cov":      (dict(seed=10, genome_len=4_000_000, n_reads=2000, len_min=100, len_max=100, p_sub=0.01), 100),
}


def main():
    assert O.have_reference(), "build oracle/_ref first: make -C oracle ref"
    for name, (kw, L) in SHAPES.items():
        cfg = synth.SynthConfig(**kw)
        g = synth.make_genome(cfg)
        b = synth.make_reads(cfg, g)
        with tempfile.TemporaryDirectory() as d:
            fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
            synth.write_fasta(fa, g)
            synth.write_sam(sam, b, g)
            stream, trace, _ = O.run_reference(sam, fa, d, trace=True)
            plain, _, _ = O.run_reference(sam, fa, d, trace=False)
            assert plain == stream, "tracer changed the bytes"
            sp = os.path.join(d, "s.cbc")
            with open(sp, "wb") as f:
                f.write(stream)
            decoded, _ = O.run_reference_decode(sp, fa, d)
        assert decoded == b.seq_lines(), f"{name}: reference did not round-trip"
        with open(os.path.join(HERE, name + ".cbc"), "wb") as f:
            f.write(stream)
        meta = dict(name=name, synth=kw, read_len_header=L, n_reads=b.n_reads, stream_bytes=len(stream),
                    stream_sha256=hashlib.sha256(stream).hexdigest(),
                    trace_symbols=int(len(trace)), trace_sha256=hashlib.sha256(trace.tobytes()).hexdigest(),
                    decoded_sha256=hashlib.sha256(decoded).hexdigest())
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(meta, f, indent=1)
        print(name, len(stream), len(trace))


if __name__ == "__main__":
    main()
