"""The `cbc` command line (host C ingest + C ABI) on the GPU box: README interface, reference-compatible streams."""
import os
import subprocess
import tempfile

import pytest

import oracle_lib as O
from cbc_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cbc_b200", "_build", "cbc")


def _run(*args):
    p = subprocess.run([CLI, *args], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


@pytest.mark.parametrize("kw,L", [
    (dict(seed=41, genome_len=300_000, n_reads=30_000, len_min=100, len_max=100, p_sub=0.005, p_indel=0.002, p_clip=0.05), 100),
    (dict(seed=42, genome_len=600_000, n_chr=2, n_reads=20_000, len_min=150, len_max=150, p_sub=0.005), 150),
])
def test_cli_roundtrip_and_reference_compatibility(kw, L):
    cfg = synth.SynthConfig(**kw)
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    with tempfile.TemporaryDirectory() as d:
        fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
        synth.write_fasta(fa, g); synth.write_sam(sam, b, g)
        # README interface, blocked container
        out = _run("-c", sam, os.path.join(d, "a.cbc"), fa)
        assert "Final Size:" in out and "Compression took" in out
        _run("-d", os.path.join(d, "a.cbc"), os.path.join(d, "a.txt"), fa)
        with open(os.path.join(d, "a.txt"), "rb") as f:
            assert f.read() == b.seq_lines()
        # single-block mode == the reference's stream; checked-in spelling `-c 1` / `-x`
        _run("-c", "1", "-1", sam, os.path.join(d, "s.cbc"), fa)
        with open(os.path.join(d, "s.cbc"), "rb") as f:
            stream = f.read()
        ostream, _ = O.encode_legacy(b, g, L)
        assert stream == ostream
        if O.have_reference():
            ref_stream, _, _ = O.run_reference(sam, fa, d)
            assert stream == ref_stream                              # byte-identical to the unmodified reference
            ref_text, _ = O.run_reference_decode(os.path.join(d, "s.cbc"), fa, d)   # the reference decodes our stream
            assert ref_text == b.seq_lines()
        _run("-x", os.path.join(d, "s.cbc"), os.path.join(d, "s.txt"), fa)          # we decode a reference-format stream
        with open(os.path.join(d, "s.txt"), "rb") as f:
            assert f.read() == b.seq_lines()


def test_cli_errors():
    p = subprocess.run([CLI, "-c", "only_one_file"], capture_output=True, text=True)
    assert p.returncode != 0 and "Missing required filenames" in p.stderr


def test_cli_streams_batches_to_several_contexts():
    """`-B 1 -g 0,0`: the SAM text is cut into 1 MB batches at line starts, parsed and packed on an ingest thread while two
    device threads (here two contexts on GPU 0: the multi-GPU path) code them, and written in order as one "CBCS" file
    of self-contained containers; `cbc -d -g 0,0` reads it back, and so does the Python side (cbc_b200.shard)."""
    from cbc_b200 import shard
    from cbc_b200.codec import Codec
    cfg = synth.SynthConfig(seed=43, genome_len=900_000, n_chr=3, n_reads=40_000, len_min=100, len_max=100, p_sub=0.005, p_indel=0.002)
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    with tempfile.TemporaryDirectory() as d:
        fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
        synth.write_fasta(fa, g); synth.write_sam(sam, b, g)
        assert os.path.getsize(sam) > 8 << 20
        out = _run("-c", "-B", "1", "-g", "0,0", sam, os.path.join(d, "a.cbcs"), fa)
        assert "on 2 device(s)" in out
        with open(os.path.join(d, "a.cbcs"), "rb") as f:
            data = f.read()
        shards = shard.read_shards(data)
        assert len(shards) >= 8
        c = Codec(0); c.set_reference(g)
        text = b"".join(c.decompress(s)[0] for s in shards)
        c.close()
        assert text == b.seq_lines()
        _run("-d", "-g", "0,0", os.path.join(d, "a.cbcs"), os.path.join(d, "a.txt"), fa)
        with open(os.path.join(d, "a.txt"), "rb") as f:
            assert f.read() == b.seq_lines()
        one = _run("-c", sam, os.path.join(d, "one.cbc"), fa)                     # one batch: a plain CBCB container
        assert "batches 1 " in one
        assert os.path.getsize(os.path.join(d, "one.cbc")) < len(data)           # a few thousand reads per shard: every shard pays its own start-up


def test_cli_cigar_recovery():
    """`-C`: `cbc -c -C` writes the CIGAR side sections to <out>.cig, `cbc -d -C` returns
    "CIGAR<TAB>SEQ" per read -- the input's CIGAR text byte for byte (variable-length reads with indels and soft clips,
    several batches on two contexts)."""
    cfg = synth.SynthConfig(seed=44, genome_len=500_000, n_reads=30_000, len_min=50, len_max=250, p_sub=0.005, p_indel=0.02, p_clip=0.3)
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    cig = [b.cigar[int(b.cigar_off[r]):int(b.cigar_off[r + 1])].tobytes() for r in range(b.n_reads)]
    seqs = b.seq_lines().split(b"\n")[:-1]
    want = b"".join(c + b"\t" + s + b"\n" for c, s in zip(cig, seqs))
    with tempfile.TemporaryDirectory() as d:
        fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
        synth.write_fasta(fa, g); synth.write_sam(sam, b, g)
        for extra in ((), ("-B", "1", "-g", "0,0")):
            out = _run("-c", "-l", "-C", *extra, sam, os.path.join(d, "a.cbc"), fa)
            assert "CIGAR sections:" in out
            assert os.path.getsize(os.path.join(d, "a.cbc.cig")) < sum(len(c) for c in cig) // 3
            _run("-d", "-C", *extra[2:], os.path.join(d, "a.cbc"), os.path.join(d, "a.txt"), fa)
            with open(os.path.join(d, "a.txt"), "rb") as f:
                assert f.read() == want
