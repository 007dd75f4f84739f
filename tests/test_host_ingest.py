"""Host C ingest (cbc_b200/csrc/host/sam_ingest.c): SAM + FASTA text -> the SoA batch of include/cbcg.h.
Checked against the generator, which writes the same reads as text and as a batch. CPU only."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from cbc_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cbc_b200", "_build", "libcbchost.so")


class Fasta(C.Structure):
    _fields_ = [("n", C.c_uint32), ("names", C.POINTER(C.c_char_p)), ("bases", C.POINTER(C.POINTER(C.c_uint8))),
                ("len", C.POINTER(C.c_uint64))]


class HBatch(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("cap", C.c_uint64),
                ("pos", C.POINTER(C.c_uint32)), ("flag", C.POINTER(C.c_uint16)), ("seq_len", C.POINTER(C.c_uint16)),
                ("chr", C.POINTER(C.c_uint32)), ("seq_off", C.POINTER(C.c_uint64)), ("cigar_off", C.POINTER(C.c_uint64)),
                ("md_off", C.POINTER(C.c_uint64)), ("seq", C.POINTER(C.c_uint8)), ("cigar", C.POINTER(C.c_uint8)),
                ("md", C.POINTER(C.c_uint8)), ("seq_cap", C.c_uint64), ("cigar_cap", C.c_uint64), ("md_cap", C.c_uint64),
                ("read_len_header", C.c_uint32), ("max_len", C.c_uint32), ("n_unmapped", C.c_uint64), ("n_lines", C.c_uint64)]


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-s", "-C", ROOT, "host"], check=True)
    l = C.CDLL(LIB)
    l.cbch_read_fasta.argtypes = [C.c_char_p, C.POINTER(Fasta), C.c_char_p, C.c_size_t]
    l.cbch_read_sam.argtypes = [C.c_char_p, C.POINTER(Fasta), C.c_int, C.POINTER(HBatch), C.c_char_p, C.c_size_t]
    l.cbch_read_sam_mt.argtypes = [C.c_char_p, C.POINTER(Fasta), C.c_int, C.c_int, C.POINTER(HBatch), C.c_char_p, C.c_size_t]
    return l


def _arr(ptr, n, dt):
    return np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].astype(dt, copy=True)


def _parse(lib, sam, fa, var_length=0):
    f, b = Fasta(), HBatch()
    err = C.create_string_buffer(256)
    assert lib.cbch_read_fasta(fa.encode(), C.byref(f), err, 256) == 0, err.value
    rc = lib.cbch_read_sam(sam.encode(), C.byref(f), var_length, C.byref(b), err, 256)
    return rc, f, b, err.value.decode()


@pytest.mark.parametrize("kw", [
    dict(seed=31, genome_len=200_000, n_reads=5000, len_min=100, len_max=100, p_sub=0.01, p_indel=0.01, p_clip=0.2),
    dict(seed=32, genome_len=400_000, n_chr=3, n_reads=4000, len_min=50, len_max=250, p_sub=0.01, p_indel=0.02, p_clip=0.3),
])
def test_sam_text_parses_to_the_generators_batch(lib, kw):
    cfg = synth.SynthConfig(**kw)
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    with tempfile.TemporaryDirectory() as d:
        fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
        synth.write_fasta(fa, g); synth.write_sam(sam, b, g)
        rc, f, hb, err = _parse(lib, sam, fa, var_length=int(cfg.len_min != cfg.len_max))
        assert rc == 0, err
        n = hb.n_reads
        assert n == b.n_reads and f.n == g.n_chr
        for c in range(g.n_chr):
            assert f.names[c].decode() == g.names[c] and f.len[c] == len(g.bases[c])
            assert np.array_equal(_arr(f.bases[c], f.len[c], np.uint8), g.bases[c])
        assert np.array_equal(_arr(hb.pos, n, np.uint32), b.pos) and np.array_equal(_arr(hb.flag, n, np.uint16), b.flag)
        assert np.array_equal(_arr(hb.seq_len, n, np.uint16), b.seq_len) and np.array_equal(_arr(hb.chr, n, np.uint32), b.chr)
        for name in ("seq", "cigar", "md"):
            off = _arr(getattr(hb, name + "_off"), n + 1, np.uint64)
            assert np.array_equal(off, getattr(b, name + "_off"))
            assert np.array_equal(_arr(getattr(hb, name), int(off[-1]), np.uint8), getattr(b, name)[:int(off[-1])])
        want = int(b.seq_len.max()) if cfg.len_min != cfg.len_max else int(b.seq_len[1])
        assert hb.read_len_header == want                         # get_read_length, src/sam_file_allocation.c:26-79


def test_edge_records(lib):
    ref = "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"
    with tempfile.TemporaryDirectory() as d:
        fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
        with open(fa, "w") as f:
            f.write(">chrA some description\nACGTACGTAC\nGTACGTACGT\r\nACGTACGTACGTACGTACGT\n>chrB\nTTTT\n")
        lines = ["@HD\tVN:1.6", "@SQ\tSN:chrA\tLN:40",
                 "r0\t0\tchrA\t1\t60\t8M\t*\t0\t0\tACGTACGT\tIIIIIIII\tNM:i:0\tMD:Z:8",            # MD last field, no trailing tab
                 "r1\t4\t*\t0\t0\t*\t*\t0\t0\tACGTAC\tIIIIII",                                      # unmapped: skipped
                 "r2\t16\tchrA\t5\t60\t6M\t*\t0\t0\tACGTAC\tIIIIII\tMD:Z:6\tAS:i:0\r",            # CRLF
                 "r3\t0\tchrB\t1\t60\t4M\t*\t0\t0\tTTTA\tIIII\tXX:Z:q\tMD:Z:3T0"]
        with open(sam, "w") as f:
            f.write("\n".join(lines) + "\n")
        rc, f, hb, err = _parse(lib, sam, fa)
        assert rc == 0, err
        assert f.n == 2 and f.names[0] == b"chrA" and f.len[0] == 40 and bytes(_arr(f.bases[0], 40, np.uint8)) == ref.encode()
        assert hb.n_reads == 3 and hb.n_unmapped == 1
        assert _arr(hb.pos, 3, np.uint32).tolist() == [1, 5, 1] and _arr(hb.chr, 3, np.uint32).tolist() == [0, 0, 1]
        md_off = _arr(hb.md_off, 4, np.uint64); md = bytes(_arr(hb.md, int(md_off[-1]), np.uint8))
        assert md == b"863T0" and md_off.tolist() == [0, 1, 2, 5]
        assert hb.read_len_header == 6                              # second record's SEQ length, mapped or not
        # errors are reported, not asserted
        with open(sam, "w") as f:
            f.write("r0\t0\tchrZ\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\tMD:Z:4\n")
        rc, _, _, err = _parse(lib, sam, fa)
        assert rc == -4 and "chrZ" in err
        with open(sam, "w") as f:
            f.write("r0\t0\tchrA\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\n")
        rc, _, _, err = _parse(lib, sam, fa)
        assert rc == -5


def _snapshot(hb):
    n = hb.n_reads
    d = {"n": n, "L": hb.read_len_header, "max": hb.max_len, "unmapped": hb.n_unmapped, "lines": hb.n_lines}
    for name, dt in (("pos", np.uint32), ("flag", np.uint16), ("seq_len", np.uint16), ("chr", np.uint32)):
        d[name] = _arr(getattr(hb, name), n, dt)
    for name in ("seq", "cigar", "md"):
        off = _arr(getattr(hb, name + "_off"), n + 1, np.uint64)
        d[name + "_off"] = off
        d[name] = _arr(getattr(hb, name), int(off[-1]), np.uint8)
    return d


def test_threaded_ingest_does_not_depend_on_the_worker_count(lib):
    """cbch_read_sam_mt cuts the file at line starts, parses the ranges in parallel and merges them: same batch, same
    whole-file facts (second record's length, line and unmapped counts), same error line for any worker count."""
    cfg = synth.SynthConfig(seed=33, genome_len=300_000, n_chr=3, n_reads=6000, len_min=50, len_max=250, p_sub=0.01, p_indel=0.02, p_clip=0.3)
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    with tempfile.TemporaryDirectory() as d:
        fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
        synth.write_fasta(fa, g); synth.write_sam(sam, b, g)
        text = open(sam).read().split("\n")
        # headers, an unmapped record as the file's second record, and unmapped records sprinkled through the file
        body = [l for l in text if l and not l.startswith("@")]
        unm = "u\t4\t*\t0\t0\t*\t*\t0\t0\tACGTACG\tIIIIIII"
        lines = ["@HD\tVN:1.6", body[0], unm] + [x for i, l in enumerate(body[1:]) for x in ([l, unm] if i % 97 == 0 else [l])]
        with open(sam, "w") as f:
            f.write("\n".join(lines) + "\n")
        f_, ref = Fasta(), None
        err = C.create_string_buffer(256)
        assert lib.cbch_read_fasta(fa.encode(), C.byref(f_), err, 256) == 0
        for T in (1, 2, 3, 7, 16, 64):
            hb = HBatch()
            assert lib.cbch_read_sam_mt(sam.encode(), C.byref(f_), 0, T, C.byref(hb), err, 256) == 0, err.value
            snap = _snapshot(hb)
            if ref is None:
                ref = snap
                assert snap["n"] == b.n_reads and snap["L"] == 7 and snap["unmapped"] == lines.count(unm) and snap["lines"] == len(lines)
                assert np.array_equal(snap["pos"], b.pos) and np.array_equal(snap["seq"], b.seq[:len(snap["seq"])])
            else:
                for k, v in ref.items():
                    assert np.array_equal(v, snap[k]) if isinstance(v, np.ndarray) else v == snap[k], (T, k)
        # a record without MD:Z three quarters into the file: every worker count reports the same whole-file line number
        bad_at = 3 * len(lines) // 4
        lines[bad_at] = "\t".join(lines[bad_at].split("\t")[:11])
        with open(sam, "w") as f:
            f.write("\n".join(lines) + "\n")
        for T in (1, 5, 32):
            hb = HBatch()
            assert lib.cbch_read_sam_mt(sam.encode(), C.byref(f_), 0, T, C.byref(hb), err, 256) == -5
            assert err.value.decode().startswith(f"line {bad_at + 1}:"), (T, err.value)


def test_compact_batch_packer_round_trips():
    """cbch_pack_batch (2 bits per base, text lengths, chromosome runs, a list for whatever is not A/C/G/T): unpacked on
    the host it is the batch again, whatever the worker count."""
    import ctypes as C
    from cbc_b200 import synth
    from cbc_b200.codec import CompactBatch
    cfg = synth.SynthConfig(seed=3, genome_len=200_000, n_chr=3, n_reads=5000, len_min=50, len_max=250, p_sub=0.01, p_indel=0.01,
                            p_clip=0.2, p_n=0.01)
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    images = []
    for threads in (1, 3, 8):
        c = CompactBatch(b, pinned=False, threads=threads)
        v = c.c.v
        n = v.n_reads
        assert n == b.n_reads and v.n_runs == 3

        def arr(ptr, ct, m):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), (m,)).copy() if m else np.zeros(0, np.uint64)
        so = np.concatenate([[0], np.cumsum((b.seq_len.astype(np.uint64) + 3) // 4)]).astype(np.uint64)
        tb = arr(v.tile_base, C.c_uint64, 4 * ((n + 127) // 128 + 1)).reshape(-1, 4)
        cut = np.minimum(np.arange(tb.shape[0]) * 128, n)
        assert np.array_equal(tb[:, 0], b.seq_off[cut]) and np.array_equal(tb[:, 1], so[cut])
        assert np.array_equal(tb[:, 2], b.cigar_off[cut]) and np.array_equal(tb[:, 3], b.md_off[cut])
        assert v.max_len == int(b.seq_len.max()) and v.min_len == int(b.seq_len.min())
        seq2 = arr(v.seq2, C.c_uint8, int(so[n]))
        lut = np.frombuffer(b"ACGT", np.uint8)
        out = []
        for r in range(n):
            by = seq2[int(so[r]):int(so[r + 1])]
            out.append(lut[np.stack([(by >> (2 * k)) & 3 for k in range(4)], axis=1).reshape(-1)[:int(b.seq_len[r])]])
        seq = np.concatenate(out)
        er, eb, ec = arr(v.exc_read, C.c_uint32, v.n_exc), arr(v.exc_base, C.c_uint16, v.n_exc), arr(v.exc_char, C.c_uint8, v.n_exc)
        assert v.n_exc > 0 and np.all(np.diff(er.astype(np.int64)) >= 0)
        seq[b.seq_off[er].astype(np.int64) + eb] = ec
        assert np.array_equal(seq, b.seq[:len(seq)])
        assert np.array_equal(arr(v.cigar_len, C.c_uint16, n), np.diff(b.cigar_off).astype(np.uint16))
        assert np.array_equal(arr(v.md_len, C.c_uint16, n), np.diff(b.md_off).astype(np.uint16))
        assert np.array_equal(arr(v.run_chr, C.c_uint32, 3), np.array([0, 1, 2], np.uint32))
        # the fixed fields and the CIGAR / MD text travel as they are (copied by the workers, each its own reads)
        assert np.array_equal(arr(v.pos, C.c_uint32, n), b.pos) and np.array_equal(arr(v.flag, C.c_uint16, n), b.flag)
        assert np.array_equal(arr(v.seq_len, C.c_uint16, n), b.seq_len)
        assert np.array_equal(arr(v.cigar, C.c_uint8, int(b.cigar_off[n])), b.cigar[:int(b.cigar_off[n])])
        assert np.array_equal(arr(v.md, C.c_uint8, int(b.md_off[n])), b.md[:int(b.md_off[n])])
        run_first = arr(v.run_first, C.c_uint64, 3)
        assert run_first[0] == 0 and np.array_equal(b.chr[run_first.astype(np.int64)], np.array([0, 1, 2], np.uint32))
        assert c.link_bytes < 0.45 * (b.seq.nbytes + b.cigar.nbytes + b.md.nbytes + n * 36)
        images.append((seq2.tobytes(), er.tobytes(), eb.tobytes(), ec.tobytes()))
        c.close()
    assert images[0] == images[1] == images[2]
