"""Size-independent properties at BASELINE.json's full sizes (the oracle checks the small cases): encode -> decode
round trips, idempotence, automatic block sizing, and the <= 1 % bits/base budget of the blocked container against
the reference's single stream (whose size comes from the CPU restatement, ~3 s per 3 M reads)."""
import struct

import numpy as np
import pytest

import oracle_lib as O
from cbc_b200 import synth

pytestmark = pytest.mark.gpu

AUTO = 0xffffffff


@pytest.fixture(scope="module")
def codec():
    from cbc_b200.codec import Codec
    c = Codec(0)
    yield c
    c.close()


def _roundtrip(codec, cfg, L, gen_mode=1, block_reads=AUTO):
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    codec.set_reference(g)
    codec.upload(b)
    codec.encode_resident(L, block_reads, gen_mode)
    cont = codec.fetch_container().tobytes()
    codec.decode_resident()
    assert codec.fetch_decoded().tobytes() == b.seq_lines()
    codec.encode_resident(L, block_reads, gen_mode)                       # idempotent
    assert codec.fetch_container().tobytes() == cont
    text, n = codec.decompress(cont)                                      # host-buffer path agrees with the resident one
    assert n == b.n_reads and text == b.seq_lines()
    return g, b, cont


def test_config2_full_size_roundtrip_and_bits_per_base_budget(codec):
    cfg = synth.SynthConfig.named("config2")                              # 3 014 484 reads x 150 bp, 30x
    g, b, cont = _roundtrip(codec, cfg, 150)
    single, _ = O.encode_legacy(b, g, 150)                                # the reference's single stream (size only)
    overhead = (len(cont) - len(single)) / len(single)
    assert 0.0 < overhead <= 0.01, overhead                               # north_star: <= 1 % from blocking
    chosen = struct.unpack_from("<I", cont, 32)[0]
    assert 64 <= chosen <= 16384
    # the block size the library chose is an ordinary block size: the CPU restatement writes the same container
    assert cont == O.encode_blocked(b, g, 150, chosen, 1)


def test_config1_shape_full_size(codec):
    cfg = synth.SynthConfig.named("config1")                              # 1 M x 100 bp, 0.5 % sub, 0.1 % indel
    g, b, cont = _roundtrip(codec, cfg, 100)
    otext, on = O.decode_blocked(cont, g)                                 # the CPU restatement decodes the GPU's container
    assert on == b.n_reads and otext == b.seq_lines()
    single, _ = O.encode_legacy(b, g, 100)
    overhead = (len(cont) - len(single)) / len(single)
    assert 0.0 < overhead <= 0.01, overhead                               # the <= 1 % budget on the smallest named shape (round 1: 2.7 %)


def test_config3_shape_half_size_budget(codec):
    cfg = synth.SynthConfig.named("config3", scale=0.5)                   # 6.4 M x 150 bp: several waves of last-generation blocks
    g, b, cont = _roundtrip(codec, cfg, 150)
    single, _ = O.encode_legacy(b, g, 150)
    overhead = (len(cont) - len(single)) / len(single)
    assert 0.0 < overhead <= 0.01, overhead                               # (round 1: 1.24 % on config 3)


@pytest.mark.parametrize("name,L", [("config2", 150), ("config1", 100)])
def test_default_layout_full_size_budget(codec, name, L):
    """substreams = 0, what the command line and bench.py use: four substreams in the narrow early generations, one
    stream in the wide ones. Same bytes as the CPU restatement given the cut, inside the budget, decodes to the input."""
    cfg = synth.SynthConfig.named(name)
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    codec.set_reference(g)
    codec.upload(b)
    codec.encode_resident(L, AUTO, 1, substreams=0)
    cont = codec.fetch_container().tobytes()
    codec.decode_resident()
    assert codec.fetch_decoded().tobytes() == b.seq_lines()
    mode = struct.unpack_from("<I", cont, 36)[0]
    assert (mode >> 16) & 0xff >= 4 and not mode & 0x200
    assert cont == O.encode_like(cont, b, g)
    single, _ = O.encode_legacy(b, g, L)
    assert 0.0 < (len(cont) - len(single)) / len(single) <= 0.01
    text, n = codec.decompress(cont)                                      # host-buffer (pipelined) decode
    assert n == b.n_reads and text == b.seq_lines()


def test_four_substream_container_full_size(codec):
    """CBCG_MODE_SPLIT4 at config-2 size: a CTA per block, a warp per substream, pipelined decode; the CPU restatement
    writes the same bytes for the same cut, and the budget holds."""
    cfg = synth.SynthConfig.named("config2")
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    codec.set_reference(g)
    codec.upload(b)
    codec.encode_resident(150, AUTO, 1, substreams=4)
    cont = codec.fetch_container().tobytes()
    codec.decode_resident()
    assert codec.fetch_decoded().tobytes() == b.seq_lines()
    assert struct.unpack_from("<I", cont, 36)[0] & 0x200
    assert cont == O.encode_like(cont, b, g)
    single, _ = O.encode_legacy(b, g, 150)
    assert 0.0 < (len(cont) - len(single)) / len(single) <= 0.01


def test_config5_shape_variable_length_indel_heavy(codec):
    cfg = synth.SynthConfig.named("config5", scale=0.25)                  # 50-250 bp, 2 % indels, soft clips: ~750 k reads
    g, b, cont = _roundtrip(codec, cfg, 250)
    recs, edits = codec.extract(b)
    orecs, oedits = O.extract(b, g)
    assert np.array_equal(recs, orecs) and np.array_equal(edits, oedits)


def test_config4_shape_many_chromosomes(codec):
    cfg = synth.SynthConfig.named("config4", scale=0.002)                 # 24 records, ~1.2 M reads
    g, b, cont = _roundtrip(codec, cfg, 150)
    assert struct.unpack_from("<I", cont, 28)[0] == 24


def test_resident_encode_orders_agree(codec, monkeypatch):
    """Large gen_mode-1 batches extract the tail of the batch on a side stream while the early generations are coded
    (api.cu, encode_resident_overlapped); CBCG_NO_OVERLAP=1 keeps everything on one stream. Same cut, same bytes;
    an indel-heavy variable-length batch exercises the workspace projection from the head's edit density."""
    for name, scale, L in (("config1", 1.0, 100), ("config5", 0.35, 250)):
        cfg = synth.SynthConfig.named(name, scale=scale)
        g = synth.make_genome(cfg)
        b = synth.make_reads(cfg, g)
        codec.set_reference(g)
        codec.upload(b)
        monkeypatch.delenv("CBCG_NO_OVERLAP", raising=False)
        codec.encode_resident(L, AUTO, 1)
        a = codec.fetch_container().tobytes()
        sa = codec.stats()
        assert sa["retried"] == 0                                         # the projection from the head held
        monkeypatch.setenv("CBCG_NO_OVERLAP", "1")
        codec.encode_resident(L, AUTO, 1)
        assert codec.fetch_container().tobytes() == a
        sb = codec.stats()
        assert sa["n_edits"] == sb["n_edits"] and sa["n_symbols"] == sb["n_symbols"] and sa["n_blocks"] == sb["n_blocks"]
        monkeypatch.delenv("CBCG_NO_OVERLAP")
        codec.decode_resident()
        assert codec.fetch_decoded().tobytes() == b.seq_lines()


def _concat(a, b):
    """Reads of a followed by reads of b (both on the same genome, a's positions below b's)."""
    from cbc_b200.batch import Batch

    def pool(oa, da, ob, db):
        return np.concatenate([oa, ob[1:] + oa[-1]]).astype(np.uint64), np.concatenate([da[:int(oa[-1])], db[:int(ob[-1])]])
    so, s = pool(a.seq_off, a.seq, b.seq_off, b.seq)
    co, c = pool(a.cigar_off, a.cigar, b.cigar_off, b.cigar)
    mo, m = pool(a.md_off, a.md, b.md_off, b.md)
    return Batch(np.concatenate([a.pos, b.pos]), np.concatenate([a.flag, b.flag]), np.concatenate([a.seq_len, b.seq_len]),
                 np.concatenate([a.chr, b.chr]), so, s, co, c, mo, m)


def test_resident_encode_falls_back_when_the_head_misjudges_the_tail(codec, monkeypatch):
    """The overlapped resident encode sizes the coder's workspace from the edit density of the batch's head. A batch whose
    first 300 k reads are perfect matches and whose other 900 k are indel-heavy makes that projection far too small: the
    plan kernel reports it, the coder and merge kernels stand down, and the call takes the one-stream order (exact sizes).
    The container must be the one-stream order's, and decode to the input."""
    import dataclasses
    clean = synth.SynthConfig(seed=91, genome_len=6_000_000, n_reads=1_200_000, len_min=150, len_max=150, p_sub=0.0)
    dirty = dataclasses.replace(clean, p_sub=0.02, p_indel=0.02)
    g = synth.make_genome(clean)
    a, d = synth.make_reads(clean, g), synth.make_reads(dirty, g)
    mid = 1_500_000
    ia, id_ = int(np.searchsorted(a.pos, mid)), int(np.searchsorted(d.pos, mid))
    b = _concat(a.slice(0, ia), d.slice(id_, d.n_reads))
    assert ia > 280_000 and b.n_reads > 1_100_000 and np.all(np.diff(b.pos.astype(np.int64)) >= 0)
    codec.set_reference(g)
    codec.upload(b)
    monkeypatch.delenv("CBCG_NO_OVERLAP", raising=False)
    codec.encode_resident(150, AUTO, 1)
    assert codec.stats()["retried"] in (0, 1)                             # 1 when the cut leaves a head small enough to misjudge the tail
    got = codec.fetch_container().tobytes()
    monkeypatch.setenv("CBCG_NO_OVERLAP", "1")
    codec.encode_resident(150, AUTO, 1)
    assert codec.stats()["retried"] == 0
    assert codec.fetch_container().tobytes() == got
    monkeypatch.delenv("CBCG_NO_OVERLAP")
    codec.decode_resident()
    assert codec.fetch_decoded().tobytes() == b.seq_lines()
    # a fresh context (its edit array starts at the default guess of one entry per 16 bases) and 8 % edited bases: K1 on
    # the tail runs out of edit entries WHILE the early generations and their merges are in flight on the other stream
    from cbc_b200.codec import Codec
    worse = dataclasses.replace(clean, n_reads=1_000_000, p_sub=0.06, p_indel=0.02)
    w = synth.make_reads(worse, g)
    c2 = Codec(0)
    try:
        c2.set_reference(g)
        c2.upload(w)
        c2.encode_resident(150, AUTO, 1)
        assert c2.stats()["retried"] in (0, 1)
        got = c2.fetch_container().tobytes()
        c2.decode_resident()
        assert c2.fetch_decoded().tobytes() == w.seq_lines()
        monkeypatch.setenv("CBCG_NO_OVERLAP", "1")
        c2.encode_resident(150, AUTO, 1)
        assert c2.fetch_container().tobytes() == got
    finally:
        monkeypatch.delenv("CBCG_NO_OVERLAP", raising=False)
        c2.close()


def test_pipelined_host_buffer_encode_and_decode(codec, monkeypatch):
    """cbcg_encode / cbcg_decode on a large batch overlap the PCIe copies with the kernels and size last-generation
    blocks by their place in the batch. Whatever cut the encoder chose, the CPU restatement given the same cut writes
    the same bytes, both decoders return the input, and the <= 1 % budget holds."""
    monkeypatch.setenv("CBCG_PIPE_MIN_READS", "200000")
    cfg = synth.SynthConfig.named("config2", scale=0.25)                  # ~750 k reads
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    codec.set_reference(g)
    cont = codec.compress(b, 150, block_reads=AUTO, gen_mode=1)           # host buffers: the pipelined path
    codec.upload(b)
    codec.encode_resident(150, AUTO, 1)
    resident = codec.fetch_container().tobytes()
    assert cont != resident                                               # a different cut (the ramp) ...
    assert struct.unpack_from("<Q", cont, 16)[0] == b.n_reads
    assert cont == O.encode_like(cont, b, g)                              # ... that the restatement reproduces byte for byte
    assert resident == O.encode_like(resident, b, g)
    for c in (cont, resident):
        text, n = codec.decompress(c)                                     # pipelined decode
        assert n == b.n_reads and text == b.seq_lines()
    otext, on = O.decode_blocked(cont, g)
    assert on == b.n_reads and otext == b.seq_lines()
    monkeypatch.setenv("CBCG_PIPE_MIN_READS", "1000000000")               # same containers through the one-stream decoder
    text, n = codec.decompress(cont)
    assert n == b.n_reads and text == b.seq_lines()


def test_pipelined_encode_many_chromosomes_and_variable_length(codec, monkeypatch):
    monkeypatch.setenv("CBCG_PIPE_MIN_READS", "200000")
    for name, scale, L in (("config4", 0.001, 150), ("config5", 0.25, 250)):     # 24 records; 50-250 bp with indels and clips
        cfg = synth.SynthConfig.named(name, scale=scale)
        g = synth.make_genome(cfg)
        b = synth.make_reads(cfg, g)
        codec.set_reference(g)
        cont = codec.compress(b, L, block_reads=AUTO, gen_mode=1)
        assert cont == O.encode_like(cont, b, g)
        text, n = codec.decompress(cont)                                  # variable length: the one-stream decoder
        assert n == b.n_reads and text == b.seq_lines()


def test_pipelined_decode_of_damaged_containers_reports_and_returns(codec, monkeypatch):
    """The pipelined decoder on a container with a damaged payload, a truncated one and too small an output buffer:
    a status (or, for damage the coder cannot see, different text), never a hang, and the context stays usable."""
    from cbc_b200.codec import CbcgError
    monkeypatch.setenv("CBCG_PIPE_MIN_READS", "200000")
    cfg = synth.SynthConfig.named("config2", scale=0.1)                   # ~300 k reads
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    codec.set_reference(g)
    cont = codec.compress(b, 150, block_reads=AUTO, gen_mode=1)
    good, n = codec.decompress(cont)
    assert n == b.n_reads and good == b.seq_lines()
    bad = bytearray(cont)
    for k in range(len(bad) // 2, len(bad) // 2 + 64):
        bad[k] ^= 0x5a
    try:
        text, n2 = codec.decompress(bytes(bad))
        assert text != good or n2 != n
    except CbcgError as e:
        assert e.status in (-9, -5, -6, -8, -10, -11)                     # corrupt, capacity, input, format, limit, internal
    with pytest.raises(CbcgError):
        codec.decompress(cont[:len(cont) - 1000])                          # payload truncated
    out = np.empty(1000, np.uint8)
    with pytest.raises(CbcgError) as e:
        codec.decompress_into(np.frombuffer(cont, np.uint8), out)          # room for 1000 bytes of text
    assert e.value.status == -5
    text, n3 = codec.decompress(cont)                                      # and the context still works
    assert n3 == b.n_reads and text == good


def test_pipelined_full_size_budget(codec):
    cfg = synth.SynthConfig.named("config2")
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    codec.set_reference(g)
    cont = codec.compress(b, 150, block_reads=AUTO, gen_mode=1)
    single, _ = O.encode_legacy(b, g, 150)
    overhead = (len(cont) - len(single)) / len(single)
    assert 0.0 < overhead <= 0.01, overhead
    text, n = codec.decompress(cont)
    assert n == b.n_reads and text == b.seq_lines()
