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
    assert 64 <= chosen <= 1280
    # the block size the library chose is an ordinary block size: the CPU restatement writes the same container
    assert cont == O.encode_blocked(b, g, 150, chosen, 1)


def test_config1_shape_full_size(codec):
    cfg = synth.SynthConfig.named("config1")                              # 1 M x 100 bp, 0.5 % sub, 0.1 % indel
    g, b, cont = _roundtrip(codec, cfg, 100)
    otext, on = O.decode_blocked(cont, g)                                 # the CPU restatement decodes the GPU's container
    assert on == b.n_reads and otext == b.seq_lines()


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
