"""Region shards on the GPU: two contexts on device 0 stand for two ranks. Each codes its contiguous range of ONE
position-sorted input with the CUDA coder, the shard containers go into one "CBCS" file at the scanned offsets
(cbc_b200.shard), and each context decodes the OTHER context's shard from that file: the concatenation
of the decoded shards is the input. Also: the state of "the last encode" does not survive a decode of another
container (encode_resident -> decode(other) -> fetch_container must not glue two containers)."""
import os
import tempfile

import numpy as np
import pytest

from cbc_b200 import shard, synth
from cbc_b200.codec import CbcgError, Codec

pytestmark = pytest.mark.gpu

AUTO = 0xffffffff


@pytest.mark.parametrize("name,scale,world", [("config3", 0.02, 2), ("config4", 0.0005, 3), ("config5", 0.02, 2)])
def test_cbcs_round_trip_with_the_gpu_coder(name, scale, world):
    cfg = synth.SynthConfig.named(name, scale=scale)
    g = synth.make_genome(cfg)
    whole = synth.make_reads(cfg, g)
    L = cfg.len_max
    ranges = shard.shard_ranges(cfg.n_reads, world)
    codecs = [Codec(0) for _ in range(world)]
    conts, heads, payloads, parts = [], [], [], []
    for r, (r0, r1) in enumerate(ranges):
        mine = synth.make_reads(cfg, g, r0, r1)              # each "rank" generates only its region
        parts.append(mine)
        codecs[r].set_reference(g)
        codecs[r].upload(mine)
        codecs[r].encode_resident(L, AUTO, 1)
        head, pb = codecs[r].fetch_index()
        cont = codecs[r].fetch_container().tobytes()
        assert cont[:len(head)] == head and len(cont) == len(head) + pb
        assert shard.container_head_len(cont) == len(head)
        conts.append(cont); heads.append(head); payloads.append(pb)
    assert b"".join(p.seq_lines() for p in parts) == whole.seq_lines()
    layout = shard.layout_from_sizes([len(h) for h in heads], payloads, heads)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.cbcs")
        for r in reversed(range(world)):                     # any write order gives the same file
            shard.write_shard(path, r, layout, conts[r])
        with open(path, "rb") as f:
            data = f.read()
    assert len(data) == layout.total
    shards = shard.read_shards(data)
    text = b""
    for r in range(world):
        t, n = codecs[(r + 1) % world].decompress(shards[r])  # decoded by another context: shards are self-contained
        assert n == parts[r].n_reads
        text += t
    assert text == whole.seq_lines()
    for c in codecs:
        c.close()


def test_last_encode_state_does_not_survive_a_foreign_decode():
    cfg = synth.SynthConfig(seed=9, genome_len=150_000, n_reads=9_000, len_min=100, len_max=100, p_sub=0.01)
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    a, other = b.slice(0, 5000), b.slice(5000, 9000)
    c = Codec(0)
    c.set_reference(g)
    foreign = c.compress(other, 100, 256, 1)
    c.upload(a)
    c.encode_resident(100, 256, 1)
    mine = c.fetch_container().tobytes()
    text, n = c.decompress(foreign)                          # overwrites the payload and descriptors of "the last encode"
    assert n == 4000 and text == other.seq_lines()
    with pytest.raises(CbcgError):
        c.fetch_container()
    with pytest.raises(CbcgError):
        c.decode_resident()
    c.encode_resident(100, 256, 1)                           # the batch is still resident
    assert c.fetch_container().tobytes() == mine
    c.close()


def test_batch_validation_reports_input_errors():
    cfg = synth.SynthConfig(seed=3, genome_len=50_000, n_reads=500, len_min=100, len_max=100)
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    c = Codec(0)
    c.set_reference(g)
    bad = b.slice(0, 500)
    bad.seq_len[7] = 0
    with pytest.raises(CbcgError) as e:
        c.upload(bad)
    assert e.value.status == -6
    bad = b.slice(0, 500)
    bad.md_off[100] = bad.md_off[99] - np.uint64(1) if bad.md_off[99] else np.uint64(5)
    bad.md_off[99] = bad.md_off[100] + np.uint64(3)
    with pytest.raises(CbcgError) as e:
        c.upload(bad)
    assert e.value.status == -6
    c.upload(b)                                              # the context is still usable
    c.encode_resident(100, 128, 0)
    c.close()


@pytest.mark.parametrize("name,scale,L", [("config2", 0.01, 150), ("config5", 0.01, 250), ("config4", 0.0005, 150)])
def test_compact_batch_gives_the_same_container(name, scale, L):
    """cbcg_encode_compact: the batch crosses the link at 2 bits per base and is unpacked on the device (k0_unpack.cu); the
    container is the cbcg_encode one byte for byte -- small batches (one-stream path), bases that are not A/C/G/T, several
    chromosomes, variable length."""
    from cbc_b200.codec import CompactBatch
    cfg = synth.SynthConfig.named(name, scale=scale)
    cfg.p_n = 0.003
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    c = Codec(0)
    c.set_reference(g)
    cb = CompactBatch(b)
    assert cb.link_bytes < 0.45 * (b.seq.nbytes + b.cigar.nbytes + b.md.nbytes + b.n_reads * 36)
    for R, gm in ((AUTO, 1), (300, 0), (0, 0)):
        want = c.compress(b, L, R, gm)
        out = np.empty(len(want) + 4096, np.uint8)
        n = c.compress_compact_into(cb, L, R, out, gm)
        assert out[:n].tobytes() == want
    c.upload_compact(cb)
    recs, edits = c.extract(b)                                # K1 over a normally uploaded batch ...
    c.upload_compact(cb)
    c.encode_resident(L, 256, 1)
    text, nr = c.decompress(c.fetch_container().tobytes())
    assert nr == b.n_reads and text == b.seq_lines()
    cb.close(); c.close()


def test_compact_batch_through_the_pipelined_encode(monkeypatch):
    from cbc_b200.codec import CompactBatch
    monkeypatch.setenv("CBCG_PIPE_MIN_READS", "200000")
    cfg = synth.SynthConfig.named("config2", scale=0.2)       # ~600 k reads: the chunked, overlapped path
    cfg.p_n = 0.001
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    c = Codec(0)
    c.set_reference(g)
    want = c.compress(b, 150, AUTO, 1)
    cb = CompactBatch(b)
    out = np.empty(len(want) + 65536, np.uint8)
    fresh = Codec(0)                                          # a context whose buffers have never held this batch in any form
    fresh.set_reference(g)
    n = fresh.compress_compact_into(cb, 150, AUTO, out, 1)
    got = out[:n].tobytes()
    fresh.close()
    n = c.compress_compact_into(cb, 150, AUTO, out, 1)
    assert out[:n].tobytes() == got
    # the encoder sizes last-generation blocks by how fast the batch arrives (a flatter ramp for the compact form): another
    # cut than the plain batch's, and the CPU restatement given that cut writes the same bytes
    import oracle_lib as O
    assert got == O.encode_like(got, b, g)
    assert abs(len(got) - len(want)) < 0.002 * len(want)
    st = c.stats()
    assert st["h2d_bytes"] < 0.45 * (b.seq.nbytes + b.cigar.nbytes + b.md.nbytes + b.n_reads * 36)
    text, nr = c.decompress(got)
    assert nr == b.n_reads and text == b.seq_lines()
    # lengths that disagree with the tile offsets are caught on the device
    import ctypes as C
    sl = np.ctypeslib.as_array(C.cast(cb.c.v.seq_len, C.POINTER(C.c_uint16)), (b.n_reads,))
    sl[1000] += 1
    with pytest.raises(CbcgError) as e:
        c.compress_compact_into(cb, 150, AUTO, out, 1)
    assert e.value.status == -6
    sl[1000] -= 1
    n2 = c.compress_compact_into(cb, 150, AUTO, out, 1)     # and the context is still usable
    assert out[:n2].tobytes() == got
    cb.close(); c.close()


def test_contexts_in_flight_from_their_own_host_threads():
    """cbcg.h: one host thread per context, no globals. Three contexts on device 0, each driven by its own thread, code
    DIFFERENT batches at the same time (resident round trips and host-buffer calls interleaved): every container equals
    the one the same batch gives on a quiet device, every decode returns its own input. This is how bench.py keeps
    several batches in flight per GPU."""
    import threading
    cfgs = [synth.SynthConfig.named("config2", scale=0.05), synth.SynthConfig.named("config5", scale=0.03),
            synth.SynthConfig.named("config1", scale=0.2)]
    jobs = []
    for cfg in cfgs:
        g = synth.make_genome(cfg)
        b = synth.make_reads(cfg, g)
        c = Codec(0)
        c.set_reference(g)
        jobs.append((cfg, g, b, c))
    quiet = []
    for cfg, g, b, c in jobs:                                 # one at a time first
        c.upload(b)
        c.encode_resident(cfg.len_max, AUTO, 1, 0)
        quiet.append(c.fetch_container().tobytes())
    quiet_host = [c.compress(b, cfg.len_max, AUTO, 1, None, 0) for cfg, g, b, c in jobs]
    errors = []

    def work(k):
        cfg, g, b, c = jobs[k]
        try:
            for it in range(4):
                c.upload(b)
                c.encode_resident(cfg.len_max, AUTO, 1, 0)
                assert c.fetch_container().tobytes() == quiet[k], f"context {k}: container differs under concurrency"
                c.decode_resident()
                assert c.fetch_decoded().tobytes() == b.seq_lines(), f"context {k}: resident decode differs"
                cont = c.compress(b, cfg.len_max, AUTO, 1, None, 0)
                assert cont == quiet_host[k], f"context {k}: host-buffer container differs"
                text, n = c.decompress(cont)
                assert n == b.n_reads and text == b.seq_lines(), f"context {k}: host-buffer decode differs"
        except Exception as e:                               # noqa: BLE001 -- reported by the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    for t in threads: t.start()
    for t in threads: t.join()
    for _, _, _, c in jobs:
        c.close()
    assert not errors, errors
