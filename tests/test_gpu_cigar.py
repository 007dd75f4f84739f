"""CIGAR recovery (cbc_b200/csrc/k4_cigar.cu, cbcg_cigar_pack / cbcg_cigar_unpack). Upstream declares
reconstructCigar / cigarFlags (include/sam_block.h:179,443) and sketches decompress_cigar (src/read_decompression.c:91-113)
without implementing them, so there is no reference output to compare with; the checks are
  * the round trip: container + side section -> every read's CIGAR text, byte for byte the input's;
  * a plain-Python restatement of "the CIGAR the indels imply" (from the edit records the ORACLE extracts) and of the
    classes: the section lists exactly the reads the restatement says it must, verbatim exactly where it says so, and
    without a section the GPU emits the restatement's implied text."""
import re

import numpy as np
import pytest

import oracle_lib as O
from cbc_b200 import synth
from cbc_b200.batch import Batch
from cbc_b200.codec import CbcgError, Codec

pytestmark = pytest.mark.gpu

AUTO = 0xffffffff


@pytest.fixture(scope="module")
def codec():
    c = Codec(0)
    yield c
    c.close()


def _cigars(b: Batch):
    return [b.cigar[int(b.cigar_off[r]):int(b.cigar_off[r + 1])].tobytes() for r in range(b.n_reads)]


def _with_cigars(b: Batch, cig) -> Batch:
    off = np.zeros(b.n_reads + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in cig])
    return Batch(b.pos, b.flag, b.seq_len, b.chr, b.seq_off, b.seq, off, np.frombuffer(b"".join(cig) + b"\0", np.uint8)[:-1].copy(), b.md_off, b.md)


def implied_ops(rec, edits):
    """Deletions and insertions are stored in M coordinates as deltas against the previous event of their kind
    (src/read_compression.c:321-352): merge the two lists; insertions before deletions at one coordinate."""
    e = edits[int(rec["edit_off"]):]
    nd, ns, ni = int(rec["n_dels"]), int(rec["n_snps"]), int(rec["n_ins"])
    dels = np.cumsum(e[:nd] & 0xff).tolist()
    ins = np.cumsum(e[nd + ns:nd + ns + ni] & 0xff).tolist()
    events = sorted(set(dels) | set(ins))
    ops, cur = [], 0
    for at in events:
        if at > cur:
            ops.append((at - cur, "M")); cur = at
        if at in ins:
            ops.append((ins.count(at), "I"))
        if at in dels:
            ops.append((dels.count(at), "D"))
    total_m = int(rec["len"]) - ni
    if total_m > cur:
        ops.append((total_m - cur, "M"))
    return ops


def classify(text: bytes, ops):
    """0 implied, 1 / 2 / 3 first / last / both end operations are soft clips, 4 verbatim."""
    toks = re.findall(rb"(\d+)([A-Za-z=*])", text)
    if b"".join(a + b for a, b in toks) != text or len(toks) != len(ops) or len(ops) > 48:
        return 4
    c = 0
    for k, ((num, op), (n, want)) in enumerate(zip(toks, ops)):
        if num.startswith(b"0") or int(num) != n:
            return 4
        if op.decode() != want:
            if op == b"S" and want == "I" and (k == 0 or k == len(ops) - 1):
                c |= 1 if k == 0 else 2
            else:
                return 4
    return c


def text_of(ops, c=0):
    out = []
    for k, (n, op) in enumerate(ops):
        if op == "I" and ((k == 0 and c & 1) or (k == len(ops) - 1 and k != 0 and c & 2)):
            op = "S"
        out.append(b"%d%s" % (n, op.encode()))
    return b"".join(out)


def parse_section(sec: bytes):
    assert sec[:4] == b"CBCC" and int.from_bytes(sec[4:8], "little") == 1
    n, entries = int.from_bytes(sec[8:16], "little"), int.from_bytes(sec[16:24], "little")
    o, prev, out = 24, 0, {}

    def varint():
        nonlocal o
        v, sh = 0, 0
        while True:
            c = sec[o]; o += 1
            v |= (c & 0x7f) << sh; sh += 7
            if not c & 0x80:
                return v
    for _ in range(entries):
        prev += varint()
        c = sec[o]; o += 1
        if c == 4:
            ln = varint()
            out[prev] = (4, sec[o:o + ln]); o += ln
        else:
            out[prev] = (c, None)
    assert o == len(sec)
    return n, out


def _check(codec, g, b, L, block_reads=AUTO, gen_mode=1, legacy=False):
    codec.set_reference(g)
    cig = _cigars(b)
    sec = codec.cigar_pack(b)
    cont = codec.compress(b, L, 0 if legacy else block_reads, 0 if legacy else gen_mode, None, 1 if legacy else 0)
    text, n = codec.cigar_unpack(cont, sec, legacy=legacy)
    assert n == b.n_reads
    assert text == b"".join(x + b"\n" for x in cig)                       # the round trip, byte for byte
    # the restatement: classes from the oracle's edit records
    recs, edits = O.extract(b, g)
    n_sec, listed = parse_section(sec)
    assert n_sec == b.n_reads
    want = {}
    for r in range(b.n_reads):
        c = classify(cig[r], implied_ops(recs[r], edits))
        if c:
            want[r] = (c, cig[r] if c == 4 else None)
    assert listed == want
    # without the section: the implied CIGARs alone
    text0, _ = codec.cigar_unpack(cont, None, legacy=legacy)
    assert text0 == b"".join(text_of(implied_ops(recs[r], edits)) + b"\n" for r in range(b.n_reads))
    return sec, listed


def test_match_only_cigars_cost_the_header(codec):
    cfg = synth.SynthConfig.named("config2", scale=0.01)                   # 150 bp, substitutions only: every CIGAR is "150M"
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    sec, listed = _check(codec, g, b, 150)
    assert len(sec) == 24 and not listed


def test_indels_and_soft_clips(codec):
    cfg = synth.SynthConfig.named("config5", scale=0.01)                   # 50-250 bp, 2 % indels, soft clips
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    sec, listed = _check(codec, g, b, 250)
    classes = [c for c, _ in listed.values()]
    assert any(c in (1, 2, 3) for c in classes)                           # clipped reads: a class byte, no text
    total_cigar = int(b.cigar_off[-1])
    assert len(sec) < total_cigar // 4                                     # far below the CIGAR text itself
    cfg = synth.SynthConfig(seed=77, genome_len=200_000, n_reads=20_000, len_min=100, len_max=100, p_sub=0.01, p_indel=0.01)
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    _check(codec, g, b, 100, block_reads=512, gen_mode=0)
    _check(codec, g, b, 100, legacy=True)                                  # beside the reference's own stream too


def test_operations_the_indels_do_not_imply_are_kept_verbatim(codec):
    """=/X spellings, hard clips, a padded number: all decode to the same SEQ, none is what the indels imply."""
    cfg = synth.SynthConfig(seed=78, genome_len=100_000, n_reads=6_000, len_min=100, len_max=100, p_sub=0.01, p_indel=0.0)
    g = synth.make_genome(cfg); b0 = synth.make_reads(cfg, g)
    mds = [b0.md[int(b0.md_off[r]):int(b0.md_off[r + 1])].tobytes() for r in range(b0.n_reads)]

    def eqx(md):
        out = []
        for tok in re.findall(rb"\d+|[A-Z]", md):
            if tok.isdigit():
                if int(tok):
                    out.append(b"%d=" % int(tok))
            else:
                out.append(b"1X")
        return b"".join(out) or b"100="
    cig = []
    for r, c in enumerate(_cigars(b0)):
        k = r % 5
        cig.append(eqx(mds[r]) if k == 1 else b"7H" + c if k == 2 else c + b"3H" if k == 3 else b"0" + c if k == 4 else c)
    b = _with_cigars(b0, cig)
    sec, listed = _check(codec, g, b, 100)
    assert sum(1 for c, _ in listed.values() if c == 4) == sum(1 for r in range(b.n_reads) if r % 5)
    # a damaged section is reported, not decoded
    cont = codec.compress(b, 100, AUTO, 1, None, 0)
    for bad in (sec[:20], b"XXXX" + sec[4:], sec[:8] + (b.n_reads + 1).to_bytes(8, "little") + sec[16:], sec[:-3]):
        with pytest.raises(CbcgError) as e:
            codec.cigar_unpack(cont, bad)
        assert e.value.status == -8
