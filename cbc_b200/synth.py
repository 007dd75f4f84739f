"""ctypes front end of the synthetic workload generator (csrc/host/synth.c)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .batch import Batch, Genome

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_uint64), ("n_chr", C.c_uint32),
                ("len_min", C.c_uint32), ("len_max", C.c_uint32),
                ("p_sub", C.c_double), ("p_indel", C.c_double), ("p_clip", C.c_double),
                ("p_rev", C.c_double), ("p_n", C.c_double),
                ("flag_mode", C.c_uint32), ("avoid_b3", C.c_uint32)]


class _Out(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("reads_cap", C.c_uint64),
                ("pos", C.c_void_p), ("flag", C.c_void_p), ("seq_len", C.c_void_p), ("chr", C.c_void_p),
                ("seq_off", C.c_void_p), ("seq", C.c_void_p), ("seq_size", C.c_uint64), ("seq_cap", C.c_uint64),
                ("cigar_off", C.c_void_p), ("cigar", C.c_void_p), ("cigar_size", C.c_uint64), ("cigar_cap", C.c_uint64),
                ("md_off", C.c_void_p), ("md", C.c_void_p), ("md_size", C.c_uint64), ("md_cap", C.c_uint64)]


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libcbcsynth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _LIB = C.CDLL(path)
        _LIB.cbcs_genome.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64]
        _LIB.cbcs_genome.restype = None
        _LIB.cbcs_reads.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.POINTER(_Out)]
        _LIB.cbcs_reads_range.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(_Out)]
        _LIB.cbcs_write_fasta.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.cbcs_write_sam.argtypes = [C.c_char_p, C.POINTER(_Out), C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    return _LIB


@dataclass
class SynthConfig:
    seed: int = 42
    genome_len: int = 1_000_000      # total over chromosomes
    n_chr: int = 1
    n_reads: int = 10_000
    len_min: int = 100
    len_max: int = 100
    p_sub: float = 0.005
    p_indel: float = 0.0
    p_clip: float = 0.0
    p_rev: float = 0.5
    p_n: float = 0.0
    flag_mode: int = 0
    avoid_b3: int = 1

    # The named shapes of BASELINE.json. `scale` shrinks genome and read count together.
    @staticmethod
    def named(name: str, scale: float = 1.0) -> "SynthConfig":
        g1 = 15_072_423
        if name == "config1":      # 1 M x 100 bp, 0.5 % sub, 0.1 % indel
            return SynthConfig(seed=42, genome_len=int(g1 * scale), n_reads=int(1_000_000 * scale),
                               len_min=100, len_max=100, p_sub=0.005, p_indel=0.001)
        if name == "config2":      # 30x 150 bp over CHROMOSOME_I size, 0.5 % sub
            g = int(g1 * scale)
            return SynthConfig(seed=77, genome_len=g, n_reads=g * 30 // 150, len_min=150, len_max=150, p_sub=0.005)
        if name == "config3":      # 64 Mbp, 30x 150 bp
            g = int(64_000_000 * scale)
            return SynthConfig(seed=78, genome_len=g, n_reads=g * 30 // 150, len_min=150, len_max=150, p_sub=0.005)
        if name == "config4":      # GRCh38-shaped: 24 records, ~600 M reads at scale 1
            g = int(3_100_000_000 * scale)
            return SynthConfig(seed=79, genome_len=g, n_chr=24, n_reads=int(600_000_000 * scale),
                               len_min=150, len_max=150, p_sub=0.005)
        if name == "config5":      # indel-heavy variable length with soft clips
            g = int(g1 * scale)
            return SynthConfig(seed=80, genome_len=g, n_reads=g * 30 // 150, len_min=50, len_max=250,
                               p_sub=0.005, p_indel=0.02, p_clip=0.3)
        raise ValueError(name)


_ROMAN = ["I", "II", "III", "IV", "V", "VI", "VII", "VIII", "IX", "X", "XI", "XII", "XIII", "XIV", "XV", "XVI",
          "XVII", "XVIII", "XIX", "XX", "XXI", "XXII", "XXIII", "XXIV"]


def make_genome(cfg: SynthConfig) -> Genome:
    lib = _lib()
    per = cfg.genome_len // cfg.n_chr
    names, bases = [], []
    for c in range(cfg.n_chr):
        n = per if c + 1 < cfg.n_chr else cfg.genome_len - per * (cfg.n_chr - 1)
        a = np.empty(n, dtype=np.uint8)
        lib.cbcs_genome(cfg.seed, c, a.ctypes.data, n)
        names.append("chr" + (_ROMAN[c] if c < len(_ROMAN) else str(c + 1)))
        bases.append(a)
    return Genome(names, bases)


def _out_struct(n, seq_cap, cig_cap, md_cap):
    arrs = dict(pos=np.empty(n, np.uint32), flag=np.empty(n, np.uint16), seq_len=np.empty(n, np.uint16),
                chr=np.empty(n, np.uint32), seq_off=np.zeros(n + 1, np.uint64), seq=np.empty(seq_cap, np.uint8),
                cigar_off=np.zeros(n + 1, np.uint64), cigar=np.empty(cig_cap, np.uint8),
                md_off=np.zeros(n + 1, np.uint64), md=np.empty(md_cap, np.uint8))
    o = _Out()
    o.n_reads, o.reads_cap = 0, n
    for k, a in arrs.items():
        setattr(o, k, a.ctypes.data)
    o.seq_cap, o.cigar_cap, o.md_cap = seq_cap, cig_cap, md_cap
    return o, arrs


def make_reads(cfg: SynthConfig, genome: Genome, r0: int = 0, r1: int | None = None) -> Batch:
    """Reads [r0, r1) of the position-sorted input `cfg` describes (default: all of it). A sub-range holds exactly the
    reads the whole input holds there: every read has its own RNG stream keyed by (seed, ordinal)."""
    lib = _lib()
    r1 = cfg.n_reads if r1 is None else min(r1, cfg.n_reads)
    n = max(r1 - r0, 0)
    per_read = 16 + int(cfg.len_max * (4 * cfg.p_indel + 3 * (cfg.p_sub + cfg.p_n)) * 4) + (8 if cfg.p_clip else 0)
    o, arrs = _out_struct(n, n * cfg.len_max + 64, n * per_read + 4096, n * per_read + 4096)
    p = _Params(cfg.seed, cfg.n_reads, cfg.n_chr, cfg.len_min, cfg.len_max, cfg.p_sub, cfg.p_indel, cfg.p_clip,
                cfg.p_rev, cfg.p_n, cfg.flag_mode, cfg.avoid_b3)
    ptrs, lens, _ = genome.c_arrays()
    rc = lib.cbcs_reads_range(C.byref(p), ptrs, lens, r0, r1, C.byref(o))
    if rc:
        raise RuntimeError(f"cbcs_reads failed: {rc}")
    assert o.n_reads == n
    return Batch(arrs["pos"], arrs["flag"], arrs["seq_len"], arrs["chr"],
                 arrs["seq_off"], arrs["seq"][:o.seq_size].copy(),
                 arrs["cigar_off"], arrs["cigar"][:o.cigar_size].copy(),
                 arrs["md_off"], arrs["md"][:o.md_size].copy())


def write_fasta(path: str, genome: Genome) -> None:
    ptrs, lens, names = genome.c_arrays()
    if _lib().cbcs_write_fasta(path.encode(), genome.n_chr, names, ptrs, lens):
        raise OSError(path)


def write_sam(path: str, batch: Batch, genome: Genome, header: bool = True) -> None:
    o = _Out()
    o.n_reads = o.reads_cap = batch.n_reads
    for k in ("pos", "flag", "seq_len", "chr", "seq_off", "seq", "cigar_off", "cigar", "md_off", "md"):
        setattr(o, k, getattr(batch, k).ctypes.data)
    _, lens, names = genome.c_arrays()
    if _lib().cbcs_write_sam(path.encode(), C.byref(o), genome.n_chr, names, lens, int(header)):
        raise OSError(path)
