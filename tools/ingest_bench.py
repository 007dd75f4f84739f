"""Host SAM ingest throughput (cbch_read_sam_mt, cbc_b200/csrc/host/sam_ingest.c) against the worker count.
usage: ingest_bench.py [n_reads] [threads ...]   (CPU only)"""
import ctypes as C
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cbc_b200 import synth                                  # noqa: E402
from test_host_ingest import Fasta, HBatch                  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
threads = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8, 16]
lib = C.CDLL(os.path.join(ROOT, "cbc_b200", "_build", "libcbchost.so"))
lib.cbch_read_fasta.argtypes = [C.c_char_p, C.POINTER(Fasta), C.c_char_p, C.c_size_t]
lib.cbch_read_sam_mt.argtypes = [C.c_char_p, C.POINTER(Fasta), C.c_int, C.c_int, C.POINTER(HBatch), C.c_char_p, C.c_size_t]
lib.cbch_free_batch.argtypes = [C.POINTER(HBatch)]
cfg = synth.SynthConfig(seed=42, genome_len=15_072_423, n_reads=n_reads, len_min=100, len_max=100, p_sub=0.005, p_indel=0.001, p_clip=0.0)
g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
with tempfile.TemporaryDirectory() as d:
    fa, sam = os.path.join(d, "r.fa"), os.path.join(d, "r.sam")
    synth.write_fasta(fa, g); synth.write_sam(sam, b, g)
    size = os.path.getsize(sam)
    f = Fasta(); err = C.create_string_buffer(256)
    assert lib.cbch_read_fasta(fa.encode(), C.byref(f), err, 256) == 0
    for T in threads:
        best = 1e9
        for _ in range(3):
            hb = HBatch()
            t0 = time.perf_counter()
            rc = lib.cbch_read_sam_mt(sam.encode(), C.byref(f), 0, T, C.byref(hb), err, 256)
            dt = time.perf_counter() - t0
            assert rc == 0 and hb.n_reads == n_reads
            lib.cbch_free_batch(C.byref(hb))
            best = min(best, dt)
        print(f"threads {T:2d}: {best * 1e3:8.1f} ms  {size / best / 1e6:8.0f} MB/s  {n_reads / best / 1e6:6.2f} M reads/s  ({size / 1e6:.0f} MB of SAM text, page cache warm)")
