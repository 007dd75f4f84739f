#!/usr/bin/env python
"""bench.py -- compress + decompress throughput of the aligned-read coding path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--block-reads R]

A step is one pass of the hot path over one batch of synthetic reads: compress the batch (K1 edit
extraction -> K2 block coder -> container index), then decompress it (K2 block decoder -> K3 read
reconstruction). `value` times the step with the batch already resident in HBM (CUDA events on the
library's own stream); `e2e` times the same step through the host-buffer C-ABI calls (cbcg_encode /
cbcg_decode) with pinned host buffers, host<->device copies inside the timed region.

N = 1: BASELINE.json configs[1] (150 bp reads at 30x over a 15.07 Mbp chromosome, 0.5 % substitutions).
N > 1 (torchrun): every rank codes its own config-2-sized genomic region (weak scaling, no collective
on the coding path) and the ranks all-gather their block-length tables over NCCL for the container index.

--impl reference: the UNMODIFIED reference encoder/decoder (oracle/_ref/cbc_ref, built from
/root/reference by oracle/Makefile) on the host CPU, one thread (it has no threading), on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from cbc_b200 import synth                                    # noqa: E402

METRIC = "compress+decompress round-trip reads/s"
UNIT = "reads/s"
CPU_SAMPLE_READS = 500_000


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload(rank: int, scale: float):
    cfg = synth.SynthConfig.named("config2", scale=scale)
    cfg.seed += 1000 * rank                                   # every rank: its own region of a larger genome
    g = synth.make_genome(cfg)
    b = synth.make_reads(cfg, g)
    return cfg, g, b


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "utilization.gpu")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, busy, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
                busy.append(float(f[9]) if len(f) > 9 else 100.0)
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        loaded = [c for c, u in zip(sm, busy) if u >= 10.0] or sm       # samples taken while the GPU was working
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "samples_under_load": len(loaded) if len(loaded) != len(sm) or busy and min(busy) >= 10.0 else 0}


# ---------------------------------------------------------------------------------------------- CPU reference

def cpu_reference_once(sample, genome, workdir, write_inputs=True):
    """One encode + decode of `sample` by oracle/_ref/cbc_ref. Returns (enc_s, dec_s, stream_bytes)."""
    import oracle_lib as O
    fa, sam = os.path.join(workdir, "r.fa"), os.path.join(workdir, "r.sam")
    if write_inputs:
        synth.write_fasta(fa, genome)
        synth.write_sam(sam, sample, genome)
    stream, _, enc_s = O.run_reference(sam, fa, workdir)
    decoded, dec_s = O.run_reference_decode(os.path.join(workdir, "ref.cbc"), fa, workdir)
    if decoded != sample.seq_lines():
        raise RuntimeError("reference decoder output != input SEQ")
    return enc_s, dec_s, len(stream), stream


def cpu_port_once(sample, genome):
    """Fallback when oracle/_ref is absent: the plain-C restatement (single stream)."""
    import oracle_lib as O
    t0 = time.perf_counter()
    stream, _ = O.encode_legacy(sample, genome, int(sample.seq_len[1] if sample.n_reads > 1 else sample.seq_len[0]))
    t1 = time.perf_counter()
    decoded, _ = O.decode_legacy(stream, genome)
    t2 = time.perf_counter()
    if decoded != sample.seq_lines():
        raise RuntimeError("oracle decode != input SEQ")
    return t1 - t0, t2 - t1, len(stream), stream


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_arm(args):
    import oracle_lib as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, g, b = workload(0, args.scale)
    n = min(CPU_SAMPLE_READS, b.n_reads)
    sample = b.slice(0, n)
    kind = "reference" if O.have_reference() else "port"
    times = []
    with tempfile.TemporaryDirectory() as d:
        first = True
        for _ in range(args.warmup + args.steps):
            if kind == "reference":
                e, dd, sz, _ = cpu_reference_once(sample, g, d, write_inputs=first)
            else:
                e, dd, sz, _ = cpu_port_once(sample, g)
            first = False
            times.append((e, dd))
    timed = times[args.warmup:]
    step_s = float(np.mean([e + dd for e, dd in timed]))
    value = n / step_s
    desc = (f"first {n} reads of the workload through oracle/_ref/cbc_ref -c 1 / -x (program's own clock() lines)"
            if kind == "reference" else f"first {n} reads through the C restatement oracle/cbc_oracle.c")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"config2: 150bp reads at 30x over {cfg.genome_len} bp, 0.5% substitutions (scale {args.scale})",
                   "n_reads_per_step": n, "read_len": 150},
        "compress_reads_per_s": n / float(np.mean([e for e, _ in timed])),
        "decompress_reads_per_s": n / float(np.mean([dd for _, dd in timed])),
        "bits_per_base": 8.0 * sz / sample.total_bases(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": desc, "cpu": cpu_model(),
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------- B200 arm

def run_b200_arm(args):
    import torch
    from cbc_b200.codec import Codec, pin_batch, pinned_empty
    from cbc_b200 import shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner goes to stdout otherwise: stdout is the JSON line
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU implementation of this path")
    dev = torch.device("cuda", local)

    cfg, g, b = workload(rank, args.scale)
    L = 150
    R = args.block_reads
    G = args.gen_mode
    codec = Codec(local)
    codec.set_reference(g)
    pb = pin_batch(b)
    n = b.n_reads
    bases = b.total_bases()

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- device-resident step (value)
    codec.upload(pb)
    stage = {k: [] for k in ("k1", "plan", "k2e", "gather", "k2d", "k3", "enc_total", "dec_total")}
    launches = 0
    index_bytes = 0
    block_reads_used = R

    def resident_step(record: bool):
        nonlocal launches, index_bytes, block_reads_used
        codec.encode_resident(L, R, G)
        se = codec.stats()
        head, payload = codec.fetch_index()
        if dist is not None:                                  # container index: all-gather of per-shard block tables
            shard.gather_index(head, payload, dist, dev)
        codec.decode_resident()
        sd = codec.stats()
        if record:
            stage["k1"].append(se["ms_k1"]); stage["plan"].append(se["ms_plan"]); stage["k2e"].append(se["ms_code"])
            stage["gather"].append(se["ms_gather"]); stage["enc_total"].append(se["ms_total"])
            stage["k2d"].append(sd["ms_code"]); stage["k3"].append(sd["ms_k3"]); stage["dec_total"].append(sd["ms_total"])
            launches += se["kernel_launches"] + sd["kernel_launches"]
        index_bytes = len(head)
        block_reads_used = int.from_bytes(head[32:36], "little")
        return se, sd

    # nvidia-smi needs ~0.1 s to come up and the timed region of the resident leg is ~0.1 s long: the sampler starts
    # before the warm-up steps (the same work) and runs until the end of the e2e leg, 20 ms apart
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        resident_step(False)
    barrier()
    codec.mark(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        se, sd = resident_step(True)
    codec.mark(1)
    dev_ms = codec.elapsed_ms(0, 1)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    step_ms = max_over_ranks(dev_ms) / args.steps
    total_reads = sum_over_ranks(float(n))
    value = total_reads / (step_ms * 1e-3)

    # correctness of what was timed: decoded text == input SEQ (every rank)
    text = codec.fetch_decoded()
    if text.tobytes() != b.seq_lines():
        raise RuntimeError("round trip mismatch: decoded reads != input SEQ")
    container_bytes = se["container_bytes"]
    n_edits = se["n_edits"]
    # K1 alone, for its roofline entry: in the timed steps above the tail of K1 runs on a side stream beside the early
    # generations of the block coder (api.cu, encode_resident_overlapped), which stretches its own launch time;
    # CBCG_NO_OVERLAP=1 is the one-stream order (same container). Outside the timed region.
    os.environ["CBCG_NO_OVERLAP"] = "1"
    k1_alone = []
    for _ in range(3):
        codec.encode_resident(L, R, G)
        k1_alone.append(codec.stats()["ms_k1"])
    del os.environ["CBCG_NO_OVERLAP"]
    if codec.stats()["container_bytes"] != container_bytes:
        raise RuntimeError("one-stream and overlapped resident encodes disagree")

    # ---------------- end-to-end step through the host-buffer C ABI (e2e): pinned host buffers in, pinned host buffers
    # out, every copy inside the timed region. The batch goes through `--inflight` contexts (one host thread and one
    # CUDA stream each, the ABI's threading model): sub-batch k+1 is on the PCIe link while sub-batch k is being coded,
    # which is how a streaming caller keeps both busy. Each sub-batch is a self-contained container (a shard).
    from concurrent.futures import ThreadPoolExecutor
    K = max(1, args.inflight)
    cuts = shard.shard_ranges(n, K)
    codecs = [codec] + [Codec(local) for _ in range(K - 1)]
    for c2 in codecs[1:]:
        c2.set_reference(g)
    subs = [pin_batch(b.slice(r0_, r1_)) if K > 1 else pb for r0_, r1_ in cuts]
    outs_c = [pinned_empty(int(container_bytes * 1.5 / K) + 65536, np.uint8) for _ in range(K)]
    outs_t = [pinned_empty(sb_.total_bases() + sb_.n_reads + 64, np.uint8) for sb_ in subs]

    def e2e_one(k):
        c2 = codecs[k]
        nc = c2.compress_into(subs[k], L, R, outs_c[k], G)
        s1 = c2.stats()
        head, payload = c2.fetch_index()
        nt, nr = c2.decompress_into(outs_c[k][:nc], outs_t[k])
        s2 = c2.stats()
        return nc, nt, s1, s2, head, payload

    pool = ThreadPoolExecutor(K)

    def e2e_step():
        res = list(pool.map(e2e_one, range(K)))
        if dist is not None:                                  # container index across ranks (first sub-batch's table stands for the shard)
            shard.gather_index(res[0][4], sum(r_[5] for r_ in res), dist, dev)
        return res

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    clocks = sampler.stop()
    text_all = b"".join(outs_t[k][:res[k][1]].tobytes() for k in range(K))
    if text_all != b.seq_lines():
        raise RuntimeError("e2e round trip mismatch")
    e2e_container = sum(r_[0] for r_ in res)
    e2e = {"value": total_reads / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(sum(r_[2]["h2d_bytes"] + r_[3]["h2d_bytes"] for r_ in res)),
           "d2h_bytes_per_step": int(sum(r_[2]["d2h_bytes"] + r_[3]["d2h_bytes"] for r_ in res)),
           "contexts_in_flight": K, "container_bytes": int(e2e_container),
           "bits_per_base": 8.0 * e2e_container / bases}
    s1 = {"h2d_bytes": sum(r_[2]["h2d_bytes"] for r_ in res)}
    pool.shutdown()
    for c2 in codecs[1:]:
        c2.close()

    # ---------------- roofline (SURVEY.md 8d figures, DESIGN.md "Measurement")
    peak, peak_src = peaks()
    cov = bases / max(cfg.genome_len, 1)
    C_ = float(b.cigar_off[-1]) / n
    D_ = float(b.md_off[-1]) / n
    E_ = n_edits / n
    k1_bytes = (L + C_ + D_ + 24 + L / cov + 12 + 4 * E_) * n
    k3_bytes = (12 + 4 * E_ + L / cov + (L + 1)) * n
    syms = se["n_symbols"]
    k2_bytes = 4.0 * syms + se["payload_bytes"]
    med = {k: float(np.median(v)) for k, v in stage.items()}
    med["k1_alone"] = float(np.median(k1_alone))

    def roof(name, alg_bytes, ms):
        a = alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"kernel": name, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                "traffic": None, "ms": ms, "algorithmic_bytes": alg_bytes}
    kernels = [roof("k1_extract_kernel", k1_bytes, med["k1_alone"]), roof("k2_coder_kernel<encode>", k2_bytes, med["k2e"]),
               roof("k2_coder_kernel<decode>", k2_bytes, med["k2d"]), roof("k3_reconstruct_kernel", k3_bytes, med["k3"])]
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        with open(tr_path) as f:
            tr = json.load(f)
        for k in kernels:
            if k["kernel"] in tr and abs(cfg.n_reads - 3014484) < 10 and world == 1:   # captured on this exact workload
                k["traffic"] = tr[k["kernel"]]
    dominant = max(kernels, key=lambda k: k["ms"])
    roofline = {k: dominant[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
    roofline.update({"kernel": dominant["kernel"], "peak_source": peak_src,
                     "note": "the block coder is serial integer work per block (latency-bound); K1/K3 are the HBM-bound kernels"})

    # ---------------- CPU baseline + bits/base overhead on a bounded sample (rank 0, N = 1 only)
    cpu_baseline, overhead = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_lib as O
        ns = min(CPU_SAMPLE_READS, n)
        sample = b.slice(0, ns)
        with tempfile.TemporaryDirectory() as d:
            if O.have_reference():
                enc_s, dec_s, sz, ref_stream = cpu_reference_once(sample, g, d)
                kind = "reference"
                desc = f"first {ns} reads of the workload through oracle/_ref/cbc_ref -c 1 / -x, program's own clock() lines"
            else:
                enc_s, dec_s, sz, ref_stream = cpu_port_once(sample, g)
                kind = "port"
                desc = f"first {ns} reads of the workload through oracle/cbc_oracle.c"
        cpu_baseline = {"value": ns / (enc_s + dec_s), "unit": UNIT, "cores": 1, "kind": kind, "sample": desc,
                        "compress_reads_per_s": ns / enc_s, "decompress_reads_per_s": ns / dec_s,
                        "bits_per_base": 8.0 * sz / sample.total_bases(), "cpu": cpu_model(), "host_cores": os.cpu_count()}
        # single-block mode on the sample must be the reference's bytes (parity definition 2, at scale)
        single = codec.compress(sample, L, 0)
        # blocking overhead on the WHOLE workload: blocked container vs the reference's single stream, whose size
        # comes from the CPU restatement (pinned byte for byte to cbc_ref; ~3 s for 3 M reads, outside any timed region)
        full_single, _ = O.encode_legacy(b, g, L)
        overhead = {"single_stream_bytes": len(full_single), "blocked_bytes": int(container_bytes),
                    "single_bits_per_base": 8.0 * len(full_single) / bases,
                    "blocked_bits_per_base": 8.0 * container_bytes / bases,
                    "overhead_pct": 100.0 * (container_bytes - len(full_single)) / len(full_single),
                    "sample_reads": ns, "sample_single_stream_bytes": len(single),
                    "single_stream_byte_identical_to_reference": bool(single == ref_stream)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": f"config2: 150bp reads at 30x over {cfg.genome_len} bp, 0.5% substitutions, one such region per GPU",
                   "n_reads_per_gpu": n, "read_len": 150, "block_reads": block_reads_used, "block_reads_auto": R == 0xffffffff, "gen_mode": G, "blocks_per_gpu": int(se["n_blocks"]),
                   "l2": "inputs larger than L2 (batch %.0f MB, decoded text %.0f MB per GPU)" % (s1["h2d_bytes"] / 1e6, (bases + n) / 1e6),
                   "parallelism": f"{world} region shard(s), no collective on the coding path"},
        "compress_reads_per_s": total_reads / (max_over_ranks(med["enc_total"]) * 1e-3),
        "decompress_reads_per_s": total_reads / (max_over_ranks(med["dec_total"]) * 1e-3),
        "bits_per_base": 8.0 * container_bytes / bases,
        "blocking": overhead,
        "stage_ms": med, "wall_ms_per_step": wall_ms / args.steps,
        "roofline": roofline, "roofline_kernels": kernels,
        "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
        "symbols_per_s_encode": syms / (med["k2e"] * 1e-3) if med["k2e"] > 0 else None,
        "symbols_per_s_decode": syms / (med["k2d"] * 1e-3) if med["k2d"] > 0 else None,
        "index_bytes_per_gpu": index_bytes,
    }
    if rank == 0:
        emit(json.dumps(line))
    codec.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = None


def guard_stdout():
    """stdout carries exactly one JSON line. Libraries write there behind Python's back (NCCL's version banner did, under
    torchrun, NCCL_DEBUG_FILE notwithstanding): file descriptor 1 is pointed at stderr for the rest of the process and
    the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: str):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--block-reads", type=int, default=0xffffffff, help="reads per block; default: sized to whole waves (CBCG_BLOCK_AUTO)")
    ap.add_argument("--inflight", type=int, default=1, help="contexts (host threads / streams) the e2e leg keeps in flight")
    ap.add_argument("--gen-mode", type=int, default=1, help="1: generation-primed blocks (default), 0: cold blocks")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (tests only; the bench line needs 1.0)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    subprocess.run(["make", "-s", "-C", ROOT, "host"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
