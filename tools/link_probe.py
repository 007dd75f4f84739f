"""Host <-> device link ceiling with N ranks at once (run under torchrun on the GPU box).

Every rank moves what one end-to-end step of config 2 moves -- plain pinned cudaMemcpyAsync, nothing else -- all ranks
at the same time: H2D alone, D2H alone, both directions at once. The per-rank GB/s is the ceiling any host-buffer
path can reach with N GPUs sharing the host's memory and PCIe root complexes; bench.py's `e2e` is to be read against it.
    python -m torch.distributed.run --nproc-per-node N tools/link_probe.py [h2d_MB d2h_MB]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

h2d_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 593.0
d2h_mb = float(sys.argv[2]) if len(sys.argv) > 2 else 461.0
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
src = torch.empty(int(h2d_mb * 1e6), dtype=torch.uint8).pin_memory(); dst_d = torch.empty_like(src, device=dev)
out_d = torch.empty(int(d2h_mb * 1e6), dtype=torch.uint8, device=dev); out_h = torch.empty(int(d2h_mb * 1e6), dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = None
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        s1.synchronize(); s2.synchronize()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best


def h2d():
    with torch.cuda.stream(s1):
        dst_d.copy_(src, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        out_h.copy_(out_d, non_blocking=True)


def both():
    h2d(); d2h()


res = {"h2d_ms": timed(h2d), "d2h_ms": timed(d2h), "both_ms": timed(both)}
res["h2d_gbs"] = h2d_mb / res["h2d_ms"]; res["d2h_gbs"] = d2h_mb / res["d2h_ms"]; res["both_gbs"] = (h2d_mb + d2h_mb) / res["both_ms"]
if world > 1:
    allr = [None] * world
    dist.all_gather_object(allr, res)
else:
    allr = [res]
if rank == 0:
    print(json.dumps({"n_ranks": world, "h2d_mb": h2d_mb, "d2h_mb": d2h_mb,
                      "per_rank": allr,
                      "slowest": {k: max(r[k] for r in allr) for k in ("h2d_ms", "d2h_ms", "both_ms")},
                      "aggregate_gbs": {"h2d": world * h2d_mb / max(r["h2d_ms"] for r in allr), "d2h": world * d2h_mb / max(r["d2h_ms"] for r in allr),
                                        "both": world * (h2d_mb + d2h_mb) / max(r["both_ms"] for r in allr)}}))
if world > 1:
    dist.destroy_process_group()
