"""Resident round trips with K contexts in flight on ONE GPU (a host thread and a set of streams each, every context
with its own copy of the batch): does the work of one batch fill the slots the other's narrow early generations and
merges leave idle? Prints reads/s of the whole device per K. usage: inflight_probe.py [K list] [steps] (under gpurun)"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                     # noqa: E402
from cbc_b200 import synth                      # noqa: E402
from cbc_b200.codec import Codec, pin_batch     # noqa: E402

Ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 3]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
name = os.environ.get("CBC_CONFIG", "config2")
cfg = synth.SynthConfig.named(name, scale=float(os.environ.get("CBC_SCALE", "1")))
L_HDR = {"config1": 100, "config5": 250}.get(name, 150)
g = synth.make_genome(cfg)
b = synth.make_reads(cfg, g)
pb = pin_batch(b)
ref = b.seq_lines()
codecs = []
for K in Ks:
    while len(codecs) < K:
        c = Codec(0)
        c.set_reference(g)
        c.upload(pb)
        codecs.append(c)

    def loop(c, n):
        for _ in range(n):
            c.encode_resident(L_HDR, 0xffffffff, 1, 0)
            c.fetch_index()
            c.decode_resident()

    def run(n):
        th = [threading.Thread(target=loop, args=(codecs[k], n)) for k in range(K)]
        for t in th: t.start()
        for t in th: t.join()

    run(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    ok = all(c.fetch_decoded().tobytes() == ref for c in codecs[:K])
    print(json.dumps({"config": name, "contexts": K, "steps_each": steps, "ok": ok, "event_ms": round(ms, 3), "wall_ms": round(wall, 3),
                      "ms_per_batch": round(ms / (K * steps), 3), "reads_per_s": round(b.n_reads * K * steps / (ms * 1e-3))}), flush=True)
