"""N > 1 host logic on CPU: two gloo ranks shard a position-sorted batch, code their shards (with the
oracle standing in for the GPU coder, which cannot run here), all-gather the index and write one
sharded file with a pwrite each; decoding the shards in order gives back the input."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, path, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from cbc_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.SynthConfig(seed=21, genome_len=300_000, n_reads=6000, len_min=100, len_max=100, p_sub=0.005, p_indel=0.002)
    g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
    r0, r1 = shard.shard_ranges(b.n_reads, world)[rank]
    mine = b.slice(r0, r1)
    cont = O.encode_blocked(mine, g, 100, 256)
    head_len = shard.container_head_len(cont)               # header + names + varint block index: what cbcg_fetch_index returns
    layout = shard.gather_index(cont[:head_len], len(cont) - head_len, dist, torch.device("cpu"))
    assert layout.heads[rank] == cont[:head_len]
    shard.write_shard(path, rank, layout, cont)
    dist.barrier()
    q.put((rank, layout.offsets, layout.total, r0, r1))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_container():
    import oracle_lib as O
    from cbc_b200 import shard, synth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.cbcs")
        port = 29500 + (os.getpid() % 2000)
        procs = [ctx.Process(target=_worker, args=(r, world, port, path, q)) for r in range(world)]
        for p in procs: p.start()
        res = sorted(q.get(timeout=240) for _ in range(world))
        for p in procs:
            p.join(timeout=60); assert p.exitcode == 0
        assert res[0][1] == res[1][1] and res[0][2] == res[1][2]          # same scan on every rank
        with open(path, "rb") as f:
            data = f.read()
        assert len(data) == res[0][2]
        shards = shard.read_shards(data)
        assert len(shards) == world
        cfg = synth.SynthConfig(seed=21, genome_len=300_000, n_reads=6000, len_min=100, len_max=100, p_sub=0.005, p_indel=0.002)
        g = synth.make_genome(cfg); b = synth.make_reads(cfg, g)
        text = b""
        for s in shards:
            t, _ = O.decode_blocked(s, g)
            text += t
        assert text == b.seq_lines()


def test_shard_ranges_cover_input():
    from cbc_b200 import shard
    for n, w in ((10, 3), (0, 2), (7, 8), (1000003, 8)):
        r = shard.shard_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
