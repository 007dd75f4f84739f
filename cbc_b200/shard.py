"""Multi-GPU sharding of the aligned-read coding path.

Position-sorted reads split into contiguous ordinal ranges, one per rank; every rank codes its range
into a standalone "CBCB" container with no communication. The only collective is an all-gather of the
per-shard container lengths and block tables (NCCL over NVLink on the GPU box, gloo in the CPU tests):
every rank then runs the same exclusive scan and knows the byte offset of every shard in the sharded
file ("CBCS": u32 magic, u32 version, u32 n_shards, u32 0, then n_shards x {u64 offset, u64 bytes},
then the shard containers back to back), so ranks write their shard with one pwrite each.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

CBCS_MAGIC = 0x53434243   # "CBCS"
CBCS_VERSION = 1


def container_head_len(container: bytes) -> int:
    """Bytes of a "CBCB" container before its payload: the 40-byte header, the chromosome names, the u32 index length and
    the varint block index itself (what cbcg_fetch_index returns)."""
    n_chr, = struct.unpack_from("<I", container, 28)
    o = 40
    for _ in range(n_chr):
        nl, = struct.unpack_from("<I", container, o)
        o += 4 + nl + ((4 - (nl & 3)) & 3)
    index_bytes, = struct.unpack_from("<I", container, o)
    return o + 4 + index_bytes


def shard_ranges(n_reads: int, world: int) -> List[Tuple[int, int]]:
    """Equal contiguous ordinal ranges [r0, r1) per rank (reads are position-sorted, so these are
    genomic regions). Blocks are cut inside each shard, so shard cuts are block boundaries by construction."""
    cuts = [(n_reads * r) // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


@dataclass
class ShardLayout:
    head_bytes: List[int]      # per rank: container header + block index
    payload_bytes: List[int]   # per rank
    offsets: List[int]         # per rank: byte offset of its container in the sharded file
    total: int                 # sharded file size
    heads: List[bytes]         # every rank's header + block table (the global index)

    def superheader(self) -> bytes:
        out = struct.pack("<IIII", CBCS_MAGIC, CBCS_VERSION, len(self.offsets), 0)
        for o, h, p in zip(self.offsets, self.head_bytes, self.payload_bytes):
            out += struct.pack("<QQ", o, h + p)
        return out


def layout_from_sizes(head_bytes: List[int], payload_bytes: List[int], heads: List[bytes]) -> ShardLayout:
    world = len(head_bytes)
    off = 16 + 16 * world
    offsets = []
    for h, p in zip(head_bytes, payload_bytes):
        offsets.append(off)
        off += h + p
    return ShardLayout(list(head_bytes), list(payload_bytes), offsets, off, heads)


def gather_index(head: bytes, payload_bytes: int, dist, device) -> ShardLayout:
    """All-gather of (index length, payload length) and of the block tables themselves."""
    import torch
    world = dist.get_world_size()
    mine = torch.tensor([len(head), int(payload_bytes)], dtype=torch.int64, device=device)
    sizes = torch.empty(2 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, mine)
    sizes = sizes.view(world, 2).cpu().tolist()
    hb = [int(s[0]) for s in sizes]
    pb = [int(s[1]) for s in sizes]
    width = (max(hb) + 15) & ~15
    buf = torch.zeros(width, dtype=torch.uint8)
    buf[:len(head)] = torch.frombuffer(bytearray(head), dtype=torch.uint8)
    buf = buf.to(device)
    allh = torch.empty(width * world, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allh, buf)
    allh = allh.cpu().numpy().reshape(world, width)
    heads = [allh[r, :hb[r]].tobytes() for r in range(world)]
    return layout_from_sizes(hb, pb, heads)


def write_shard(path: str, rank: int, layout: ShardLayout, container: bytes) -> None:
    """Each rank writes its own shard at its scanned offset; rank 0 also writes the super-header."""
    assert len(container) == layout.head_bytes[rank] + layout.payload_bytes[rank]
    fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o644)
    try:
        if rank == 0:
            os.pwrite(fd, layout.superheader(), 0)
        os.pwrite(fd, container, layout.offsets[rank])
    finally:
        os.close(fd)


def read_shards(data: bytes) -> List[bytes]:
    magic, ver, n, _ = struct.unpack_from("<IIII", data, 0)
    if magic != CBCS_MAGIC or ver != CBCS_VERSION:
        raise ValueError("not a CBCS file")
    out = []
    for r in range(n):
        off, size = struct.unpack_from("<QQ", data, 16 + 16 * r)
        if off + size > len(data):
            raise ValueError("truncated CBCS file")
        out.append(data[off:off + size])
    return out
