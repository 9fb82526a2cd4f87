"""Multi-GPU plumbing for the block codec: one process per GPU, blocks sharded by contiguous ranges.

Blocks are independent (fresh hash table per block, reference tsq_threads.cpp:176; no cross-block
references; the TSQ1 container is a concatenation, turbosqueeze.cpp:64-84), so the hot path needs no
collective: rank r encodes / decodes its own contiguous block range.  The only exchange step is the
assembly of ONE container from the per-rank streams (north_star: "NCCL over NVLink only to gather
the per-GPU output streams"): an all-gather of the per-rank byte counts, then a variable-length
gather to the destination rank at the prefix-summed offsets.  torch.distributed is the transport
(NCCL on the GPU box, gloo in the CPU tests); tensors stay on whatever device they are on.

Contiguous ranges keep the reference's input over-read local: the encoder of block b reads up to 19
bytes past the block (tsq_encode.cpp:74,126-128), i.e. into block b+1, so a rank needs its shard
plus INPUT_PAD bytes of the following shard (zeros after the very last block).
"""
import struct

import torch
import torch.distributed as dist

INPUT_PAD = 128
HEADER = 16          # "TSQ1" | n_blocks u32 | total_uncompressed u64 (turbosqueeze.cpp:64-67)


def block_range(n_blocks, rank, world):
    """Blocks [lo, hi) of rank `rank`: floor(r*n/W) .. floor((r+1)*n/W)."""
    return rank * n_blocks // world, (rank + 1) * n_blocks // world


def byte_range(total, block, rank, world):
    """(lo, hi, hi_with_tail): the rank's input bytes and how far its encoder may read."""
    nb = (total + block - 1) // block
    b0, b1 = block_range(nb, rank, world)
    lo, hi = min(total, b0 * block), min(total, b1 * block)
    return lo, hi, min(total, hi + INPUT_PAD)


def container_header(n_blocks, total_uncompressed):
    return b"TSQ1" + struct.pack("<IQ", n_blocks, total_uncompressed)


def gather_bodies(body, dst=0, group=None):
    """Variable-length gather of one uint8 tensor per rank to `dst`, in rank order.

    body: 1-D uint8 tensor (this rank's container body: u24 length prefixes + block streams).
    Returns (concatenated tensor on dst | None elsewhere, list of per-rank byte counts).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = body.device
    n = torch.tensor([body.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    if rank == dst:
        out = torch.empty(sum(counts), dtype=torch.uint8, device=dev)
        at, ops = 0, []
        for r, c in enumerate(counts):
            if r == dst:
                out[at:at + c].copy_(body)
            elif c:
                ops.append(dist.P2POp(dist.irecv, out[at:at + c], r, group))
            at += c
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return out, counts
    if body.numel():
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, body.contiguous(), dst, group)]):
            w.wait()
    return None, counts


def gather_container(local_container, total_uncompressed, n_blocks_total, dst=0, group=None):
    """Assemble one TSQ1 container on `dst` from per-rank containers (each rank's own header is dropped).

    local_container: uint8 tensor holding a TSQ1 container of this rank's blocks (what
    Context.pack_container / tsqb_pack_container produce).  Returns the uint8 tensor on dst, else None.
    """
    body, _ = gather_bodies(local_container[HEADER:], dst=dst, group=group)
    if body is None:
        return None
    hdr = torch.frombuffer(bytearray(container_header(n_blocks_total, total_uncompressed)), dtype=torch.uint8).to(body.device)
    return torch.cat([hdr, body])


def all_gather_decoded(local_out, group=None):
    """Decoded shards back into one buffer on every rank (shards are contiguous block ranges, so this
    is a plain variable-count all-gather)."""
    world = dist.get_world_size(group)
    dev = local_out.device
    n = torch.tensor([local_out.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(counts) if counts else 0
    padded = torch.zeros(cap, dtype=torch.uint8, device=dev)
    padded[: local_out.numel()].copy_(local_out)
    parts = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])
