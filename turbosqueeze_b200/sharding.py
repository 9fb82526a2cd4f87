"""Multi-GPU plumbing for the block codec: one process per GPU, blocks sharded by contiguous ranges.

Blocks are independent (fresh hash table per block, reference tsq_threads.cpp:176; no cross-block
references; the TSQ1 container is a concatenation, turbosqueeze.cpp:64-84), so the hot path needs no
collective: rank r encodes / decodes its own contiguous block range.  The only exchange step is the
assembly of ONE container from the per-rank streams (north_star: "NCCL over NVLink only to gather
the per-GPU output streams"): an all-gather of the per-rank byte counts, then a variable-length
gather to the destination rank at the prefix-summed offsets -- on GPUs by direct peer-memory writes
over NVLink / NVSwitch into the destination's buffer (CUDA IPC mapping, gather_bodies_peer), with
grouped NCCL send/recv as the fallback; gloo send/recv in the CPU tests.  torch.distributed is the
plumbing; tensors stay on whatever device they are on.

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


def gather_bodies(body, dst=0, group=None, length=None, out=None, out_offset=0):
    """Variable-length gather of one uint8 tensor per rank to `dst`, in rank order.

    body:   1-D uint8 tensor; this rank's bytes are body[:length] (length: 1-element int64 tensor on the same device,
            e.g. straight from tsqb_pack_container, or None = all of body).  Passing the device-side length avoids a
            host synchronisation per rank: the byte counts are all-gathered as they are and read back ONCE.
    out:    optional destination buffer on dst (the bytes land at out[out_offset:]); allocated when None.
    Returns (destination tensor holding all bodies back to back from out_offset | None off dst, list of byte counts).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = body.device
    if length is None:
        n = torch.tensor([body.numel()], dtype=torch.int64, device=dev)
    else:
        n = length.reshape(1).to(torch.int64)
    counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, n, group=group)
    counts = [int(c) for c in counts_t.cpu().tolist()]                # the one host synchronisation of the gather
    mine = counts[rank]
    if rank == dst:
        if out is None:
            out = torch.empty(out_offset + sum(counts), dtype=torch.uint8, device=dev)
        at, ops = out_offset, []
        for r, c in enumerate(counts):
            if r == dst:
                out[at:at + c].copy_(body[:c])
            elif c:
                ops.append(dist.P2POp(dist.irecv, out[at:at + c], r, group))
            at += c
        for w in (dist.batch_isend_irecv(ops) if ops else []):       # one grouped NCCL launch straight into `out`
            w.wait()
        return out, counts
    if mine:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, body[:mine], dst, group)]):
            w.wait()
    return None, counts


def gather_bodies_peer(body, dst=0, group=None, length=None, out_offset=0):
    """The same gather through PEER MEMORY instead of NCCL send/recv: `dst` allocates the destination, exports it as a CUDA
    IPC handle (tsqb_ipc_export), every other rank maps it and writes its bytes straight to its prefix-summed offset with one
    device-to-device copy over NVLink / NVSwitch (tsqb_copy_d2d); a tiny all-reduce behind the copies orders them before
    whatever `dst` does next.  NCCL's send/recv stages through its channel buffers (~320 GB/s into the root with seven
    senders); direct peer writes are bound by the root's NVLink ingress.  CUDA tensors + NCCL process group only.
    Returns (destination tensor | None, counts) like gather_bodies, or None when peer memory cannot be used (all ranks agree).
    """
    from . import api
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = body.device
    n = torch.tensor([body.numel()], dtype=torch.int64, device=dev) if length is None else length.reshape(1).to(torch.int64)
    counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, n, group=group)
    counts = [int(c) for c in counts_t.cpu().tolist()]
    msg = torch.zeros(80, dtype=torch.uint8, device=dev)                    # 64-byte handle | u64 offset | u64 ok
    out = None
    if rank == dst:
        out = torch.empty(out_offset + sum(counts), dtype=torch.uint8, device=dev)
        try:
            handle, off = api.ipc_export(out.data_ptr())
            msg.copy_(torch.frombuffer(bytearray(handle + struct.pack("<QQ", off, 1)), dtype=torch.uint8))
        except Exception:
            pass                                                            # ok stays 0: everybody falls back
    dist.broadcast(msg, src=dst, group=group)
    raw = bytes(msg.cpu().numpy())
    off, ok = struct.unpack("<QQ", raw[64:80])
    if not ok:
        return None
    at = out_offset + sum(counts[:rank])
    if rank == dst:
        out[at:at + counts[rank]].copy_(body[:counts[rank]])
    elif counts[rank]:
        base = _peer_mapping(raw[:64])
        api.copy_d2d(base + off + at, body.data_ptr(), counts[rank])
    done = torch.zeros(1, dtype=torch.int32, device=dev)
    dist.all_reduce(done, group=group)                                      # stream-ordered behind every rank's copy
    return out, counts


# Opening an IPC handle maps the WHOLE allocation the destination lies in (with torch's caching allocator: the whole
# segment, often tens of GB) -- ~4 ms per GB, 96 ms measured for a 5.3 GB container inside a large segment -- so the
# mappings are kept: the allocator hands the same segment to the next container and the handle is found here again.
_PEER_MAPPINGS = {}
_PEER_MAPPINGS_MAX = 8


def _peer_mapping(handle):
    base = _PEER_MAPPINGS.get(handle)
    if base is None:
        from . import api
        if len(_PEER_MAPPINGS) >= _PEER_MAPPINGS_MAX:
            release_peer_mappings()
        base = _PEER_MAPPINGS[handle] = api.ipc_open(handle)
    return base


def release_peer_mappings():
    """Unmap every peer allocation this process holds (local, not a collective).  Call it on every rank before the
    destination rank gives the memory back to the driver (torch.cuda.empty_cache()): CUDA requires an exported
    allocation to stay alive while a peer has it mapped."""
    from . import api
    if _PEER_MAPPINGS:
        torch.cuda.synchronize()
        for base in _PEER_MAPPINGS.values():
            api.ipc_close(base)
        _PEER_MAPPINGS.clear()


def gather_container(local_container, total_uncompressed, n_blocks_total, dst=0, group=None, length=None, transport="auto"):
    """Assemble one TSQ1 container on `dst` from per-rank containers (each rank's own header is dropped).

    local_container: uint8 tensor holding a TSQ1 container of this rank's blocks (what Context.pack_container /
    tsqb_pack_container produce); `length`: its device-side length tensor (None = the whole tensor).
    transport: "peer" = direct peer-memory writes (gather_bodies_peer), "nccl" = grouped send/recv (gather_bodies),
    "auto" = peer memory for CUDA tensors under NCCL, else send/recv.
    The header is written in place in front of the gathered bodies -- no second copy of the container.
    Returns the uint8 tensor on dst, else None.
    """
    body_len = None if length is None else length.reshape(1).to(torch.int64) - HEADER
    res = None
    if transport != "nccl" and local_container.is_cuda and dist.get_backend(group) == "nccl":
        res = gather_bodies_peer(local_container[HEADER:], dst=dst, group=group, length=body_len, out_offset=HEADER)
    if res is None:
        res = gather_bodies(local_container[HEADER:], dst=dst, group=group, length=body_len, out_offset=HEADER)
    out = res[0]
    if out is None:
        return None
    hdr = torch.frombuffer(bytearray(container_header(n_blocks_total, total_uncompressed)), dtype=torch.uint8)
    out[:HEADER].copy_(hdr, non_blocking=True)
    return out


def encode_sharded(ctx, host_buf, total, block, ext=0, dst=0, group=None, transport="auto"):
    """The N-GPU encode path: ONE input stream of `total` bytes, rank r encodes the contiguous block
    range byte_range() gives it -- its shard plus INPUT_PAD bytes of the following shard, because the last block of a
    shard reads a few bytes past itself (tsq_encode.cpp:74,126-128; zeros behind the very end) -- frames its streams
    as a TSQ1 body on its own GPU (tsqb_pack_container) and the bodies are gathered to `dst` over NCCL.

    ctx: turbosqueeze_b200.Context on this rank's device; host_buf: numpy uint8 array holding the whole stream (every
    rank reads only its own range).  Returns (container tensor on dst | None, dict of timings-free bookkeeping).
    """
    import numpy as np
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi, hi_tail = byte_range(total, block, rank, world)
    n = hi - lo
    dev = torch.device("cuda", ctx.device)
    d_in = torch.zeros(n + INPUT_PAD, dtype=torch.uint8, device=dev)
    if hi_tail > lo:
        d_in[: hi_tail - lo].copy_(torch.from_numpy(np.ascontiguousarray(host_buf[lo:hi_tail])))
    nb_total = (total + block - 1) // block
    if n:
        slots, sizes = ctx.encode_blocks(d_in, n, block, ext)
        cont, clen = ctx.pack_container(slots, sizes, block, n, ext)
    else:                                                             # more ranks than blocks: an empty body
        cont = torch.zeros(HEADER, dtype=torch.uint8, device=dev)
        clen = torch.tensor([HEADER], dtype=torch.int64, device=dev)
    out = gather_container(cont, total, nb_total, dst=dst, group=group, length=clen, transport=transport)
    return out, {"lo": lo, "hi": hi, "hi_tail": hi_tail, "blocks": (n + block - 1) // block}


def all_gather_decoded(local_out, group=None):
    """Decoded shards back into one buffer on every rank (shards are contiguous block ranges, so this
    is a plain variable-count all-gather)."""
    world = dist.get_world_size(group)
    dev = local_out.device
    n = torch.tensor([local_out.numel()], dtype=torch.int64, device=dev)
    counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, n, group=group)
    counts = [int(c) for c in counts_t.cpu().tolist()]
    cap = max(counts) if counts else 0
    padded = torch.zeros(cap, dtype=torch.uint8, device=dev)
    padded[: local_out.numel()].copy_(local_out)
    parts = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(parts, padded, group=group)
    if all(c == cap for c in counts):
        return parts
    return torch.cat([parts[r * cap: r * cap + c] for r, c in enumerate(counts)])
