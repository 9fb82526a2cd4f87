"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: contiguous block ranges per rank and
the variable-length gather that assembles one TSQ1 container (turbosqueeze_b200/sharding.py).

The per-rank "kernel" here is the CPU oracle -- this test is about the host-side sharding / gather
logic, the CUDA kernels are covered by test_gpu_parity.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oraclelib import Oracle, PAD, slot_stride
from turbosqueeze_b200 import sharding as S
from turbosqueeze_b200 import workloads as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _container_of(buf, lo, hi, block, tail):
    """TSQ1 container (numpy) of bytes [lo, hi) of buf; the encoder may read up to `tail`."""
    n = hi - lo
    shard = np.zeros(n + PAD, dtype=np.uint8)
    shard[: tail - lo] = buf[lo:tail]
    slots, sizes, _ = Oracle().encode_blocks(shard, n, block, 0)
    stride = slot_stride(block)
    parts = [S.container_header(len(sizes), n)]
    for b, c in enumerate(sizes):
        c = int(c)
        parts.append(bytes([c & 0xFF, (c >> 8) & 0xFF, (c >> 16) & 0xFF]))
        parts.append(bytes(slots[b * stride: b * stride + c]))
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()


def _worker(rank, world, port, total, block, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        buf = W.fill("text", total, seed=11)
        lo, hi, tail = S.byte_range(total, block, rank, world)
        local = torch.from_numpy(_container_of(buf, lo, hi, block, tail))
        nb = (total + block - 1) // block
        cont = S.gather_container(local, total, nb, dst=0)
        # decode side: every rank decodes its own blocks, then an all-gather restores the buffer
        mine = torch.from_numpy(buf[lo:hi].copy())
        whole = S.all_gather_decoded(mine)
        ok_whole = bool(np.array_equal(whole.numpy(), buf[:total]))
        if rank == 0:
            q.put((cont.numpy().tobytes(), ok_whole))
        else:
            q.put((None, ok_whole))
    finally:
        dist.destroy_process_group()


def test_block_and_byte_ranges_partition_the_input():
    for total, block in [(1, 1), (1000, 64), (10 ** 6 + 3, 4096), (1 << 20, 1 << 18)]:
        nb = (total + block - 1) // block
        for world in (1, 2, 3, 8):
            prev_b, prev_y = 0, 0
            for r in range(world):
                b0, b1 = S.block_range(nb, r, world)
                lo, hi, tail = S.byte_range(total, block, r, world)
                assert b0 == prev_b and lo == prev_y and lo == min(total, b0 * block)
                assert hi <= tail <= min(total, hi + S.INPUT_PAD)
                prev_b, prev_y = b1, hi
            assert prev_b == nb and prev_y == total


@pytest.mark.timeout(120)
def test_two_ranks_assemble_the_single_process_container():
    total, block, world = 300000 + 17, 16384, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, block, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(ok for _, ok in got)
    cont = next(c for c, _ in got if c is not None)
    buf = W.fill("text", total, seed=11)
    want = _container_of(buf, 0, total, block, total).tobytes()
    assert cont == want, "sharded encode + gather differs from the single-process container"
