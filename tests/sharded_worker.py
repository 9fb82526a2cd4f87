"""Worker of tests/test_sharded_gpu.py, launched with torch.distributed.run (one rank per GPU, NCCL).

The real N-GPU encode path of SURVEY.md 8(e): ONE stream, sharded by contiguous block ranges with INPUT_PAD bytes of the
following shard replicated (turbosqueeze_b200.sharding.encode_sharded), encoded on every rank's GPU, the TSQ1 bodies
gathered to rank 0 over NCCL.  Rank 0 then checks the gathered container
  (a) byte for byte against the container ONE GPU produces for the whole stream (tsqb_encode_blocks + tsqb_pack_container),
  (b) block by block against the streams of the unmodified reference (oracle/_ref; else the pinned C restatement),
  (c) by decoding it, sharded again, and all-gathering the decoded shards.
Prints one JSON line; exit code 0 = all checks passed.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    kind, total, block, ext = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    import turbosqueeze_b200 as T
    from turbosqueeze_b200 import sharding as S
    from turbosqueeze_b200 import workloads as W
    from oraclelib import best_cpu_codec, slot_stride

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = T.Context(local)
    buf = W.fill(kind, total, seed=4242)                          # every rank generates the same stream, reads its own range
    nb = (total + block - 1) // block
    # both transports of the gather: peer-memory writes (the default under NCCL) and grouped NCCL send/recv
    cont_nccl, _ = S.encode_sharded(ctx, buf, total, block, ext, dst=0, transport="nccl")
    cont, info = S.encode_sharded(ctx, buf, total, block, ext, dst=0, transport="peer")
    torch.cuda.synchronize()
    res = {"world": world, "kind": kind, "total": total, "block": block, "ext": ext, "n_blocks": nb}
    ok = True
    if rank == 0:
        got = cont.cpu().numpy()
        # (a) the single-GPU container of the whole stream
        d = torch.from_numpy(buf).cuda()
        slots, sizes = ctx.encode_blocks(d, total, block, ext)
        one, n1 = ctx.pack_container(slots, sizes, block, total, ext)
        one = one[: int(n1.item())].cpu().numpy()
        res["equals_single_gpu_container"] = bool(got.size == one.size and np.array_equal(got, one))
        res["transports_agree"] = bool(torch.equal(cont, cont_nccl))
        # (b) the reference's streams
        codec = best_cpu_codec()
        res["checker"] = codec.name
        want_slots, want_sizes, _ = codec.encode_blocks(buf, total, block, ext, threads=os.cpu_count() or 8)
        stride = slot_stride(block)
        at, same = 16, got[:4].tobytes() == b"TSQ1" and int.from_bytes(got[4:8].tobytes(), "little") == nb and \
            int.from_bytes(got[8:16].tobytes(), "little") == total
        for b in range(nb):
            ln = int(got[at]) | int(got[at + 1]) << 8 | int(got[at + 2]) << 16
            at += 3
            same = same and bool(ln & 0x800000) == bool(ext)
            ln &= 0x7FFFFF
            same = same and ln == int(want_sizes[b]) and np.array_equal(got[at:at + ln], want_slots[b * stride: b * stride + ln])
            at += ln
            if not same:
                res["first_bad_block"] = b
                break
        res["equals_reference_streams"] = bool(same and at == got.size)
        ok = res["equals_single_gpu_container"] and res["equals_reference_streams"] and res["transports_agree"]
    # (c) decode, sharded: rank r indexes the container (rank 0 broadcasts it) and decodes its own block range
    n_t = torch.tensor([cont.numel() if rank == 0 else 0], dtype=torch.int64, device="cuda")
    dist.broadcast(n_t, 0)
    full = cont if rank == 0 else torch.empty(int(n_t.item()), dtype=torch.uint8, device="cuda")
    dist.broadcast(full, 0)
    offs, csz, _e, n_idx = ctx.index_container(full, full.numel(), nb)
    b0, b1 = S.block_range(nb, rank, world)
    if b1 > b0:
        out, osz = ctx.decode_blocks(full, b1 - b0, block, ext, offsets=offs[b0:b1].contiguous(), comp_sizes=csz[b0:b1].contiguous())
        mine = out[: int(osz.sum().item())]
    else:
        mine = torch.empty(0, dtype=torch.uint8, device="cuda")
    whole = S.all_gather_decoded(mine)
    rt = bool(int(n_idx.item()) == nb and whole.numel() == total and torch.equal(whole.cpu(), torch.from_numpy(buf[:total])))
    flag = torch.tensor([1 if (rt and ok) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        res["sharded_decode_round_trip"] = rt
        res["ok"] = bool(flag.item())
        print(json.dumps(res), flush=True)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
