#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on the per-block encode/decode hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input: encode every block, then
decode every block, with the input already resident in HBM (`value`).  At N = 1 the workload is
BASELINE.json configs[1]: "enwik9 (1 GB text), 256 KiB blocks, --no-ext" (enwik9 itself is not on
the image; the enwik9-shape generator in turbosqueeze_b200/csrc/tsq_workload.c is calibrated to the
reference's ratio on enwik9).  At N > 1 every rank runs the same shape on its own
shard of the stream (weak scaling, no data-path collective; blocks are independent).

`e2e` is the same round trip through the host-buffer C-ABI calls (tsqb_compress_into /
tsqb_decompress_into: what tsqCompress_MT / tsqDecompress_MT memory->memory map to) with pinned
HOST buffers: H2D + kernels + TSQ1 framing + D2H inside the timed region.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified
reference compiled by oracle/Makefile; else the C port) with all host threads on the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "uncompressed GB/s encode+decode on enwik9-shape blocks"
WORKLOADS = {
    # name: (kind, total bytes, block)
    "enwik9-shape-1GB-256KiB": ("text", 10 ** 9, 262144),
    "random-4GiB-256KiB": ("random", 4 << 30, 262144),
    "rep8-16GiB-1MiB": ("rep8", 16 << 30, 1 << 20),
    "smoke-64MiB-256KiB": ("text", 64 << 20, 262144),
}
SEED = 20240917


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_codec():
    from oraclelib import Oracle, Reference
    if Reference.available():
        return Reference(), "reference"
    return Oracle(), "port"


def cpu_round_trip(codec, kind, buf, n, block, threads):
    """One encode + decode pass of the CPU codec; returns (t_enc, t_dec, compressed bytes)."""
    from oraclelib import slot_stride
    if kind == "reference":
        slots, sizes, t_enc = codec.encode_blocks(buf, n, block, 0, threads=threads)
        _, _, t_dec = codec.decode_blocks(slots, slot_stride(block), len(sizes), block, 0, threads=threads, comp_sizes=sizes)
    else:
        t0 = time.perf_counter()
        slots, sizes, _ = codec.encode_blocks(buf, n, block, 0)
        t1 = time.perf_counter()
        codec.decode_blocks(slots, slot_stride(block), len(sizes), block, 0)
        t_enc, t_dec = t1 - t0, time.perf_counter() - t1
    return t_enc, t_dec, int(sizes.sum())


def run_reference(args):
    """--impl reference: the reference's CPU path on the host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from turbosqueeze_b200 import workloads as W
    name = args.workload
    kind, total, block = WORKLOADS[name]
    sample = min(total, args.cpu_sample_mb << 20)
    buf = W.fill(kind, sample, seed=SEED)
    codec, ckind = cpu_codec()
    threads = (os.cpu_count() or 1) if ckind == "reference" else 1
    for _ in range(args.warmup):
        cpu_round_trip(codec, ckind, buf, sample, block, threads)
    t0 = time.perf_counter()
    te = td = 0.0
    for _ in range(args.steps):
        a, b, comp = cpu_round_trip(codec, ckind, buf, sample, block, threads)
        te += a; td += b
    wall = time.perf_counter() - t0
    t_step = (te + td) / args.steps
    value = sample / t_step / 1e9
    desc = f"first {sample} bytes of {name} ({(sample + block - 1) // block} blocks of {block}), encode+decode per step"
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(t_step * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": name, "block": block, "bytes_per_step": sample, "ratio": round(comp / sample, 4)},
            "encode_gbs": round(sample * args.steps / te / 1e9, 4), "decode_gbs": round(sample * args.steps / td / 1e9, 4),
            "cpu_baseline": {"value": round(value, 4), "unit": "GB/s", "cores": threads, "kind": ckind, "sample": desc},
            "e2e": {"value": round(value, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(wall, 2)}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import turbosqueeze_b200 as T
    from turbosqueeze_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    name = args.workload
    kind, total, block = WORKLOADS[name]
    nb = (total + block - 1) // block
    stride = T.slot_stride(block)
    # rank r works on bytes [r*total, (r+1)*total) of the stream: same shape, different data
    buf = W.fill(kind, total, seed=SEED, offset=rank * total if kind != "rep8" else 0)
    hbm_peak, peak_src = peaks()

    ctx = T.Context(local)
    d_in = torch.empty(total + T.INPUT_PAD, dtype=torch.uint8, device="cuda")
    pinned_in = torch.empty(total + T.INPUT_PAD, dtype=torch.uint8, pin_memory=True)
    pinned_in.numpy()[:] = buf
    d_in.copy_(pinned_in, non_blocking=True)
    slots = torch.zeros(nb * stride, dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(nb, dtype=torch.int32, device="cuda")
    out = torch.empty(nb * block, dtype=torch.uint8, device="cuda")
    osz = torch.zeros(nb, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    L = T.library()

    def step():
        ctx.encode_blocks(d_in, total, block, 0, slots=slots, sizes=sizes)
        ctx.decode_blocks(slots, nb, block, 0, comp_sizes=sizes, out=out, out_sizes=osz)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    comp = int(sizes.sum().item())
    ok = bool(torch.equal(out[:total], d_in[:total])) and int(osz.sum().item()) == total

    # ---- timed region: K steps, CUDA events on the launching stream; per-kernel events inside
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = L.tsqb_launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record()
        ctx.encode_blocks(d_in, total, block, 0, slots=slots, sizes=sizes)
        ev[k][1].record()
        ctx.decode_blocks(slots, nb, block, 0, comp_sizes=sizes, out=out, out_sizes=osz)
        ev[k][2].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = L.tsqb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev[0][0].elapsed_time(ev[-1][2])
    ms_enc = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    ms_dec = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t = torch.tensor([ms_total, ms_enc, ms_dec, float(comp)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, ms_enc, ms_dec = (float(x) for x in tmax[:3])
        comp_all = float(tsum[3])
    else:
        comp_all = float(comp)
    ms_step = ms_total / args.steps
    total_all = total * world
    value = total_all / (ms_step * 1e-3) / 1e9

    # ---- e2e: host buffers through the C-ABI (H2D + kernels + framing + D2H), same N GPUs
    cap = 16 + nb * (stride + 3)
    pinned_cont = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    pinned_out = torch.empty(total + 128, dtype=torch.uint8, pin_memory=True)

    def e2e_step():
        n = ctx.compress_into(pinned_in.data_ptr(), total, block, 0, pinned_cont.data_ptr(), cap)
        m = ctx.decompress_into(pinned_cont.data_ptr(), n, pinned_out.data_ptr(), total + 128)
        return n, m

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    clen, m = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_ok = m == total and bool((pinned_out.numpy()[:total] == buf[:total]).all())
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    e2e_value = total_all / e2e_s / 1e9

    # ---- N > 1: the one exchange step of the path -- every rank frames its streams as a TSQ1 body on
    # its GPU, then the bodies are gathered into ONE container on rank 0 (NCCL over NVLink).  Reported
    # next to `value`, never inside it (the root's NVLink ingress bounds it).
    gather = None
    if world > 1:
        from turbosqueeze_b200 import sharding as S
        body_bytes = [0]

        def gather_step():
            cont, n = ctx.pack_container(slots, sizes, block, total)
            clen_local = int(n.item())
            body_bytes[0] = clen_local - 16
            return S.gather_container(cont[:clen_local], total_all, nb * world, dst=0)
        gather_step()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gathered = gather_step()
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        tb = torch.tensor([float(body_bytes[0])], dtype=torch.float64, device="cuda")
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        if rank == 0:
            hdr = bytes(gathered[:16].cpu().numpy())
            gather_ok = (hdr[:4] == b"TSQ1" and int.from_bytes(hdr[4:8], "little") == nb * world and
                         int.from_bytes(hdr[8:16], "little") == total_all and int(gathered.numel()) == 16 + int(tb[0]) and
                         bool(torch.equal(gathered[16:16 + body_bytes[0]], ctx.pack_container(slots, sizes, block, total)[0][16:16 + body_bytes[0]])))
            gather = {"ms": round(float(tg[0]), 3), "container_bytes": int(gathered.numel()), "ok": gather_ok,
                      "what": "tsqb_pack_container per rank + all_gather of byte counts + variable-length gather to rank 0 (NCCL)"}

    if rank == 0:
        r = comp_all / total_all
        enc_gbs = total_all / (ms_enc * 1e-3) / 1e9
        dec_gbs = total_all / (ms_dec * 1e-3) / 1e9
        # roofline of the dominant kernel (the encoder): algorithmic bytes = U read + C written per launch
        # (SURVEY.md 8(d)); per GPU, against the measured copy bandwidth
        alg = (total + comp_all / world)
        enc_ach = alg / (ms_enc * 1e-3) / 1e9
        dec_ach = alg / (ms_dec * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": name, "block": block, "n_blocks_per_gpu": nb, "bytes_per_gpu": total, "ratio": round(r, 4),
                       "with_extensions": 0, "l2": "inputs (1 GB in, 0.6 GB streams, 1 GB out per GPU) are larger than the 126 MB L2",
                       "step": "encode all blocks then decode all blocks, device-resident"},
            "encode_gbs": round(enc_gbs, 3), "decode_gbs": round(dec_gbs, 3), "ms_encode": round(ms_enc, 4), "ms_decode": round(ms_dec, 4),
            "bit_exact_round_trip": ok,
            "roofline": {"kernel": "encode_batch_kernel", "bound": "hbm", "achieved": round(enc_ach, 2), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(enc_ach / hbm_peak, 5), "traffic": args.traffic_encode if name == "enwik9-shape-1GB-256KiB" else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg), "share_of_step": round(ms_enc / (ms_enc + ms_dec), 4)},
            "roofline_decode": {"kernel": "decode_split_kernel", "bound": "hbm", "achieved": round(dec_ach, 2), "peak": hbm_peak, "unit": "GB/s",
                                "frac": round(dec_ach / hbm_peak, 5), "traffic": args.traffic_decode if name == "enwik9-shape-1GB-256KiB" else None,
                                "algorithmic_bytes_per_launch": int(alg), "share_of_step": round(ms_dec / (ms_enc + ms_dec), 4)},
            "e2e": {"value": round(e2e_value, 3), "unit": "GB/s", "h2d_bytes_per_step": int(total + clen), "d2h_bytes_per_step": int(clen + total),
                    "call": "tsqb_compress_into + tsqb_decompress_into (pinned host buffers)", "steps": e2e_steps, "round_trip_ok": e2e_ok,
                    "ms_per_step": round(e2e_s * 1e3, 3)},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": round(t_wall, 3),
        }
        if gather is not None:
            gather["encode_with_gather_gbs"] = round(total_all / ((ms_enc + gather["ms"]) * 1e-3) / 1e9, 3)
            line["gather"] = gather
        if world == 1 and not args.no_cpu:
            codec, ckind = cpu_codec()
            threads = (os.cpu_count() or 1) if ckind == "reference" else 1
            sample = min(total, args.cpu_sample_mb << 20)
            best = None
            for _ in range(2):
                a, b, _c = cpu_round_trip(codec, ckind, buf, sample, block, threads)
                if best is None or a + b < best[0] + best[1]:
                    best = (a, b)
            line["cpu_baseline"] = {"value": round(sample / (best[0] + best[1]) / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": ckind,
                                    "encode_gbs": round(sample / best[0] / 1e9, 4), "decode_gbs": round(sample / best[1] / 1e9, 4),
                                    "sample": f"first {sample} bytes of {name} ({(sample + block - 1) // block} blocks), encode+decode, best of 2"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="enwik9-shape-1GB-256KiB", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-mb", type=int, default=1024, help="bytes of the workload the CPU baseline is timed on")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the 1 GB workload, from the committed ncu --set full
    # capture profiles/r01_v6_ncu_full_summary.txt (a profiler figure cannot be measured inside a timed run)
    ap.add_argument("--traffic-encode", type=float, default=55.30e9, help="dram bytes per launch from an ncu --set full capture")
    ap.add_argument("--traffic-decode", type=float, default=5.86e9)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
