#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on the per-block encode/decode hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic input: encode every block, then decode every
block, with the input already resident in HBM (`value`).  The default workload is BASELINE.json configs[1]:
"enwik9 (1 GB text), 256 KiB blocks, --no-ext" (enwik9 itself is not on the image; the enwik9-shape generator in
turbosqueeze_b200/csrc/tsq_workload.c is calibrated to the reference's ratio on enwik9).

Workloads (BASELINE.json configs[1..4]):
  enwik9-shape-1GB-256KiB   cfg 2; N > 1: every rank runs the same shape on its own part of the stream (weak scaling)
  random-4GiB-256KiB        cfg 3; weak like cfg 2
  rep8-16GiB-1MiB           cfg 4; ONE 16 GiB stream SHARDED over the N ranks by contiguous block ranges, with the
                            bytes the last block of a shard reads past itself replicated (sharding.byte_range)
  text-8GiB-sweep           cfg 5; the cfg-2 text replicated to 8 GiB, sharded like cfg 4, at 4 KiB ... 4 MiB blocks
No collective sits on the data path (blocks are independent); the one exchange step -- gathering the
per-rank streams into ONE container on rank 0 over NCCL -- is timed separately and reported as `gather`.

The default N = 1 run also measures cfg 3 / 4 / 5 (fewer steps, device-resident legs only) and reports them under
`other_configs`, so that one driver-run line covers every configuration BASELINE.json names.

`e2e` is the round trip through the host-buffer C-ABI calls (tsqb_compress_into / tsqb_decompress_into: what
tsqCompress_MT / tsqDecompress_MT memory->memory map to) with pinned HOST buffers: H2D + kernels + TSQ1 framing + D2H
inside the timed region; `pcie_floor_ms` is a bare pinned H2D + D2H of the same byte counts on the same box.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified reference compiled by
oracle/Makefile; else the C port) with all host threads on the same workload.
"""
import argparse
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
TEXT_PERIOD = 10 ** 9          # cfg 5: "enwik9 replicated": byte i of the stream = byte i % 10^9 of the cfg-2 text
WORKLOADS = {
    # name: kind, total bytes, block sizes, how N ranks split it
    "enwik9-shape-1GB-256KiB": dict(kind="text", total=10 ** 9, blocks=[262144], split="weak", cfg=2),
    "random-4GiB-256KiB": dict(kind="random", total=4 << 30, blocks=[262144], split="weak", cfg=3),
    "rep8-16GiB-1MiB": dict(kind="rep8", total=16 << 30, blocks=[1 << 20], split="sharded", cfg=4),
    "text-8GiB-sweep": dict(kind="text", total=8 << 30, blocks=[4096, 16384, 65536, 262144, 1 << 20, 4 << 20], split="sharded", cfg=5,
                            period=TEXT_PERIOD),
    "smoke-64MiB-256KiB": dict(kind="text", total=64 << 20, blocks=[262144], split="weak", cfg=0),
}
SEED = 20240917
PAD = 128


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_record():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the cfg-2 workload: a profiler figure, so it cannot be
    measured inside a timed run; it is read from profiles/traffic.json, which names the ncu --set full summary (of the
    kernels at a stated commit) it was copied from."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


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


# ------------------------------------------------------------------------------------------- synthetic input
def stream_bytes(wl, lo, n):
    """Bytes [lo, lo + n) of the workload's stream, followed by PAD zero bytes (host, numpy)."""
    from turbosqueeze_b200 import workloads as W
    period = wl.get("period")
    if not period:
        return W.fill(wl["kind"], n, seed=SEED, offset=lo)
    out = np.zeros(n + PAD, dtype=np.uint8)
    at = 0
    while at < n:
        off = (lo + at) % period
        m = min(n - at, period - off)
        out[at:at + m] = W.fill(wl["kind"], m, seed=SEED, offset=off)[:m]
        at += m
    return out


def config_of(name, wl, block, n_gpus):
    """The `config` object: identical (keys AND values) from both arms -- the driver compares them.  What differs between
    the arms by construction (bytes per step: the CPU arm times a bounded sample; the measured ratio) sits beside it."""
    return {"workload": name, "baseline_config": wl["cfg"], "block": block,
            "with_extensions": 0, "split": wl["split"] if n_gpus > 1 else "single",
            "l2": "inputs and outputs are far larger than the 126 MB L2 (no flush needed)",
            "step": "encode all blocks then decode all blocks"}


# ------------------------------------------------------------------------------------------------ CPU arm
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


def cpu_baseline(buf, sample, block, name, extra=True):
    """The reference's CPU path on `sample` bytes: per-block pool on all host threads (the work of tsq_threads.cpp:176-177,590),
    and -- extra -- one thread (README.md:93's setting) and the reference's own MT pipeline at its native 4 MiB blocks
    (tsqCompress_MT / tsqDecompress_MT memory -> memory, tsq_threads.cpp:413-441,862-890)."""
    codec, ckind = cpu_codec()
    threads = (os.cpu_count() or 1) if ckind == "reference" else 1
    best = None
    for _ in range(2):
        a, b, _c = cpu_round_trip(codec, ckind, buf, sample, block, threads)
        if best is None or a + b < best[0] + best[1]:
            best = (a, b)
    out = {"value": round(sample / (best[0] + best[1]) / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": ckind,
           "encode_gbs": round(sample / best[0] / 1e9, 4), "decode_gbs": round(sample / best[1] / 1e9, 4),
           "sample": f"first {sample} bytes of {name} ({(sample + block - 1) // block} blocks of {block}), encode+decode, best of 2"}
    if extra and ckind == "reference":
        s1 = min(sample, 64 << 20)
        a, b, _c = cpu_round_trip(codec, ckind, buf, s1, block, 1)
        out["one_thread"] = {"encode_gbs": round(s1 / a / 1e9, 4), "decode_gbs": round(s1 / b / 1e9, 4), "sample_bytes": s1,
                             "note": "README.md:93 quotes 0.305 / 2.503 GB/s on a Ryzen 7 3700U"}
        s2 = min(sample, 512 << 20)
        t0 = time.perf_counter()
        blob = codec.compress_mt(buf[:s2], 0)
        t1 = time.perf_counter()
        back = codec.decompress_mt(blob) if blob is not None else None
        t2 = time.perf_counter()
        if back is not None and len(back) == s2:
            out["mt_pipeline_4MiB"] = {"compress_gbs": round(s2 / (t1 - t0) / 1e9, 4), "decompress_gbs": round(s2 / (t2 - t1) / 1e9, 4),
                                       "sample_bytes": s2, "ratio": round(len(blob) / s2, 4),
                                       "what": "the reference's own tsqCompress_MT / tsqDecompress_MT, memory -> memory, native 4 MiB blocks, "
                                               "hardware_concurrency() workers (incl. context setup and its writer's memcpy)"}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path on the host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.workload
    wl = WORKLOADS[name]
    block = args.block or wl["blocks"][min(len(wl["blocks"]) - 1, 3)]
    sample = min(wl["total"], args.cpu_sample_mb << 20)
    buf = stream_bytes(wl, 0, sample)
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
            "config": config_of(name, wl, block, args.gpus), "bytes_per_step": int(sample), "ratio": round(comp / sample, 4),
            "encode_gbs": round(sample * args.steps / te / 1e9, 4), "decode_gbs": round(sample * args.steps / td / 1e9, 4),
            "cpu_baseline": {"value": round(value, 4), "unit": "GB/s", "cores": threads, "kind": ckind, "sample": desc},
            "e2e": {"value": round(value, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(wall, 2)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t]


def shard_of(env, wl, block):
    """(lo, n, n_with_tail, split) of this rank's bytes of the workload's stream."""
    from turbosqueeze_b200 import sharding as S
    total = wl["total"]
    if env.world == 1:
        return 0, total, total
    if wl["split"] == "weak":                                    # same shape per rank, its own part of the stream
        return env.rank * total, total, total
    lo, hi, hi_tail = S.byte_range(total, block, env.rank, env.world)
    return lo, hi - lo, hi_tail - lo


def device_input(env, wl, lo, n, n_tail, cache):
    """This rank's input in HBM: n bytes + whatever follows them in the stream (n_tail - n bytes) + zeros to PAD."""
    torch = env.torch
    d_in = torch.zeros(n + PAD, dtype=torch.uint8, device="cuda")
    if wl["kind"] == "rep8":                                        # built on the device: in[i] = pat[i & 7]
        from turbosqueeze_b200 import workloads as W
        pat = torch.from_numpy(W.fill("rep8", 8, seed=SEED, offset=lo & ~7)[:8].copy()).cuda()
        reps = (n_tail + 15) // 8 + 1
        tiled = pat.repeat(reps)[(lo & 7):(lo & 7) + n_tail]
        d_in[:n_tail].copy_(tiled)
        return d_in
    period = wl.get("period")
    if period:                                                      # the cfg-2 text, replicated on the device
        if "text" not in cache:
            cache["text"] = torch.from_numpy(stream_bytes(dict(kind="text"), 0, period)[:period]).cuda()
        base, at = cache["text"], 0
        while at < n_tail:
            off = (lo + at) % period
            m = min(n_tail - at, period - off)
            d_in[at:at + m].copy_(base[off:off + m])
            at += m
        return d_in
    step = 1 << 30
    for at in range(0, n_tail, step):                               # staged through pageable host memory in 1 GiB pieces
        m = min(step, n_tail - at)
        d_in[at:at + m].copy_(torch.from_numpy(stream_bytes(wl, lo + at, m)[:m]))
    return d_in


def time_device(env, ctx, T, d_in, n, block, steps, warmup):
    """K timed steps of encode-all + decode-all with the input resident in HBM; CUDA events on the launching stream."""
    torch = env.torch
    nb = (n + block - 1) // block
    stride = T.slot_stride(block)
    slots = torch.zeros(max(nb, 1) * stride, dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(max(nb, 1), dtype=torch.int32, device="cuda")
    out = torch.empty(max(nb, 1) * block, dtype=torch.uint8, device="cuda")
    osz = torch.zeros(max(nb, 1), dtype=torch.int32, device="cuda")

    def step():
        if n:
            ctx.encode_blocks(d_in, n, block, 0, slots=slots, sizes=sizes)
            ctx.decode_blocks(slots, nb, block, 0, comp_sizes=sizes, out=out, out_sizes=osz)

    for _ in range(warmup):
        step()
    env.barrier()
    comp = int(sizes[:nb].sum().item()) if n else 0
    ok = (not n) or (bool(torch.equal(out[:n], d_in[:n])) and int(osz[:nb].sum().item()) == n)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    env.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        ev[k][0].record()
        if n:
            ctx.encode_blocks(d_in, n, block, 0, slots=slots, sizes=sizes)
        ev[k][1].record()
        if n:
            ctx.decode_blocks(slots, nb, block, 0, comp_sizes=sizes, out=out, out_sizes=osz)
        ev[k][2].record()
    env.barrier()
    wall = time.perf_counter() - t0
    ms_total = ev[0][0].elapsed_time(ev[-1][2])
    ms_enc = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    ms_dec = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    return dict(ms_total=ms_total, ms_enc=ms_enc, ms_dec=ms_dec, comp=comp, ok=ok, wall=wall, slots=slots, sizes=sizes, nb=nb, out=out)


def pcie_floor(env, h2d_bytes, d2h_bytes, reps=2):
    """A bare pinned H2D of h2d_bytes and D2H of d2h_bytes, concurrently on two streams (PCIe is full duplex): the
    floor of any host-buffer path moving that many bytes.  Returns ms (best of reps) and the per-direction GB/s."""
    torch = env.torch
    cap = 1 << 30
    hb = torch.empty(min(h2d_bytes, cap), dtype=torch.uint8, pin_memory=True)
    hd = torch.empty(min(d2h_bytes, cap), dtype=torch.uint8, pin_memory=True)
    db = torch.empty(hb.numel(), dtype=torch.uint8, device="cuda")
    dd = torch.empty(hd.numel(), dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for _ in range(reps + 1):
        env.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            left = h2d_bytes
            while left > 0:
                m = min(left, hb.numel()); db[:m].copy_(hb[:m], non_blocking=True); left -= m
        with torch.cuda.stream(s2):
            left = d2h_bytes
            while left > 0:
                m = min(left, hd.numel()); hd[:m].copy_(dd[:m], non_blocking=True); left -= m
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        best = t if best is None or t < best else best
    return best * 1e3


def time_e2e(env, ctx, T, wl, lo, n, block, steps):
    """Round trip through tsqb_compress_into + tsqb_decompress_into with pinned host buffers."""
    torch = env.torch
    nb = (n + block - 1) // block
    stride = T.slot_stride(block)
    cap = 16 + nb * (stride + 3)
    pinned_in = torch.empty(n + PAD, dtype=torch.uint8, pin_memory=True)
    host = stream_bytes(wl, lo, n)
    pinned_in.numpy()[:] = host[: n + PAD]
    pinned_cont = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    pinned_out = torch.empty(n + 128, dtype=torch.uint8, pin_memory=True)

    def one():
        t0 = time.perf_counter()
        c = ctx.compress_into(pinned_in.data_ptr(), n, block, 0, pinned_cont.data_ptr(), cap)
        t1 = time.perf_counter()
        m = ctx.decompress_into(pinned_cont.data_ptr(), c, pinned_out.data_ptr(), n + 128)
        return c, m, t1 - t0, time.perf_counter() - t1

    clen, m, _, _ = one()
    env.barrier()
    t0 = time.perf_counter()
    tc = td = 0.0
    for _ in range(steps):
        _c, _m, a, b = one()
        tc += a; td += b
    env.barrier()
    e2e_s = (time.perf_counter() - t0) / steps
    ok = m == n and bool((pinned_out.numpy()[:n] == host[:n]).all())
    floor_ms = pcie_floor(env, n + clen, clen + n)
    return dict(s=e2e_s, clen=clen, ok=ok, compress_ms=tc / steps * 1e3, decompress_ms=td / steps * 1e3, floor_ms=floor_ms,
                container=pinned_cont, host=host)


def byte_sum(torch, t):
    """Sum of a uint8 tensor's bytes, 64 MiB at a time (sum(dtype=int64) materialises an 8x copy of its input)."""
    return sum(int(c.sum(dtype=torch.int64).item()) for c in t.split(1 << 26)) if t.numel() else 0


def time_gather(env, ctx, wl, res, n, block, nb_total, total_all):
    """N > 1: every rank frames its streams as a TSQ1 body on its GPU, the bodies are gathered into ONE container on
    rank 0 (NCCL over NVLink).  Reported next to `value`, never inside it (the root's NVLink ingress bounds it)."""
    from turbosqueeze_b200 import sharding as S
    torch, dist = env.torch, env.dist
    state = {}

    def gather_step(transport):
        cont, clen = ctx.pack_container(res["slots"], res["sizes"], block, n)
        state["cont"], state["clen"] = cont, clen
        return S.gather_container(cont, total_all, nb_total, dst=0, length=clen, transport=transport)

    times, gathered = {}, None
    for transport in ("nccl", "peer"):                              # "peer" last: its container is the one checked below
        del gathered
        gathered = gather_step(transport)
        env.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        del gathered
        gathered = gather_step(transport)
        g1.record()
        env.barrier()
        times[transport] = env.reduce([g0.elapsed_time(g1)], "MAX")[0]
    ms = times["peer"]
    mine = int(state["clen"].item()) - 16
    body = env.reduce([float(mine)], "SUM")[0]
    # every rank's body must have arrived: sum of all bytes (the byte-exact check is tests/test_sharded_gpu.py)
    t = torch.tensor([byte_sum(torch, state["cont"][16:16 + mine])], dtype=torch.int64, device=state["cont"].device)
    if env.world > 1:
        dist.all_reduce(t)
    out = None
    if env.rank == 0:
        hdr = bytes(gathered[:16].cpu().numpy())
        ok = (hdr[:4] == b"TSQ1" and int.from_bytes(hdr[4:8], "little") == nb_total and int.from_bytes(hdr[8:16], "little") == total_all and
              int(gathered.numel()) == 16 + int(body) and bool(torch.equal(gathered[16:16 + mine], state["cont"][16:16 + mine])) and
              byte_sum(torch, gathered[16:]) == int(t[0].item()))
        out = {"ms": round(ms, 3), "container_bytes": int(gathered.numel()), "ok": ok, "root_ingress_gbs": round((body - mine) / (ms * 1e-3) / 1e9, 1),
               "ms_nccl_send_recv": round(times["nccl"], 3), "root_ingress_gbs_nccl_send_recv": round((body - mine) / (times["nccl"] * 1e-3) / 1e9, 1),
               "what": "tsqb_pack_container per rank + one all_gather of the device-side byte counts (one host sync) + every rank writes its "
                       "body straight into the root's container through peer memory (CUDA IPC mapping, one device-to-device copy over "
                       "NVLink / NVSwitch per rank); ms_nccl_send_recv = the same with grouped NCCL send/recv"}
        if wl["split"] == "sharded":
            out["note"] = ("sharded stream: the gathered container IS the container of the whole stream "
                           "(tests/test_sharded_gpu.py checks it against the single-GPU container and the reference)")
    del gathered
    return out


def measure(env, ctx, T, name, wl, block, steps, warmup, cache, full):
    """One (workload, block size) measurement on this rank's shard.  full: also e2e, gather, cpu_baseline."""
    torch = env.torch
    lo, n, n_tail = shard_of(env, wl, block)
    d_in = device_input(env, wl, lo, n, n_tail, cache)
    torch.cuda.synchronize()
    res = time_device(env, ctx, T, d_in, n, block, steps, warmup)
    ms_total, ms_enc, ms_dec = env.reduce([res["ms_total"], res["ms_enc"], res["ms_dec"]], "MAX")
    comp_all, n_all, ok_all = env.reduce([float(res["comp"]), float(n), 1.0 if res["ok"] else 0.0], "SUM")
    total_all = int(n_all)
    ms_step = ms_total / steps
    hbm_peak, peak_src = peaks()
    alg = n + res["comp"]                                          # this rank's algorithmic bytes per launch (U + C)
    alg_max = env.reduce([float(alg)], "MAX")[0]
    r = dict(name=name, block=block, total_all=total_all, ratio=comp_all / max(total_all, 1), ms_step=ms_step, ms_enc=ms_enc, ms_dec=ms_dec,
             value=total_all / (ms_step * 1e-3) / 1e9, enc_gbs=total_all / (ms_enc * 1e-3) / 1e9, dec_gbs=total_all / (ms_dec * 1e-3) / 1e9,
             ok=ok_all == env.world, alg=int(alg_max), enc_ach=alg_max / (ms_enc * 1e-3) / 1e9, dec_ach=alg_max / (ms_dec * 1e-3) / 1e9,
             peak=hbm_peak, peak_src=peak_src, wall=res["wall"], n_local=n, nb_local=res["nb"])
    if full:
        if env.world > 1:
            nb_total = int(env.reduce([float(res["nb"])], "SUM")[0])
            r["gather"] = time_gather(env, ctx, wl, res, n, block, nb_total, total_all)
        del res["slots"], res["out"]
        torch.cuda.empty_cache()
        e_n = min(n, 1 << 30)                                      # e2e on at most the first 1 GiB of the rank's shard
        if e_n < n:
            e_n -= e_n % block
        e2e = time_e2e(env, ctx, T, wl, lo, e_n, block, max(1, min(steps, 3)))
        e2e_s = env.reduce([e2e["s"]], "MAX")[0]
        e_all = env.reduce([float(e_n)], "SUM")[0]
        floor_ms = env.reduce([e2e["floor_ms"]], "MAX")[0]
        r["e2e"] = {"value": round(e_all / e2e_s / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": int(e_n + e2e["clen"]),
                    "d2h_bytes_per_step": int(e2e["clen"] + e_n), "call": "tsqb_compress_into + tsqb_decompress_into (pinned host buffers)",
                    "steps": max(1, min(steps, 3)), "round_trip_ok": e2e["ok"], "ms_per_step": round(e2e_s * 1e3, 3),
                    "compress_ms": round(e2e["compress_ms"], 3), "decompress_ms": round(e2e["decompress_ms"], 3),
                    "bytes_per_gpu": int(e_n), "pcie_floor_ms": round(floor_ms, 3),
                    "pcie_floor_note": "bare pinned H2D of (U + C) bytes and D2H of (C + U) bytes on two streams at once, same box, max over ranks; "
                                       f"per rank that is {round((2 * e_n + 2 * e2e['clen']) / (floor_ms * 1e-3) / 1e9, 1)} GB/s of PCIe traffic"}
        if env.rank == 0 and reference_present():
            # the 1 GiB container of the host path against the reference's own bytes (first blocks; the whole container's
            # block streams are covered by tests/test_gpu_parity.py::test_large_buffers_bit_exact_and_round_trip)
            r["e2e"]["container_matches_reference"] = container_head_matches(e2e, e_n, block)
        del e2e
    del d_in
    torch.cuda.empty_cache()
    return r


def reference_present():
    from oraclelib import Reference
    return Reference.available()


def container_head_matches(e2e, n, block, blocks=64):
    """The first `blocks` block streams of the host path's container == the unmodified reference's streams of the same
    bytes.  Both encode in place inside the same buffer, so every compared block sees the same bytes behind it."""
    from oraclelib import Reference, slot_stride
    buf = e2e["host"]
    m = min(n, blocks * block)
    want_slots, want_sizes, _ = Reference().encode_blocks(buf, m, block, 0, threads=os.cpu_count() or 8)
    cont = e2e["container"].numpy()
    stride = slot_stride(block)
    at = 16
    for b in range(len(want_sizes)):
        ln = int(cont[at]) | int(cont[at + 1]) << 8 | int(cont[at + 2]) << 16
        at += 3
        if ln != int(want_sizes[b]) or not np.array_equal(cont[at:at + ln], want_slots[b * stride: b * stride + ln]):
            return False
        at += ln
    return True


def run_ours(args):
    env = Env()
    torch, dist = env.torch, env.dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU baseline)")
    if os.environ.get("TSQB_LIBRARY") and not args.allow_variant_library:
        raise SystemExit("bench.py: TSQB_LIBRARY is set (a development kernel variant); refusing to benchmark it as the product "
                         "(pass --allow-variant-library for A/B runs)")
    torch.cuda.set_device(env.local)
    if env.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", env.local))
    import turbosqueeze_b200 as T
    L = T.library()
    ctx = T.Context(env.local)
    name = args.workload
    wl = WORKLOADS[name]
    warmup = max(args.warmup, 3)
    cache = {}
    blocks = [args.block] if args.block else wl["blocks"]
    main_block = blocks[min(len(blocks) - 1, 3)] if len(blocks) > 1 else blocks[0]   # the sweep's headline row is 256 KiB

    sampler = ClockSampler(env.local)
    if env.rank == 0:
        sampler.start()
    launches0 = L.tsqb_launch_count()
    r = measure(env, ctx, T, name, wl, main_block, args.steps, warmup, cache, full=True)
    launches_main = 2 * args.steps                                 # timed region: K x (encode_batch_kernel + decode_split_kernel)
    clocks = sampler.stop() if env.rank == 0 else None

    sweep = []
    for b in blocks:
        if b != main_block:
            q = measure(env, ctx, T, name, wl, b, max(2, min(args.steps, 3)), 2, cache, full=False)
            sweep.append(q)

    others = []
    if env.world == 1 and name == "enwik9-shape-1GB-256KiB" and not args.no_other_configs:
        for oname in ("random-4GiB-256KiB", "rep8-16GiB-1MiB", "text-8GiB-sweep"):
            owl = WORKLOADS[oname]
            for b in owl["blocks"]:
                try:
                    q = measure(env, ctx, T, oname, owl, b, 2, 2, cache, full=False)
                    others.append(q)
                except Exception as e:                              # e.g. a smaller GPU: say so instead of dying
                    others.append(dict(name=oname, block=b, error=str(e)[:200]))
                    torch.cuda.empty_cache()
    cache.clear()
    torch.cuda.empty_cache()

    def compact(q):
        if "error" in q:
            return q
        return {"workload": q["name"], "block": q["block"], "bytes": q["total_all"], "ratio": round(q["ratio"], 4),
                "value": round(q["value"], 3), "encode_gbs": round(q["enc_gbs"], 3), "decode_gbs": round(q["dec_gbs"], 3),
                "ms_encode": round(q["ms_enc"], 3), "ms_decode": round(q["ms_dec"], 3), "bit_exact_round_trip": q["ok"],
                "roofline_frac_encode": round(q["enc_ach"] / q["peak"], 5), "roofline_frac_decode": round(q["dec_ach"] / q["peak"], 5)}

    if env.rank == 0:
        tr = traffic_record() if name == "enwik9-shape-1GB-256KiB" and main_block == 262144 else {}
        line = {
            "metric": METRIC, "value": round(r["value"], 3), "unit": "GB/s", "n_gpus": env.world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": round(r["ms_step"], 4), "higher_is_better": True, "scaling": "strong" if (wl["split"] == "sharded" and env.world > 1) else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_of(name, wl, main_block, env.world), "bytes_per_step": int(r["total_all"]), "ratio": round(r["ratio"], 4),
            "encode_gbs": round(r["enc_gbs"], 3), "decode_gbs": round(r["dec_gbs"], 3), "ms_encode": round(r["ms_enc"], 4),
            "ms_decode": round(r["ms_dec"], 4), "bit_exact_round_trip": r["ok"],
            "roofline": {"kernel": "encode_batch_kernel", "bound": "hbm", "achieved": round(r["enc_ach"], 2), "peak": r["peak"], "unit": "GB/s",
                         "frac": round(r["enc_ach"] / r["peak"], 5), "traffic": tr.get("encode_batch_kernel"), "traffic_source": tr.get("source"),
                         "peak_source": r["peak_src"], "algorithmic_bytes_per_launch": r["alg"],
                         "share_of_step": round(r["ms_enc"] / (r["ms_enc"] + r["ms_dec"]), 4), "ceiling_note": tr.get("ceiling_note")},
            "roofline_decode": {"kernel": "decode_split_kernel", "bound": "hbm", "achieved": round(r["dec_ach"], 2), "peak": r["peak"], "unit": "GB/s",
                                "frac": round(r["dec_ach"] / r["peak"], 5), "traffic": tr.get("decode_split_kernel"), "traffic_source": tr.get("source"),
                                "algorithmic_bytes_per_launch": r["alg"], "share_of_step": round(r["ms_dec"] / (r["ms_enc"] + r["ms_dec"]), 4)},
            "e2e": r["e2e"], "gpu_launches": int(launches_main), "library": os.path.relpath(T.library_path(), ROOT),
            "library_launches_whole_run": int(L.tsqb_launch_count() - launches0), "clocks": clocks, "wall_s_timed_region": round(r["wall"], 3),
        }
        line["n_blocks_per_gpu"] = r["nb_local"]
        line["bytes_per_gpu"] = r["n_local"]
        if r.get("gather") is not None:
            g = r["gather"]
            g["encode_with_gather_gbs"] = round(r["total_all"] / ((r["ms_enc"] + g["ms"]) * 1e-3) / 1e9, 3)
            line["gather"] = g
        if sweep:
            line["block_sweep"] = [compact(r)] + [compact(q) for q in sweep]
        if others:
            line["other_configs"] = [compact(q) for q in others]
            line["other_configs_note"] = "BASELINE.json configs 3, 4, 5 on this GPU: device-resident legs, 2 warm-ups + 2 timed steps each"
        if env.world == 1 and not args.no_cpu:
            sample = min(wl["total"], args.cpu_sample_mb << 20)
            buf = stream_bytes(wl, 0, sample)
            line["cpu_baseline"] = cpu_baseline(buf, sample, main_block, name)
        print(json.dumps(line), flush=True)
    ctx.close()
    if env.world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="enwik9-shape-1GB-256KiB", choices=sorted(WORKLOADS))
    ap.add_argument("--block", type=int, default=0, help="one block size instead of the workload's own (sweep: all of them)")
    ap.add_argument("--cpu-sample-mb", type=int, default=1024, help="bytes of the workload the CPU baseline is timed on")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="default run: skip the cfg 3 / 4 / 5 measurements")
    ap.add_argument("--allow-variant-library", action="store_true", help="development: benchmark the library TSQB_LIBRARY names")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
