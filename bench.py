#!/usr/bin/env python
"""bench.py -- filtered-scan rows/s (BASELINE.json metric) on N B200s of one node.

One "step" = one pass of the hot path over the resident table at each of the three selectivities
BASELINE config C2 names: `{(I1,<,t1)}^{(R,<,t2)}` at joint selectivity 1 %, 10 % and 50 %, project
[I1,I2,R,S], COUNT / SUM(I2) / SUM(R) / MIN(I1) / MAX(I1).  So a step scans 3 x rows-per-GPU rows on
every rank.  N=1: the C2 table (100 M rows, 2.8 GB in HBM).  N>1: every rank holds its own 100 M-row
position range of the same table (TID-range sharding, weak scaling), scans it with no data-path
collective, then NCCL all-reduces the aggregates, all-gathers the counts and gathers the 1 % query's
positions + projected values on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

`--impl reference` times the reference's CPU algorithm (oracle/: the literal C++ restatement; the Java
original cannot run here -- no JDK in the image) on the box's host cores on a bounded sample of the same
workload.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20260101
SELECTIVITIES = (0.01, 0.10, 0.50)
DESCS = [(1, 4), (1, 4), (2, 4), (0, 16)]                # I1 int, I2 int, R real, S char(16)
AGGS = [(0, 0), (1, 1), (1, 2), (2, 0), (3, 0)]          # COUNT, SUM(I2), SUM(R), MIN(I1), MAX(I1)
ROW_BYTES_IN = 28                                        # every referenced column read once
ROW_BYTES_PRED = 8                                       # the two compared columns (pass 1 reads only these)
ROW_BYTES_OUT = 36                                       # 8 B position + 28 B projected values


def algorithmic_bytes(nrows: int, sel: float) -> float:
    """N * (28 + 36 * s)."""
    return nrows * (ROW_BYTES_IN + ROW_BYTES_OUT * sel)


def c2_terms(term_cls, sel: float):
    r = float(np.sqrt(sel))
    t1 = int(np.ceil(r * (1 << 20)))
    t2 = float(np.float32(r * 1000.0))
    return [term_cls(1, ("col", 0), ("int", t1), 0), term_cls(1, ("col", 2), ("real", t2), 1)]    # op 1 = aopLT


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(nrows_sample: int, nthreads: int, repeats: int = 1):
    """Time the oracle (CPU restatement of TupleScan -> PredEval -> Projection) on a sample of the workload."""
    from oracle import oracle as orc
    orc.build()
    cols = [orc.synth_int(SEED, 0, nrows_sample, 1 << 20), orc.synth_int(SEED, 1, nrows_sample, 1 << 20),
            orc.synth_real(SEED, 2, nrows_sample), orc.synth_str(SEED, 3, nrows_sample, 16)]
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        for s in SELECTIVITIES:
            orc.scan(DESCS, cols, c2_terms(orc.Term, s), proj=[0, 1, 2, 3], aggs=AGGS, nthreads=nthreads)
        times.append(time.perf_counter() - t0)
    return 3 * nrows_sample / min(times), times


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = orc.max_threads()
    sample = args.cpu_rows
    cols = [orc.synth_int(SEED, 0, sample, 1 << 20), orc.synth_int(SEED, 1, sample, 1 << 20),
            orc.synth_real(SEED, 2, sample), orc.synth_str(SEED, 3, sample, 16)]

    def step():
        for s in SELECTIVITIES:
            orc.scan(DESCS, cols, c2_terms(orc.Term, s), proj=[0, 1, 2, 3], aggs=AGGS, nthreads=threads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = 3 * sample * args.steps / dt
    sample_desc = (f"{sample} rows of the C2 table per scan (3 scans/step), oracle/mbc_oracle.cpp orc_scan with "
                   f"{threads} std::threads; the Java reference itself cannot run (no JDK in the image)")
    print(json.dumps({
        "impl": "reference", "metric": "filtered_scan_rows_per_s", "value": value, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/float32/u8x16 compare, int64/float64 aggregate",
        "data": "synthetic", "config": workload_config(args, sample),
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def measured_traffic(rows_per_gpu):
    """DRAM bytes per scan (mean of the three selectivities) from the committed ncu --set full capture of this workload
    (profiles/r1/traffic.json); null for any other table size or when the file is absent."""
    path = os.path.join(ROOT, "profiles", "r1", "traffic.json")
    if rows_per_gpu != 100_000_000 or not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)["per_scan_mean"]


def workload_config(args, rows_per_gpu):
    return {"workload": "C2: synthetic 4-column table (I1,I2 int in [0,2^20); R real in [0,1000); S char(16)), "
                        "{(I1,<,t1)}^{(R,<,t2)} at 1%/10%/50% selectivity, project [I1,I2,R,S], COUNT/SUM(I2)/SUM(R)/MIN(I1)/MAX(I1)",
            "rows_per_gpu": rows_per_gpu, "scans_per_step": 3, "selectivities": list(SELECTIVITIES),
            "sharding": "TID range per rank, no data-path collective; per step ONE NCCL all-gather of every rank's aggregate/count "
                        "blocks and a gather of the 1% query's positions+values on rank 0, software-pipelined: the exchange of "
                        "step i-1 runs on a side stream while step i's scans run; the last one is drained inside the timed region"
                        if args.gpus > 1 else "single GPU",
            "l2": "inputs (2.8 GB per GPU) are larger than the 126 MB L2; no flush needed"}


class _CudaArray:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    import mbcol
    N = mbcol._native

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)       # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    ctx = mbcol.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    rows = args.rows
    table = ctx.create_table(DESCS, rows, position_base=rank * rows)
    table.generate(0, 0, SEED, 1 << 20)
    table.generate(1, 0, SEED, 1 << 20)
    table.generate(2, 1, SEED)
    table.generate(3, 2, SEED)
    terms = {s: c2_terms(mbcol.Term, s) for s in SELECTIVITIES}
    want_dev = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG
    kernel_ms = {s: [] for s in SELECTIVITIES}
    phase_ms = {s: [] for s in SELECTIVITIES}
    counts = {}

    views = {}
    # how the 1% result reaches rank 0: NCCL all-gather of padded blocks (default), grouped send/recv ("p2p"), or copies into
    # rank 0's IPC-exported buffer ("peer": measured 10.3 ms per step at N=2 on this pool -- the cross-process peer copies
    # ran at ~5 GB/s, i.e. staged through the host, not NVLink DMA -- so it is opt-in)
    gather_mode = [os.environ.get("MBC_BENCH_GATHER", "allgather")]
    peer = [None]
    if world > 1 and gather_mode[0] == "peer":
        from mbcol import sharding as _sh
        try:                                                    # 3 % of the whole table's rows x 36 B: three times the 1 % result
            peer[0] = _sh.PeerGather(int(0.03 * rows * world) * ROW_BYTES_OUT, dev)
        except Exception as e:                                  # every rank raises alike (agreed by an all-reduce inside)
            if rank == 0:
                print(f"[bench] peer-memory gather unavailable ({e}); using the NCCL all-gather form", file=sys.stderr)
            gather_mode[0] = "allgather"

    def dev_bytes(ptr, nbytes):
        """uint8 view of device memory owned by the library.  The pool hands the same buffers back step after step, so
        the torch view of (pointer, size) is built once."""
        key = (ptr, nbytes)
        v = views.get(key)
        if v is None:
            v = views[key] = torch.as_tensor(_CudaArray(ptr, nbytes), device=dev)
        return v

    def exchange_step(results):
        """mbcol.sharding over NCCL, once per step: ONE all-gather of every rank's [aggregates..., count] blocks of the
        three scans (each rank then folds COUNT/SUM/MIN/MAX on the host, which is the all-reduce), then the 1% query's
        positions + projected values are gathered on rank 0 in rank (= position) order through one all-gather of
        padded blocks (mbcol.sharding.allgather_rows); MBC_BENCH_GATHER=p2p selects the grouped send/recv form, =peer the
        copies into rank 0's IPC-exported buffer (mbcol.sharding.PeerGather)."""
        from mbcol import sharding
        mine = torch.cat([dev_bytes(r.device_pointers()["aggs"], 9 * 8) for r in results]).view(torch.int64)
        blocks = sharding.allgather_blocks(mine).view(world, len(results), 9).cpu().numpy()   # every rank sees every rank's partials
        folded = (blocks.sum(0), blocks[:, :, 2].copy().view(np.float64).sum(0), blocks[:, :, 3].min(0), blocks[:, :, 4].max(0))
        res = results[0]                                        # the 1% scan
        cnts = [int(c) for c in blocks[:, 0, 8]]
        bufs = [(res.device_pointers()["positions"], 8)] + [res.column_device(i) for i in range(4)]
        locals_ = [(dev_bytes(ptr, rows * stride)[:res.count * stride], stride) for ptr, stride in bufs]
        mode = gather_mode[0]
        if mode == "peer":                                      # rank 0: the whole table's 1% result, in position order
            gathered = peer[0].gather(locals_, cnts)            # DMA into rank 0's IPC-exported buffer over NVLink
        elif mode == "p2p":
            gathered = sharding.gather_rows_multi(locals_, cnts)
        else:
            gathered = sharding.allgather_rows(locals_, cnts)
        return folded, gathered

    dbg = {"scan": 0.0, "exchange": 0.0} if os.environ.get("MBC_BENCH_DEBUG") else None

    side = torch.cuda.Stream(device=dev) if world > 1 else None
    pending = []                                                # [(results of the previous step, event after its scans)]
    pipelined = world > 1 and os.environ.get("MBC_BENCH_PIPELINE", "1") != "0"

    def run_exchange():
        """Exchange of the PREVIOUS step's results, on a side stream: the host-side NCCL / torch work and the transfers
        overlap the scans of the current step, which are already queued on the main stream.  The results are released
        on the main stream once it has waited for the exchange."""
        while pending:
            results, scans_done = pending.pop(0)
            side.wait_event(scans_done)
            with torch.cuda.stream(side):
                exchange_step(results)
                done = torch.cuda.Event()
                done.record(side)
            stream.wait_event(done)                             # frees are stream-ordered on the main stream
            for res in results:
                res.close()

    def step(record=False):
        t0 = time.perf_counter()
        results = []
        for s in SELECTIVITIES:                                 # device-resident results complete asynchronously:
            results.append(table.scan(terms[s], proj=[0, 1, 2, 3], want=want_dev, aggs=AGGS))   # the three scans queue back to back
        tq = time.perf_counter()
        if pipelined:
            scans_done = torch.cuda.Event()
            scans_done.record(stream)
            run_exchange()                                      # step i-1's exchange while step i's scans run
        tx = time.perf_counter()
        for s, res in zip(SELECTIVITIES, results):              # the host reads every count (this is the wait)
            counts[s] = res.count
            if record:
                kernel_ms[s].append(res.kernel_ms)
                phase_ms[s].append(res.phase_ms)
        t1 = time.perf_counter()
        if pipelined:
            pending.append((results, scans_done))
        else:
            if world > 1:
                exchange_step(results)
            for res in results:
                res.close()
        if dbg is not None and record:
            dbg["enqueue"] = dbg.get("enqueue", 0.0) + tq - t0
            dbg["scan"] += t1 - tx + tq - t0
            dbg["exchange"] += (tx - tq) if pipelined else (time.perf_counter() - t1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    if pipelined:
        run_exchange()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(record=True)
    if pipelined:
        run_exchange()                                          # the last step's exchange is inside the timed region
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - launches0
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    value = 3.0 * rows * world * args.steps / (ms_total * 1e-3)

    # ---- end to end: host-resident columns -> mbc_scan_host -> host-resident results -------------------
    host_cols = []
    for c, (t, w) in enumerate(DESCS):
        dt = np.int32 if t == 1 else np.float32 if t == 2 else np.uint8
        shape = (rows,) if t != 0 else (rows, w)
        pinned = torch.empty(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=torch.uint8, pin_memory=True).numpy()
        arr = pinned.view(dt).reshape(shape)
        arr[...] = table.read_column(c)
        host_cols.append(arr)
    want_host = N.WANT_POSITIONS | N.WANT_COLUMNS | N.WANT_AGG | N.WANT_HOST
    h2d = d2h = 0

    def e2e_step(count_bytes=False):
        nonlocal d2h, h2d
        moved0 = ctx.h2d_bytes
        for s in SELECTIVITIES:
            res = ctx.scan_host(DESCS, host_cols, terms[s], proj=[0, 1, 2, 3], want=want_host, aggs=AGGS,
                                position_base=rank * rows)
            if count_bytes:
                d2h += res.count * ROW_BYTES_OUT + 8 * (len(AGGS) + 1)
            res.close()
        if count_bytes:
            h2d = ctx.h2d_bytes - moved0           # counted by the library: selective scans upload the predicate columns only

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(e2e_steps):
        e2e_step(count_bytes=(i == 0))
    e1.record(stream)
    barrier()
    e2e_wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    e2e_t = torch.tensor([max(e2e_ms, e2e_wall_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = 3.0 * rows * world * e2e_steps / (float(e2e_t.item()) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    if dbg is not None:
        print(f"[bench debug] rank {rank}: per step ms: " + ", ".join(f"{k}={1e3 * v / args.steps:.3f}" for k, v in dbg.items()), file=sys.stderr)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        per_sel = {}
        tot_bytes = tot_ms = 0.0
        for s in SELECTIVITIES:
            m = statistics.mean(kernel_ms[s])
            b = algorithmic_bytes(rows, s)
            per_sel[str(s)] = {"kernel_ms": m, "rows_per_s": rows / (m * 1e-3), "achieved_gbs": b / (m * 1e-3) / 1e9,
                               "frac_of_measured_peak": b / (m * 1e-3) / 1e9 / peak,
                               "frac_of_8000": b / (m * 1e-3) / 1e9 / 8000.0, "count": counts[s]}
            tot_bytes += b
            tot_ms += m
        achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
        # per kernel (CUDA events between the launches of every scan): its own algorithmic bytes over its own time
        kernels = {}
        names = ("filter_kernel", "tile_offsets_kernel", "write_kernel", "agg_finish_kernel")
        for s in SELECTIVITIES:
            ph = np.asarray(phase_ms[s], dtype=np.float64).mean(0)
            alg = {"filter_kernel": rows * (ROW_BYTES_PRED + 1 / 8),                       # predicate columns in, bitmap out
                   "write_kernel": rows * (1 / 8 + s * (ROW_BYTES_IN + ROW_BYTES_OUT))}    # bitmap + survivors' values in, rows out
            for i, name in enumerate(names):
                k = kernels.setdefault(name, {"ms_per_step": 0.0, "algorithmic_bytes_per_step": 0.0, "per_selectivity_ms": {}})
                k["ms_per_step"] += float(ph[i])
                k["per_selectivity_ms"][str(s)] = float(ph[i])
                k["algorithmic_bytes_per_step"] += alg.get(name, 0.0)
        for name, k in kernels.items():
            k["achieved_gbs"] = k["algorithmic_bytes_per_step"] / (k["ms_per_step"] * 1e-3) / 1e9 if k["ms_per_step"] > 0 else None
            k["frac"] = k["achieved_gbs"] / peak if k["achieved_gbs"] else None
            k["share_of_step"] = k["ms_per_step"] / tot_ms
        dominant = max(kernels, key=lambda n: kernels[n]["ms_per_step"])
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import oracle as orc
            v1, _ = cpu_reference_run(args.cpu_rows // 4, 1)
            vn, _ = cpu_reference_run(args.cpu_rows, orc.max_threads())
            cpu = {"value": vn, "unit": "rows/s", "cores": orc.max_threads(), "kind": "port",
                   "single_thread_rows_per_s": v1,
                   "sample": f"{args.cpu_rows} rows of the C2 table per scan x 3 selectivities (all threads) and "
                             f"{args.cpu_rows // 4} rows (1 thread); oracle/mbc_oracle.cpp orc_scan"}
        out = {
            "metric": "filtered_scan_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32/float32/u8x16 compare, int64/float64 aggregate", "data": "synthetic",
            "config": workload_config(args, rows),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "mbc_scan_host (pinned host columns -> chunked H2D + scan -> D2H results; selective scans read the survivors' projected values in place)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "mbc scan = filter_kernel + tile_offsets_kernel + write_kernel + agg_finish_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0, "peak_source": peak_src,
                         "traffic": measured_traffic(rows), "algorithmic_bytes_per_launch": tot_bytes / 3,
                         "per_selectivity": per_sel, "dominant_kernel": dominant, "kernels": kernels},
            "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    table.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rows", type=int, default=100_000_000, help="rows per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rows", type=int, default=8_000_000, help="rows of the CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29577",
                                   os.path.abspath(__file__)] + sys.argv[1:])
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
