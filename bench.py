#!/usr/bin/env python
"""Benchmark of the windowed spectral front-end on B200 (BASELINE.json metric: analysis windows/sec
and track-pairs/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the CPU path on the host cores

A step = one full analysis (strip → slice → gate → pitch → tempo src → prior → tempo nc → hop-64
IBI pass → bootstraps → result assembly) of a batch of synthetic track pairs of BASELINE config 5
(180 s source at 22 050 Hz + 1.25× nightcore).  `value` is measured with the batch resident in HBM;
`e2e` goes through the public batch API with host buffers (pinned → H2D inside the timed region,
results back on the host).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nightcore-to-flac-analyzer_b200"))

import numpy as np  # noqa: E402

SR = 22050
PAIR_SEC = 180.0
N_DISTINCT = 8          # distinct synthetic pairs, tiled to the batch size (stated in `data`)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def load_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/ncu_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)[kernel]
        return int(t["dram_bytes_per_launch"]), t["source"]
    except Exception:
        return None, None


def make_pairs(n_distinct: int, dur: float):
    from oracle import synth
    pairs = []
    for i in range(n_distinct):
        src, nc = synth.make_pair(5000 + i, dur, SR)
        pairs.append((nc, src))
    return pairs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------- CPU arm
def _cpu_one_pair(args):
    """One pair through the CPU port in a worker process.  BLAS/OpenMP pools are pinned to one thread so that P
    processes use P cores (without this the einsum/BLAS pools oversubscribe the box and the baseline is ~3x slower)."""
    nc, src = args
    from oracle import pipeline_port
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        threadpool_limits = None
    t0 = time.perf_counter()
    if threadpool_limits is not None:
        with threadpool_limits(limits=1):
            res, n_windows = pipeline_port.run_arrays(nc, src, SR, log=None, return_window_count=True)
    else:
        res, n_windows = pipeline_port.run_arrays(nc, src, SR, log=None, return_window_count=True)
    return n_windows, time.perf_counter() - t0


def cpu_throughput(n_pairs: int, dur: float, procs: int):
    """windows/s and pairs/s of the CPU restatement (oracle port) on `procs` host processes."""
    import multiprocessing as mp
    pairs = make_pairs(min(n_pairs, N_DISTINCT), dur)
    work = [pairs[i % len(pairs)] for i in range(n_pairs)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    if procs > 1:
        with ctx.Pool(procs) as pool:
            out = pool.map(_cpu_one_pair, work, chunksize=1)
    else:
        out = [_cpu_one_pair(w) for w in work]
    wall = time.perf_counter() - t0
    windows = sum(o[0] for o in out)
    return windows / wall, n_pairs / wall, wall, windows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.cpu_procs or cores))
    n_pairs = procs
    dur = args.cpu_dur
    vals, pps, walls = [], [], []
    for s in range(args.warmup + args.steps):
        w, p, wall, nwin = cpu_throughput(n_pairs, dur, procs)
        if s >= args.warmup:
            vals.append(w)
            pps.append(p)
            walls.append(wall)
    v = float(np.mean(vals))
    sample = f"{n_pairs} pairs of {dur:.0f}s source + 1.25x nightcore per step, one pair per process"
    line = {
        "impl": "reference", "metric": "analysis_windows_per_sec", "value": v, "unit": "windows/s",
        "pairs_per_sec": float(np.mean(pps)), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs": n_pairs, "pair_sec": dur, "sr": SR},
        "cpu_baseline": {"value": v, "unit": "windows/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


WORKLOAD = ("config 5: batch of 1000 synthetic 180 s track pairs (22050 Hz source + 1.25x nightcore), "
            "full windowed tempo+pitch+IBI+bootstrap analysis, pairs sharded across ranks")


# --------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from nightcore_analyzer import _engine, batch as nbatch, parallel as npar

    eng = _engine.get_engine()
    total_pairs = args.pairs
    my_ids = npar.shard_indices(total_pairs, rank, world)
    distinct = make_pairs(min(N_DISTINCT, max(1, total_pairs)), args.pair_sec)
    kw = dict(compute_pitch=not args.no_pitch, compute_ibi=not args.no_ibi)

    # The rank's pairs are analysed in sub-batches (the batch scheduler of DESIGN.md): every sub-batch has the
    # same composition (distinct pairs tiled), so ONE pinned host buffer feeds all of them.
    sub = max(1, min(args.sub_batch, len(my_ids))) if my_ids else 1
    sizes = [min(sub, len(my_ids) - s) for s in range(0, len(my_ids), sub)]        # resident sub-batches
    sub_e2e = min(sub, max(16, -(-len(my_ids) // 4)))                            # at least ~4 pieces per rank
    sizes_e2e = nbatch.plan_subbatches(len(my_ids), sub_e2e, args.workers)       # short head: first kernels start early
    pairs_sub = [distinct[j % len(distinct)] for j in range(sub)]
    pinned = nbatch.pin_pairs(pairs_sub, SR)
    resident = [nbatch.upload(pinned, k) for k in sizes]       # `value`: inputs already in HBM
    torch.cuda.synchronize()
    resident_bytes = sum(st.h2d_bytes for st in resident)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def merge(total, stats):
        for k, v in stats.items():
            total[k] = total.get(k, 0) + v

    def step_resident():
        stats = {}
        res = nbatch.analyse_resident(resident, stats=stats, workers=args.workers, **kw)
        return res, stats

    def step_e2e():
        # pinned host → HBM inside the timed region; results come back as host objects.  Sub-batches are dealt to
        # `--workers` host threads / CUDA streams, so copies and host stages of one overlap kernels of another.
        stats = {}
        res = nbatch.analyse_pinned(pinned, sizes_e2e, stats=stats, workers=args.workers, **kw)
        return res, stats, stats["h2d_bytes"]

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    l0 = _engine.total_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record()
        for _ in range(args.steps):
            res, stats = step_resident()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = (_engine.total_launches() - l0) // max(1, args.steps)
    n_ok = sum(1 for r in res if not isinstance(r, Exception))
    windows = stats["windows"]

    # ---- per-kernel shares and the front-end roofline: one extra step with the library's event profiler on
    from nightcore_analyzer import _native
    # (single host worker here: with two streams the event brackets of one stream would include the other's kernels)
    _native.lib.ncfa_profile_enable(1)
    pstats = {}
    nbatch.analyse_resident(resident, stats=pstats, workers=1, **kw)
    torch.cuda.synchronize()
    prof = _native.profile_report()
    _native.lib.ncfa_profile_enable(0)
    roof = npar.frontend_roofline(prof, pstats)

    # ---- end-to-end timing through the host-buffer API
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        res_e, stats_e, h2d = step_e2e()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / args.steps
    d2h = npar.result_bytes(stats_e)

    # ---- max over ranks, totals over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        c = torch.tensor([windows, len(my_ids), n_ok, h2d, d2h, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        windows, n_pairs_all, n_ok, h2d, d2h, launches = (int(x) for x in c.tolist())
        gathered = npar.gather_result_records(res, my_ids, total_pairs)   # the NCCL gather of per-pair records
    else:
        n_pairs_all = len(my_ids)

    if rank == 0:
        peak, peak_kind = load_peaks()
        traffic, traffic_src = load_traffic("stft_logmel_kernel[hop<=128]")
        value = windows / (ms / 1e3)
        line = {
            "metric": "analysis_windows_per_sec", "value": value, "unit": "windows/s",
            "pairs_per_sec": n_pairs_all / (ms / 1e3), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": f"synthetic ({len(distinct)} distinct pairs tiled to {total_pairs}; oracle/synth.py)",
            "config": {"workload": WORKLOAD, "pairs": total_pairs, "pair_sec": args.pair_sec, "sr": SR,
                       "sub_batch_pairs": sub, "e2e_sub_batches": sizes_e2e, "host_workers": args.workers, "windows_per_step": windows, "pairs_ok": n_ok,
                       "pitch": not args.no_pitch, "ibi": not args.no_ibi,
                       "l2": "inputs larger than L2 (resident audio per rank %.0f MB > 126 MB)" % (resident_bytes / 1e6)
                       if resident_bytes > (126 << 20) else "inputs smaller than L2 (reduced --pairs run)"},
            "e2e": {"value": windows / (ms_e2e / 1e3), "unit": "windows/s", "ms_per_step": ms_e2e,
                    "pairs_per_sec": n_pairs_all / (ms_e2e / 1e3), "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": roof["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": roof["gbs"] / peak, "traffic": traffic, "peak_kind": peak_kind,
                         "kernel": roof["kernel"], "launches": roof["launches"],
                         "algorithmic_bytes_per_launch": roof["bytes_per_launch"],
                         "ms_per_launch": roof["ms_per_launch"], "traffic_source": traffic_src,
                         "fp32_tflops": roof["tflops"], "fp32_peak_tflops": 74.5,
                         "fp32_frac": roof["tflops"] / 74.5, "share_of_step": roof["share"],
                         "note": "compute bound (AI 249 flop/B vs ridge 11.5): the FP32 fraction is the binding one",
                         "front_end": roof["front_end"]},
            "kernels": roof["table"],
            "clocks": clk.summary(),
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            procs = max(1, min(cores, args.cpu_procs or cores))
            w, p, wall, nwin = cpu_throughput(procs, args.cpu_dur, procs)
            line["cpu_baseline"] = {"value": w, "unit": "windows/s", "cores": procs, "kind": "port",
                                    "pairs_per_sec": p,
                                    "sample": f"{procs} pairs of {args.cpu_dur:.0f}s source + 1.25x nightcore, one pair "
                                              f"per process, {wall:.1f}s wall"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_single_pair(args):
    """BASELINE config 1 (the reference's own CPU-runnable case): latency of ONE 3-minute pair through the drop-in
    `pipeline.run_arrays` (host arrays in, AnalysisResult out) beside the CPU port on one core.  Not the contract line."""
    import torch
    from nightcore_analyzer import pipeline as npipe
    from oracle import pipeline_port
    nc, src = make_pairs(1, args.pair_sec)[0]
    for _ in range(2):
        res = npipe.run_arrays(nc, src, SR, log=None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = npipe.run_arrays(nc, src, SR, log=None)
    torch.cuda.synchronize()
    gpu_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    t0 = time.perf_counter()
    want = pipeline_port.run_arrays(nc, src, SR, log=None)
    cpu_s = time.perf_counter() - t0
    same = (res.tempo_ratio, res.tempo_ci) == want["tempo"] and (res.ibi_ratio, res.ibi_ci) == want["ibi"] and \
        res.nc_pitches_raw == want["nc_hz"]
    print(json.dumps({"metric": "single_pair_latency_ms", "value": gpu_ms, "unit": "ms", "config": {
        "workload": "config 1: single synthetic 3-min pair, full pipeline through pipeline.run_arrays", "pair_sec": args.pair_sec},
        "cpu_port_seconds_1core": cpu_s, "speedup_vs_1core": cpu_s * 1e3 / gpu_ms, "identical_to_cpu_port": bool(same),
        "tempo_ratio": res.tempo_ratio, "ibi_ratio": res.ibi_ratio, "classification": res.classification}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--pairs", type=int, default=1000, help="total track pairs per step (all ranks)")
    ap.add_argument("--pair-sec", type=float, default=PAIR_SEC)
    ap.add_argument("--sub-batch", type=int, default=125, help="pairs analysed per device pass (per rank)")
    ap.add_argument("--workers", type=int, default=2, help="host threads / CUDA streams the sub-batches are dealt to")
    ap.add_argument("--no-pitch", action="store_true")
    ap.add_argument("--no-ibi", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-dur", type=float, default=30.0, help="source seconds per CPU-baseline pair")
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--single-pair", action="store_true", help="config 1: latency of one pair through pipeline.run_arrays")
    args = ap.parse_args()
    if args.single_pair:
        return run_single_pair(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
