"""Sharding of a pair batch over the GPUs of one box, and the single NCCL gather of the per-pair
result records.  Pairs are independent: rank r of G analyses pairs {i : i mod G == r}
with no data-path collective; the only exchange is one all_gather of fixed-size records at the end."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

RECORD_F64 = 16   # float64 slots per pair record


def bind_to_gpu_cpus(device_index: int) -> List[int]:
    """Restrict the calling process to the host cores NVML reports as local to GPU ``device_index`` (its NUMA node) and
    return them ([] when NVML or the affinity call is unavailable — nothing changes then).  Call it before the first
    pinned allocation: pinned pages are then first-touched on the GPU's own node, and with one process per GPU the
    host→device copies of the ranks stop crossing the socket interconnect."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 — an optimisation only
        return []


def use_blocking_sync(device_index: int) -> bool:
    """Make host threads that wait for ``device_index`` sleep instead of spin (cudaDeviceScheduleBlockingSync).  Must run
    before the process creates its CUDA context on that device; returns False (and changes nothing) if the runtime
    refuses.  NOT used by default: the hypothesis that spinning waits starve the ranks' Python threads at N = 8 did not
    hold — a rank restricted to 4 (2) host cores runs its 125-pair share end to end at 920 (835) pairs/s with spinning
    waits and at 866 (733) with blocking ones (profiles/r3d_blocking_sync.log)."""
    import ctypes
    try:
        import torch  # noqa: F401 — loads the CUDA runtime this process will use
        rt = ctypes.CDLL("libcudart.so.12")
        if rt.cudaSetDevice(int(device_index)) != 0:
            return False
        return rt.cudaSetDeviceFlags(4) == 0          # cudaDeviceScheduleBlockingSync
    except Exception:  # noqa: BLE001 — an optimisation only
        return False


def shard_indices(n_pairs: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_pairs, world))


def result_record(res) -> np.ndarray:
    """Fixed-size numeric record of one pair's result (NaN = absent; status 1 = failed)."""
    r = np.full(RECORD_F64, np.nan)
    if isinstance(res, Exception) or res is None:
        r[0] = 1.0
        return r
    r[0] = 0.0
    r[1], r[2], r[3] = res.tempo_ratio, res.tempo_ci[0], res.tempo_ci[1]
    r[4], r[5], r[6] = res.pitch_ratio, res.pitch_ci[0], res.pitch_ci[1]
    if res.ibi_ratio is not None:
        r[7], r[8], r[9] = res.ibi_ratio, res.ibi_ci[0], res.ibi_ci[1]
    r[10], r[11] = res.n_source_tempo_windows, res.n_nc_tempo_windows
    r[12], r[13] = res.n_source_pitch_windows, res.n_nc_pitch_windows
    r[14] = res.nc_median_bpm if res.nc_median_bpm is not None else np.nan
    r[15] = res.src_median_bpm if res.src_median_bpm is not None else np.nan
    return r


def records_of(results: Sequence) -> np.ndarray:
    return np.stack([result_record(r) for r in results]) if len(results) else np.zeros((0, RECORD_F64))


def result_bytes(stats: dict) -> int:
    """Device→host bytes of one step: per-window lags and beat counts, hop-64 beat frames, bootstrap triples."""
    return int(stats.get("d2h_bytes", 0))


def gather_result_records(results: Sequence, my_ids: Sequence[int], n_pairs: int, backend_device=None) -> np.ndarray:
    """all_gather the per-pair records of every rank into pair order → float64 [n_pairs, RECORD_F64].
    Uses the default process group (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    per_rank = (n_pairs + world - 1) // world
    use_cuda = dist.get_backend() == "nccl"
    dev = backend_device if backend_device is not None else (torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu"))
    mine = torch.full((per_rank, RECORD_F64 + 1), float("nan"), dtype=torch.float64)
    rec = records_of(results)
    for j, i in enumerate(my_ids):
        mine[j, 0] = float(i)
        mine[j, 1:] = torch.from_numpy(rec[j])
    mine = mine.to(dev)
    out = torch.empty((world * per_rank, RECORD_F64 + 1), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, mine)
    out = out.cpu().numpy()
    table = np.full((n_pairs, RECORD_F64), np.nan)
    for row in out:
        if not np.isnan(row[0]):
            table[int(row[0])] = row[1:]
    return table


# algorithmic figures of SURVEY.md §8(d) / DESIGN.md
BYTES_PER_WINDOW = 220500 * 4 + 431 * 4      # one 10 s analysis window → f32[431] onset envelope
BYTES_PER_HOP64_FRAME = 64 * 4 + 4           # one hop-64 frame of the whole-track pass
FLOP_PER_FRAME = 64.7e3                      # rFFT-2048 + power + sparse mel + log + flux


def frontend_roofline(prof: dict, stats: dict) -> dict:
    """Roofline of the dominant kernel — the hop-64 launches of ``stft_logmel_kernel`` (the whole-track onset front-end,
    tempo.py:158) — from one profiled step: achieved GB/s = algorithmic bytes per launch (frames × 260 B) ÷
    average launch duration (CUDA events on the launching stream).  The hop-512 window launches and ``flux_kernel`` are
    reported in the per-kernel table and in ``front_end`` (all STFT + flux launches together)."""
    total_ms = sum(v[1] for v in prof.values()) or 1e-9
    table = {k: {"launches": v[0], "ms": round(v[1], 4), "share": round(v[1] / total_ms, 4)}
             for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    windows = stats.get("windows", 0)
    frames64 = stats.get("hop64_frames", 0)
    k64 = prof.get("stft_logmel_kernel[hop<=128]", (0, 0.0))
    k512 = prof.get("stft_logmel_kernel[hop>128]", (0, 0.0))
    flux = prof.get("flux_kernel", (0, 0.0))
    n64, ms64 = max(k64[0], 1), max(k64[1], 1e-9)
    bytes64 = frames64 * BYTES_PER_HOP64_FRAME
    fe_ms = max(k64[1] + k512[1] + flux[1], 1e-9)
    fe_bytes = windows * BYTES_PER_WINDOW + bytes64
    return {
        "kernel": "stft_logmel_kernel, hop-64 whole-track launches",
        "launches": k64[0],
        "bytes_per_launch": bytes64 / n64,
        "ms_per_launch": ms64 / n64,
        "gbs": bytes64 / (ms64 * 1e-3) / 1e9,
        "tflops": frames64 * FLOP_PER_FRAME / (ms64 * 1e-3) / 1e12,
        "share": ms64 / total_ms,
        "front_end": {"kernels": "stft_logmel_kernel (both hops) + flux_kernel", "ms": fe_ms, "bytes": int(fe_bytes),
                      "gbs": fe_bytes / (fe_ms * 1e-3) / 1e9,
                      "tflops": (windows * 431 + frames64) * FLOP_PER_FRAME / (fe_ms * 1e-3) / 1e12,
                      "share": fe_ms / total_ms},
        "table": table,
    }
