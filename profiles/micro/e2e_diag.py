import sys, os, time, threading
sys.path[:0] = ["/root/repo", "/root/repo/nightcore-to-flac-analyzer_b200"]
import numpy as np, torch
import bench
from nightcore_analyzer import _engine, batch as nbatch
eng = _engine.get_engine()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
sub = 125
distinct = bench.make_pairs(8, 180.0)
pairs_sub = [distinct[j % len(distinct)] for j in range(sub)]
pinned = nbatch.pin_pairs(pairs_sub, 22050)
sizes = nbatch.plan_subbatches(n, sub, 2)
log = []
lock = threading.Lock()
g0 = torch.cuda.Event(enable_timing=True)
def run(sizes, workers=2):
    bufs = [[torch.empty(int(pinned.pinned.numel()), dtype=torch.float32, device=eng.device) for _ in range(2)] for _ in range(workers)]
    def prepare(k, w, i):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); th = time.perf_counter()
        st = nbatch.upload(pinned, k, out=bufs[w][i % 2])
        b.record()
        with lock: log.append(["up", w, k, a, b, th, time.perf_counter()])
        return st
    def one(st):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        th = time.perf_counter(); a.record()
        r = nbatch.analyse_staged(st)
        b.record()
        with lock: log.append(["fn", threading.get_ident() % 1000, st.n_pairs, a, b, th, time.perf_counter()])
        return r
    return nbatch.run_subbatches(list(sizes), one, workers, prepare=prepare)
for rep in range(3):
    log.clear()
    torch.cuda.synchronize(); g0.record(); T0 = time.perf_counter()
    run(sizes)
    torch.cuda.synchronize(); T1 = time.perf_counter()
    print("rep", rep, "wall ms", round(1e3 * (T1 - T0), 1), "sizes", sizes)
for kind, w, k, a, b, th0, th1 in sorted(log, key=lambda r: g0.elapsed_time(r[3])):
    print(f"{kind} w={w:3d} k={k:4d} gpu[{g0.elapsed_time(a):8.1f} -> {g0.elapsed_time(b):8.1f}] = {a.elapsed_time(b):7.1f} ms   host[{1e3*(th0-T0):8.1f} -> {1e3*(th1-T0):8.1f}]")
# resident reference
res = [nbatch.upload(pinned, k) for k in [125] * (n // 125)]
torch.cuda.synchronize()
for rep in range(3):
    T0 = time.perf_counter(); nbatch.analyse_resident(res, workers=2); torch.cuda.synchronize(); print("resident wall ms", round(1e3 * (time.perf_counter() - T0), 1))
