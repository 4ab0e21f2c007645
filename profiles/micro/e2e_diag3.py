import sys, os, time, threading
sys.path[:0] = ["/root/repo", "/root/repo/nightcore-to-flac-analyzer_b200"]
import numpy as np, torch
import bench
from nightcore_analyzer import _engine, batch as nbatch, _native
eng = _engine.get_engine()
sub = 125
distinct = bench.make_pairs(8, 180.0)
pairs_sub = [distinct[j % len(distinct)] for j in range(sub)]
pinned = nbatch.pin_pairs(pairs_sub, 22050)
res = [nbatch.upload(pinned, 125) for _ in range(4)]
torch.cuda.synchronize()
def timeit(f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return round(1e3 * best, 1)
print("resident alone", timeit(lambda: nbatch.analyse_resident(res, workers=2)))
stop = False
dst = torch.empty_like(res[0].audio)
cs = torch.cuda.Stream()
def bg(kind):
    torch.cuda.set_device(0)
    n = 0
    while not stop:
        with torch.cuda.stream(cs):
            if kind == "h2d":
                dst.copy_(pinned.pinned[: dst.numel()], non_blocking=True)
            elif kind == "d2d":
                dst.copy_(res[1].audio, non_blocking=True)
        cs.synchronize(); n += 1
    print("  bg", kind, "copies", n)
for kind in ("h2d", "d2d"):
    stop = False
    th = threading.Thread(target=bg, args=(kind,)); th.start()
    time.sleep(0.2)
    print("resident with background", kind, timeit(lambda: nbatch.analyse_resident(res, workers=2)))
    stop = True; th.join()
# per-kernel profile with and without background h2d (single worker so event brackets are clean)
for kind in (None, "h2d"):
    stop = False
    if kind:
        th = threading.Thread(target=bg, args=(kind,)); th.start(); time.sleep(0.2)
    _native.lib.ncfa_profile_enable(1)
    nbatch.analyse_resident(res[:2], workers=1)
    torch.cuda.synchronize()
    prof = _native.profile_report()
    _native.lib.ncfa_profile_enable(0)
    if kind:
        stop = True; th.join()
    print("profile bg=", kind, {k: round(v[1], 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:10]})
