"""Where does the device idle inside a resident step?  One profiled step of analyse_resident with W host workers; every
launch's (start, end) on the device clock → union busy time, per-stream busy time, the largest idle gaps and what ran
before / after them.    python profiles/micro/timeline.py [pairs] [workers]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nightcore-to-flac-analyzer_b200"))
import numpy as np, torch
import bench
from nightcore_analyzer import _native, batch as nbatch

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 500
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.cuda.set_device(0)
distinct = bench.make_pairs(range(32), 180.0)
pairs_np = [distinct[i % 32] for i in range(pairs)]
pinned = nbatch.pin_pairs(pairs_np, 22050)
sizes = nbatch.stagger_sizes(pairs, min(125, max(16, -(-pairs // 4))), workers)
starts = [sum(sizes[:j]) for j in range(len(sizes))]
resident = [nbatch.upload(pinned, k, start_pair=s) for s, k in zip(starts, sizes)]
for _ in range(2):
    nbatch.analyse_resident(resident, workers=workers)
torch.cuda.synchronize()
t0 = time.perf_counter(); nbatch.analyse_resident(resident, workers=workers); torch.cuda.synchronize()
plain = 1e3 * (time.perf_counter() - t0)
_native.lib.ncfa_profile_enable(1)
t0 = time.perf_counter(); nbatch.analyse_resident(resident, workers=workers); torch.cuda.synchronize()
wall = 1e3 * (time.perf_counter() - t0)
tl = _native.profile_timeline(); _native.lib.ncfa_profile_enable(0)
tl.sort(key=lambda r: r[2])
lo, hi = min(r[2] for r in tl), max(r[3] for r in tl)
# union of busy intervals
busy, cur_s, cur_e, gaps = 0.0, None, None, []
last_name = None
for name, st, a, b in tl:
    if cur_e is None:
        cur_s, cur_e, last_name = a, b, name
    elif a <= cur_e:
        if b > cur_e: cur_e, last_name = b, name
    else:
        gaps.append((a - cur_e, cur_e - lo, last_name, name)); busy += cur_e - cur_s; cur_s, cur_e, last_name = a, b, name
busy += cur_e - cur_s
print(f"pairs {pairs} workers {workers}: wall unprofiled {plain:.1f} ms, profiled {wall:.1f} ms, device span {hi-lo:.1f} ms, union busy {busy:.1f} ms, "
      f"idle {hi-lo-busy:.1f} ms in {len(gaps)} gaps; launches {len(tl)}; sum of kernel times {sum(r[3]-r[2] for r in tl):.1f} ms")
for st in sorted(set(r[1] for r in tl)):
    rows = [r for r in tl if r[1] == st]
    print(f"  stream {st}: {len(rows)} launches, busy {sum(r[3]-r[2] for r in rows):.1f} ms")
hist = {}
for g, at, before, after in gaps:
    k = (before, after); hist[k] = (hist.get(k, (0, 0.0))[0] + 1, hist.get(k, (0, 0.0))[1] + g)
print("idle time by (kernel before gap -> kernel after gap), top 25:")
for (b, a), (n, ms) in sorted(hist.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {ms:8.2f} ms in {n:4d} gaps   {b} -> {a}")
print("largest single gaps:")
for g, at, before, after in sorted(gaps, reverse=True)[:15]:
    print(f"  {g:7.3f} ms at t={at:8.2f}  {before} -> {after}")
