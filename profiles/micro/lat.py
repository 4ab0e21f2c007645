import sys, time, threading
sys.path[:0] = ["/root/repo", "/root/repo/nightcore-to-flac-analyzer_b200"]
import numpy as np, torch
from nightcore_analyzer import _engine
eng = _engine.get_engine()
big_h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
big_d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
cs = torch.cuda.Stream()
stop = False
def bg():
    torch.cuda.set_device(0)
    while not stop:
        with torch.cuda.stream(cs):
            for _ in range(4): big_d.copy_(big_h, non_blocking=True)
        cs.synchronize()
a = np.arange(256, dtype=np.int64)
d = torch.zeros(256, dtype=torch.float64, device="cuda")
pin_out = torch.empty(256, dtype=torch.float64, pin_memory=True)
def bench(name, f, n=200):
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    print(f"  {name:38s} median {np.median(ts):7.3f} ms   p95 {np.percentile(ts,95):7.3f}   max {ts.max():7.3f}")
def run_all():
    bench("to_dev (ring+kernel) + sync", lambda: (eng.to_dev(a), torch.cuda.current_stream().synchronize()))
    bench("pageable .to(device) + sync", lambda: (torch.from_numpy(a).to("cuda", non_blocking=True), torch.cuda.current_stream().synchronize()))
    bench("to_host (.cpu())", lambda: d.cpu())
    bench("D2H into pinned + sync", lambda: (pin_out.copy_(d, non_blocking=True), torch.cuda.current_stream().synchronize()))
    bench("kernel only (d.add_) + sync", lambda: (d.add_(1.0), torch.cuda.current_stream().synchronize()))
    bench("torch.zeros small + sync", lambda: (torch.zeros(1000, device="cuda"), torch.cuda.current_stream().synchronize()))
    bench("cudaMalloc-ish empty 300MB + sync", lambda: (torch.empty(300 << 20, dtype=torch.uint8, device="cuda"), torch.cuda.current_stream().synchronize()), n=50)
print("idle link"); run_all()
th = threading.Thread(target=bg); th.start(); time.sleep(0.3)
print("with background bulk H2D"); run_all()
stop = True; th.join()
