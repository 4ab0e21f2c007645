import torch, time
n = 1<<30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
s = torch.cuda.Stream()
for rep in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(4): d.copy_(h, non_blocking=True)
        e1.record(s)
    e1.synchronize()
    print("H2D pinned GB/s", 4*n/1e9/(e0.elapsed_time(e1)/1e3))
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device='cuda'); s2 = torch.cuda.Stream()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
s.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(4): d.copy_(h, non_blocking=True)
    e1.record(s)
with torch.cuda.stream(s2):
    for _ in range(4): d2.copy_(h2, non_blocking=True)
    f1.record(s2)
torch.cuda.synchronize()
print("two concurrent H2D streams: total GB/s", 8*n/1e9/(max(e0.elapsed_time(e1), e0.elapsed_time(f1))/1e3))
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv", shell=True, capture_output=True, text=True).stdout)
print(subprocess.run("nproc; free -g | head -2", shell=True, capture_output=True, text=True).stdout)
