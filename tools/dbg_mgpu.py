import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nightcore-to-flac-analyzer_b200"))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from nightcore_analyzer import batch as nbatch, _engine
from oracle import synth
pairs = []
for i in range(4):
    src, nc = synth.make_pair(5000 + i, 180.0, 22050); pairs.append((nc, src))
sub = 63
pinned = nbatch.pin_pairs([pairs[j % 4] for j in range(sub)], 22050)
def bar(tag):
    try:
        dist.barrier(); torch.cuda.synchronize(); print(rank, tag, "barrier ok", flush=True)
    except Exception as e:
        print(rank, tag, "barrier FAILED:", str(e).splitlines()[0], flush=True)
        try:
            dist.barrier(); torch.cuda.synchronize(); print(rank, tag, "retry ok", flush=True)
        except Exception as e2:
            print(rank, tag, "retry FAILED", str(e2).splitlines()[0], flush=True)
bar("start")
for sizes, workers in (([16, 16], 2), ([63], 1), ([63, 63], 1), ([63, 63], 2), ([15, 15, 63, 63], 2)):
    st = {}
    res = nbatch.analyse_pinned(pinned, sizes, stats=st, workers=workers, compute_pitch=True, compute_ibi=True)
    torch.cuda.synchronize()
    print(rank, sizes, workers, "ok", sum(1 for r in res if not isinstance(r, Exception)), flush=True)
    bar(f"after {sizes} w{workers}")
dist.destroy_process_group()
