#!/usr/bin/env python
"""Registers / spills / static shared memory of every kernel (nvcc -Xptxas=-v), as a markdown table.
    python profiles/ptxas_resources.py > profiles/r2_ptxas_resources.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
print("# ptxas resource usage per kernel (sm_100a, `nvcc -Xptxas=-v`, flags of `__graft_entry__.build()`)\n")
print("| source | kernel | registers | spill stores / loads (B) | stack (B) | static smem (B) | barriers |")
print("|---|---|---|---|---|---|---|")
for src, extra in ge.SOURCES.items():
    cmd = [ge.NVCC] + ge.ARCH + ge.COMMON + extra + ["-Xptxas=-v", "-c", os.path.join(ge.CSRC, src), "-o", "/dev/null"]
    err = subprocess.run(cmd, capture_output=True, text=True).stderr
    names = []
    blocks = err.split("Compiling entry function '")[1:]
    for b in blocks:
        mangled = b.split("'")[0]
        m1 = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
        m2 = re.search(r"Used (\d+) registers", b)
        bars = re.search(r"used (\d+) barriers", b)
        smem = re.search(r"(\d+) bytes smem", b)
        if not (m1 and m2):
            continue
        names.append((mangled, m1.group(1), m1.group(2), m1.group(3), m2.group(1), bars.group(1) if bars else "0", None,
                      smem.group(1) if smem else "0"))
    dem = subprocess.run(["c++filt"], input="\n".join(n[0] for n in names), capture_output=True, text=True).stdout.split("\n")
    for (mangled, stack, sst, sld, regs, bars, cum, smem), d in sorted(zip(names, dem), key=lambda x: x[1]):
        d = re.sub(r"\(.*", "", d).replace("void ", "").replace("ncfa::", "")
        print(f"| {src} | `{d}` | {regs} | {sst} / {sld} | {stack} | {smem or 0} | {bars or 0} |")
