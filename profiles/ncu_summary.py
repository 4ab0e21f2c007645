#!/usr/bin/env python
"""Key metrics of one or more .ncu-rep files (ncu -i … --page raw --csv) as a markdown table.
    python profiles/ncu_summary.py gpurun_out/r2f_*.ncu-rep > profiles/r2f_ncu_summary.md"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp64_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
STALL = "smsp__average_warp_latency_issue_stalled_"   # ..._<reason>.ratio  / warps_issue_stalled
def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    return hdr, units, r[2:]
for rep in sys.argv[1:]:
    hdr, units, data = rows(rep)
    for d in data:
        m = dict(zip(hdr, d)); u = dict(zip(hdr, units))
        print(f"\n## {rep.split('/')[-1]} — `{m.get('Kernel Name','?')}` (launch id {m.get('ID')})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in m and m[k] != "":
                print(f"| {k} | {m[k]} | {u.get(k,'')} |")
        st = sorted(((float(m[k].replace(',', '')), k) for k in m if k.startswith("smsp__average_warp") and k.endswith("_stalled_" + k.split("_stalled_")[-1]) and "issue_stalled" in k and m[k] not in ("", "n/a") and k.endswith(".ratio")), reverse=True)
        if st:
            print("\nstall reasons (warp-cycles per issued instruction): " + ", ".join(f"{k.split('issue_stalled_')[-1].replace('.ratio','')} {v:.2f}" for v, k in st[:8]))
