# last capture of the round: the hop-64 tempogram kernel (tg_lag_kernel[long]) — all tg_lag launches of a 125-pair job
set -x
B="python bench.py --pairs 125 --steps 1 --warmup 1 --no-cpu-baseline --no-pageable --workers 1"
timeout 400 ncu --section SpeedOfLight --section WarpStateStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -f -k regex:tg_lag -s 6 -c 12 -o gpurun_out/r3e_tglag $B > gpurun_out/r3e_ncu_tglag.log 2>&1; echo "tglag rc=$?"
