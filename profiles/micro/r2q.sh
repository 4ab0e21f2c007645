# a rank's share at N = 8 is 125 pairs: job size and worker count (one GPU, --pairs 125; --sub-batch now honoured as given)
set -x
B="python bench.py --pairs 125 --steps 4 --warmup 3 --no-cpu-baseline --no-pageable"
for sb in 24 42 63 125; do $B --sub-batch $sb > gpurun_out/r2q_125_sub$sb.json 2> gpurun_out/r2q_125_sub$sb.err; echo "sub$sb rc=$?"; done
$B --sub-batch 32 --workers 3 > gpurun_out/r2q_125_sub32_w3.json 2>&1; echo "w3 rc=$?"
$B --sub-batch 63 --workers 3 > gpurun_out/r2q_125_sub63_w3.json 2>&1; echo "w3b rc=$?"
NCFA_E2E_FIRST=16 NCFA_E2E_GROWTH=2 $B --sub-batch 63 > gpurun_out/r2q_125_sub63_first16.json 2>&1; echo "first16 rc=$?"
python profiles/micro/timeline.py 125 2 > gpurun_out/r2q_timeline_125.log 2>&1; echo "timeline rc=$?"
