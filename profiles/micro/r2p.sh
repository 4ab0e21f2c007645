# a rank's share at N = 8 is 125 pairs: which job plan serves it best?  (one GPU, --pairs 125)
set -x
B="python bench.py --pairs 125 --steps 4 --warmup 3 --no-cpu-baseline --no-pageable"
for sb in 16 32 42 63 125; do $B --sub-batch $sb > gpurun_out/r2p_125_sub$sb.json 2> gpurun_out/r2p_125_sub$sb.err; echo "sub$sb rc=$?"; done
for f in 16 32; do NCFA_E2E_FIRST=$f NCFA_E2E_GROWTH=2 $B --sub-batch 63 > gpurun_out/r2p_125_sub63_first$f.json 2>&1; echo "first$f rc=$?"; done
NCFA_E2E_GROWTH=2 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2p_1000_growth2.json 2>&1; echo "growth2 rc=$?"
NCFA_E2E_GROWTH=2.5 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2p_1000_growth25.json 2>&1; echo "growth2.5 rc=$?"
