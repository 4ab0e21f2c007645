# fewer, larger streamed jobs for a 125-pair batch (one GPU)
set -x
B="python bench.py --pairs 125 --steps 4 --warmup 3 --no-cpu-baseline --no-pageable"
for plan in "12,24,40,29,20" "20,55,50" "16,45,64" "25,50,50" "40,85" "16,36,43,30"; do
  NCFA_E2E_SIZES=$plan $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$plan', round(d['e2e']['pairs_per_sec'],1), round(d['e2e']['ms_per_step'],1), d['schedule']['e2e_sub_batches'])"
done > gpurun_out/r2w_plans_125.log 2>&1
grep -v "^+" gpurun_out/r2w_plans_125.log
