# ramp-down tails for the streamed plan (one GPU)
set -x
B="python bench.py --pairs 125 --steps 4 --warmup 3 --no-cpu-baseline --no-pageable"
for plan in "8,12,18,27,60" "8,12,18,27,36,16,8" "8,12,18,27,30,18,12" "8,14,24,34,25,12,8" "8,16,32,37,20,12" "12,24,40,29,20"; do
  NCFA_E2E_SIZES=$plan $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$plan', round(d['e2e']['pairs_per_sec'],1), round(d['e2e']['ms_per_step'],1), d['schedule']['e2e_sub_batches'])"
done > gpurun_out/r2t_plans_125.log 2>&1
B2="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable"
for plan in "8,12,18,27,40,60,91,124,124,124,124,124,124" "8,12,18,27,40,60,91,124,124,124,124,100,72,44,24,8" "8,12,18,27,40,60,91,124,124,124,124,124,64,36,16,8"; do
  NCFA_E2E_SIZES=$plan $B2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$plan', round(d['e2e']['pairs_per_sec'],1), round(d['e2e']['ms_per_step'],1))"
done > gpurun_out/r2t_plans_1000.log 2>&1
cat gpurun_out/r2t_plans_125.log gpurun_out/r2t_plans_1000.log
