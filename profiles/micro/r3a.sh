set -x
for sl in 3 4 6; do NCFA_E2E_SLOTS=$sl python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-pageable 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('slots $sl resident', round(d['pairs_per_sec'],1), 'e2e', round(d['e2e']['pairs_per_sec'],1), round(d['e2e']['ms_per_step'],1))"; done > gpurun_out/r3a_slots.log 2>&1
for g in 1.5 3; do NCFA_E2E_SLOTS=4 NCFA_E2E_GROWTH=$g python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-pageable 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('slots 4 growth $g e2e', round(d['e2e']['pairs_per_sec'],1), d['schedule']['e2e_sub_batches'])"; done >> gpurun_out/r3a_slots.log 2>&1
grep -v "^+" gpurun_out/r3a_slots.log
