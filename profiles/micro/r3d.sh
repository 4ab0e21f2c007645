# does a rank with only 4 host cores (its share of the 32-core box at N = 8) suffer from spinning waits?  (one GPU)
set -x
B="python bench.py --pairs 125 --steps 4 --warmup 3 --no-cpu-baseline --no-pageable"
P='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], "resident", round(d["pairs_per_sec"],1), "e2e", round(d["e2e"]["pairs_per_sec"],1), round(d["e2e"]["ms_per_step"],1), "blocking", d["schedule"].get("blocking_sync"))'
{
taskset -c 0-3 $B 2>/dev/null | python -c "$P" "4cores_spin"
NCFA_BLOCKING_SYNC=1 taskset -c 0-3 $B 2>/dev/null | python -c "$P" "4cores_blocking"
taskset -c 0-1 $B 2>/dev/null | python -c "$P" "2cores_spin"
NCFA_BLOCKING_SYNC=1 taskset -c 0-1 $B 2>/dev/null | python -c "$P" "2cores_blocking"
NCFA_BLOCKING_SYNC=1 $B 2>/dev/null | python -c "$P" "16cores_blocking"
NCFA_BLOCKING_SYNC=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable 2>/dev/null | python -c "$P" "16cores_blocking_1000pairs"
} > gpurun_out/r3d_blocking.log 2>&1
cat gpurun_out/r3d_blocking.log
