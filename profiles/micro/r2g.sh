# round-2 step g: ring-form xcorr, coalesced CQT producer, 8-output decimator, trimmed STFT un-pack; host worker count
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -4 gpurun_out/r2g_pytest.log
python bench.py --config 4 --steps 5 > gpurun_out/r2g_cfg4.json 2> gpurun_out/r2g_cfg4.err; echo "cfg4 rc=$?"
NCFA_XCORR_IMPL=tma3 python bench.py --config 4 --steps 5 > gpurun_out/r2g_cfg4_tma3.json 2>&1; echo "cfg4 tma3 rc=$?"
python bench.py --config 3 --steps 5 > gpurun_out/r2g_cfg3.json 2> gpurun_out/r2g_cfg3.err; echo "cfg3 rc=$?"
for w in 2 3 4; do
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable --workers $w > gpurun_out/r2g_bench_w$w.json 2> gpurun_out/r2g_bench_w$w.err; echo "bench w$w rc=$?"
done
python bench.py --pairs 125 --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2g_bench_125.json 2>&1; echo "bench 125 rc=$?"
