set -x
python -m pytest tests/test_gpu_xcorr.py tests/test_gpu_tempo.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -3 gpurun_out/r2l_pytest.log
for d in 1 2 4 7 3; do NCFA_TC_DEBUG=$d python bench.py --config 3 --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('TC_DEBUG=$d', d['kernels']['cqt_tc_kernel'])" ; done > gpurun_out/r2l_cqt_debug.log 2>&1
NCFA_XCORR_IMPL=x2 python bench.py --config 4 --steps 5 > gpurun_out/r2l_cfg4_x2.json 2>&1; echo "cfg4 x2 rc=$?"
python bench.py --config 4 --steps 5 > gpurun_out/r2l_cfg4.json 2>&1; echo "cfg4 rc=$?"
for f in 8 24 48 83; do NCFA_E2E_FIRST=$f python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2l_bench_first$f.json 2> gpurun_out/r2l_bench_first$f.err; echo "bench first=$f rc=$?"; done
