set -x
python -m pytest tests/test_gpu_xcorr.py tests/test_gpu_tempo.py tests/test_gpu_fullsize.py tests/test_gpu_stress.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -4 gpurun_out/r2j_pytest.log
python bench.py --config 4 --steps 5 > gpurun_out/r2j_cfg4.json 2> gpurun_out/r2j_cfg4.err; echo "cfg4 rc=$?"
NCFA_XCORR_IMPL=ring8 python bench.py --config 4 --steps 5 > gpurun_out/r2j_cfg4_ring8.json 2>&1; echo "cfg4 ring8 rc=$?"
python bench.py --config 2 --steps 3 > gpurun_out/r2j_cfg2.json 2> gpurun_out/r2j_cfg2.err; echo "cfg2 rc=$?"
NCFA_BEAT_DP=scan python bench.py --config 2 --steps 3 > gpurun_out/r2j_cfg2_scan.json 2>&1; echo "cfg2 scan rc=$?"
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
