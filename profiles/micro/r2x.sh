set -x
python -m pytest tests/test_gpu_tempo.py tests/test_gpu_fullsize.py tests/test_gpu_stress.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log; tail -3 gpurun_out/r2x_pytest.log
python bench.py --config 2 --steps 3 > gpurun_out/r2x_cfg2.json 2> gpurun_out/r2x_cfg2.err; echo "cfg2 rc=$?"
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python bench.py --pairs 125 --steps 4 --warmup 3 --no-cpu-baseline --no-pageable > gpurun_out/r2x_bench_125.json 2>&1; echo "bench125 rc=$?"
