set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; echo "bench rc=$?"
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3b_pytest.log; tail -2 gpurun_out/r3b_pytest.log
python bench.py --pairs 125 --steps 4 --warmup 3 --no-cpu-baseline --no-pageable > gpurun_out/r3b_bench_125.json 2>&1; echo "bench125 rc=$?"
