set -x
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; echo "bench rc=$?"
python -m pytest tests/test_gpu_tempo.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3f_pytest.log; tail -3 gpurun_out/r3f_pytest.log
