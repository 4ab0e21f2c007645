set -x
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_io.py -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log; tail -3 gpurun_out/r2z_pytest.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
NCFA_STAGE_THREADS=4 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2z_bench_st4.json 2> gpurun_out/r2z_bench_st4.err; echo "bench st4 rc=$?"
NCFA_STAGE_THREADS=12 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2z_bench_st12.json 2> gpurun_out/r2z_bench_st12.err; echo "bench st12 rc=$?"
nproc
