# two GPUs: the 2-rank gather-identity test and the contract bench at N = 2
set -x
nvidia-smi -L
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_io.py -m gpu -x -q -k "two_rank or engines_of_ended" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log; tail -4 gpurun_out/r2o_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2o_bench_2gpu.json 2> gpurun_out/r2o_bench_2gpu.err; echo "bench2 rc=$?"
python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_1gpu.json 2> gpurun_out/r2o_bench_1gpu.err; echo "bench1 rc=$?"
nproc
