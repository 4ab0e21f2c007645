# eight GPUs: the contract bench at N = 8 (torchrun, NCCL gather), plus N = 4 on the same box
set -x
nvidia-smi -L | wc -l; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2s_bench_8gpu.json 2> gpurun_out/r2s_bench_8gpu.err; echo "bench8 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 --no-pageable > gpurun_out/r2s_bench_4gpu.json 2> gpurun_out/r2s_bench_4gpu.err; echo "bench4 rc=$?"
