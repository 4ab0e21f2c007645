# eight GPUs again: the up-then-down plan for a 125-pair share
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 5 --warmup 3 --no-pageable > gpurun_out/r2u_bench_8gpu.json 2> gpurun_out/r2u_bench_8gpu.err; echo "bench8 rc=$?"
NCFA_E2E_SIZES=8,12,18,27,60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-pageable > gpurun_out/r2u_bench_8gpu_oldplan.json 2> gpurun_out/r2u_bench_8gpu_oldplan.err; echo "bench8 old rc=$?"
