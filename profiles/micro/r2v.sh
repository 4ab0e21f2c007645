# eight GPUs: NUMA binding of every rank to its GPU's cores (on / off)
set -x
nvidia-smi topo -m > gpurun_out/r2v_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --no-pageable > gpurun_out/r2v_bench_8gpu_bind.json 2> gpurun_out/r2v_bench_8gpu_bind.err; echo "bind rc=$?"
NCFA_NO_NUMA_BIND=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --no-pageable > gpurun_out/r2v_bench_8gpu_nobind.json 2> gpurun_out/r2v_bench_8gpu_nobind.err; echo "nobind rc=$?"
