# eight GPUs with the final scheduler (two uploads ahead)
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 8 --steps 5 --warmup 3 --no-pageable > gpurun_out/r3c_bench_8gpu.json 2> gpurun_out/r3c_bench_8gpu.err; echo "bench8 rc=$?"
