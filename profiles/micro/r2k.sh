set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -4 gpurun_out/r2k_pytest.log
timeout 300 build/micro/tma_bw > gpurun_out/r2k_tma_bw.log 2>&1; echo "tma_bw rc=$?"
python bench.py --config 3 --steps 5 > gpurun_out/r2k_cfg3.json 2> gpurun_out/r2k_cfg3.err; echo "cfg3 rc=$?"
NCFA_CQT_IMPL=tc1 python bench.py --config 3 --steps 5 > gpurun_out/r2k_cfg3_tc1.json 2>&1; echo "cfg3 tc1 rc=$?"
python bench.py --config 4 --steps 5 > gpurun_out/r2k_cfg4.json 2> gpurun_out/r2k_cfg4.err; echo "cfg4 rc=$?"
python bench.py --config 2 --steps 3 > gpurun_out/r2k_cfg2.json 2> gpurun_out/r2k_cfg2.err; echo "cfg2 rc=$?"
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
