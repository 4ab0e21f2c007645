set -x
timeout 600 python -m pytest tests/test_gpu_pitch.py tests/test_gpu_stress.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -3 gpurun_out/r2m_pytest.log
timeout 200 python bench.py --config 3 --steps 5 > gpurun_out/r2m_cfg3.json 2> gpurun_out/r2m_cfg3.err; echo "cfg3 rc=$?"
NCFA_CQT_IMPL=tc1 timeout 200 python bench.py --config 3 --steps 5 > gpurun_out/r2m_cfg3_tc1.json 2>&1; echo "cfg3 tc1 rc=$?"
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
