set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -4 gpurun_out/r2c_pytest.log
python bench.py --config 4 --steps 5 > gpurun_out/r2c_cfg4.log 2> gpurun_out/r2c_cfg4.err; echo "cfg4 rc=$?"
NCFA_XCORR_IMPL=regs python bench.py --config 4 --steps 5 > gpurun_out/r2c_cfg4_regs.log 2>&1; echo "cfg4 regs rc=$?"
python bench.py --config 2 --steps 3 > gpurun_out/r2c_cfg2.log 2> gpurun_out/r2c_cfg2.err; echo "cfg2 rc=$?"
python bench.py --pairs 250 --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2c_250.log 2> gpurun_out/r2c_250.err; echo "250 rc=$?"
NCFA_STFT_IMPL=warp1 python bench.py --pairs 250 --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2c_250_warp1.log 2>&1; echo "250 warp1 rc=$?"
NCFA_STFT_WARPS=10 python bench.py --pairs 250 --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2c_250_w10.log 2>&1; echo "250 w10 rc=$?"
