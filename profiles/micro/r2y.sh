# final verification of the round-2 build: smoke, every GPU test, the contract bench with its CPU baseline
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2y_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log; tail -3 gpurun_out/r2y_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
