set -x
python -m pytest tests/test_gpu_xcorr.py tests/test_gpu_pitch.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -4 gpurun_out/r2h_pytest.log
python bench.py --config 4 --steps 5 > gpurun_out/r2h_cfg4.json 2> gpurun_out/r2h_cfg4.err; echo "cfg4 rc=$?"
python bench.py --config 3 --steps 5 > gpurun_out/r2h_cfg3.json 2> gpurun_out/r2h_cfg3.err; echo "cfg3 rc=$?"
python profiles/micro/timeline.py 500 2 > gpurun_out/r2h_timeline_w2.log 2>&1; echo "timeline rc=$?"
python profiles/micro/timeline.py 500 1 > gpurun_out/r2h_timeline_w1.log 2>&1; echo "timeline w1 rc=$?"
