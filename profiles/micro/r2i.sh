set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -4 gpurun_out/r2i_pytest.log
python bench.py --config 4 --steps 5 > gpurun_out/r2i_cfg4.json 2> gpurun_out/r2i_cfg4.err; echo "cfg4 rc=$?"
NCFA_XCORR_IMPL=ring4 python bench.py --config 4 --steps 5 > gpurun_out/r2i_cfg4_ring4.json 2>&1; echo "cfg4 ring4 rc=$?"
python bench.py --config 3 --steps 5 > gpurun_out/r2i_cfg3.json 2> gpurun_out/r2i_cfg3.err; echo "cfg3 rc=$?"
python profiles/micro/timeline.py 500 2 > gpurun_out/r2i_timeline_w2.log 2>&1; echo "timeline rc=$?"
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable --workers 3 > gpurun_out/r2i_bench_w3.json 2> gpurun_out/r2i_bench_w3.err; echo "bench w3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:cqt_tc -s 1 -c 1 -o gpurun_out/r2i_cqt python bench.py --config 3 --steps 1 --warmup 1 --family-pairs 32 > gpurun_out/r2i_ncu_cqt.log 2>&1; echo "ncu cqt rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:xcorr_blocks -c 1 -o gpurun_out/r2i_xcorr python bench.py --config 4 --steps 1 --warmup 1 --family-pairs 4 > gpurun_out/r2i_ncu_xcorr.log 2>&1; echo "ncu xcorr rc=$?"
