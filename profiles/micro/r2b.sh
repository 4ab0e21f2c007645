set -x
python -m pytest tests/test_gpu_xcorr.py tests/test_gpu_fullsize.py -m gpu -x -q -k "xcorr" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log; tail -3 gpurun_out/r2b_pytest.log
python profiles/micro/chroma_diag.py > gpurun_out/r2b_chroma_tc.log 2>&1
NCFA_CQT_IMPL=simt python profiles/micro/chroma_diag.py > gpurun_out/r2b_chroma_simt.log 2>&1
python bench.py --config 4 --steps 5 > gpurun_out/r2b_cfg4.log 2> gpurun_out/r2b_cfg4.err; echo "cfg4 rc=$?"
NCFA_XCORR_IMPL=direct python bench.py --config 4 --steps 5 > gpurun_out/r2b_cfg4_direct.log 2> gpurun_out/r2b_cfg4_direct.err; echo "cfg4 direct rc=$?"
python bench.py --config 2 --steps 3 > gpurun_out/r2b_cfg2.log 2> gpurun_out/r2b_cfg2.err; echo "cfg2 rc=$?"
python bench.py --config 3 --steps 5 > gpurun_out/r2b_cfg3.log 2> gpurun_out/r2b_cfg3.err; echo "cfg3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xcorr_blocks -c 1 -f -o gpurun_out/r2b_xcorr python bench.py --config 4 --steps 1 --warmup 1 --family-pairs 4 > gpurun_out/r2b_ncu_xcorr.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
