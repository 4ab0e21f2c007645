# round-2 final-state evidence on one GPU: tests, contract bench (+ reference arm), family lines, ncu
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -4 gpurun_out/r2n_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2n_reference.json 2> gpurun_out/r2n_reference.err; echo "reference rc=$?"
for c in 2 3 4; do python bench.py --config $c --steps 5 > gpurun_out/r2n_cfg$c.json 2> gpurun_out/r2n_cfg$c.err; echo "cfg$c rc=$?"; done
python bench.py --single-pair --steps 5 > gpurun_out/r2n_single_pair.json 2>&1; echo "single rc=$?"
B="python bench.py --pairs 125 --steps 1 --warmup 1 --no-cpu-baseline --no-pageable --workers 1"
N="timeout 900 ncu --set full --clock-control none --import-source on -f"
$N -k regex:stft_logmel2 -s 60 -c 1 -o gpurun_out/r2n_stft64 $B > gpurun_out/r2n_ncu_stft64.log 2>&1; echo "stft64 rc=$?"
$N -k regex:cqt_tc -s 1 -c 1 -o gpurun_out/r2n_cqt $B > gpurun_out/r2n_ncu_cqt.log 2>&1; echo "cqt rc=$?"
$N -k regex:decimate2_kernel -s 6 -c 1 -o gpurun_out/r2n_decim $B > gpurun_out/r2n_ncu_decim.log 2>&1; echo "decim rc=$?"
$N -k regex:beat_track -s 5 -c 1 -o gpurun_out/r2n_beat $B > gpurun_out/r2n_ncu_beat.log 2>&1; echo "beat rc=$?"
$N -k regex:xcorr_blocks -c 1 -o gpurun_out/r2n_xcorr python bench.py --config 4 --steps 1 --warmup 1 --family-pairs 4 > gpurun_out/r2n_ncu_xcorr.log 2>&1; echo "xcorr rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --pairs 32 --sub-batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-pageable --workers 1 > gpurun_out/r2n_ncu_list.log 2>&1; echo "list rc=$?"
