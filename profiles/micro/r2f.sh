# round-2 ncu evidence: one --set full capture per kernel family (125-pair sub-batch command) + the launch list
set -x
B="python bench.py --pairs 125 --steps 1 --warmup 1 --no-cpu-baseline --no-pageable --workers 1"
$B > gpurun_out/r2f_plain.log 2>&1; echo "plain rc=$?"
N="timeout 900 ncu --set full --clock-control none --import-source on -f"
$N -k regex:stft_logmel2 -s 60 -c 1 -o gpurun_out/r2f_stft64 $B > gpurun_out/r2f_ncu_stft64.log 2>&1; echo "stft64 rc=$?"
$N -k regex:cqt_tc -s 1 -c 1 -o gpurun_out/r2f_cqt $B > gpurun_out/r2f_ncu_cqt.log 2>&1; echo "cqt rc=$?"
$N -k regex:decimate2_kernel -s 6 -c 1 -o gpurun_out/r2f_decim $B > gpurun_out/r2f_ncu_decim.log 2>&1; echo "decim rc=$?"
$N -k regex:tuning_peaks -s 1 -c 1 -o gpurun_out/r2f_tuning $B > gpurun_out/r2f_ncu_tuning.log 2>&1; echo "tuning rc=$?"
$N -k regex:tg_lag -s 8 -c 2 -o gpurun_out/r2f_tglag $B > gpurun_out/r2f_ncu_tglag.log 2>&1; echo "tglag rc=$?"
$N -k regex:beat_track -s 5 -c 1 -o gpurun_out/r2f_beat $B > gpurun_out/r2f_ncu_beat.log 2>&1; echo "beat rc=$?"
$N -k regex:xcorr_blocks -c 1 -o gpurun_out/r2f_xcorr python bench.py --config 4 --steps 1 --warmup 1 --family-pairs 4 > gpurun_out/r2f_ncu_xcorr.log 2>&1; echo "xcorr rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --pairs 32 --sub-batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-pageable --workers 1 > gpurun_out/r2f_ncu_list.log 2>&1; echo "list rc=$?"
ls -la gpurun_out/*.ncu-rep
