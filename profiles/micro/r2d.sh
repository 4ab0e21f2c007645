set -x
B="python bench.py --pairs 125 --steps 1 --warmup 1 --no-cpu-baseline --no-pageable --no-pitch --workers 1"
$B > gpurun_out/r2d_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft_logmel2 -s 40 -c 1 -f -o gpurun_out/r2d_stft $B > gpurun_out/r2d_ncu_stft.log 2>&1; echo "ncu stft rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beat_track -s 3 -c 1 -f -o gpurun_out/r2d_beat $B > gpurun_out/r2d_ncu_beat.log 2>&1; echo "ncu beat rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xcorr_blocks -c 1 -f -o gpurun_out/r2d_xcorr python bench.py --config 4 --steps 1 --warmup 1 --family-pairs 4 > gpurun_out/r2d_ncu_xcorr.log 2>&1; echo "ncu xcorr rc=$?"
ls -la gpurun_out/*.ncu-rep
