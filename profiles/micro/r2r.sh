# 125-pair share with 125-pair jobs: streamed plans; H2D copy rate while kernels run
set -x
B="python bench.py --pairs 125 --steps 4 --warmup 3 --no-cpu-baseline --no-pageable"
$B > gpurun_out/r2r_125_default.json 2> gpurun_out/r2r_125_default.err; echo "default rc=$?"
NCFA_E2E_FIRST=16 NCFA_E2E_GROWTH=2.5 $B > gpurun_out/r2r_125_f16_g25.json 2>&1; echo "f16g25 rc=$?"
NCFA_E2E_FIRST=16 NCFA_E2E_GROWTH=2 $B > gpurun_out/r2r_125_f16_g2.json 2>&1; echo "f16g2 rc=$?"
NCFA_E2E_FIRST=24 NCFA_E2E_GROWTH=2 $B > gpurun_out/r2r_125_f24_g2.json 2>&1; echo "f24g2 rc=$?"
NCFA_E2E_FIRST=32 NCFA_E2E_GROWTH=3 $B > gpurun_out/r2r_125_f32_g3.json 2>&1; echo "f32g3 rc=$?"
$B --workers 3 > gpurun_out/r2r_125_w3.json 2>&1; echo "w3 rc=$?"
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2r_1000.json 2>&1; echo "1000 rc=$?"
NCFA_E2E_FIRST=24 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-pageable > gpurun_out/r2r_1000_f24.json 2>&1; echo "1000 f24 rc=$?"
