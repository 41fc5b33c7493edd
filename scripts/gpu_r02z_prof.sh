# r02z (1 GPU): ncu --set full of every launch of the TMA-staged fused CGS2 kernel (THCM_FUSED_CGS2=3) in one GMRES(50) cycle at 1 degree
# (nv = 17 .. 49), raw metrics + the source page of one launch at nv ~ 45
TAG=${1:-r02z}
mkdir -p gpurun_out /tmp/ncu
THCM_FUSED_CGS2=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fused3" -c 34 -o /tmp/ncu/f3 -f python scripts/prof_kernels.py krylov > gpurun_out/ncu_full_f3_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_f3_$TAG.log
ncu -i /tmp/ncu/f3.ncu-rep --page raw --csv > gpurun_out/ncu_raw_f3_$TAG.csv 2>/dev/null
ncu -i /tmp/ncu/f3.ncu-rep --page source --csv --launch-skip 28 --launch-count 1 2>/dev/null | head -c 4000000 > gpurun_out/ncu_source_f3_$TAG.csv
ncu -i /tmp/ncu/f3.ncu-rep --page details --launch-skip 28 --launch-count 1 2>/dev/null | head -c 200000 > gpurun_out/ncu_details_f3_$TAG.txt
ls -la /tmp/ncu/f3.ncu-rep gpurun_out | head -20
