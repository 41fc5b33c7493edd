export THCM_ASM_PIPE=${1:-1}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${2:-jac_tma} -s 2 -c ${3:-2} -o gpurun_out/prof_jac_$THCM_ASM_PIPE -f python scripts/jac_only.py 3 > gpurun_out/ncu_jac.log 2>&1
tail -3 gpurun_out/ncu_jac.log
