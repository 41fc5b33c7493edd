# First GPU call of the next round (1 GPU, ~12 min worst case: steps are individually bounded; split into two calls if the budget is tight): everything that was written after round 1's GPU budget was spent, then the
# captures the next optimisation steps need.  Usage: gpurun --timeout 900 -- 'bash scripts/gpu_r02_first.sh'
TAG=${1:-r02a}
# 1. the gated tests (integral condition / pressure rows, pattern-compressed SpMV) + the whole suite
THCM_RUN_UNVERIFIED=1 timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -8 gpurun_out/pytest_gpu_$TAG.log
# 2. A/B of the candidates on the same box: pattern SpMV, fused CGS2 occupancy
for v in "" "THCM_KRYLOV_COMPACT=1" "THCM_SPMV_SKIP_LAND=1" "THCM_SPMV_PATTERN=1" "THCM_ASM_PIPE=5" "THCM_FUSED2_BPS=38" "THCM_FUSED2_BPS=2" ; do
  name=$(echo "${v:-default}" | tr '=' '_')
  env $v timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_$name.json'):
    if l.startswith('{'):
        d = json.loads(l); print('$name', 'step_ms', round(d['ms_per_step'], 3), 'resid', d['gmres']['resid'], {k: (v['launches_per_step'], round(v['avg_ms'], 4), round(v['frac_of_peak'], 3)) for k, v in d['kernels'].items() if 'multi' in k or 'JAC' in k or 'spmv' in k})
PY
done
# 3. launch list of the bench command + source-level capture of the Jacobian kernels (the 44 % kernel) and the L2-tiled fused kernel
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"jac_tma|spmv_csr" -s 4 -c 6 -o gpurun_out/prof_${TAG}_asm -f python scripts/prof_kernels.py asm > gpurun_out/ncu_full_asm_$TAG.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"fused|multi_dot" -s 120 -c 4 -o gpurun_out/prof_${TAG}_krylov -f python scripts/prof_kernels.py krylov > gpurun_out/ncu_full_krylov_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_krylov_$TAG.log
