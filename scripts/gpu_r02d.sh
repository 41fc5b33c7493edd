# r02d (1 GPU): suite + bench with the fused Arnoldi head, launch list, ncu --set full of today's dominant kernels
TAG=${1:-r02d}
timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
for v in "" "THCM_NO_FUSED_HEAD=1"; do
  name=$(echo "${v:-default}" | tr '=' '_')
  env $v timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  tail -3 gpurun_out/bench_${TAG}_$name.err
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_$name.json'):
    if l.startswith('{'):
        d = json.loads(l); print('$name', 'step_ms', round(d['ms_per_step'], 3), 'e2e', d['e2e']['value'], 'resid', d['gmres']['resid'], {k: (v['launches_per_step'], round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
done
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"jac_tma|thcm_assemble|blockdiag_build" -s 3 -c 4 -o gpurun_out/prof_${TAG}_asm -f python scripts/prof_kernels.py asm > gpurun_out/ncu_full_asm_$TAG.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"multi_dot|fused|multi_axpy_dot|spmv_compact|scale_precon" -s 122 -c 10 -o gpurun_out/prof_${TAG}_krylov25 -f python scripts/prof_kernels.py krylov > gpurun_out/ncu_full_krylov25_$TAG.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"multi_dot|fused|multi_axpy_dot|spmv_compact|scale_precon" -s 242 -c 5 -o gpurun_out/prof_${TAG}_krylov50 -f python scripts/prof_kernels.py krylov > gpurun_out/ncu_full_krylov50_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_krylov50_$TAG.log
