# final verification of round 1 (templated coupled kernels, fused CGS2 default, L2-tiled fused variant, C++ mirror / reference
# Krylov templates on the device); most valuable first, every step bounded
TAG=r01d
timeout 150 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -8 gpurun_out/pytest_gpu_$TAG.log
THCM_FUSED_CGS2=2 timeout 60 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "gmres" > gpurun_out/pytest_gpu_${TAG}_fused2.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_${TAG}_fused2.log; tail -3 gpurun_out/pytest_gpu_${TAG}_fused2.log
for f in 2 1; do THCM_FUSED_CGS2=$f timeout 70 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_fused$f.json 2> gpurun_out/bench_${TAG}_fused$f.err; done
python - <<PY
import json
for f in (2, 1):
    try:
        for l in open(f'gpurun_out/bench_r01d_fused{f}.json'):
            if l.startswith('{'):
                d = json.loads(l); print('fused', f, 'step_ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['gmres'], {k: (v['launches_per_step'], round(v['avg_ms'], 4)) for k, v in d['kernels'].items() if 'multi' in k or 'asm' in k or 'spmv' in k})
    except Exception as e:
        print('fused', f, 'no result', e)
PY
THCM_FUSED_CGS2=2 timeout 80 ncu --set full --clock-control none --import-source on -k regex:"fused" -s 40 -c 2 -o gpurun_out/prof_${TAG}_fused2 -f python scripts/prof_kernels.py krylov > gpurun_out/ncu_full_fused2_$TAG.log 2>&1; tail -2 gpurun_out/ncu_full_fused2_$TAG.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:"fused" -s 40 -c 2 -o gpurun_out/prof_${TAG}_fused1 -f python scripts/prof_kernels.py krylov > gpurun_out/ncu_full_fused1_$TAG.log 2>&1; tail -2 gpurun_out/ncu_full_fused1_$TAG.log
