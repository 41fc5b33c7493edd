# one-shot verification of round 1c (coupled mode, B1 diagnostics, theta stepping, fused CGS2): most valuable first
TAG=r01c
timeout 240 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
THCM_FUSED_CGS2=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "gmres" > gpurun_out/pytest_gpu_${TAG}_fused.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_${TAG}_fused.log; tail -3 gpurun_out/pytest_gpu_${TAG}_fused.log
THCM_FUSED_CGS2=1 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_fused1.json 2> gpurun_out/bench_${TAG}_fused1.err; tail -c 200 gpurun_out/bench_${TAG}_fused1.err
THCM_FUSED_CGS2=0 timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_fused0.json 2> gpurun_out/bench_${TAG}_fused0.err; tail -c 200 gpurun_out/bench_${TAG}_fused0.err
python - <<PY
import json
for f in (1, 0):
    try:
        for l in open(f'gpurun_out/bench_r01c_fused{f}.json'):
            if l.startswith('{'):
                d = json.loads(l); print('fused', f, 'step_ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['gmres'], {k: (v['launches_per_step'], round(v['avg_ms'], 4)) for k, v in d['kernels'].items() if 'multi' in k or 'asm' in k or 'spmv' in k})
    except Exception as e:
        print('fused', f, 'no result', e)
PY
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
