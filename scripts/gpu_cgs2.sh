timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gmres or idrs or full_size or vector_kernels" > gpurun_out/pytest_cgs2.log 2>&1; echo rc=$? >> gpurun_out/pytest_cgs2.log; tail -5 gpurun_out/pytest_cgs2.log
for f in 0 1; do THCM_FUSED_CGS2=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cgs2_$f.json 2> gpurun_out/bench_cgs2_$f.err; done
python - <<PY
import json
for f in (0,1):
    for l in open(f'gpurun_out/bench_cgs2_{f}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['gmres'], {k:(v['launches_per_step'],round(v['avg_ms'],4)) for k,v in d['kernels'].items() if 'multi' in k})
PY
