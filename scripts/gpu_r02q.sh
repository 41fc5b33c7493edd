# r02q (1 GPU): whole -m gpu suite, then the bench line without the CPU sample
TAG=${1:-r02q}
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
for l in open('gpurun_out/bench_$TAG.json'):
    if l.startswith('{'):
        d = json.loads(l); print('step_ms', round(d['ms_per_step'], 3), 'e2e', d['e2e']['value'], 'b1', d.get('e2e_b1'), 'mixing1', d.get('mixing1'))
        print({k: (v.get('launches_per_step'), round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
