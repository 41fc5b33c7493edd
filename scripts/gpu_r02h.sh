TAG=${1:-r02h}
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -8 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-b1 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}.json'):
    if l.startswith('{'):
        d = json.loads(l); print('step_ms', round(d['ms_per_step'], 3), 'e2e', d['e2e']['value'], 'asm+spmv', d.get('assembly_plus_spmv'))
        print({k: (v.get('launches_per_step'), round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3), round(v.get('graph_equivalent_frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
