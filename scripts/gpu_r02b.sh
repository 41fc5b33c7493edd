# r02b (1 GPU): whole GPU suite with the ocean-only Krylov space as the default + bench
TAG=${1:-r02b}
timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -15 gpurun_out/pytest_gpu_$TAG.log
for v in "" "THCM_KRYLOV_COMPACT=0"; do
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
