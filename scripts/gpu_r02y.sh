# r02y (1 GPU): the whole -m gpu suite at the head (parameter-list constructors, device applyMassMat, default-run root, TMA-staged fused
# CGS2 as a parametrised variant), smoke(), the bench line, and the A/B of THCM_FUSED_CGS2=3 against the default
TAG=${1:-r02y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo rc=$? >> gpurun_out/smoke_$TAG.log; tail -3 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err
THCM_FUSED_CGS2=3 timeout 600 python bench.py --no-cpu-baseline --no-mixing --no-b1 > gpurun_out/bench_${TAG}_cgs3.json 2> gpurun_out/bench_${TAG}_cgs3.err; tail -3 gpurun_out/bench_${TAG}_cgs3.err
python - <<PY
import json
for f in ('gpurun_out/bench_$TAG.json', 'gpurun_out/bench_${TAG}_cgs3.json'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f, 'step_ms', round(d['ms_per_step'], 3), 'e2e', d['e2e']['value'], 'resid', d['gmres']['resid'])
            print({k: (v.get('launches_per_step'), round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
