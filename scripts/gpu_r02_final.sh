# final evidence of round 2 (1 GPU): the whole -m gpu suite at the head, smoke(), both bench arms as the driver runs them, SASS excerpt
TAG=${1:-r02final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo rc=$? >> gpurun_out/smoke_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -3 gpurun_out/bench_${TAG}_reference.err
cuobjdump -sass i-emic_b200/libthcm_b200.so 2>/dev/null | grep -E "Function :|UBLKCP|SYNCS|UTMA" | grep -B1 -E "UBLKCP|SYNCS|UTMA" | head -80 > gpurun_out/sass_tma_$TAG.txt
python - <<PY
import json
for f in ('gpurun_out/bench_$TAG.json', 'gpurun_out/bench_${TAG}_reference.json'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f, 'value', d.get('value'), 'e2e', d.get('e2e', {}).get('value'), 'roofline', d.get('roofline', {}).get('frac'), 'cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'), 'asm+spmv', d.get('assembly_plus_spmv', {}).get('streamed_frac_of_peak'), d.get('assembly_plus_spmv', {}).get('graph_equivalent_frac_of_peak'))
PY
