# 1 GPU on the per-rank problem size of the 8-GPU run (90x76x24): intrinsic small-kernel cost without communication
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --grid 90 76 24 --no-cpu-baseline > gpurun_out/bench_small1.json 2> gpurun_out/bench_small1.err
python - <<PY
import json
for l in open('gpurun_out/bench_small1.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value']); print({k:(v['launches_per_step'],round(v['avg_ms'],4)) for k,v in d['kernels'].items()})
PY
