# r02e (2 GPUs): whole suite (multi tests at world = 2) after the prune, N = 1 bench, N = 2 benches with the A/B switches
TAG=${1:-r02e}
timeout 500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g1.json 2> gpurun_out/bench_${TAG}_g1.err
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_g1.json'):
    if l.startswith('{'):
        d = json.loads(l); print('N=1 step_ms', round(d['ms_per_step'], 3), 'e2e', d['e2e']['value'], 'resid', d['gmres']['resid'], {k: (v['launches_per_step'], round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
N=2
for v in "" "THCM_BALANCE=0" "THCM_NO_FUSED_HEAD=1" "THCM_FUSED_CGS2=0"; do
  name=$(echo "${v:-default}" | tr '= ' '__')
  env $v timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g${N}_$name.json 2> gpurun_out/bench_${TAG}_g${N}_$name.err
  tail -2 gpurun_out/bench_${TAG}_g${N}_$name.err | cut -c1-300
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_g${N}_$name.json'):
    if l.startswith('{'):
        d = json.loads(l); print('$name', 'N=$N step_ms', round(d['ms_per_step'], 3), 'resid', d['gmres']['resid'], {k: (v['launches_per_step'], round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
done
