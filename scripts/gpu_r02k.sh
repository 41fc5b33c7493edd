# r02k (2 GPUs): suite (multi tests at world = 2), smoke, N = 1 / N = 2 benches with the explicit-norm A/B
TAG=${1:-r02k}
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
run() { # gpus name env
  local g=$1 name=$2 envx=$3
  if [ "$g" = "1" ]; then env $envx timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-b1 > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  else env $envx timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $g --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; fi
  tail -2 gpurun_out/bench_${TAG}_$name.err | cut -c1-300
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_$name.json'):
    if l.startswith('{'):
        d = json.loads(l); print('$name', 'step_ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']*1e3, 3), 'resid', d['gmres']['resid'], {k: (v.get('launches_per_step'), round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
}
run 1 g1 ""
run 1 g1_explicit_norm "THCM_EXPLICIT_NORM=1"
run 2 g2 ""
run 2 g2_explicit_norm "THCM_EXPLICIT_NORM=1"
