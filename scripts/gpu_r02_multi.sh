# Multi-GPU pass (gpurun --gpus N -- bash scripts/gpu_r02_multi.sh TAG N): the multi-rank parity tests on N ranks, then the bench on N ranks
# with the A/B switches that matter there.  Every step is bounded so a lost hand-off cannot hang the box.
TAG=${1:-r02m}
N=${2:-2}
shift 2
THCM_TEST_WORLD=$N timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_multi_g${N}_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi_g${N}_$TAG.log; tail -12 gpurun_out/pytest_multi_g${N}_$TAG.log
for v in "" "$@"; do
  name=$(echo "${v:-default}" | tr '= ' '__')
  env $v timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g${N}_$name.json 2> gpurun_out/bench_${TAG}_g${N}_$name.err
  tail -3 gpurun_out/bench_${TAG}_g${N}_$name.err | cut -c1-300
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_g${N}_$name.json'):
    if l.startswith('{'):
        d = json.loads(l); print('$name', 'N=$N step_ms', round(d['ms_per_step'], 3), 'resid', d['gmres']['resid'], {k: (v['launches_per_step'], round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
done
