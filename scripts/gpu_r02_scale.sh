# Scaling pass on an N-GPU box (gpurun --gpus N -- bash scripts/gpu_r02_scale.sh TAG N "n1 n2 ..." [weak]): multi-rank parity tests on N ranks
# (log kept), then the bench at each listed GPU count, optionally the weak-scaling configuration at N.  Every step is bounded.
TAG=${1:-r02s}
N=${2:-4}
LIST=${3:-"1 2 4"}
WEAK=${4:-}
THCM_TEST_WORLD=$N timeout 500 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_multi_g${N}_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi_g${N}_$TAG.log; tail -6 gpurun_out/pytest_multi_g${N}_$TAG.log
run() { # gpus, name, extra args, env
  local g=$1 name=$2; shift 2
  if [ "$g" = "1" ]; then
    env $ENVX timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-b1 "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  else
    env $ENVX timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $g --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  fi
  tail -2 gpurun_out/bench_${TAG}_$name.err | cut -c1-300
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_$name.json'):
    if l.startswith('{'):
        d = json.loads(l); print('$name', 'step_ms', round(d['ms_per_step'], 3), 'grid', d['config']['grid'], 'resid', d['gmres']['resid'], 'ocean_cells', d.get('ocean_cells_local'), {k: (v.get('launches_per_step'), round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
}
for g in $LIST; do ENVX="" run $g g$g; done
if [ -n "$WEAK" ]; then ENVX="" run $N g${N}_weak --weak; fi
if [ "$N" != "1" ] && [ -n "$5" ]; then ENVX="THCM_BALANCE=0" run $N g${N}_uniform_cuts; fi
