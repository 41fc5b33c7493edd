# final evidence of round 2 on 2 GPUs: multi-rank parity at the head, the bench line as the driver launches it
TAG=${1:-r02final}
mkdir -p gpurun_out
THCM_TEST_WORLD=2 timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_multi_g2_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi_g2_$TAG.log; tail -3 gpurun_out/pytest_multi_g2_$TAG.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-mixing --no-b1 > gpurun_out/bench_${TAG}_g2.json 2> gpurun_out/bench_${TAG}_g2.err
tail -2 gpurun_out/bench_${TAG}_g2.err | cut -c1-200
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_g2.json'):
    if l.startswith('{'):
        d = json.loads(l); print('g2 step_ms', round(d['ms_per_step'], 3), 'resid', d['gmres']['resid'], {k: (v.get('launches_per_step'), round(v['avg_ms'], 4)) for k, v in d['kernels'].items()})
PY
