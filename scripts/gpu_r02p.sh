# r02p (2 GPUs): multi-rank parity on 2 ranks, then N = 2 bench with per-rank kernel tables (load-balance diagnosis)
TAG=${1:-r02p}
THCM_TEST_WORLD=2 timeout 500 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_multi_g2_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi_g2_$TAG.log; tail -4 gpurun_out/pytest_multi_g2_$TAG.log
THCM_BENCH_RANK_TABLES=gpurun_out timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g2.json 2> gpurun_out/bench_${TAG}_g2.err
tail -2 gpurun_out/bench_${TAG}_g2.err | cut -c1-300
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_g2.json'):
    if l.startswith('{'):
        d = json.loads(l); print('g2 step_ms', round(d['ms_per_step'], 3), 'resid', d['gmres']['resid'])
for r in (0, 1):
    d = json.load(open(f'gpurun_out/kernels_g2_rank{r}.json'))
    print(r, d['ndim'], d['nnz'], {k: (v[0], round(v[1], 4)) for k, v in d['kernels'].items()})
PY
