# r02t (2 GPUs): halo landing in the head kernel + unified-gather SpMV: multi-rank parity, then the bench with and without it
TAG=${1:-r02t}
THCM_TEST_WORLD=2 timeout 500 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_multi_g2_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi_g2_$TAG.log; tail -3 gpurun_out/pytest_multi_g2_$TAG.log
for mode in landing polling; do
if [ $mode = polling ]; then export THCM_NO_HALO_LANDING=1; else unset THCM_NO_HALO_LANDING; fi
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g2_$mode.json 2> gpurun_out/bench_${TAG}_g2_$mode.err
tail -2 gpurun_out/bench_${TAG}_g2_$mode.err | cut -c1-200
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_g2_$mode.json'):
    if l.startswith('{'):
        d = json.loads(l); print('$mode g2 step_ms', round(d['ms_per_step'], 3), 'resid', d['gmres']['resid'], {k: (v.get('launches_per_step'), round(v['avg_ms'], 4)) for k, v in d['kernels'].items()})
PY
done
