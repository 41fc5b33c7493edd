# r02s (2 GPUs): timing experiment on the halo path of the compact SpMV (results of modes 4 / 3 are numerically wrong on purpose)
TAG=${1:-r02s}
for bps in 4 3; do
THCM_SPMV_HALO_BPS=$bps timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g2_bps$bps.json 2> gpurun_out/bench_${TAG}_g2_bps$bps.err
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_g2_bps$bps.json'):
    if l.startswith('{'):
        d = json.loads(l); print('mode $bps g2 step_ms', round(d['ms_per_step'], 3), 'resid', d['gmres']['resid'], {k: (v.get('launches_per_step'), round(v['avg_ms'], 4)) for k, v in d['kernels'].items()})
PY
done
