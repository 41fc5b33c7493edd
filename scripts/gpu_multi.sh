N=${1:-2}
if [ "${2:-test}" = "test" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi.log; tail -15 gpurun_out/pytest_multi.log; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_p2p_g$N.json 2> gpurun_out/bench_p2p_g$N.err; tail -c 600 gpurun_out/bench_p2p_g$N.err
python - <<PY
import json
for l in open('gpurun_out/bench_p2p_g$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value']); print({k:(v['launches_per_step'],round(v['avg_ms'],4)) for k,v in d['kernels'].items()})
PY
