# Multi-GPU follow-up (gpurun --gpus 2 ...): the multi-rank tests, then the fused CGS2 kernels on 2 ranks (off by default there: the
# all-reduce tail is shared with multi_dot but has not run fused yet) -- bounded, so a lost hand-off cannot hang the box
TAG=${1:-r02m}
N=${2:-2}
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_multi_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi_$TAG.log; tail -5 gpurun_out/pytest_multi_$TAG.log
for f in 0 2; do
  THCM_FUSED_CGS2=$f timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g${N}_fused$f.json 2> gpurun_out/bench_${TAG}_g${N}_fused$f.err
  grep '^{' gpurun_out/bench_${TAG}_g${N}_fused$f.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused $f', d['ms_per_step'], d['gmres'])"
done
