for b in 4 3; do THCM_FUSED2_BPS=$b timeout 20 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01g_bps$b.json 2> gpurun_out/bench_r01g_bps$b.err; python -c "
import json
for l in open('gpurun_out/bench_r01g_bps$b.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bps $b', d['ms_per_step'], d['gmres']['resid'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items() if 'multi' in k})
"; done
