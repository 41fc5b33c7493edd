# r02f (1 GPU): the two bench arms exactly as the driver runs them, the GPU suite, launch list + ncu --set full (exported as CSV on the box:
# the .ncu-rep files are too large to travel back)
TAG=${1:-r02f}
( time python bench.py ) > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -4 gpurun_out/bench_${TAG}.err
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}.json'):
    if l.startswith('{'):
        d = json.loads(l); print('ours step_ms', round(d['ms_per_step'], 3), 'e2e', d['e2e']['value'], 'b1', d.get('e2e_b1'), 'cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'), 'asm+spmv', d.get('assembly_plus_spmv'), 'roof', d['roofline'])
        print({k: (v.get('launches_per_step'), round(v['avg_ms'], 4), round(v.get('frac_of_peak', 0), 3), round(v.get('graph_equivalent_frac_of_peak', 0), 3)) for k, v in d['kernels'].items()})
PY
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -4 gpurun_out/bench_${TAG}_reference.err; cut -c1-400 gpurun_out/bench_${TAG}_reference.json
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-b1 > gpurun_out/ncu_bench_$TAG.log 2>&1
mkdir -p /tmp/ncu
cap() { # name, kernel regex, skip, count, workload
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/ncu/$1 -f python scripts/prof_kernels.py $5 > gpurun_out/ncu_full_$1_$TAG.log 2>&1
  ncu -i /tmp/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1_$TAG.csv 2>/dev/null
  ls -la /tmp/ncu/$1.ncu-rep | awk '{print $5}'
}
cap asm "jac_tma|thcm_assemble|rhs_tma|blockdiag_build|spmv_csr" 5 7 asm
cap krylov25 "multi_dot|fused|multi_axpy_dot|spmv_compact|scale_precon" 118 8 krylov
cap krylov50 "multi_dot|fused|multi_axpy_dot|spmv_compact|scale_precon" 238 5 krylov
ncu -i /tmp/ncu/asm.ncu-rep --page source --csv --kernel-name regex:jac_tma 2>/dev/null | head -c 3000000 > gpurun_out/ncu_source_jac_$TAG.csv
cuobjdump -sass i-emic_b200/libthcm_b200.so 2>/dev/null | grep -E "Function :|UBLKCP|SYNCS|UTMA" | grep -B1 -E "UBLKCP|SYNCS|UTMA" | head -80 > gpurun_out/sass_tma_$TAG.txt
du -sh gpurun_out
