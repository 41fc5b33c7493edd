set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "variants" > gpurun_out/pytest_variants.log 2>&1; echo rc=$? >> gpurun_out/pytest_variants.log
tail -5 gpurun_out/pytest_variants.log
rm -f gpurun_out/perf_pipe.jsonl
for v in 0 1 4; do THCM_ASM_PIPE=$v timeout 300 python tests/perf_kernels.py >> gpurun_out/perf_pipe.jsonl 2>gpurun_out/perf_pipe_$v.err; done
python - <<'PY'
import json
for l in open('gpurun_out/perf_pipe.jsonl'):
    d=json.loads(l); print(d['env'], {k:v for k,v in d['kernels'].items() if 'assemble' in k})
PY
