# r02n (1 GPU): the whole -m gpu suite first (consistent vertical mixing, row-group residual kernels, one-wave multi_dot), then the r02f
# measurement script (both bench arms as the driver runs them, smoke, launch list, ncu --set full exports)
TAG=${1:-r02n}
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -5 gpurun_out/pytest_gpu_$TAG.log
bash scripts/gpu_r02f.sh $TAG
