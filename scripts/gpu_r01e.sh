# last GPU seconds of round 1: the C++ mirror / reference Krylov templates on the device, the default (L2-tiled fused CGS2) GMRES path
TAG=r01e
timeout 70 python -m pytest tests/test_zz_cpp_mirror.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "gmres or templates_run" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -12 gpurun_out/pytest_gpu_$TAG.log
timeout 45 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.json | cut -c1-300
