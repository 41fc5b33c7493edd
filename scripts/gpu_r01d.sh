# final verification of round 1 (templated coupled kernels, fused CGS2 default, C++ mirror / reference Krylov templates on the device)
TAG=r01d
timeout 200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -8 gpurun_out/pytest_gpu_$TAG.log
timeout 90 ncu --set full --clock-control none --import-source on -k regex:"fused_axpy_dot" -s 40 -c 2 -o gpurun_out/prof_${TAG}_fused -f python scripts/prof_kernels.py krylov > gpurun_out/ncu_full_fused_$TAG.log 2>&1; tail -2 gpurun_out/ncu_full_fused_$TAG.log
timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; tail -c 300 gpurun_out/ncu_bench_$TAG.log
