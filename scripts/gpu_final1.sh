set -x
TAG=${1:-r01b}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_$TAG.log; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jac_tma|thcm_assemble|spmv_csr|blockdiag_apply" -s 8 -c 6 -o gpurun_out/prof_${TAG}_asm -f python scripts/prof_kernels.py asm > gpurun_out/ncu_full_asm_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"multi_dot|multi_axpy" -s 100 -c 4 -o gpurun_out/prof_${TAG}_krylov -f python scripts/prof_kernels.py krylov > gpurun_out/ncu_full_krylov_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_krylov_$TAG.log
