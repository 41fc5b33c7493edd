./tests/cpp/_bin/drop_in_krylov tests/golden/masks/mask_natl8 > gpurun_out/drop_in_krylov_r01f.json 2> gpurun_out/drop_in_krylov_r01f.err; echo rc=$?
timeout 50 python -m pytest tests/test_zz_cpp_mirror.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_r01f.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_r01f.log; tail -5 gpurun_out/pytest_gpu_r01f.log
