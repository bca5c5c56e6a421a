timeout 300 python -m pytest tests/test_gpu_tgemm.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -n 3 > gpurun_out/r3c_tests.log
timeout 100 python scripts/profile_step.py c4 5 2>&1 | head -n 14 > gpurun_out/r3c_c4.log
