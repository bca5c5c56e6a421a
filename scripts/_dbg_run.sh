timeout 200 python -m pytest tests/test_gpu_fattn.py -x -q 2>&1 | tail -n 5 > gpurun_out/r4p_fattn.log
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r4p_tests.log
timeout 100 python scripts/profile_step.py c4 8 > gpurun_out/r4p_c4.log 2>&1
timeout 100 python scripts/profile_step.py c3 10 > gpurun_out/r4p_c3.log 2>&1
