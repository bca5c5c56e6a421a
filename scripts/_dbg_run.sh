timeout 300 python -m pytest tests/test_gpu_tablegrad.py -x -q 2>&1 | tail -n 15 > gpurun_out/r4b_tg.log
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/r4b_tests.log
timeout 100 python scripts/profile_step.py c4 5 > gpurun_out/r4b_c4.log 2>&1
timeout 100 python scripts/profile_step.py c2 20 > gpurun_out/r4b_c2.log 2>&1
timeout 100 python scripts/profile_step.py c3 10 > gpurun_out/r4b_c3.log 2>&1
