timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4 > gpurun_out/r4w_tests.log
timeout 100 python scripts/profile_step.py c4 8 > gpurun_out/r4w_c4.log 2>&1
timeout 100 python scripts/profile_step.py c3 10 > gpurun_out/r4w_c3.log 2>&1
