timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4 > gpurun_out/r4y_tests.log
timeout 100 python scripts/profile_step.py c4 8 > gpurun_out/r4y_c4.log 2>&1
timeout 100 python scripts/profile_step.py c1 10 > gpurun_out/r4y_c1.log 2>&1
