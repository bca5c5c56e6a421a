timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/r4c_tests.log
timeout 100 python scripts/profile_step.py c4 8 > gpurun_out/r4c_c4.log 2>&1
