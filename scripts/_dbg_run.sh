timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r4s_tests.log
timeout 100 python scripts/profile_step.py c4 8 > gpurun_out/r4s_c4.log 2>&1
timeout 100 python scripts/profile_step.py c3 10 > gpurun_out/r4s_c3.log 2>&1
timeout 100 python scripts/profile_step.py c2 20 > gpurun_out/r4s_c2.log 2>&1
