timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r4t_tests.log
timeout 100 python scripts/profile_step.py c1 10 2>&1 | head -2 > gpurun_out/r4t_c1.log
timeout 100 python scripts/profile_step.py c3 10 2>&1 | head -2 > gpurun_out/r4t_c3.log
timeout 100 python scripts/profile_step.py c4 8 2>&1 | head -2 > gpurun_out/r4t_c4.log
timeout 100 python scripts/profile_step.py c2 20 2>&1 | head -2 > gpurun_out/r4t_c2.log
