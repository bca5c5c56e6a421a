timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r4u_tests.log
timeout 100 python scripts/profile_step.py c3 10 > gpurun_out/r4u_c3.log 2>&1 
timeout 100 python scripts/profile_step.py c1 10 2>&1 | head -2 > gpurun_out/r4u_c1.log
timeout 100 python scripts/profile_step.py c2 20 2>&1 | head -2 > gpurun_out/r4u_c2.log
