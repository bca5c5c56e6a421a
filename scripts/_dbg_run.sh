timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 > gpurun_out/r3b_tests.log
timeout 100 python scripts/profile_step.py c4 5 2>&1 | grep -E "us/step|grad_reduce|embed|colsum" > gpurun_out/r3b_c4.log
timeout 100 python scripts/profile_step.py c3 10 2>&1 | grep -E "us/step|grad_reduce" > gpurun_out/r3b_c3.log
timeout 100 python scripts/profile_step.py c2 20 2>&1 | grep -E "us/step|grad_reduce" > gpurun_out/r3b_c2.log
