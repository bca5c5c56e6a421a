timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 > gpurun_out/r3g_tests.log
timeout 100 python scripts/profile_step.py c4 5 2>&1 | grep -E "ce_|us/step|head" > gpurun_out/r3g_c4.log
timeout 100 python scripts/profile_step.py c3 10 2>&1 | grep -E "ce_|us/step|head" > gpurun_out/r3g_c3.log
timeout 100 python scripts/profile_step.py c1 10 2>&1 | grep -E "ce_|us/step|head" > gpurun_out/r3g_c1.log
timeout 100 python scripts/profile_step.py c2 20 2>&1 | grep -E "ce_|us/step|head" > gpurun_out/r3g_c2.log
