timeout 200 python -m pytest tests/test_gpu_fattn.py -x -q 2>&1 | tail -n 3 > gpurun_out/r3e_fattn.log
timeout 100 python scripts/profile_step.py c4 5 2>&1 | grep -E "attn_|us/step" > gpurun_out/r3e_c4.log
timeout 100 python scripts/profile_step.py c3 10 2>&1 | grep -E "attn_|us/step" > gpurun_out/r3e_c3.log
