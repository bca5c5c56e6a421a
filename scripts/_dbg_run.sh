timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_shard.py -x -q 2>&1 | tail -n 3 > gpurun_out/r3f_tests.log
timeout 100 python scripts/profile_step.py c4 5 2>&1 | grep -E "ce_|us/step|grad_reduce" > gpurun_out/r3f_c4.log
