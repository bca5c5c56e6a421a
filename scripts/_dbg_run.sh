timeout 200 python -m pytest tests/test_gpu_engine.py -x -q -k "ce_backward or backward_grads" 2>&1 | tail -n 5 > gpurun_out/r4h_ce.log
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r4h_tests.log
timeout 100 python scripts/profile_step.py c4 8 > gpurun_out/r4h_c4.log 2>&1
