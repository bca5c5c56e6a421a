timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 > gpurun_out/r2o_all.log
timeout 100 python scripts/profile_step.py c4 5 2>&1 | head -n 16 > gpurun_out/r2o_c4.log
timeout 100 python scripts/profile_step.py c2 20 2>&1 | head -n 6 > gpurun_out/r2o_c2.log
