timeout 60 python scripts/_dbg_c4.py 1024 ctr 2>&1 | head -n 8 > gpurun_out/r2j_ctr.log
timeout 100 python scripts/profile_step.py c4 5 > gpurun_out/r2k_c4.log 2>&1
timeout 100 python scripts/profile_step.py c3 10 > gpurun_out/r2k_c3.log 2>&1
timeout 100 python scripts/profile_step.py c1 10 > gpurun_out/r2k_c1.log 2>&1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r2k_all.log
