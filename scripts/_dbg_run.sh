# scratch driver for `gpurun -- 'bash scripts/_dbg_run.sh'`: GPU test suite + smoke (+ two step timings)
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 > gpurun_out/last_gpu_tests.log
timeout 60 python scripts/profile_step.py c2 20 2>&1 | head -1 > gpurun_out/last_c2.log
timeout 60 python scripts/profile_step.py c4 5 2>&1 | head -1 > gpurun_out/last_c4.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1
