timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4 > gpurun_out/r4x_tests.log
timeout 100 python scripts/profile_step.py c4 8 > gpurun_out/r4x_c4.log 2>&1
timeout 100 python scripts/profile_step.py c3 10 > gpurun_out/r4x_c3.log 2>&1
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4x_smoke.log 2>&1
