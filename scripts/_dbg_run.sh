# scratch driver for `gpurun -- 'bash scripts/_dbg_run.sh'`: GPU test suite + smoke
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 > gpurun_out/last_gpu_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1
