timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/r4d_tests.log
timeout 100 python scripts/profile_step.py c4 8 > gpurun_out/r4d_c4.log 2>&1
timeout 100 python scripts/profile_step.py c3 10 > gpurun_out/r4d_c3.log 2>&1
timeout 100 python scripts/micro/hbm_rw.py > gpurun_out/r4d_hbm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tgemm_kernel -s 2 -c 1 -o gpurun_out/r4d_ffn1 -f python scripts/ncu_step.py c4 2 > gpurun_out/r4d_ncu.log 2>&1
