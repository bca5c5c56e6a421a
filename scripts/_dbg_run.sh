for ns in 0 1000 2000 4000; do
B4R_FATTN_STAGGER_NS=$ns timeout 100 python scripts/profile_step.py c4 5 2>&1 | grep -E "attn_bwd|us/step" > gpurun_out/r2w_c4_$ns.log
done
timeout 200 python -m pytest tests/test_gpu_fattn.py -x -q 2>&1 | tail -n 3 > gpurun_out/r2w_fattn.log
