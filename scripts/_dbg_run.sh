timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4 > gpurun_out/r02f_gpu_tests.log
timeout 700 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err
for c in c1 c2 c3 c4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches_$c.csv python scripts/ncu_step.py $c 2 > gpurun_out/r02f_ncu_$c.log 2>&1
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"fattn|tg_chunk|tg_boundary|trowln|twgrad|ln_bwd" -s 20 -c 12 -o gpurun_out/r02f_ncu_c4 -f python scripts/ncu_step.py c4 2 > gpurun_out/r02f_ncu_full.log 2>&1
