set -e
for w in c4 c3 c1; do
  python scripts/ncu_step.py $w 3 > gpurun_out/plain_$w.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$w.csv python scripts/ncu_step.py $w 3 > gpurun_out/ncu_$w.log 2>&1
done
python scripts/ncu_step.py c4 2 > gpurun_out/plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fattn_bwd|fattn_fwd" -s 8 -c 2 -o gpurun_out/r02_fattn_c4 python scripts/ncu_step.py c4 2 > gpurun_out/ncu.log 2>&1
