timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err
timeout 600 python bench.py --gpus 1 --workload c3 --no-secondary --steps 20 --warmup 5 > gpurun_out/r02f_bench_c3_n1.json 2> gpurun_out/r02f_bench_c3_n1.err
