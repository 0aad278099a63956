#!/bin/bash
# the bench at N = 8 as the driver launches it (one gpurun --gpus 8 call)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/r2_bench_8gpu_final.json 2> gpurun_out/r2_bench_8gpu_final.err
tail -c 300 gpurun_out/r2_bench_8gpu_final.json
