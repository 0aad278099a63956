#!/bin/bash
# one gpurun --gpus 2 call: the multi-rank GPU tests, then the bench at N = 2 as the driver launches it
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_comm.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gpu_tests_2gpu_final.log
tail -2 gpurun_out/r2_gpu_tests_2gpu_final.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/r2_bench_2gpu_final.json 2> gpurun_out/r2_bench_2gpu_final.err
tail -c 400 gpurun_out/r2_bench_2gpu_final.json
