#!/bin/bash
# one gpurun call: the whole GPU suite, smoke(), the default bench line
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu_tests_final.log
tail -3 gpurun_out/r2_gpu_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 600 gpurun_out/r2_bench_final.json
