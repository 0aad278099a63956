#!/bin/bash
# one gpurun call: the reacting-eqnset GPU tests, then the reacting objects of the bench (Jacobian refresh timing)
timeout 500 python -m pytest tests/test_gpu_fr.py tests/test_gpu_nsfr.py tests/test_gpu_variants.py tests/test_gpu_unsteady.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_frjac_tests.log
tail -4 gpurun_out/r2_frjac_tests.log
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-sgs --no-ns --no-fma --no-parity-check > gpurun_out/r2_bench_frjac.json 2> gpurun_out/r2_bench_frjac.err
python - <<PY
import json
t = open("gpurun_out/r2_bench_frjac.json").read()
j = json.loads(t[t.index('{"metric"'):])
for k in ("reacting", "reacting_viscous"):
    r = j[k]
    print(k, r["jacobian_refresh_ms"], r["iteration_with_refresh_ms"],
          {a: round(b, 2) for a, b in r["kernels_ms"].items() if "jac" in a or "lu" in a})
PY
