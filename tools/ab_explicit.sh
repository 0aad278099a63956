#!/usr/bin/env bash
# A/B of the explicit-iteration variants on one box (env switches of libpcfd_b200.so / bench options), one line per variant
set -u
out=gpurun_out
run() {
  name=$1; shift
  env "$@" > "$out/ab_$name.json" 2> "$out/ab_$name.err"
  python - "$out/ab_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    k = d["phases"]["kernels"]
    print(sys.argv[2], "ms/step %.4f" % d["ms_per_step"], "parity", (d.get("parity_check") or {}).get("ok"), " ".join(f"{n}={v['ms_per_step']:.4f}" for n, v in k.items()))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
B="timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fr --no-sgs --no-ns --no-fma"
run lex $B
run brick $B --order brick
