#!/usr/bin/env bash
# Round-2 bounded profiling pass (one gpurun call, 1 GPU, ~5 GPU-minutes).  Every ncu invocation carries a launch cap and
# a wall-clock timeout.  Outputs under gpurun_out/ (copied into profiles/ by hand afterwards).
#   gpurun --timeout 900 -- 'bash tools/gpu_profile_round2.sh'
set -u
out=gpurun_out
mkdir -p "$out"
# 1. launch list of the bench command itself (explicit iteration + implicit 5x5 object), duration only
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file "$out/r2_launches_bench.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fr --no-ns --no-fma --no-parity-check > "$out/r2_launches_bench.log" 2>&1
# 2. --set full of the kernels of the explicit iteration (bench numbering) and of one SGS level
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_flux_edges|k_gradient|k_limiter$|k_limiter\(|k_residual_gather|k_explicit|k_eig_bedges|k_update_bcs_edges" -c 8 \
    -o "$out/r2_prof_explicit" python tools/profile_run.py --n 118 --explicit-only --natural > "$out/r2_prof_explicit.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_sgs_tile_t" -c 2 -o "$out/r2_prof_sgs5" python tools/profile_run.py --n 118 --nsgs 1 > "$out/r2_prof_sgs5.log" 2>&1
python tools/ncu_summary.py "$out/r2_prof_explicit.ncu-rep" "$out/r2_prof_sgs5.ncu-rep" > "$out/r2_ncu_tables.md" 2>&1
cat "$out/r2_ncu_tables.md"
ls -la "$out" | tail -12
