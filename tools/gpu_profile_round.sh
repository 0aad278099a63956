#!/usr/bin/env bash
# Bounded profiling pass for one gpurun call (B200).  Every ncu invocation carries a launch cap (-c) and a wall-clock
# timeout: an UNBOUNDED `ncu ... python bench.py` serialises ~4000 launches and does not finish in 25 minutes (round 1,
# session 5).  Outputs go to gpurun_out/ (kept under 64 MiB: duration-only CSVs are tiny, each --set full report of a dozen
# launches is 15-30 MB -- keep the -k filters narrow).
#
#   gpurun --timeout 900 -- 'bash tools/gpu_profile_round.sh'
set -u
out=gpurun_out
mkdir -p "$out"

# 0. first B200 run of the variants written after round 1's GPU minutes were spent (Green-Gauss gradients, central-difference
#    Jacobians, their drop-in runs): XPASS = verified, then drop the xfail marker of tests/test_zz_gpu_pending.py.  ~1 min
timeout 300 python -m pytest tests/test_zz_gpu_pending.py -m gpu -q -rxX > "$out/pending_variants.log" 2>&1

# 1. launch list of the bench command itself (explicit iteration + 5x5 SGS), duration only: ~1 min
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fr > "$out/launches_bench.log" 2>&1

# 2. launch lists of one iteration of each eqnset family between cudaProfilerStart/Stop: ~40 s each
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_fr.csv" \
    --profile-from-start off python tools/profile_fr.py --n 118 --viscous --nsgs 2 > "$out/launches_fr.log" 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_pg.csv" \
    --profile-from-start off python tools/profile_run.py --n 118 --nsgs 2 > "$out/launches_pg.log" 2>&1

# 2b. the session-6 variants (Green-Gauss gradient, central-difference Jacobian refresh): duration-only launch list, ~40 s
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_pg_variants.csv" \
    --profile-from-start off python tools/profile_run.py --n 118 --nsgs 2 --green-gauss --jac-central > "$out/launches_pg_variants.log" 2>&1

# 3. --set full of the kernels that carry the headline numbers (one launch each; the SGS tile kernel twice): ~2 min
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_flux_edges|k_gradient|k_limiter|k_residual_gather|k_timestep" -c 6 -o "$out/prof_explicit" \
    python tools/profile_run.py --n 118 --explicit-only --natural > "$out/prof_explicit.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_sgs_tile_t" -c 2 -o "$out/prof_sgs5" python tools/profile_run.py --n 118 --nsgs 1 > "$out/prof_sgs5.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"kfr_flux_edges|kfr_vflux_edges|kfr_clip_neg_edges|k_sgs_tile_t" -c 5 -o "$out/prof_fr" \
    python tools/profile_fr.py --n 118 --viscous --nsgs 1 > "$out/prof_fr.log" 2>&1

ls -la "$out" | tail -20
du -sh "$out"
