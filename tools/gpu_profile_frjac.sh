#!/usr/bin/env bash
# ncu --set full of the reacting Jacobian refresh kernels (one launch each) at bench size; one gpurun call, 1 GPU.
#   gpurun --timeout 900 -- 'bash tools/gpu_profile_frjac.sh'
set -u
out=gpurun_out
mkdir -p "$out"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"kfr_jac_edges|kfr_jac_bedges|kfr_jac_node|kfr_jac_diag|k_lu_diag_lanes|kfr_vjac_edges" -c 6 \
    -o "$out/r2_prof_frjac" python tools/profile_run.py --n 118 --fr --fr-viscous > "$out/r2_prof_frjac.log" 2>&1
tail -3 "$out/r2_prof_frjac.log"
python tools/ncu_summary.py "$out/r2_prof_frjac.ncu-rep" > "$out/r2_ncu_frjac.md" 2>&1
cat "$out/r2_ncu_frjac.md"
