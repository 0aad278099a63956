#!/usr/bin/env python
"""Driver for ncu: one reacting-eqnset implicit iteration (Jacobian refresh + 1 SGS sweep) at bench size, bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees only those launches.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_fr \
        python tools/profile_fr.py --n 118
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=118)
    ap.add_argument("--nsgs", type=int, default=1)
    ap.add_argument("--viscous", action="store_true", help="compressibleNSFR: viscous flux / Jacobian with species transport")
    args = ap.parse_args()
    import torch
    from bench import fr_params_from_fixture
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_box_case
    mesh, params, q, beta = fr_box_case(args.n, fr_params_from_fixture(args.viscous), device="cuda:0")
    ctx = capi.Context(mesh, params, device=0)
    ctx.set_field(capi.F_BETA, beta)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q)
    ctx.implicit_iterate(args.nsgs, refresh_jac=True)      # warm-up
    ctx.set_field(capi.F_Q, q)
    ctx.synchronize()
    torch.cuda.profiler.start()
    ctx.implicit_iterate(args.nsgs, refresh_jac=True)
    ctx.synchronize()
    torch.cuda.profiler.stop()
    print("profiled launches done; total launches", ctx.launch_count())


if __name__ == "__main__":
    main()
