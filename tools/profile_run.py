#!/usr/bin/env python
"""Small driver for ncu: one explicit and one implicit iteration of the hot path at bench size,
bracketed by cudaProfilerStart/Stop so `ncu --profile-from-start off` sees only those launches.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof \
        python tools/profile_run.py --n 118
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
    ap.add_argument("--explicit-only", action="store_true")
    ap.add_argument("--natural", action="store_true", help="lexicographic node numbering (the explicit bench case)")
    ap.add_argument("--green-gauss", action="store_true", help="Param::gradType = 1 (k_gradient_gg)")
    ap.add_argument("--jac-central", action="store_true", help="fieldJacType = boundaryJacType = 1 (k_jac_*_central)")
    ap.add_argument("--fr", action="store_true", help="reacting eqnset: one Jacobian refresh (kfr_jac_*), nothing else")
    ap.add_argument("--fr-viscous", action="store_true", help="with --fr: compressibleNSFR")
    args = ap.parse_args()
    import torch
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    if args.fr:
        import bench
        from proteuscfd_b200.cases import fr_box_case
        mesh, params, q, beta = fr_box_case(args.n, bench.fr_params_from_fixture(args.fr_viscous), device="cuda:0")
        ctx = capi.Context(mesh, params, device=0)
        ctx.set_field(capi.F_BETA, beta)
        ctx.lsq_coefficients()
        ctx.set_field(capi.F_Q, q)
        ctx.timestep(want_min=False)
        ctx.jacobian()           # warm-up (allocates A)
        ctx.synchronize()
        torch.cuda.profiler.start()
        ctx.jacobian()
        ctx.prepare_sgs()
        ctx.synchronize()
        torch.cuda.profiler.stop()
        print("profiled launches done; total launches", ctx.launch_count())
        return
    mesh, params, q = box_case(args.n, colored=not args.natural, device="cuda:0")
    ctx = capi.Context(mesh, params, device=0)
    if args.green_gauss:
        ctx.set_gradient_type(1)
    if args.jac_central:
        ctx.set_jacobian_type(1, 1)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q)
    ctx.explicit_iterate(refresh_dt=True)      # warm-up
    ctx.synchronize()
    torch.cuda.profiler.start()
    ctx.explicit_iterate(refresh_dt=True)
    if not args.explicit_only:
        ctx.set_cfl(5.0)
        ctx.implicit_iterate(args.nsgs, refresh_jac=True)
    ctx.synchronize()
    torch.cuda.profiler.stop()
    print("profiled launches done; total launches", ctx.launch_count())


if __name__ == "__main__":
    main()
