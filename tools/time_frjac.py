#!/usr/bin/env python
"""Per-kernel times of one reacting Jacobian refresh at bench size (pcfd profile events), for A/B runs of kernel variants
selected by environment variables.    python tools/time_frjac.py [--n 118] [--viscous]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=118)
    ap.add_argument("--viscous", action="store_true")
    args = ap.parse_args()
    import bench
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_box_case
    mesh, params, q, beta = fr_box_case(args.n, bench.fr_params_from_fixture(args.viscous), device="cuda:0")
    ctx = capi.Context(mesh, params, device=0)
    ctx.set_field(capi.F_BETA, beta)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q)
    ctx.timestep(want_min=False)
    ctx.jacobian()
    ctx.prepare_sgs()
    ctx.synchronize()
    ctx.profile(on=True, reset=True)
    for _ in range(3):
        ctx.timestep(want_min=False)
        ctx.jacobian()
        ctx.prepare_sgs()
    ctx.synchronize()
    for _ in range(3):
        ctx.implicit_iterate(1, refresh_jac=False)
    ctx.synchronize()
    prof = ctx.profile_table()
    out = {k: round(ms / max(cnt, 1), 3) for k, (ms, cnt) in prof.items()}
    out["total"] = round(sum(ms for ms, _ in prof.values()) / 3, 3)
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PCFD_")}, "ms": out}))


if __name__ == "__main__":
    main()
