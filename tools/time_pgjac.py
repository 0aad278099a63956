#!/usr/bin/env python
"""Per-kernel times of the perfect-gas Jacobian refresh + one implicit iteration at bench size (pcfd profile events), for
A/B runs of kernel variants selected by environment variables.    python tools/time_pgjac.py [--n 118]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=118)
    args = ap.parse_args()
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(args.n, colored=True, device="cuda:0")
    ctx = capi.Context(mesh, params, device=0)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q)
    ctx.set_cfl(5.0)
    ctx.implicit_iterate(2, refresh_jac=True)
    ctx.synchronize()
    ctx.profile(on=True, reset=True)
    for _ in range(3):
        ctx.implicit_iterate(2, refresh_jac=True)
    ctx.synchronize()
    prof = ctx.profile_table()
    out = {k: round(ms / max(cnt, 1), 3) for k, (ms, cnt) in prof.items()}
    out["total"] = round(sum(ms for ms, _ in prof.values()) / 3, 3)
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PCFD_")}, "ms": out}))


if __name__ == "__main__":
    main()
