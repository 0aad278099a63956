#!/usr/bin/env python
"""Driver for ncu on the SGS sweep: builds the implicit case at bench size, warms up, then runs ONE sweep between
cudaProfilerStart/Stop.

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_sgs -c 4 \
        -o gpurun_out/prof_sgs python tools/profile_sgs.py 118
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 118
    import torch
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(n, cfl=5.0, colored=True, device="cuda:0")
    c = capi.Context(mesh, params, device=0)
    c.lsq_coefficients()
    c.set_field(capi.F_Q, q)
    c.implicit_iterate(1, refresh_jac=True)
    c.blank_x()
    c.sgs(1, want_ddq=False)
    c.synchronize()
    torch.cuda.profiler.start()
    c.sgs(1, want_ddq=False)
    c.synchronize()
    torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()
