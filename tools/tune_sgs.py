#!/usr/bin/env python
"""Time CRS::SGS sweeps for the tuning variants of k_sgs_level (PCFD_SGS_UNROLL) at bench size."""
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
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    variants = [(0, 5, 0)] + [(w, lpr, pf) for (w, lpr) in ((1, 5), (2, 10), (4, 16), (1, 16)) for pf in (0, -1, 1000, 4000, 16000)]
    xref = None
    for w, lpr, pf in variants:
        os.environ["PCFD_SGS_TILE_WARPS"] = str(w)
        os.environ["PCFD_SGS_TILE_LPR"] = str(lpr)
        os.environ["PCFD_SGS_PREFETCH_TILES"] = str(pf)
        var, u = "tile warps / lanes per row / L2 prefetch distance", (w, lpr, pf)
        c = capi.Context(mesh, params, device=0)
        c.set_stream(stream.cuda_stream)
        c.lsq_coefficients()
        c.set_field(capi.F_Q, q)
        c.implicit_iterate(2, refresh_jac=True)
        c.blank_x()
        c.sgs(2, want_ddq=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        c.sgs(10, want_ddq=False)
        e1.record(stream)
        torch.cuda.synchronize()
        x = c.get_field(capi.F_X)
        if xref is None:
            xref = x
        same = bool((x == xref).all())
        print(f"{var}={u}: {e0.elapsed_time(e1) / 10:.4f} ms/sweep, x bit-identical to the first variant: {same}", flush=True)
        c.close()


if __name__ == "__main__":
    main()
