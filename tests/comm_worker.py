"""torchrun worker of tests/test_gpu_comm.py::test_comm_two_processes_over_cuda_ipc: one process per rank, gloo for the
control plane (halo maps, the exchange of the comm blobs), pcfd_comm_* for the data: fields and flag pages mapped through
CUDA IPC.  Two composite iterations on z-slabs, bit-compared with the per-rank oracle replay."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import slab_case
    from proteuscfd_b200.parallel import CommExchange, PObj, TorchGroup
    from tests.oracle_lib import load_oracle
    from tests.partition_oracle import replay_perfect_gas
    implicit = sys.argv[1] == "implicit"
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = lr % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    parts = [slab_case(7, r, world, colored=implicit, cfl=5.0 if implicit else 0.5) for r in range(world)]
    sws, ref = replay_perfect_gas(load_oracle(), parts, implicit, iters=2, nsweeps=3)
    mesh, params, q = parts[rank]
    ctx = capi.Context(mesh, params, device=dev)
    group = TorchGroup(dist)
    pobj = PObj(rank, world).BuildCommMaps(mesh["gNodeOwner"], mesh["gNodeLocalId"], group)
    x = CommExchange(ctx, pobj, group)
    ctx.lsq_coefficients()
    assert np.array_equal(ctx.get_field(capi.F_LSQ_SW), sws[rank]), "sw (ghost rows included)"
    ctx.set_field(capi.F_Q, q)
    for it in range(2):
        if implicit:
            ctx.implicit_iterate(3, refresh_jac=True)
        else:
            ctx.explicit_iterate(refresh_dt=True)
        ctx.synchronize()
        for k, f in (("qgrad", capi.F_QGRAD), ("limiter", capi.F_LIMITER), ("b", capi.F_B), ("q", capi.F_Q)):
            assert np.array_equal(ctx.get_field(f), ref[it][k][rank]), f"{k} it {it}"
        if implicit:
            assert np.array_equal(ctx.get_field(capi.F_X), ref[it]["x"][rank]), f"x it {it}"
    g = x.allgather([float(rank)])
    assert np.array_equal(g[:, 0], np.arange(world))
    dist.barrier()
    x.close()
    ctx.close()
    print(f"RANK_OK {rank} dev={dev}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
