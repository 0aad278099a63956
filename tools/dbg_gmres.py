import os, sys, time
sys.path.insert(0, "/root/repo")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ["PCFD_COMM_SPIN_SECONDS"] = "3"
import numpy as np
from proteuscfd_b200 import capi
from proteuscfd_b200.cases import slab_case
from proteuscfd_b200.parallel import CommExchange, PObj, ThreadGroup
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 2
parts = [slab_case(6, r, nr, colored=True, cfl=5.0) for r in range(nr)]
t00 = time.time()
def log(rank, msg):
    print(f"[{time.time()-t00:7.3f}] r{rank}: {msg}", flush=True)
def fn(rank, group):
    mesh, params, q = parts[rank]
    ctx = capi.Context(mesh, params)
    ctx.device_ptr(capi.F_A)
    ctx.gmres(1, 20, 0)
    pobj = PObj(rank, nr).BuildCommMaps(mesh["gNodeOwner"], mesh["gNodeLocalId"], group)
    x = CommExchange(ctx, pobj, group)
    ctx.lsq_coefficients(); ctx.set_field(capi.F_Q, q)
    ctx.timestep(want_min=False); ctx.jacobian(); ctx.update_bcs(); x.update(capi.F_Q)
    ctx.gradient(); x.update(capi.F_QGRAD); ctx.limiter(); x.update(capi.F_LIMITER); ctx.residual(); ctx.blank_x()
    ctx.lib.pcfd_synchronize(ctx.h); log(rank, f"assembled {ctx.comm_debug_flags(nr)}")
    for k in range(4):
        a = x.allgather([rank + k]); log(rank, f"allgather {k} {a[:,0]}")
    try:
        dq = ctx.gmres(int(sys.argv[2]) if len(sys.argv) > 2 else 1, int(sys.argv[3]) if len(sys.argv) > 3 else 3, 2); log(rank, f"gmres ok {dq}")
    except Exception as e:
        log(rank, f"gmres failed {e} {ctx.comm_debug_flags(nr)}")
    group.allgather(None)
    x.close(); ctx.close()
ThreadGroup(nr).run(fn)
print("DONE")
