import os, sys, time, threading
sys.path.insert(0, "/root/repo")
os.environ["PCFD_COMM_SPIN_SECONDS"] = "3"
import numpy as np
from proteuscfd_b200 import capi
from proteuscfd_b200.cases import slab_case
from proteuscfd_b200.parallel import CommExchange, PObj, ThreadGroup
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 2
barrier_after = len(sys.argv) > 2 and sys.argv[2] == "barrier"
parts = [slab_case(5, r, nr) for r in range(nr)]
t00 = time.time()
def log(rank, msg):
    print(f"[{time.time()-t00:7.3f}] r{rank}: {msg}", flush=True)
def fn(rank, group):
    mesh, params, q = parts[rank]
    ctx = capi.Context(mesh, params)
    pobj = PObj(rank, nr).BuildCommMaps(mesh["gNodeOwner"], mesh["gNodeLocalId"], group)
    x = CommExchange(ctx, pobj, group)
    log(rank, "connected")
    if barrier_after:
        group.allgather(None)
    for k in range(3):
        ctx.comm_post(capi.F_LSQ_S); log(rank, f"post {k} launched")
        ctx.comm_wait(capi.F_LSQ_S); log(rank, f"wait {k} launched")
        ctx.lib.pcfd_synchronize(ctx.h); log(rank, f"sync {k} done {ctx.comm_debug_flags(nr)}")
    ctx.lsq_coefficients(); ctx.set_field(capi.F_Q, q)
    for it in range(2):
        ctx.explicit_iterate(); log(rank, f"iter {it} launched")
        ctx.lib.pcfd_synchronize(ctx.h); log(rank, f"iter {it} done {ctx.comm_debug_flags(nr)}")
    group.allgather(None)
    x.close(); ctx.close()
ThreadGroup(nr).run(fn)
print("DONE")
