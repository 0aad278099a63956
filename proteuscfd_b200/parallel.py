"""Host side of the halo exchange: the PObj of the reference (ucs/parallel.{h,tcc}), one process per GPU.

`PObj.BuildCommMaps` reproduces `PObj<Type>::BuildCommMaps` (ucs/parallel.tcc:461-554): from the ghost tables of
the parted mesh (`Ghost Nodes Owning Process`, `Ghost Nodes Local Id`, ucs/decomp.cpp:355-388) it derives
`commCountsRecv/commOffsetsRecv`, asks every owner for its nodes and receives `commCountsSend` and
`nodePackingList` -- bit-for-bit the reference's maps (tests/test_parallel_maps.py checks that against maps dumped
from the reference itself).  The lists are built ONCE and live on the device (the reference re-sends them on
every call, parallel.tcc:809-827).

`PObj.UpdateGeneralVectors(field)` is the exchange (parallel.tcc:779-873), owner -> ghost.  Two data paths:

* "nccl":  k_halo_pack into a staging buffer, then grouped ncclSend/ncclRecv (torch.distributed
  batch_isend_irecv) straight into the contiguous ghost segment of the field (no unpack kernel);
* "put":   k_halo_pack writes DIRECTLY into the peer GPU's ghost segment through a CUDA-IPC mapping (one
  kernel per peer, stores travel over NVLink), bracketed by two stream-ordered barriers.

Groups: `TorchGroup` (torch.distributed: nccl on GPUs, gloo in CPU tests) and `LocalGroup` (several ranks inside
one process, for single-GPU loopback tests).
"""
import numpy as np

from . import capi


def recv_maps(nranks, g_node_owner):
    """commCountsRecv / commOffsetsRecv (parallel.tcc:482-511).  Ghost nodes are grouped by ascending owner."""
    owner = np.asarray(g_node_owner, dtype=np.int64)
    if owner.size and np.any(np.diff(owner) < 0):
        raise ValueError("ghost nodes are not grouped by owning process")
    counts = np.bincount(owner, minlength=nranks).astype(np.int32)
    offsets = np.zeros(nranks, dtype=np.int32)
    offsets[1:] = np.cumsum(counts)[:-1]
    return counts, offsets


class LocalGroup:
    """All ranks live in this process (loopback).  Collective calls take one entry per rank."""

    def __init__(self, nranks):
        self.nranks = nranks

    def alltoall_lists(self, per_rank_wants):
        """per_rank_wants[r][p] = ids rank r requests from owner p  ->  out[r][p] = ids rank p requests from r."""
        n = self.nranks
        return [[per_rank_wants[p][r] for p in range(n)] for r in range(n)]


class TorchGroup:
    """One rank per process over torch.distributed (nccl or gloo)."""

    def __init__(self, dist):
        self.dist = dist
        self.rank = dist.get_rank()
        self.nranks = dist.get_world_size()

    def alltoall_lists(self, wants):
        """wants[p] = ids this rank requests from owner p -> list of ids each peer requests from this rank."""
        gathered = [None] * self.nranks
        self.dist.all_gather_object(gathered, [np.asarray(w, dtype=np.int32) for w in wants])
        return [gathered[p][self.rank] for p in range(self.nranks)]

    def allgather(self, obj):
        out = [None] * self.nranks
        self.dist.all_gather_object(out, obj)
        return out


class PObj:
    """Halo maps + exchange of one rank (the reference's PObj<Type>)."""

    def __init__(self, rank, nranks):
        self.rank, self.np = int(rank), int(nranks)
        self.commCountsRecv = self.commOffsetsRecv = self.commCountsSend = self.commOffsetsSend = None
        self.nodePackingList = None
        self.ctx = None

    # ---- parallel.tcc:461-554
    def wants(self, g_node_owner, g_node_local_id):
        self.commCountsRecv, self.commOffsetsRecv = recv_maps(self.np, g_node_owner)
        ids = np.asarray(g_node_local_id, dtype=np.int32)
        return [ids[self.commOffsetsRecv[p]: self.commOffsetsRecv[p] + self.commCountsRecv[p]] for p in range(self.np)]

    def set_requests(self, requested_by_peer):
        """requested_by_peer[p] = local ids peer p wants, in p's ghost order."""
        self.commCountsSend = np.array([len(requested_by_peer[p]) if p != self.rank else 0 for p in range(self.np)],
                                       dtype=np.int32)
        self.commOffsetsSend = np.zeros(self.np, dtype=np.int32)
        self.commOffsetsSend[1:] = np.cumsum(self.commCountsSend)[:-1]
        parts = [np.asarray(requested_by_peer[p], dtype=np.int32) for p in range(self.np) if p != self.rank]
        self.nodePackingList = np.concatenate(parts) if parts else np.zeros(0, np.int32)

    def BuildCommMaps(self, g_node_owner, g_node_local_id, group):
        """One-rank-per-process form (TorchGroup)."""
        self.set_requests(group.alltoall_lists(self.wants(g_node_owner, g_node_local_id)))
        return self

    def attach(self, ctx):
        """Hand the persistent maps to the device context."""
        self.ctx = ctx
        ctx.halo_configure(self.rank, self.np, self.commCountsSend, self.nodePackingList, self.commCountsRecv)
        return self

    # ---- numpy model of the exchange (CPU tests of the host logic; also documents the data layout)
    def pack_numpy(self, v, n):
        v = np.asarray(v).reshape(-1, n)
        return [v[self.nodePackingList[self.commOffsetsSend[p]: self.commOffsetsSend[p] + self.commCountsSend[p]]]
                for p in range(self.np)]

    def unpack_numpy(self, v, n, nnode, from_peer):
        v = v.reshape(-1, n)
        for p in range(self.np):
            if p != self.rank and self.commCountsRecv[p]:
                o = nnode + self.commOffsetsRecv[p]
                v[o: o + self.commCountsRecv[p]] = from_peer[p]


def build_local_group_maps(ghost_tables):
    """ghost_tables[r] = (gNodeOwner, gNodeLocalId) of rank r -> [PObj] with maps built (LocalGroup)."""
    n = len(ghost_tables)
    pobjs = [PObj(r, n) for r in range(n)]
    wants = [pobjs[r].wants(*ghost_tables[r]) for r in range(n)]
    reqs = LocalGroup(n).alltoall_lists(wants)
    for r in range(n):
        pobjs[r].set_requests(reqs[r])
    return pobjs


# ---------------------------------------------------------------------------------------------- device exchanges
class LoopbackExchange:
    """Several ranks' contexts in one process on one GPU: rank r's pack kernel writes straight into rank p's ghost
    segment (the same direct-put code path as the IPC exchange, minus the IPC mapping)."""

    def __init__(self, ctxs, pobjs):
        self.ctxs, self.pobjs = ctxs, pobjs
        for c, p in zip(ctxs, pobjs):
            p.attach(c)

    def update(self, field):
        for c in self.ctxs:
            c.synchronize()
        n = len(self.ctxs)
        for r in range(n):
            for p in range(n):
                if p != r and self.pobjs[r].commCountsSend[p]:
                    self.ctxs[r].halo_pack(field, self.ctxs[p].halo_recv_ptr(field, peer=r), peer=p)
        for c in self.ctxs:
            c.synchronize()


class _DevView:
    """Expose raw device memory to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, ndoubles):
        self.__cuda_array_interface__ = {"shape": (int(ndoubles),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class NcclExchange:
    """pack -> grouped send/recv; receives land directly in the ghost segment."""

    def __init__(self, ctx, pobj, dist, torch, device):
        self.ctx, self.p, self.dist, self.torch = ctx, pobj, dist, torch
        pobj.attach(ctx)
        # the pack kernel and ncclSend must share a stream: the context launches on torch's current one from here on
        cur = torch.cuda.current_stream(device).cuda_stream
        if cur == 0:
            raise RuntimeError("NcclExchange: make a non-default torch stream current first (the context cannot launch on "
                               "the legacy default stream; pcfd_set_stream(NULL) selects its own stream)")
        ctx.set_stream(cur)
        self.device = device
        self.stage = {}
        self.views = {}

    def _bufs(self, field):
        if field not in self.stage:
            n = self.ctx.halo_width(field)
            total = int(self.p.commCountsSend.sum())
            self.stage[field] = self.torch.empty(max(total * n, 1), dtype=self.torch.float64, device=self.device)
            g = int(self.p.commCountsRecv.sum())
            base = self.ctx.halo_recv_ptr(field)
            self.views[field] = self.torch.as_tensor(_DevView(base, max(g * n, 1)), device=self.device) if g else None
        return self.stage[field], self.views[field]

    def update(self, field):
        torch, dist, p = self.torch, self.dist, self.p
        n = self.ctx.halo_width(field)
        stage, ghost = self._bufs(field)
        self.ctx.halo_pack(field, stage.data_ptr())          # on the context's stream (= torch's current stream)
        ops = []
        for peer in range(p.np):
            if peer == p.rank:
                continue
            if p.commCountsSend[peer]:
                o = int(p.commOffsetsSend[peer]) * n
                ops.append(dist.P2POp(dist.isend, stage[o: o + int(p.commCountsSend[peer]) * n], peer))
            if p.commCountsRecv[peer]:
                o = int(p.commOffsetsRecv[peer]) * n
                ops.append(dist.P2POp(dist.irecv, ghost[o: o + int(p.commCountsRecv[peer]) * n], peer))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()                                     # stream-ordered for nccl: no host sync


class PutExchange:
    """Direct put over NVLink: every rank maps its peers' fields with CUDA IPC once; an exchange is one
    k_halo_pack per peer whose destination is the peer's ghost segment, between two stream-ordered barriers
    (a 1-element NCCL all-reduce each): the first makes sure every rank is done READING the ghost rows that are
    about to be overwritten, the second that every put has landed before anyone reads."""

    FIELDS = (capi.F_Q, capi.F_QGRAD, capi.F_LIMITER, capi.F_X, capi.F_LSQ_S, capi.F_LSQ_SW)
    TURB_FIELDS = (capi.F_TVAR, capi.F_TGRAD, capi.F_TURB_X)    # allocated only with a turbulence model

    def __init__(self, ctx, pobj, dist, torch, device, group):
        self.ctx, self.p, self.dist, self.torch = ctx, pobj, dist, torch
        pobj.attach(ctx)
        if ctx.field_size(capi.F_TVAR) > 0:
            self.FIELDS = self.FIELDS + self.TURB_FIELDS
        self.token = torch.zeros(1, device=device)
        mine = {"nnode": ctx.nnode, "recv_offsets": [int(o) for o in pobj.commOffsetsRecv],
                "handles": {f: ctx.ipc_export(f) for f in self.FIELDS}}
        everyone = group.allgather(mine)
        self.dst = {f: {} for f in self.FIELDS}
        self.opened = []
        for peer, info in enumerate(everyone):
            if peer == pobj.rank or pobj.commCountsSend[peer] == 0:
                continue
            for f in self.FIELDS:
                base = ctx.ipc_open(info["handles"][f])
                self.opened.append(base)
                n = ctx.halo_width(f)
                # my rows land behind the peer's own nodes, at the peer's receive offset for me
                self.dst[f][peer] = base + (info["nnode"] + info["recv_offsets"][pobj.rank]) * n * 8

    def _barrier(self):
        self.dist.all_reduce(self.token)     # on the current stream: orders the puts against the peers' kernels

    def update(self, field):
        if field not in self.dst:
            raise KeyError(f"PutExchange: field {field} was not exported (not allocated in this context)")
        self._barrier()
        for peer, dst in self.dst[field].items():
            self.ctx.halo_pack(field, dst, peer=peer)
        self._barrier()

    def close(self):
        for b in self.opened:
            self.ctx.ipc_close(b)
        self.opened = []


class ThreadGroup:
    """Several ranks as THREADS of one process (one context each, one GPU or several): the collective calls of
    TorchGroup on a threading.Barrier.  ctypes releases the GIL inside the library, so a rank blocked in a stream
    synchronisation does not stop the others -- the same relation MPI ranks have."""

    def __init__(self, nranks):
        import threading
        self.nranks = nranks
        self._bar = threading.Barrier(nranks)
        self._slots = [None] * nranks

    def allgather(self, rank, obj):
        self._slots[rank] = obj
        self._bar.wait()
        out = list(self._slots)
        self._bar.wait()
        return out

    def view(self, rank):
        return _ThreadGroupRank(self, rank)

    def run(self, fn):
        """fn(rank, group_view) on one thread per rank; re-raises the first failure."""
        import threading
        errs = [None] * self.nranks

        def body(r):
            try:
                fn(r, self.view(r))
            except BaseException as e:      # noqa: BLE001 -- reported to the caller below
                errs[r] = e
                self._bar.abort()

        ts = [threading.Thread(target=body, args=(r,)) for r in range(self.nranks)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        real = [e for e in errs if e is not None and not isinstance(e, __import__("threading").BrokenBarrierError)]
        if real:
            raise real[0]
        if any(errs):
            raise [e for e in errs if e is not None][0]


class _ThreadGroupRank:
    def __init__(self, group, rank):
        self.g, self.rank, self.nranks = group, rank, group.nranks

    def allgather(self, obj):
        return self.g.allgather(self.rank, obj)

    def alltoall_lists(self, wants):
        gathered = self.g.allgather(self.rank, [np.asarray(w, dtype=np.int32) for w in wants])
        return [gathered[p][self.rank] for p in range(self.nranks)]


class CommExchange:
    """The library's own exchange (pcfd_comm_*, csrc/pcfd_comm.cuh): every rank publishes one blob (CUDA-IPC handles of
    its fields and of its flag page), `group.allgather` moves the blobs, and from pcfd_comm_connect on an exchange is a
    put kernel + a wait kernel with epoch flags in peer memory -- no collective, no host synchronisation.  Connecting
    also switches the composite entry points (pcfd_explicit_iterate, pcfd_implicit_iterate, pcfd_turb_compute,
    pcfd_lsq_coefficients) to the reference's multi-rank sequence."""

    def __init__(self, ctx, pobj, group):
        self.ctx, self.p = ctx, pobj
        pobj.attach(ctx)
        ctx.comm_connect(group.allgather(ctx.comm_export()))
        # nobody posts before everybody is connected (with the ranks as threads of one process a synchronous CUDA call
        # of a rank that is still connecting would otherwise wait for a peer's put kernel that waits for that rank)
        group.allgather(None)

    def update(self, field):
        self.ctx.comm_update(field)

    def post(self, field):
        self.ctx.comm_post(field)

    def wait(self, field):
        self.ctx.comm_wait(field)

    def allgather(self, vals):
        return self.ctx.comm_allgather(vals, self.p.np)

    def close(self):
        self.ctx.comm_disconnect()


class TransposeMaps:
    """PObj::TransposeCommCRS (ucs/parallel.tcc:54-338) as maps: after CRSMatrix::CRSTranspose has transposed every
    block and swapped the local mirror blocks, the block of every ghost column (local row i, ghost g) is replaced by the
    block the owner of g holds for the mirrored cut edge -- its row j = gNodeLocalId[g], its ghost column of (this rank,
    i) -- which the owner has already transposed in place.  The reference sends (row, column) pairs as global ids and the
    owner maps them back (:227-287); here a request is the pair (j, i) of LOCAL ids, which the owner resolves through
    its own ghost table.  ghost_edges: the parallel half-edges bedges_n[nbedge .. nbedge+ngedge) as (local node, ghost)
    pairs -- the order pcfd_crs_ghost_blocks packs the blocks in."""

    def __init__(self, rank, nranks, nnode, ghost_edges, g_node_owner, g_node_local_id):
        self.rank, self.np = int(rank), int(nranks)
        ge = np.asarray(ghost_edges, dtype=np.int64).reshape(-1, 2)
        owner = np.asarray(g_node_owner, dtype=np.int64)
        lid = np.asarray(g_node_local_id, dtype=np.int64)
        gi = ge[:, 1] - nnode
        if ge.size and (gi.min() < 0 or ge[:, 0].max() >= nnode):
            raise ValueError("TransposeMaps: a parallel half-edge joins a local node and a ghost")
        eo = owner[gi]
        self.requests = [np.stack([lid[gi[eo == p]], ge[eo == p, 0]], axis=1).astype(np.int32) for p in range(self.np)]
        self.slots = [np.nonzero(eo == p)[0] for p in range(self.np)]
        # what a peer may ask for: (my row, peer, the peer's local id of my ghost column) -> my half-edge slot
        self.lookup = {(int(ge[e, 0]), int(eo[e]), int(lid[gi[e]])): e for e in range(ge.shape[0])}

    def serve(self, peer, pairs, blocks):
        """blocks this rank owes `peer` for its request pairs (j = row here, i = the peer's local node)"""
        idx = [self.lookup[(int(j), int(peer), int(i))] for j, i in np.asarray(pairs).reshape(-1, 2)]
        return blocks[idx]

    def place(self, blocks, from_peer):
        out = np.array(blocks, copy=True)
        for p in range(self.np):
            if self.slots[p].size:
                out[self.slots[p]] = from_peer[p]
        return out


def transpose_ghost_blocks_local(maps, blocks):
    """all ranks in this process: blocks[r] = rank r's ghost-column blocks [ngedge_r, n, n] after the local transpose;
    returns what TransposeCommCRS leaves in them"""
    n = len(maps)
    return [maps[r].place(blocks[r], [maps[p].serve(r, maps[r].requests[p], blocks[p]) if p != r else None
                                      for p in range(n)]) for r in range(n)]


def crs_transpose(ctx, mesh, group):
    """CRSMatrix::CRSTranspose of the context's assembled matrix on a partition (one rank per process or thread):
    local part on the device (pcfd_crs_transpose), ghost-column blocks routed through `group.allgather` of host
    buffers -- once per adjoint solve, off the iteration's path.  mesh: the dict the context was created from."""
    rank, nranks = group.rank, group.nranks
    nnode, nb, ng = int(mesh["nnode"]), int(mesh["nbedge"]), int(mesh["ngedge"])
    ge = np.asarray(mesh["bedges_n"]).reshape(-1, 2)[nb: nb + ng]
    maps = TransposeMaps(rank, nranks, nnode, ge, mesh["gNodeOwner"], mesh["gNodeLocalId"])
    ctx.crs_transpose()
    blocks = ctx.get_ghost_blocks()
    asked = group.allgather(maps.requests)                       # asked[r][p]: what r requests of p
    served = group.allgather([maps.serve(r, asked[r][rank], blocks) if r != rank else None for r in range(nranks)])
    ctx.set_ghost_blocks(maps.place(blocks, [served[p][rank] if p != rank else None for p in range(nranks)]))
    return maps


class DistributedHotPath:
    """SolutionSpace::NewtonIterate (ucs/solutionSpace.tcc:640-904) across ranks: the phase calls of one context
    with the reference's halo exchanges and reductions in the reference's places."""

    def __init__(self, ctx, exchange, allreduce_sum=None, allreduce_min=None, any_rank=None, fused=True):
        """any_rank(flag) -> True if the flag is set on any rank (a max all-reduce); with it the limiter / residual
        pair runs as pcfd_limiter_raw -> halo -> pcfd_residual_fused (one edge pass instead of two, no host sync
        inside the limiter) and falls back to the ordered pressure-clip path only when some rank reports a clip."""
        self.ctx, self.x = ctx, exchange
        self.sum = allreduce_sum or (lambda a: a)
        self.min = allreduce_min or (lambda a: a)
        self.any_rank = any_rank
        self.fused = fused and any_rank is not None
        self.clip_fallbacks = 0

    def setup(self):
        # Gradient::ComputeNodeLSQCoefficients ends with halos of s and sw (gradient.tcc:131-134)
        self.ctx.lsq_coefficients()
        self.x.update(capi.F_LSQ_S)
        self.x.update(capi.F_LSQ_SW)

    def head(self, want_norms=False):
        c = self.ctx
        c.update_bcs()                      # solutionSpace.tcc:662
        self.x.update(capi.F_Q)             # :665
        c.gradient()
        self.x.update(capi.F_QGRAD)         # gradient.tcc:98
        if self.fused:
            c.limiter_raw()
            self.x.update(capi.F_LIMITER)   # limiters.tcc:128 (raw values; both sides clamp)
            norms, hit = c.residual_fused(want_norms=want_norms)
            if not self.any_rank(hit):
                return norms
            self.clip_fallbacks += 1
        c.limiter()
        self.x.update(capi.F_LIMITER)       # limiters.tcc:128
        return c.residual(want_norms=want_norms)

    def turb_compute(self, nsgs, want_norm=False):
        """TurbulenceModel::Compute (ucs/turb.tcc:163-339, Spalart-Allmaras) across ranks: the phases of pcfd_turb_phase
        with the reference's exchanges in the reference's places; returns the all-reduced sum(b^2) when asked.
        Block-Jacobi across partitions, like CRS::SGS (ghost values of x lag one sweep)."""
        c = self.ctx
        c.turb_phase(0)
        self.x.update(capi.F_TVAR)          # turb.tcc:185
        c.turb_phase(1)
        self.x.update(capi.F_TGRAD)         # gradient.tcc:98
        s = c.turb_phase(2, want_norm=want_norm)
        if want_norm:
            s = self.sum(s)                 # ParallelL2Norm, turb.tcc:256
        for _ in range(nsgs):
            c.turb_phase(3)
            self.x.update(capi.F_TURB_X)    # crs.tcc:146
        c.turb_phase(4)
        self.x.update(capi.F_TVAR)          # turb.tcc:325
        c.turb_phase(5)
        return s

    def explicit_iterate(self, refresh_dt=True):
        c = self.ctx
        if refresh_dt:
            c.timestep(want_min=False)
        self.head()
        c.explicit_solve()
        self.x.update(capi.F_Q)             # solutionSpace.tcc:857

    def implicit_iterate(self, nsgs, refresh_jac=True):
        c = self.ctx
        if refresh_jac:
            c.timestep(want_min=False)
            c.jacobian()
        self.head()
        c.prepare_sgs()
        c.blank_x()
        self.x.update(capi.F_X)             # crs.tcc:88
        for _ in range(nsgs):
            c.sgs(1, want_ddq=False)
            self.x.update(capi.F_X)         # crs.tcc:146
        c.apply_dq()
        self.x.update(capi.F_Q)
