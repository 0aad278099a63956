#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (contract: see the task statement, section 4).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--n 118]

Workload (BASELINE.json configs[1]): inviscid Euler on a synthetic ~10 M-cell
tetrahedralised box (Kuhn box, n = 118 hexes per side: 9.86 M tets, 1.69 M nodes,
11.7 M edges), second-order Roe + weighted-LSQ gradient + Venkatakrishnan limiter,
explicit update.  One "step" = one explicit iteration exactly as
SolutionSpace::NewtonIterate runs it with numberSGS = 0 (ucs/solutionSpace.tcc:640-904):
ComputeTimesteps, UpdateBCs, Gradient::Compute, Limiter::Compute, ComputeResiduals,
ExplicitSolve.  `value` = interior edges x steps / time: million edges per second
through the WHOLE iteration; per-phase figures (residual-only Medges/s, SGS sweeps/s)
ride along under "phases" / "sgs".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "explicit_iteration_throughput (UpdateBCs+LSQ gradient+Venkat limiter+Roe residual+timestep+update)"
UNIT = "Medges/s"
NEQN, NVARS, NTERMS = 5, 10, 9


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# SURVEY.md 8d algorithmic (compulsory) bytes per pass, 5-equation Euler
def pass_bytes(ne, nn, nblocks=0):
    neqn, nterms = NEQN, NTERMS
    return {
        "gradient": 8 * ne + nn * (24 + 8 * nterms + 48 + 24 * nterms),
        "limiter": (8 * ne + nn * (8 * neqn + 16 * neqn)) + (8 * ne + nn * (8 * neqn + 24 * neqn + 24 + 16 * neqn) + nn * 8 * neqn)
                   + (8 * ne + nn * (8 * neqn + 24 * neqn + 24 + 8 * neqn) + nn * 8 * neqn),
        "residual": 40 * ne + nn * (8 * neqn + 24 * neqn + 8 * neqn + 24 + 8) + nn * 8 * neqn,
        "timestep": 40 * ne + nn * (8 * neqn + 8) + nn * 8,
        "sgs_sweep": 2 * (nblocks * 8 * neqn * neqn + 4 * nblocks + 4 * (nn + 1) + 4 * neqn * nn + 8 * neqn * (nn + 2 * nn)),
    }


# which kernels make up which reference pass
PASS_KERNELS = {
    "timestep": ["k_timestep", "k_eig_bedges"],      # explicit iteration: the edge terms ride along in k_flux_edges / k_residual_gather
    "update_bcs": ["k_update_bcs"],
    "gradient": ["k_gradient"],
    "limiter": ["k_limiter", "k_fill_int", "k_clip_edges", "k_clip_nodes", "k_limiter_final"],

    "residual": ["k_flux_edges", "k_flux_bedges", "k_residual_gather"],
    "update": ["k_explicit"],
}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation (oracle/_ref, built from /root/reference by
    oracle/Makefile) on all host cores, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import ref_bench
    cores = os.cpu_count() or 1
    ranks = 1
    while ranks * 2 <= min(cores, 64):
        ranks *= 2
    n = args.ref_n
    if not ref_bench.available():
        # the C port of the same loops (oracle/pcfd_oracle.c), single core
        v, sample, ms = port_baseline(min(n, 40), max(1, args.steps))
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": sample},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    case = ref_bench.ReferenceCase(n, ranks, limiter=2, nsgs=0, cfl=0.5)
    try:
        if args.warmup > 0:
            case.time(min(args.warmup, 1))
        t = case.time(max(1, args.steps))
    finally:
        case.close()
    per_it = t["t_timestep"] + t["t_update"] + t["t_gradient"] + t["t_limiter"] + t["t_residual"]
    v = t["nedge"] / per_it / 1e6
    sample = (f"Kuhn box n={n} ({6 * n ** 3} tets, {t['nnode']} nodes, {t['nedge']} edges), {ranks} reference ranks "
              f"(process-based MPI shim, block partitions cut by the reference's udecomp), {max(1, args.steps)} iterations")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_it * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "inviscid Euler, Roe 2nd order + LSQ + Venkatakrishnan, explicit; " + sample},
            "phases": {"residual_Medges_s": t["nedge"] / t["t_residual"] / 1e6,
                       "gradient_limiter_Medges_s": t["nedge"] / (t["t_gradient"] + t["t_limiter"]) / 1e6,
                       "seconds": {k: t[k] for k in t if k.startswith("t_")}},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ranks, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def port_baseline(n, iters):
    """Single-core C port (oracle/pcfd_oracle.c) of one explicit iteration on an n-box."""
    from proteuscfd_b200.cases import box_case
    from tests.oracle_lib import load_oracle
    from tests.oracle_lib import oracle_for
    mesh, params, q = box_case(n)
    o = oracle_for(load_oracle(), mesh, params)
    _, sw = o.lsq()
    beta = np.zeros(1)
    t0 = time.time()
    for _ in range(iters):
        dt, _ = o.timestep(q, beta)
        o.update_bcs(q, beta)
        grad = o.gradient(q, sw)
        lim = o.limiter(q, grad)
        b = o.residual(q, grad, lim, beta)
        o.explicit_solve(q, b, dt)
    el = (time.time() - t0) / iters
    sample = f"Kuhn box n={n} ({mesh['nnode']} nodes, {mesh['nedge']} edges), C port of the reference loops, 1 core, {iters} iterations"
    return mesh["nedge"] / el / 1e6, sample, el * 1e3


def cpu_baseline(args):
    """Bounded reference-CPU sample for the default line (rank 0, N = 1): ~10-30 s of CPU work."""
    from oracle import ref_bench
    if ref_bench.available():
        cores = os.cpu_count() or 1
        ranks = 1
        while ranks * 2 <= min(cores, 64):
            ranks *= 2
        n = args.ref_n
        case = ref_bench.ReferenceCase(n, ranks, limiter=2, nsgs=0, cfl=0.5)
        try:
            t = case.time(3)
        finally:
            case.close()
        per_it = t["t_timestep"] + t["t_update"] + t["t_gradient"] + t["t_limiter"] + t["t_residual"]
        sample = (f"Kuhn box n={n} ({t['nnode']} nodes, {t['nedge']} edges), unmodified reference (oracle/_ref/ref_harness), "
                  f"{ranks} ranks, 3 iterations")
        out = {"value": t["nedge"] / per_it / 1e6, "unit": UNIT, "cores": ranks, "kind": "reference", "sample": sample,
               "residual_Medges_s": t["nedge"] / t["t_residual"] / 1e6}
        # the second half of the metric: the reference's CRS::SGS on the implicit version of the same sample (5 sweeps per
        # iteration, Jacobian assembly and LU timed apart); sweeps/s depends on the mesh size, node-sweeps/s does not
        try:
            case = ref_bench.ReferenceCase(n, ranks, limiter=2, nsgs=5, cfl=5.0)
            try:
                ti = case.time(2, timeout=240)   # never run on the GPU box's core count yet: bounded
            finally:
                case.close()
            out["sgs"] = {"sweeps_per_s": 5.0 / ti["t_sgs"], "node_sweeps_per_s": 5.0 * ti["nnode"] / ti["t_sgs"],
                          "jacobian_s": ti["t_jacobian"], "lu_s": ti["t_lu"],
                          "implicit_iteration_s": (ti["t_update"] + ti["t_gradient"] + ti["t_limiter"] + ti["t_residual"] + ti["t_sgs"]),
                          "sample": f"same box, implicit (CFL 5, 5 sweeps), {ranks} ranks, block-Jacobi across ranks, 2 iterations"}
        except Exception as e:
            out["sgs"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        return out
    v, sample, _ = port_baseline(40, 3)
    return {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}


class RankRun:
    """One rank's context + exchange on a (possibly partitioned) case: the calls the sub-benchmarks time.

    --halo comm (default): the library's own exchange (pcfd_comm_*).  After pcfd_comm_connect the composite C entry
    points run the reference's multi-rank sequence themselves, so explicit() / implicit() are single C calls exactly
    as on one GPU.  --halo put / nccl: the host-driven exchanges of round 1 (DistributedHotPath)."""

    def __init__(self, args, mesh, params, rank, world, local_rank, torch, dist, stream, beta=None, implicit=False):
        from proteuscfd_b200 import capi
        self.capi, self.torch, self.dist = capi, torch, dist
        self.rank, self.world, self.halo = rank, world, args.halo
        self.dev = torch.device("cuda", local_rank)
        c = self.ctx = capi.Context(mesh, params, device=local_rank)
        c.set_stream(stream.cuda_stream)
        if beta is not None:
            c.set_field(capi.F_BETA, beta)
        if implicit:
            c.device_ptr(capi.F_A)        # allocate the matrix now, not inside the first timed iteration
        self.x = self.hp = None
        self.halo_rows = 0
        self.turb = int(params.get("turb_model", 0)) == 1
        if world > 1:
            from proteuscfd_b200.parallel import CommExchange, DistributedHotPath, NcclExchange, PObj, PutExchange, TorchGroup
            group = TorchGroup(dist)
            pobj = PObj(rank, world).BuildCommMaps(mesh["gNodeOwner"], mesh["gNodeLocalId"], group)
            self.halo_rows = int(pobj.commCountsSend.sum())
            self.neighbours = int((pobj.commCountsSend > 0).sum())
            if args.halo == "comm":
                self.x = CommExchange(c, pobj, group)
            else:
                self.x = (PutExchange(c, pobj, dist, torch, self.dev, group) if args.halo == "put"
                          else NcclExchange(c, pobj, dist, torch, self.dev))
                hitflag = torch.zeros(1, device=self.dev, dtype=torch.int32)

                def any_rank(hit):   # the (rare) pressure-clip fallback is a global decision: one 4-byte max all-reduce
                    hitflag.fill_(int(hit))
                    dist.all_reduce(hitflag, op=dist.ReduceOp.MAX)
                    return bool(hitflag.item())

                self.hp = DistributedHotPath(c, self.x, any_rank=any_rank)
        if self.hp is not None:
            self.hp.setup()
        else:
            c.lsq_coefficients()          # across ranks: the halos of s / sw are inside (gradient.tcc:131-134)

    def explicit(self):
        if self.hp is not None:
            self.hp.explicit_iterate(refresh_dt=True)
        else:
            self.ctx.explicit_iterate(refresh_dt=True)

    def implicit(self, nsgs, refresh):
        if self.hp is not None:
            self.hp.implicit_iterate(nsgs, refresh_jac=refresh)
            if self.turb:
                self.hp.turb_compute(nsgs)
        else:
            self.ctx.implicit_iterate(nsgs, refresh_jac=refresh)      # TurbulenceModel::Compute included

    def refresh_only(self):
        c = self.ctx
        c.timestep(want_min=False)
        c.jacobian()
        c.prepare_sgs()

    def sweep(self):
        self.ctx.sgs(1, want_ddq=False)
        if self.x is not None:
            self.x.update(self.capi.F_X)      # crs.tcc:146

    def blank_x(self):
        self.ctx.blank_x()
        if self.x is not None:
            self.x.update(self.capi.F_X)      # crs.tcc:88

    def clip_fallbacks(self):
        return self.hp.clip_fallbacks if self.hp is not None else self.ctx.clip_fallbacks()

    def sync(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxms(self, ms):
        if self.world == 1:
            return float(ms)
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def maxi(self, v):
        if self.world == 1:
            return int(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.item())

    def total(self, v):
        if self.world == 1:
            return int(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.int64)
        self.dist.all_reduce(t)
        return int(t.item())

    def timed(self, fn, reps, stream):
        """max over ranks of the device time of reps calls of fn, per call [ms]"""
        ev = lambda: self.torch.cuda.Event(enable_timing=True)
        self.sync()
        e0, e1 = ev(), ev()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        self.sync()
        return self.maxms(e0.elapsed_time(e1) / reps)

    def close(self):
        if self.world > 1:
            self.sync()
        if self.x is not None and hasattr(self.x, "close"):
            self.x.close()
        if self.world > 1:
            self.sync()          # every rank has unmapped its peers before anybody frees the exported memory
        self.ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=118, help="hexes per box side (118 -> ~10 M tets)")
    ap.add_argument("--ref-n", type=int, default=48, help="box side of the bounded CPU reference sample")
    ap.add_argument("--halo", default="comm", choices=["comm", "put", "nccl"],
                    help="N>1 halo exchange.  comm: the library's own flag-based direct puts (pcfd_comm_*: no collective, no "
                         "host sync, interior-edge flux overlaps the halos; the composite C entry points run the multi-rank "
                         "sequence); put: host-driven direct puts between two NCCL barriers (round 1); nccl: NCCL send/recv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sgs", action="store_true")
    ap.add_argument("--no-fr", action="store_true", help="skip the reacting-eqnset (compressibleEulerFR) sub-benchmark")
    ap.add_argument("--fr-n", type=int, default=0, help="box side for the reacting sub-benchmark (0: same as --n)")
    ap.add_argument("--sgs-n", type=int, default=0, help="box side for the SGS sub-benchmark (0: same as --n)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the oracle comparison outside the timed region")
    ap.add_argument("--no-ns", action="store_true", help="skip the laminar-NS (configs[2]) and RANS-SA (configs[3]) objects")
    ap.add_argument("--ns-n", type=int, default=149, help="box side of the laminar Navier-Stokes case (149 -> ~20 M tets)")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the fixed-size (strong scaling) 50 M-cell object")
    ap.add_argument("--strong-n", type=int, default=203, help="box side of the fixed-size case (203 -> ~50 M tets)")
    ap.add_argument("--no-fma", action="store_true", help="skip the FMA-contracted build's timing / drift measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W = max(3, args.warmup)
    K = max(1, args.steps)
    t_setup = time.time()
    # weak scaling: the domain grows with N (a box of n x n x (n*N) hexes cut into N z-slabs, one per GPU, in the
    # layout udecomp writes); cut edges are ghost half-edges and node states cross ranks through the halo exchange
    # at the reference's four places per iteration (q, qgrad, limiter, q)
    if world > 1:
        from proteuscfd_b200.cases import slab_case
        mesh, params, q0 = slab_case(args.n, rank, world, device=f"cuda:{local_rank}")
    else:
        mesh, params, q0 = box_case(args.n, device=f"cuda:{local_rank}")
    # a real (non-default) torch stream: the library launches on it and torch.cuda.Event times it
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    run = RankRun(args, mesh, params, rank, world, local_rank, torch, dist, stream)
    ctx = run.ctx
    ne, nn = ctx.nedge, ctx.nnode
    ctx.set_field(capi.F_Q, q0)
    step = run.explicit
    ne_global = run.total(2 * ctx.nedge + ctx.ngedge) // 2 if world > 1 else ne    # a cut edge lives on two ranks
    halo_rows = run.halo_rows
    t_setup = time.time() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing
    for _ in range(W):
        step()
    # per-kernel event records between the launches cost ~2 % of a 3 ms step: the headline region runs without them,
    # a profiled pass of the same K steps follows for the per-kernel / per-pass tables
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    ms_step = run.maxms(ms_total) / K
    value = ne_global / (ms_step * 1e-3) / 1e6
    ctx.profile(on=True, reset=True)
    Kp = min(K, 20)
    for _ in range(Kp):
        step()
    barrier()
    ctx.profile(on=False)
    table = ctx.profile_table()

    # ---------------------------------------------------------------- end to end through the C ABI with HOST buffers
    nq = ctx.field_size(capi.F_Q)
    hq = torch.empty(nq, dtype=torch.float64).pin_memory()
    hq.numpy()[:] = ctx.get_field(capi.F_Q)
    hq_np = hq.numpy()
    for _ in range(2):
        ctx.set_field(capi.F_Q, hq_np)
        step()
        ctx.get_field(capi.F_Q, out=hq_np)
    Ke = max(3, K // 2)
    barrier()
    e0.record(stream)
    for _ in range(Ke):
        ctx.set_field(capi.F_Q, hq_np)          # H2D of the step's input state (pinned)
        step()
        ctx.get_field(capi.F_Q, out=hq_np)      # D2H of the updated state
    e1.record(stream)
    barrier()
    e2e_ms = run.maxms(e0.elapsed_time(e1) / Ke)
    e2e_value = ne_global / (e2e_ms * 1e-3) / 1e6

    # ---------------------------------------------------------------- per-pass roofline from the live kernel timings
    peak, peak_src = measured_peaks()
    pb = pass_bytes(ne, nn)
    kern = {k: {"ms_avg": v[0] / max(v[1], 1), "launches_per_step": v[1] / Kp, "ms_per_step": v[0] / Kp} for k, v in table.items()}
    ms_prof = sum(v["ms_per_step"] for v in kern.values())
    passes = {}
    for pname, ks in PASS_KERNELS.items():
        ms = sum(kern[k]["ms_per_step"] for k in ks if k in kern)
        if ms <= 0:
            continue
        d = {"ms_per_step": ms, "share": ms / ms_prof}
        if pname in pb:
            d["algorithmic_bytes"] = pb[pname]
            d["GBps"] = pb[pname] / (ms * 1e-3) / 1e9
            d["frac_hbm"] = d["GBps"] / peak
        passes[pname] = d
    # ComputeTimesteps rides along in the residual pass (k_flux_edges / k_residual_gather) when no k_timestep ran: its
    # compulsory bytes are then credited to that pass
    if "k_timestep" not in kern and "residual" in passes:
        r = passes["residual"]
        r["algorithmic_bytes"] = pb["residual"] + pb["timestep"]
        r["includes"] = "ComputeTimesteps (edge terms evaluated in k_flux_edges, summed in k_residual_gather)"
        r["GBps"] = r["algorithmic_bytes"] / (r["ms_per_step"] * 1e-3) / 1e9
        r["frac_hbm"] = r["GBps"] / peak
    dominant = max((p for p in passes if "GBps" in passes[p]), key=lambda p: passes[p]["ms_per_step"])
    # measured DRAM traffic of the pass's kernels (one ncu --set full capture of this workload, committed under profiles/)
    traffic, traffic_src = None, None
    for fn_, key in (("r2_dram_traffic.json", "explicit_lexicographic"), ("r1_dram_traffic.json", "explicit_lexicographic")):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", fn_)))[key]
            if args.n == 118 and all(k in tr for k in PASS_KERNELS[dominant] if k in kern):
                traffic = sum(tr[k] for k in PASS_KERNELS[dominant] if k in kern)
                traffic_src = f"profiles/{fn_} (ncu dram__bytes_read+write.sum per launch, same workload)"
                break
        except Exception:
            continue
    roofline = {"bound": "hbm", "kernel": "+".join(k for k in PASS_KERNELS[dominant] if k in kern), "pass": dominant,
                "achieved": passes[dominant]["GBps"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": passes[dominant]["frac_hbm"], "traffic": traffic, "traffic_source": traffic_src,
                "bytes_per_launch": passes[dominant]["algorithmic_bytes"], "ms_per_launch": passes[dominant]["ms_per_step"],
                "note": "algorithmic bytes = SURVEY.md 8d compulsory bytes of the pass; the Roe flux is FP64-pipe-bound "
                        "(see profiles/), so frac is reported for completeness, not as the limiter"}
    # the dominant kernel is FP64-pipe-bound, not HBM-bound: put the pipe roofline next to the HBM one.  FP64-pipe
    # instructions per edge of k_flux_edges<true, true> (static SASS count, `cuobjdump -sass`; the rarely taken
    # raw-limiter branch included, so an upper bound on the dynamic count); B200: 148 SMs x 64 FP64 lanes, one
    # instruction per lane per clock
    try:
        ipe = FLUX_FP64_INSTR["with_timestep" if "k_timestep" not in kern else "plain"]
        sm_mhz = float((clocks or {}).get("sm_max_mhz") or 1965.0)
        bound_ms = ipe * ne / (148 * 64 * sm_mhz * 1e6) * 1e3
        ms_flux = kern["k_flux_edges"]["ms_per_step"]
        roofline["fp64_pipe"] = {"kernel": "k_flux_edges", "instr_per_edge": ipe, "bound_ms": bound_ms, "ms": ms_flux,
                                 "frac": bound_ms / ms_flux,
                                 "how": "static SASS count of FP64-pipe instructions x edges / (148 SMs x 64 lanes x SM clock); 64 lanes per SM = "
                                        "the nominal 37 TFLOP/s FMA peak at 1.965 GHz, not measured here"}
    except Exception:
        pass
    step_bytes = sum(pb[p] for p in ("gradient", "limiter", "residual", "timestep"))
    phases = {"residual_Medges_s": ne / (passes["residual"]["ms_per_step"] * 1e-3) / 1e6 if "residual" in passes else None,
              "gradient_limiter_Medges_s": ne / (sum(passes[p]["ms_per_step"] for p in ("gradient", "limiter") if p in passes) * 1e-3) / 1e6,
              "iteration_frac_hbm": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
              "profiled_ms_per_step": ms_prof, "passes": passes, "kernels": kern}
    clip_fallbacks = run.clip_fallbacks()

    # ---------------------------------------------------------------- parity outside the timed region
    parity = None
    if not args.no_parity_check:
        try:
            parity = parity_check(args, mesh, params, q0, rank, world, local_rank, torch, dist, stream)
        except Exception as e:
            parity = {"error": f"{type(e).__name__}: {e}"[:300]}

    run.close()
    torch.cuda.empty_cache()

    def guarded(fn, *a, **kw):
        try:
            return fn(*a, **kw)
        except Exception as e:      # never lose the headline line to a sub-benchmark
            return {"error": f"{type(e).__name__}: {e}"[:300]}

    common = (args, rank, world, local_rank, peak, torch, dist, stream)
    sgs = None if args.no_sgs else guarded(implicit_bench, *common, kind="euler")
    laminar = rans = None
    if not args.no_ns:
        if world <= 2:
            laminar = guarded(implicit_bench, *common, kind="laminar_ns")
        rans = guarded(implicit_bench, *common, kind="rans_sa")
    strong = None
    if world > 1 and not args.no_strong:
        strong = guarded(strong_bench, *common)
    frb = frv = None
    if not args.no_fr:
        if world > 1:
            frb = guarded(fr_bench_multi, args, rank, world, local_rank, peak, torch, dist, stream)
        elif rank == 0:
            frb = guarded(fr_bench, args, peak, torch, stream, local_rank)
            frv = guarded(fr_bench, args, peak, torch, stream, local_rank, viscous=True)
    fma = None
    if world == 1 and not args.no_fma:
        fma = guarded(fma_variant, args, mesh, params, q0, torch, stream, local_rank, ms_step)

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            base = cpu_baseline(args)
        except Exception as e:   # never lose the GPU line to a baseline hiccup
            base = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {e}"[:300]}

    if rank == 0:
        halo_txt = {"comm": "library exchange pcfd_comm_* (one put kernel with epoch-flag handshake + one wait kernel per exchange, "
                            "no collective, no host sync; the interior-edge flux kernel runs while qgrad / limiter rows are in flight)",
                    "put": "host-driven direct puts between two NCCL barriers", "nccl": "NCCL send/recv"}[args.halo]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: inviscid Euler, synthetic Kuhn box n={args.n} ({6 * args.n ** 3} tets, "
                                   f"{nn} nodes, {ne} edges) per GPU, Roe 2nd order + weighted LSQ + Venkatakrishnan, "
                                   "explicit single-stage update (the reference has no multistage RK)",
                       "parallelism": ("single partition" if world == 1 else
                                       f"{world} z-slab partitions of an n x n x {args.n * world} box, one per GPU; halos of q, qgrad, "
                                       f"limiter, q per iteration ({halo_rows} rows/rank/exchange): {halo_txt}; "
                                       f"{ne_global} edges in total"),
                       "cache": "inputs larger than L2 (q 135 MB, qgrad 364 MB, edges 470 MB per pass); no flush needed",
                       "setup_s": t_setup},
            "roofline": roofline, "cpu_baseline": base,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": nq * 8 * world,
                    "d2h_bytes_per_step": nq * 8 * world, "what": "pcfd_set_field(q) from pinned host memory + pcfd_explicit_iterate + "
                                                          "pcfd_get_field(q) per step"},
            "gpu_launches": int(launches) * world, "clocks": clocks, "clip_fallbacks": clip_fallbacks,
            "parity_check": parity, "phases": phases, "sgs": sgs, "laminar_ns": laminar, "rans_sa": rans,
            "strong_scaling": strong, "reacting": frb, "reacting_viscous": frv, "fma_variant": fma,
            "reference_arm_note": "bench.py --impl reference runs the UNMODIFIED reference (oracle/_ref/ref_harness) on the bounded "
                                  f"sample n={args.ref_n} (not n={args.n}), 16-32 ranks over the process-based MPI shim "
                                  "(oracle/mpi_shim), harness timers: a stated CPU baseline, not the same configuration",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# FP64-pipe instructions per edge of k_flux_edges<true, EIG> (static SASS counts: DADD + DMUL + DFMA + DSETP + MUFU.RCP64H/RSQ64H)
FLUX_FP64_INSTR = {"plain": 1274, "with_timestep": 1347}    # 315+451+455+53 | 331+474+488+54


def parity_check(args, mesh, params, q0, rank, world, local_rank, torch, dist, stream):
    """Correctness bit carried by the bench line, computed OUTSIDE the timed regions with the C oracle as the checker.
    N = 1: one explicit iteration through pcfd_explicit_iterate on the bench mesh itself (n = 118) against the oracle
    phase by phase.  N > 1: two explicit and two implicit composite iterations on n = 32 slabs through the SAME exchange
    the timed region used, against the oracle replayed rank by rank with a numpy halo (tests/partition_oracle.py)."""
    from proteuscfd_b200 import capi
    from tests.oracle_lib import load_oracle, oracle_for
    t0 = time.time()
    lib = load_oracle()
    eq = lambda a, b: bool(np.array_equal(a, b))
    if world == 1:
        o = oracle_for(lib, mesh, params)
        c = capi.Context(mesh, params, device=local_rank)
        c.set_stream(stream.cuda_stream)
        c.lsq_coefficients()
        c.set_field(capi.F_Q, q0)
        qo = q0.copy()
        beta = np.zeros(1)
        _, sw = o.lsq()
        dt, _ = o.timestep(qo, beta)
        o.update_bcs(qo, beta)
        grad = o.gradient(qo, sw)
        lim = o.limiter(qo, grad)
        b = o.residual(qo, grad, lim, beta)
        o.explicit_solve(qo, b, dt)
        c.explicit_iterate(refresh_dt=True)
        res = {"timestep": eq(c.get_field(capi.F_TIMESTEP), dt), "qgrad": eq(c.get_field(capi.F_QGRAD), grad),
               "limiter": eq(c.get_field(capi.F_LIMITER), lim), "b": eq(c.get_field(capi.F_B), b), "q": eq(c.get_field(capi.F_Q), qo)}
        c.close()
        return {"what": f"one explicit iteration on the bench mesh (n={args.n}) vs the C oracle (oracle/pcfd_oracle.c), bit-exact per field",
                "bit_exact": res, "ok": all(res.values()), "seconds": time.time() - t0}
    from proteuscfd_b200.cases import slab_case
    from tests.partition_oracle import replay_perfect_gas
    npar = 32
    out = {"what": f"2 explicit + 2 implicit (3 sweeps, Jacobian refreshed) composite iterations on {world} slabs of n={npar} through the "
                   f"'{args.halo}' exchange vs the per-rank oracle replay with a numpy halo, bit-exact per field (ghost rows included)"}
    ok = True
    for implicit in (False, True):
        parts = [slab_case(npar, r, world, colored=implicit, cfl=5.0 if implicit else 0.5) for r in range(world)]
        _, ref = replay_perfect_gas(lib, parts, implicit, iters=2, nsweeps=3)
        pr = RankRun(args, parts[rank][0], parts[rank][1], rank, world, local_rank, torch, dist, stream, implicit=implicit)
        pr.ctx.set_field(capi.F_Q, parts[rank][2])
        res = {}
        for it in range(2):
            if implicit:
                pr.implicit(3, True)
            else:
                pr.explicit()
            pr.sync()
            for k, f in (("qgrad", capi.F_QGRAD), ("limiter", capi.F_LIMITER), ("b", capi.F_B), ("q", capi.F_Q)) + (
                    (("x", capi.F_X),) if implicit else ()):
                res[f"{k}{it}"] = eq(pr.ctx.get_field(f), ref[it][k][rank])
        pr.close()
        bad = pr.total(sum(0 if v else 1 for v in res.values()))
        out["implicit" if implicit else "explicit"] = {"rank0_bit_exact": res, "mismatching_fields_all_ranks": bad}
        ok = ok and bad == 0
    # multi-neighbour halos: recursive-coordinate-bisection partitions of one box (udecomp layout, partition.py), every rank
    # with several peers and corner ghosts -- the same composite iterations through the same exchange
    from proteuscfd_b200.cases import partitioned_box_case
    nrcb = 20
    peers = 0
    for implicit in (False, True):
        parts = partitioned_box_case(nrcb, world, cfl=5.0 if implicit else 0.5)
        _, ref = replay_perfect_gas(lib, parts, implicit, iters=2, nsweeps=3)
        pr = RankRun(args, parts[rank][0], parts[rank][1], rank, world, local_rank, torch, dist, stream, implicit=implicit)
        pr.ctx.set_field(capi.F_Q, parts[rank][2])
        peers = int(np.unique(parts[rank][0]["gNodeOwner"]).size)
        res = {}
        for it in range(2):
            if implicit:
                pr.implicit(3, True)
            else:
                pr.explicit()
            pr.sync()
            for k, f in (("qgrad", capi.F_QGRAD), ("limiter", capi.F_LIMITER), ("b", capi.F_B), ("q", capi.F_Q)) + (
                    (("x", capi.F_X),) if implicit else ()):
                res[f"{k}{it}"] = eq(pr.ctx.get_field(f), ref[it][k][rank])
        maxp = pr.maxi(peers)
        pr.close()
        bad = pr.total(sum(0 if v else 1 for v in res.values()))
        out["rcb_implicit" if implicit else "rcb_explicit"] = {"rank0_bit_exact": res, "mismatching_fields_all_ranks": bad}
        ok = ok and bad == 0
    out["rcb"] = {"what": f"the same on {world} recursive-coordinate-bisection partitions of an n={nrcb} box (multi-neighbour halos)",
                  "peers_of_rank0": peers, "max_peers": maxp}
    out["ok"] = ok
    out["seconds"] = time.time() - t0
    return out


def implicit_bench(args, rank, world, local_rank, peak, torch, dist, stream, kind="euler"):
    """Implicit 5x5 iteration and SGS sweeps/s, colour-sorted numbering, on one GPU or on z-slab partitions:
      euler       the implicit version of the headline case (second half of the metric: SGS sweeps/s);
      laminar_ns  BASELINE configs[2]: compressibleNS, no-slip wall, ~20 M cells in total on 1-2 GPUs (strong split);
      rans_sa     BASELINE configs[3]: compressibleNS + Spalart-Allmaras (TurbulenceModel::Compute after the flow update),
                  wall distance from walldist.py, one n^3 slab per GPU.
    Reported: the iteration with the Jacobian kept, the Jacobian refresh, and the iteration WITH refresh -- the
    reference's default (jacobianUpdateFrequency = 1, ucs/param.tcc:173)."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import NS_BC, box_case, slab_case
    nsgs = 5
    viscous = kind != "euler"
    turb = kind == "rans_sa"
    n = (args.sgs_n or args.n) if kind != "laminar_ns" else args.ns_n
    nz = None
    bc = None
    if viscous:
        bc = dict(NS_BC)
        bc[5] = capi.BC_SYMMETRY
        bc[3] = capi.BC_NOSLIP           # the wall runs along y = 0: every z-slab owns a strip of it
    if kind == "laminar_ns" and world > 1:
        nz = n                           # the SAME 20 M-cell box split across the ranks
    kw = dict(cfl=5.0, colored=True, device=f"cuda:{local_rank}")
    if viscous:
        kw.update(viscous=True, turb=turb, bc=bc)
    if world > 1:
        mesh, params, q = slab_case(n, rank, world, nz=nz, **kw)
    else:
        mesh, params, q = box_case(n, **kw)
    r = RankRun(args, mesh, params, rank, world, local_rank, torch, dist, stream, implicit=True)
    c = r.ctx
    if turb:
        # ComputeWallDistOct on the device (pcfd_wall_distance): every rank's viscous wall nodes, gathered like
        # SyncParallelPoint does, against every local node
        from proteuscfd_b200.parallel import TorchGroup
        from proteuscfd_b200.walldist import wall_points
        pts = wall_points(mesh)
        if world > 1:
            pts = np.concatenate([np.asarray(p_, dtype=np.float64).reshape(-1, 3) for p_ in TorchGroup(dist).allgather(pts)])
        pts = np.unique(pts, axis=0)      # one point per wall node instead of one per half-edge: the minimum does not care
        t_wd = time.time()
        c.wall_distance(pts)
        c.synchronize()
        wall_distance_info = {"ms": (time.time() - t_wd) * 1e3, "wall_points": int(pts.shape[0]),
                              "what": "pcfd_wall_distance: exact search, every local node against every viscous wall node of every rank (set-up, once per mesh; wall clock incl. the upload of the points)"}
        tv = np.full(c.field_size(capi.F_TVAR), 1.341946)      # nu~ of the free stream (spalart.tcc:79)
        c.set_field(capi.F_TVAR, tv)
    c.set_field(capi.F_Q, q)
    r.implicit(2, True)                     # builds A and its LU; warm-up
    r.blank_x()
    for _ in range(2):
        r.sweep()
    ms_sweep = r.timed(r.sweep, 10, stream)
    c.set_field(capi.F_Q, q)
    r.implicit(nsgs, False)
    ms_it = r.timed(lambda: r.implicit(nsgs, False), 5, stream)
    ms_refresh = r.timed(r.refresh_only, 2, stream)
    c.set_field(capi.F_Q, q)
    ms_it_refresh = r.timed(lambda: r.implicit(nsgs, True), 3, stream)
    finite = bool(np.isfinite(c.get_field(capi.F_Q)).all())
    # per-kernel table of one iteration with refresh (profiling on: its event records break the PDL chain of the SGS
    # levels, so it is a separate pass)
    c.profile(on=True, reset=True)
    r.implicit(nsgs, True)
    r.sync()
    c.profile(on=False)
    tab = c.profile_table()
    nblocks = int(c.get_crs()[1].size)
    pbs = pass_bytes(c.nedge + c.ngedge, c.nnode + c.gnode, nblocks)
    bytes_sweep = r.total(pass_bytes(c.nedge, c.nnode, nblocks)["sgs_sweep"])
    head = pbs["gradient"] + pbs["limiter"] + pbs["residual"]
    ne_l, nn_l = c.nedge + c.ngedge, c.nnode + c.gnode
    if viscous:      # the viscous pass re-reads edges, q and all gradient terms, read-modify-writes b
        head += 40 * ne_l + nn_l * (8 * NTERMS + 24 * NTERMS + 24 + 8) + c.nnode * 16 * NEQN
    turb_bytes = 0
    if turb:         # scalar system on the flow's pattern: gradient of nu~, one assembly pass, nsgs scalar sweeps (8d: 1x1 blocks)
        turb_bytes = (8 * ne_l + nn_l * (24 + 8 + 48 + 24)) + (40 * ne_l + nn_l * (8 * NVARS + 24 + 8 + 8) + nblocks * 8) \
            + nsgs * 2 * (nblocks * 8 + 4 * nblocks + 4 * (c.nnode + 1) + 8 * 3 * c.nnode)
    bytes_it = r.total(head + turb_bytes) + nsgs * bytes_sweep
    bytes_jac = r.total(40 * ne_l + nn_l * 8 * NVARS + nblocks * 8 * NEQN * NEQN)
    nn_g, nb_g = r.total(c.nnode), r.total(nblocks)
    ne_g = r.total(2 * c.nedge + c.ngedge) // 2 if world > 1 else c.nedge
    names = {"euler": "implicit version of BASELINE configs[1] (inviscid Euler)",
             "laminar_ns": "BASELINE configs[2]: laminar Navier-Stokes (compressibleNS, isothermal no-slip wall, Re 400), implicit",
             "rans_sa": "BASELINE configs[3]: RANS Spalart-Allmaras (compressibleNS + SA, no-slip wall, wall distance by exact search), implicit"}
    out = {"workload": f"{names[kind]}; Kuhn box n={n}, colour-sorted numbering, {world} GPU(s)"
                       + (f", {world} z-slabs of the SAME box (strong split)" if nz else (f", one n^3 slab per GPU" if world > 1 else ""))
                       + f": {nn_g} nodes, {ne_g} edges, {nb_g} 5x5 blocks = {nb_g * 200 / 1e9:.1f} GB",
           "n_gpus": world, "nsgs": nsgs, "nodes": nn_g, "blocks": nb_g, "block": "5x5",
           "sweeps_per_s": 1e3 / ms_sweep, "ms_per_sweep": ms_sweep, "algorithmic_bytes_per_sweep": bytes_sweep,
           "GBps": bytes_sweep / (ms_sweep * 1e-3) / 1e9, "frac_hbm": bytes_sweep / (ms_sweep * 1e-3) / 1e9 / (peak * world),
           "implicit_iteration": {"what": f"UpdateBCs + gradient + limiter + residual + {nsgs} SGS sweeps + ApplyDQ"
                                          + (" + TurbulenceModel::Compute" if turb else "") + ", Jacobian kept"
                                          + (", the reference's halo exchanges in the reference's places" if world > 1 else ""),
                                  "ms": ms_it, "Medges_s": ne_g / (ms_it * 1e-3) / 1e6, "algorithmic_bytes": bytes_it,
                                  "GBps": bytes_it / (ms_it * 1e-3) / 1e9, "frac_hbm": bytes_it / (ms_it * 1e-3) / 1e9 / (peak * world)},
           "jacobian_refresh_ms": ms_refresh,
           "iteration_with_refresh": {"what": "the same + ComputeTimesteps + ComputeJacobians + LU every iteration (the reference's default, "
                                              "jacobianUpdateFrequency = 1, ucs/param.tcc:173); the finite-difference assembly is FP64-bound",
                                      "ms": ms_it_refresh, "Medges_s": ne_g / (ms_it_refresh * 1e-3) / 1e6,
                                      "algorithmic_bytes": bytes_it + bytes_jac,
                                      "frac_hbm": (bytes_it + bytes_jac) / (ms_it_refresh * 1e-3) / 1e9 / (peak * world)},
           "state_finite_after_run": finite, "clip_fallbacks": r.clip_fallbacks(),
           "kernels_ms_rank0": {k: v[0] / max(v[1], 1) for k, v in tab.items()},
           "kernel": "k_sgs_tile (cp.async.bulk + mbarrier streamed tiles, PDL-chained levels)" if "k_sgs_tile" in tab else "k_sgs_level"}
    if turb:
        out["wall_distance"] = wall_distance_info
    if kind == "euler":
        # CRS::GMRES (10 directions, block-diagonal LU right preconditioner) on the freshly assembled system
        try:
            c.timestep(want_min=False)
            c.jacobian()
            c.blank_x()
            c.gmres(1, 10, 2)                       # allocates the Krylov scratch
            c.blank_x()
            r.sync()
            t0 = time.perf_counter()
            dq = c.gmres(1, 10, 2)
            r.sync()
            ms_g = r.maxms((time.perf_counter() - t0) * 1e3)
            bytes_g = r.total(11 * (nblocks * 8 * NEQN * NEQN + 4 * nblocks))      # 11 block-CRS products stream the matrix
            out["gmres"] = {"what": "pcfd_gmres(1 restart, 10 directions, block-diagonal LU preconditioner): 11 block-CRS products, 10 "
                                    "preconditioner solves, 65 dot products (each a host round trip), wall clock incl. those",
                            "ms": ms_g, "dq_norm": dq, "matrix_stream_GBps": bytes_g / (ms_g * 1e-3) / 1e9,
                            "matrix_stream_frac_hbm": bytes_g / (ms_g * 1e-3) / 1e9 / (peak * world)}
        except Exception as e:
            out["gmres"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    r.close()
    torch.cuda.empty_cache()
    return out


def strong_bench(args, rank, world, local_rank, peak, torch, dist, stream):
    """STRONG scaling: one fixed box of n = 203 (50.2 M tets, 8.5 M nodes, 59 M edges -- the size the north-star target is
    quoted on) dealt out in z-slabs to the N GPUs; explicit and implicit (5x5, Jacobian kept / refreshed) Euler
    iterations.  The main line's `value` stays the weak-scaling figure."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import slab_case
    n, nsgs = args.strong_n, 5
    mesh, params, q = slab_case(n, rank, world, nz=n, cfl=5.0, colored=True, device=f"cuda:{local_rank}")
    r = RankRun(args, mesh, params, rank, world, local_rank, torch, dist, stream, implicit=True)
    c = r.ctx
    c.set_field(capi.F_Q, q)
    c.set_cfl(0.5)
    for _ in range(3):
        r.explicit()
    ms_exp = r.timed(r.explicit, 10, stream)
    c.set_cfl(5.0)
    c.set_field(capi.F_Q, q)
    r.implicit(nsgs, True)
    ms_it = r.timed(lambda: r.implicit(nsgs, False), 3, stream)
    ms_itr = r.timed(lambda: r.implicit(nsgs, True), 2, stream)
    finite = bool(np.isfinite(c.get_field(capi.F_Q)).all())
    nblocks = int(c.get_crs()[1].size)
    pbs = pass_bytes(c.nedge + c.ngedge, c.nnode + c.gnode, nblocks)
    bytes_it = r.total(pbs["gradient"] + pbs["limiter"] + pbs["residual"]) + nsgs * r.total(pass_bytes(c.nedge, c.nnode, nblocks)["sgs_sweep"])
    nn_g = r.total(c.nnode)
    ne_g = r.total(2 * c.nedge + c.ngedge) // 2
    nodes = [0] * world
    nodes[rank] = c.nnode
    t = torch.tensor(nodes, device=r.dev, dtype=torch.int64)
    dist.all_reduce(t)
    out = {"workload": f"fixed Kuhn box n={n} ({6 * n ** 3} tets, {nn_g} nodes, {ne_g} edges) split into {world} z-slabs "
                       f"(nodes per rank {t.tolist()}), colour-sorted numbering, halo '{args.halo}'",
           "scaling": "strong", "n_gpus": world,
           "explicit_ms": ms_exp, "explicit_Medges_s": ne_g / (ms_exp * 1e-3) / 1e6,
           "implicit_iteration_ms": ms_it, "implicit_Medges_s": ne_g / (ms_it * 1e-3) / 1e6,
           "implicit_frac_hbm": bytes_it / (ms_it * 1e-3) / 1e9 / (peak * world),
           "implicit_with_refresh_ms": ms_itr, "state_finite_after_run": finite}
    r.close()
    torch.cuda.empty_cache()
    return out


def fma_variant(args, mesh, params, q0, torch, stream, local_rank, ms_exact):
    """What bit-exactness costs: the same sources built with FMA contraction allowed (libpcfd_b200_fma.so, NOT the
    product) run the same explicit iterations; reported next to the exact build's time with the drift of the state from
    the exact (= reference) arithmetic after 10 iterations, relative per node as the north star states its bar."""
    from proteuscfd_b200 import capi
    if not os.path.exists(capi.FMA_LIB_PATH):
        return {"unavailable": "libpcfd_b200_fma.so not built"}
    res = {}
    for name, path in (("exact", None), ("fma", capi.FMA_LIB_PATH)):
        c = capi.Context(mesh, params, device=local_rank, lib_path=path)
        c.set_stream(stream.cuda_stream)
        c.lsq_coefficients()
        c.set_field(capi.F_Q, q0)
        for _ in range(10):
            c.explicit_iterate(refresh_dt=True)
        q10 = c.get_field(capi.F_Q)
        ev = lambda: torch.cuda.Event(enable_timing=True)
        e0, e1 = ev(), ev()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(20):
            c.explicit_iterate(refresh_dt=True)
        e1.record(stream)
        torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / 20, q10)
        c.close()
    nn = mesh["nnode"]
    a, b = res["fma"][1].reshape(-1, NVARS)[:nn, :NEQN], res["exact"][1].reshape(-1, NVARS)[:nn, :NEQN]
    scale = np.abs(b).max(axis=1, keepdims=True)
    drift = float((np.abs(a - b) / scale).max())
    return {"what": "explicit iteration with FMA contraction allowed (--fmad=true build of the same sources) vs the bit-exact build",
            "ms_per_step_exact": res["exact"][0], "ms_per_step_fma": res["fma"][0],
            "speedup_if_bit_exactness_were_dropped": res["exact"][0] / res["fma"][0],
            "max_relative_drift_per_node_after_10_iterations": drift,
            "within_north_star_1e-12": drift <= 1e-12,
            "note": "the finite-difference Jacobian divides flux differences by h = 1e-8, so a contracted flux is NOT within 1e-12 "
                    "for the implicit path whatever this explicit figure says (DESIGN.md section 3)"}


def fr_bench_multi(args, rank, world, local_rank, peak, torch, dist, stream):
    """BASELINE configs[4] as named: reacting 5-species air (compressibleEulerFR), implicit, on N partitions (one z-slab of
    n^3 hexes per GPU, colour-sorted numbering, 9x9 block-CRS rows with ghost columns).  One iteration = UpdateBCs,
    gradient, limiter, HLLC residual + finite-rate source, nsgs SGS sweeps, ApplyDQ with the reference's halo
    exchanges (q 21 wide, qgrad 42, limiter 9, x 9 after every sweep).  Max over ranks."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_slab_case
    n = args.fr_n or args.n
    nsgs, reps = 5, 3
    torch.cuda.empty_cache()
    mesh, params, q, beta = fr_slab_case(n, rank, world, fr_params_from_fixture(), device=f"cuda:{local_rank}")
    r = RankRun(args, mesh, params, rank, world, local_rank, torch, dist, stream, beta=beta, implicit=True)
    c = r.ctx
    c.set_field(capi.F_Q, q)
    r.implicit(nsgs, True)     # builds A and its LU; warm-up
    ms_it = r.timed(lambda: r.implicit(nsgs, False), reps, stream)
    finite = bool(np.isfinite(c.get_field(capi.F_Q)).all())
    r.blank_x()
    ms_sweep = r.timed(r.sweep, 6, stream)
    ms_jac = r.timed(r.refresh_only, 1, stream)
    c.set_field(capi.F_Q, q)
    ms_itr = r.timed(lambda: r.implicit(nsgs, True), 2, stream)
    nblocks = int(c.get_crs()[1].size)
    neqn, nterms = c.neqn, c.nterms
    ne, nn, nl = c.nedge + c.ngedge, c.nnode, c.nnode + c.gnode
    bytes_sweep = r.total(2 * (nblocks * 8 * neqn * neqn + 4 * nblocks + 4 * (nn + 1) + 4 * neqn * nn + 8 * neqn * 3 * nn))
    bytes_head = (8 * ne + nl * (24 + 8 * nterms + 48 + 24 * nterms)
                  + (8 * ne + nl * (8 * neqn + 16 * neqn)) + (8 * ne + nl * (8 * neqn + 24 * neqn + 24 + 16 * neqn) + nl * 8 * neqn)
                  + (8 * ne + nl * (8 * neqn + 24 * neqn + 24 + 8 * neqn) + nl * 8 * neqn)
                  + 40 * ne + nl * (8 * neqn + 24 * neqn + 8 * neqn + 24 + 8) + nn * 8 * neqn)
    bytes_iter = r.total(bytes_head) + nsgs * bytes_sweep
    bytes_jac = r.total(40 * ne + nl * 8 * c.nvars + nblocks * 8 * neqn * neqn)
    nn_g, nb_g = r.total(nn), r.total(nblocks)
    ne_g = r.total(2 * c.nedge + c.ngedge) // 2
    out = {"workload": f"BASELINE configs[4]: reacting 5-species air (compressibleEulerFR), implicit, {world} z-slab partitions of an "
                       f"n x n x {n * world} Kuhn box (n = {n}: {6 * n ** 3 * world} tets, {nn_g} nodes, {ne_g} edges, {nb_g} 9x9 blocks = "
                       f"{nb_g * 648 / 1e9:.1f} GB) on {world} GPUs; halo exchange '{args.halo}'",
           "n_gpus": world, "nsgs": nsgs,
           "iteration_without_refresh_ms": ms_it, "iteration_Medges_s": ne_g / (ms_it * 1e-3) / 1e6,
           "iteration_algorithmic_bytes": bytes_iter, "iteration_GBps": bytes_iter / (ms_it * 1e-3) / 1e9,
           "iteration_frac_hbm": bytes_iter / (ms_it * 1e-3) / 1e9 / (peak * world),
           "sgs_ms_per_sweep": ms_sweep, "sgs_sweeps_per_s": 1e3 / ms_sweep, "sgs_algorithmic_bytes_per_sweep": bytes_sweep,
           "sgs_GBps": bytes_sweep / (ms_sweep * 1e-3) / 1e9, "sgs_frac_hbm": bytes_sweep / (ms_sweep * 1e-3) / 1e9 / (peak * world),
           "jacobian_refresh_ms": ms_jac,
           "iteration_with_refresh_ms": ms_itr,
           "iteration_with_refresh_frac_hbm": (bytes_iter + bytes_jac) / (ms_itr * 1e-3) / 1e9 / (peak * world),
           "state_finite_after_run": finite, "clip_fallbacks": r.clip_fallbacks()}
    r.close()
    torch.cuda.empty_cache()
    return out


def fr_params_from_fixture(viscous=False):
    """Chemistry tables (the reference's chemModels/5speciesAir.rxn + NASA-7 data as its ChemModel parsed them),
    reference values and free stream of the reacting fixture the reference wrote (tools/make_golden.py); viscous: the
    compressibleNSFR fixture, which adds the species transport tables (chemdata/trans.inp), Re and PrT."""
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "box4_nsfr_implicit.npz" if viscous else "box4_fr_implicit.npz")))
    meta = dict(zip([str(k) for k in d["meta_keys"]], d["meta_vals"]))
    chem = {k: d[k] for k in ("species_mw", "species_nasa7", "rxn_A_EA_n", "rxn_flags", "rxn_species", "rxn_nup", "rxn_nupp",
                              "rxn_tbeff")}
    chem["dims"] = d["chem_dims"]
    extra = {}
    if viscous:
        extra = dict(transport={k: d[k] for k in ("species_mu_fit", "species_k_fit", "species_white", "species_fit_counts")},
                     ref_viscosity=meta["ref_viscosity"], ref_k=meta["ref_k"], Re=meta["Re"], PrT=meta["PrT"])
    return dict(chem=chem, **extra, ref_density=meta["ref_density"], ref_velocity=meta["ref_velocity"],
                ref_temperature=meta["ref_temperature"], ref_pressure=meta["ref_pressure"], ref_time=meta["ref_time"],
                ref_specific_enthalpy=meta["ref_specific_enthalpy"], pref=meta["Pref"], dt=meta["dt"],
                use_local_dt=int(meta["useLocalTimeStepping"]), rxn_on=1, qinf=d["qinf"])


def fr_bench(args, peak, torch, stream, local_rank, viscous=False):
    """Reacting 5-species air (compressibleEulerFR), implicit: 9 equations per node, 9x9 block-CRS Jacobian (FD flux +
    FD source Jacobians, dense temporal terms), SGS.  Timed: the Jacobian refresh, one implicit iteration without it
    (UpdateBCs, gradient, limiter, HLLC residual + finite-rate source, nsgs SGS sweeps, ApplyDQ) and the SGS sweep."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_box_case
    n = args.fr_n or args.n
    nsgs = 5
    torch.cuda.empty_cache()
    mesh, params, q, beta = fr_box_case(n, fr_params_from_fixture(viscous), device=f"cuda:{local_rank}")
    c = capi.Context(mesh, params, device=local_rank)
    c.set_stream(stream.cuda_stream)
    c.set_field(capi.F_BETA, beta)
    c.lsq_coefficients()
    c.set_field(capi.F_Q, q)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    c.implicit_iterate(nsgs, refresh_jac=True)    # warm-up: builds A, LU
    torch.cuda.synchronize()
    reps = 3
    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
    e0.record(stream)
    for _ in range(reps):
        c.timestep(want_min=False)
        c.jacobian()
        c.prepare_sgs()
    e1.record(stream)
    eA, eB = ev(), ev()
    eA.record(stream)
    for _ in range(reps):
        c.implicit_iterate(nsgs, refresh_jac=False)
    eB.record(stream)
    finite = bool(np.isfinite(c.get_field(capi.F_Q)).all())
    c.blank_x()
    nsw = 10
    c.sgs(2, want_ddq=False)
    e2.record(stream)
    c.sgs(nsw, want_ddq=False)
    e3.record(stream)
    torch.cuda.synchronize()
    # per-kernel tables from a separate profiled pass (event records between launches would serialise the programmatic
    # dependent launch of consecutive SGS levels, so the timed regions above run without them)
    c.profile(on=True, reset=True)
    c.timestep(want_min=False)
    c.jacobian()
    c.prepare_sgs()
    c.implicit_iterate(nsgs, refresh_jac=False)
    torch.cuda.synchronize()
    c.profile(on=False)
    tab = c.profile_table()
    ms_jac = e0.elapsed_time(e1) / reps
    ms_iter = eA.elapsed_time(eB) / reps
    ms_sweep = e2.elapsed_time(e3) / nsw
    # the reference's default: Jacobian refreshed every iteration (jacobianUpdateFrequency = 1, ucs/param.tcc:173)
    eR0, eR1 = ev(), ev()
    c.set_field(capi.F_Q, q)
    eR0.record(stream)
    for _ in range(2):
        c.implicit_iterate(nsgs, refresh_jac=True)
    eR1.record(stream)
    torch.cuda.synchronize()
    ms_iter_refresh = eR0.elapsed_time(eR1) / 2
    kern = {k: v[0] / max(v[1], 1) for k, v in tab.items()}
    it_kernels = ("kfr_update_bcs_edges", "kfr_gradient", "kfr_limiter", "kfr_fill_int", "kfr_clip_edges", "kfr_clip_nodes",
                  "kfr_limiter_final", "kfr_flux_edges", "kfr_flux_bedges", "kfr_vflux_edges", "kfr_source", "kfr_residual_gather",
                  "kfr_apply_dq")
    ms_explicit_part = sum(kern.get(k, 0.0) for k in it_kernels)
    nblocks = c.get_crs()[1].size
    neqn, nterms = c.neqn, c.nterms
    ne, nn = c.nedge, c.nnode
    bytes_sweep = 2 * (nblocks * 8 * neqn * neqn + 4 * nblocks + 4 * (nn + 1) + 4 * neqn * nn + 8 * neqn * 3 * nn)
    bytes_resid = 40 * ne + nn * (8 * neqn + 24 * neqn + 8 * neqn + 24 + 8) + nn * 8 * neqn
    if viscous:     # the viscous pass re-reads edges, q and all gradient terms, and read-modify-writes b
        bytes_resid += 40 * ne + nn * (8 * nterms + 24 * nterms + 24 + 8) + nn * 16 * neqn
    ms_resid = sum(kern.get(k, 0.0) for k in ("kfr_flux_edges", "kfr_flux_bedges", "kfr_vflux_edges", "kfr_source",
                                              "kfr_residual_gather"))
    bytes_grad = 8 * ne + nn * (24 + 8 * nterms + 48 + 24 * nterms)
    bytes_lim = ((8 * ne + nn * (8 * neqn + 16 * neqn)) + (8 * ne + nn * (8 * neqn + 24 * neqn + 24 + 16 * neqn) + nn * 8 * neqn)
                 + (8 * ne + nn * (8 * neqn + 24 * neqn + 24 + 8 * neqn) + nn * 8 * neqn))
    bytes_iter = bytes_grad + bytes_lim + bytes_resid + nsgs * bytes_sweep
    out = {"workload": f"BASELINE configs[4] on one GPU: reacting 5-species air ({'compressibleNSFR' if viscous else 'compressibleEulerFR'}), Kuhn box n={n} "
                       f"({nn} nodes, {ne} edges, {nblocks} 9x9 blocks = {nblocks * 648 / 1e9:.1f} GB), HLLC 2nd order + LSQ + "
                       "Venkatakrishnan + finite-rate source" + (" + Wilke-mixed viscous flux / analytic viscous Jacobian" if viscous else "")
                       + ", implicit",
           "jacobian_refresh_ms": ms_jac, "sgs_ms_per_sweep": ms_sweep, "sgs_sweeps_per_s": 1e3 / ms_sweep,
           "sgs_algorithmic_bytes_per_sweep": bytes_sweep, "sgs_GBps": bytes_sweep / (ms_sweep * 1e-3) / 1e9,
           "sgs_frac_hbm": bytes_sweep / (ms_sweep * 1e-3) / 1e9 / peak,
           "residual_ms": ms_resid, "residual_Medges_s": ne / (ms_resid * 1e-3) / 1e6,
           "residual_frac_hbm": bytes_resid / (ms_resid * 1e-3) / 1e9 / peak,
           "iteration_without_refresh_ms": ms_iter, "iteration_Medges_s": ne / (ms_iter * 1e-3) / 1e6, "nsgs": nsgs,
           "iteration_algorithmic_bytes": bytes_iter, "iteration_frac_hbm": bytes_iter / (ms_iter * 1e-3) / 1e9 / peak,
           "iteration_with_refresh_ms": ms_iter_refresh,
           "iteration_with_refresh_frac_hbm": (bytes_iter + 40 * ne + nn * 8 * c.nvars + nblocks * 8 * neqn * neqn)
           / (ms_iter_refresh * 1e-3) / 1e9 / peak,
           "state_finite_after_run": finite, "clip_fallbacks": c.clip_fallbacks(), "non_sgs_kernels_ms": ms_explicit_part,
           "kernels_ms": kern}
    c.close()
    return out


if __name__ == "__main__":
    main()
