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
    "timestep": ["k_timestep"],
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=118, help="hexes per box side (118 -> ~10 M tets)")
    ap.add_argument("--ref-n", type=int, default=48, help="box side of the bounded CPU reference sample")
    ap.add_argument("--halo", default="put", choices=["put", "nccl"],
                    help="N>1 halo exchange: direct puts into peer ghost segments over NVLink (CUDA IPC) or NCCL send/recv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sgs", action="store_true")
    ap.add_argument("--no-fr", action="store_true", help="skip the reacting-eqnset (compressibleEulerFR) sub-benchmark")
    ap.add_argument("--fr-n", type=int, default=0, help="box side for the reacting sub-benchmark (0: same as --n)")
    ap.add_argument("--sgs-n", type=int, default=0, help="box side for the SGS sub-benchmark (0: same as --n)")
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
        from proteuscfd_b200.parallel import DistributedHotPath, NcclExchange, PObj, PutExchange, TorchGroup
        mesh, params, q0 = slab_case(args.n, rank, world, device=f"cuda:{local_rank}")
    else:
        mesh, params, q0 = box_case(args.n, device=f"cuda:{local_rank}")
    ctx = capi.Context(mesh, params, device=local_rank)
    # a real (non-default) torch stream: the library launches on it and torch.cuda.Event times it
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    ne, nn = ctx.nedge, ctx.nnode
    ne_global = ne
    if world > 1:
        group = TorchGroup(dist)
        pobj = PObj(rank, world).BuildCommMaps(mesh["gNodeOwner"], mesh["gNodeLocalId"], group)
        dev = torch.device("cuda", local_rank)
        xch = PutExchange(ctx, pobj, dist, torch, dev, group) if args.halo == "put" else NcclExchange(ctx, pobj, dist, torch, dev)
        hitflag = torch.zeros(1, device=dev, dtype=torch.int32)

        def any_rank(hit):   # the (rare) pressure-clip fallback is a global decision: one 4-byte max all-reduce
            hitflag.fill_(int(hit))
            dist.all_reduce(hitflag, op=dist.ReduceOp.MAX)
            return bool(hitflag.item())

        hp = DistributedHotPath(ctx, xch, any_rank=any_rank)
        hp.setup()
        ctx.set_field(capi.F_Q, q0)
        step = lambda: hp.explicit_iterate(refresh_dt=True)
        t = torch.tensor([2 * ctx.nedge + ctx.ngedge], device="cuda", dtype=torch.int64)   # a cut edge lives on two ranks
        dist.all_reduce(t)
        ne_global = int(t.item()) // 2
        halo_rows = int(pobj.commCountsSend.sum())
    else:
        ctx.lsq_coefficients()
        ctx.set_field(capi.F_Q, q0)
        step = lambda: ctx.explicit_iterate(refresh_dt=True)
        halo_rows = 0
    t_setup = time.time() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing
    for _ in range(W):
        step()
    ctx.profile(on=True, reset=True)
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
    ctx.profile(on=False)
    table = ctx.profile_table()
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / K
    value = ne_global / (ms_step * 1e-3) / 1e6

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
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(Ke):
        ctx.set_field(capi.F_Q, hq_np)          # H2D of the step's input state (pinned)
        step()
        ctx.get_field(capi.F_Q, out=hq_np)      # D2H of the updated state
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1) / Ke
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = ne_global / (e2e_ms * 1e-3) / 1e6

    # ---------------------------------------------------------------- per-pass roofline from the live kernel timings
    peak, peak_src = measured_peaks()
    pb = pass_bytes(ne, nn)
    kern = {k: {"ms_avg": v[0] / max(v[1], 1), "launches_per_step": v[1] / K, "ms_per_step": v[0] / K} for k, v in table.items()}
    passes = {}
    for pname, ks in PASS_KERNELS.items():
        ms = sum(kern[k]["ms_per_step"] for k in ks if k in kern)
        if ms <= 0:
            continue
        d = {"ms_per_step": ms, "share": ms / ms_step}
        if pname in pb:
            d["algorithmic_bytes"] = pb[pname]
            d["GBps"] = pb[pname] / (ms * 1e-3) / 1e9
            d["frac_hbm"] = d["GBps"] / peak
        passes[pname] = d
    dominant = max((p for p in passes if "GBps" in passes[p]), key=lambda p: passes[p]["ms_per_step"])
    # measured DRAM traffic of the pass's kernels (one ncu --set full capture of this workload, committed under profiles/)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_dram_traffic.json")))["explicit_lexicographic"]
        if args.n == 118 and world == 1:
            traffic = sum(tr[k] for k in PASS_KERNELS[dominant] if k in kern and k in tr)
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": "+".join(k for k in PASS_KERNELS[dominant] if k in kern), "pass": dominant,
                "achieved": passes[dominant]["GBps"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": passes[dominant]["frac_hbm"], "traffic": traffic,
                "traffic_source": "profiles/r1_dram_traffic.json (ncu dram__bytes_read+write.sum per launch, same workload)" if traffic else None,
                "bytes_per_launch": passes[dominant]["algorithmic_bytes"], "ms_per_launch": passes[dominant]["ms_per_step"],
                "note": "algorithmic bytes = SURVEY.md 8d compulsory bytes of the pass; the Roe flux is FP64-pipe-bound "
                        "(see profiles/), so frac is reported for completeness, not as the limiter"}
    # the dominant kernel is FP64-pipe-bound, not HBM-bound: put the pipe roofline next to the HBM one.  1274 = FP64-pipe
    # instructions per edge of k_flux_edges<true> (static SASS count, `cuobjdump -sass`: 315 DADD + 451 DMUL + 455 DFMA
    # inside the IEEE division / sqrt sequences + 53 DSETP; the rarely taken raw-limiter branch included, so an upper
    # bound on the dynamic count); B200: 148 SMs x 64 FP64 lanes, one instruction per lane per clock
    try:
        sm_mhz = float((clocks or {}).get("sm_max_mhz") or 1965.0)
        bound_ms = 1274.0 * ne / (148 * 64 * sm_mhz * 1e6) * 1e3
        ms_flux = kern["k_flux_edges"]["ms_per_step"]
        roofline["fp64_pipe"] = {"kernel": "k_flux_edges", "instr_per_edge": 1274, "bound_ms": bound_ms, "ms": ms_flux,
                                 "frac": bound_ms / ms_flux,
                                 "how": "static SASS count of FP64-pipe instructions x edges / (148 SMs x 64 lanes x SM clock); 64 lanes per SM = "
                                        "the nominal 37 TFLOP/s FMA peak at 1.965 GHz, not measured here"}
    except Exception:
        pass
    step_bytes = sum(pb[p] for p in ("gradient", "limiter", "residual", "timestep"))
    phases = {"residual_Medges_s": ne / (passes["residual"]["ms_per_step"] * 1e-3) / 1e6 if "residual" in passes else None,
              "gradient_limiter_Medges_s": ne / ((passes["gradient"]["ms_per_step"] + passes["limiter"]["ms_per_step"]) * 1e-3) / 1e6,
              "iteration_frac_hbm": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
              "passes": passes, "kernels": kern}

    # ---------------------------------------------------------------- SGS sweeps/s (second half of the metric)
    sgs = None
    if not args.no_sgs and rank == 0 and world == 1:
        sgs = sgs_bench(args, ctx, mesh, params, q0, peak, torch, stream, local_rank)

    # ---------------------------------------------------------------- SGS sweeps/s across ranks (N > 1)
    if not args.no_sgs and world > 1:
        try:
            sgs = sgs_bench_multi(args, ctx, xch, rank, world, local_rank, peak, torch, dist, stream)
        except Exception as e:
            sgs = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---------------------------------------------------------------- reacting eqnset across ranks (BASELINE configs[4])
    frm = None
    if not args.no_fr and world > 1:
        try:
            frm = fr_bench_multi(args, rank, world, local_rank, peak, torch, dist, stream)
        except Exception as e:
            frm = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---------------------------------------------------------------- reacting eqnset (BASELINE configs[4] on one GPU)
    frb = frm
    if not args.no_fr and rank == 0 and world == 1:
        try:
            frb = fr_bench(args, peak, torch, stream, local_rank)
        except Exception as e:
            frb = {"error": f"{type(e).__name__}: {e}"[:300]}

    frv = None
    if not args.no_fr and rank == 0 and world == 1:
        try:
            frv = fr_bench(args, peak, torch, stream, local_rank, viscous=True)
        except Exception as e:
            frv = {"error": f"{type(e).__name__}: {e}"[:300]}

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            base = cpu_baseline(args)
        except Exception as e:   # never lose the GPU line to a baseline hiccup
            base = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {e}"[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: inviscid Euler, synthetic Kuhn box n={args.n} ({6 * args.n ** 3} tets, "
                                   f"{nn} nodes, {ne} edges) per GPU, Roe 2nd order + weighted LSQ + Venkatakrishnan, "
                                   "explicit single-stage update (the reference has no multistage RK)",
                       "parallelism": ("single partition" if world == 1 else
                                       f"{world} z-slab partitions of an n x n x {args.n * world} box, one per GPU; halo exchange "
                                       f"'{args.halo}' of q, qgrad, limiter, q per iteration ({halo_rows} rows/rank/exchange); "
                                       f"{ne_global} edges in total"),
                       "cache": "inputs larger than L2 (q 135 MB, qgrad 364 MB, edges 470 MB per pass); no flush needed",
                       "setup_s": t_setup},
            "roofline": roofline, "cpu_baseline": base,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": nq * 8 * world,
                    "d2h_bytes_per_step": nq * 8 * world, "what": "pcfd_set_field(q) from pinned host memory + pcfd_explicit_iterate + "
                                                          "pcfd_get_field(q) per step"},
            "gpu_launches": int(launches) * world, "clocks": clocks, "phases": phases, "sgs": sgs, "reacting": frb, "reacting_viscous": frv,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sgs_bench(args, ctx, mesh, params, q0, peak, torch, stream, local_rank):
    """SGS sweeps/s on the implicit version of the same case (5x5 blocks, colour-sorted numbering)."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    n = args.sgs_n or args.n
    ctx.close()
    torch.cuda.empty_cache()
    mesh, params, q = box_case(n, cfl=5.0, colored=True, device=f"cuda:{local_rank}")
    c = capi.Context(mesh, params, device=local_rank)
    c.set_stream(stream.cuda_stream)
    c.lsq_coefficients()
    c.set_field(capi.F_Q, q)
    c.profile(on=True, reset=True)
    c.implicit_iterate(2, refresh_jac=True)       # builds A, LU, 2 sweeps (warm-up)
    c.blank_x()
    c.sgs(2, want_ddq=False)
    torch.cuda.synchronize()
    tab0 = c.profile_table()
    # timed regions run with the per-kernel event profiling OFF (its event records sit between the launches and would
    # serialise the programmatic dependent launch of consecutive SGS levels); a profiled pass follows for the tables
    c.profile(on=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nsw = 10
    e0.record(stream)
    c.sgs(nsw, want_ddq=False)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nsw
    # the whole implicit iteration without a Jacobian refresh (UpdateBCs, gradient, limiter, residual, PrepareSGS (no-op:
    # LU kept), 5 SGS sweeps, ApplyDQ) against the HBM roofline of its compulsory bytes -- the north-star quantity
    nsgs_it, reps = 5, 5
    c.set_field(capi.F_Q, q)
    c.implicit_iterate(nsgs_it, refresh_jac=False)
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(reps):
        c.implicit_iterate(nsgs_it, refresh_jac=False)
    e3.record(stream)
    torch.cuda.synchronize()
    ms_it = e2.elapsed_time(e3) / reps
    c.profile(on=True)
    c.implicit_iterate(nsgs_it, refresh_jac=False)
    torch.cuda.synchronize()
    tab = c.profile_table()
    _, nblocks = c.get_crs()[1].size, c.get_crs()[1].size
    pbs = pass_bytes(c.nedge, c.nnode, nblocks)
    bytes_sweep = pbs["sgs_sweep"]
    bytes_it = pbs["gradient"] + pbs["limiter"] + pbs["residual"] + nsgs_it * bytes_sweep
    out = {"sweeps_per_s": 1e3 / ms, "ms_per_sweep": ms, "nodes": c.nnode, "blocks": int(nblocks), "block": "5x5",
           "levels_fwd_bwd": "colour-sorted numbering: one launch per colour per direction",
           "algorithmic_bytes_per_sweep": bytes_sweep, "GBps": bytes_sweep / (ms * 1e-3) / 1e9,
           "frac_hbm": bytes_sweep / (ms * 1e-3) / 1e9 / peak,
           "implicit_iteration": {"what": f"UpdateBCs + gradient + limiter + residual + {nsgs_it} SGS sweeps + ApplyDQ, Jacobian kept",
                                  "ms": ms_it, "Medges_s": c.nedge / (ms_it * 1e-3) / 1e6, "algorithmic_bytes": bytes_it,
                                  "GBps": bytes_it / (ms_it * 1e-3) / 1e9, "frac_hbm": bytes_it / (ms_it * 1e-3) / 1e9 / peak},
           "jacobian_ms": {k: tab0[k][0] / tab0[k][1] for k in ("k_jac_edges", "k_jac_bnodes", "k_jac_diag", "k_lu_diag") if k in tab0},
           "kernels_ms": {k: v[0] / max(v[1], 1) for k, v in tab.items()},
           "kernel": "k_sgs_tile (cp.async.bulk + mbarrier streamed tiles)" if "k_sgs_tile" in tab else "k_sgs_level",
           "launches_per_sweep": sum(tab[k][1] - tab0.get(k, (0, 0))[1] for k in ("k_sgs_level", "k_sgs_tile") if k in tab) / nsgs_it,
           "pdl": "levels of a sweep launched with programmatic stream serialization (PCFD_SGS_PDL=0 disables)"}
    c.close()
    return out


def sgs_bench_multi(args, ctx, xch, rank, world, local_rank, peak, torch, dist, stream):
    """SGS sweeps/s and the implicit iteration on N partitions (BASELINE configs[2]/[3] layout): every rank holds one
    z-slab with a colour-sorted numbering, its 5x5 block-CRS rows (ghost columns included) and runs the reference's
    multi-rank solve: block-Jacobi across partitions, one halo exchange of x after every sweep (ucs/crs.tcc:88,146).
    Times are the max over ranks; sweeps/s counts whole-domain sweeps, GB/s the bytes all ranks streamed."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import slab_case
    from proteuscfd_b200.parallel import DistributedHotPath, NcclExchange, PObj, PutExchange, TorchGroup
    n = args.sgs_n or args.n
    if hasattr(xch, "close"):
        xch.close()
    ctx.close()
    torch.cuda.empty_cache()
    dev = torch.device("cuda", local_rank)
    mesh, params, q = slab_case(n, rank, world, cfl=5.0, colored=True, device=f"cuda:{local_rank}")
    c = capi.Context(mesh, params, device=local_rank)
    c.set_stream(stream.cuda_stream)
    group = TorchGroup(dist)
    pobj = PObj(rank, world).BuildCommMaps(mesh["gNodeOwner"], mesh["gNodeLocalId"], group)
    x = PutExchange(c, pobj, dist, torch, dev, group) if args.halo == "put" else NcclExchange(c, pobj, dist, torch, dev)
    hitflag = torch.zeros(1, device=dev, dtype=torch.int32)

    def any_rank(hit):
        hitflag.fill_(int(hit))
        dist.all_reduce(hitflag, op=dist.ReduceOp.MAX)
        return bool(hitflag.item())

    hp = DistributedHotPath(c, x, any_rank=any_rank)
    hp.setup()
    c.set_field(capi.F_Q, q)
    hp.implicit_iterate(2, refresh_jac=True)      # builds A and its LU; warm-up

    def sync():
        dist.barrier()
        torch.cuda.synchronize()

    def maxms(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def total(v):
        t = torch.tensor([v], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    ev = lambda: torch.cuda.Event(enable_timing=True)
    c.blank_x()
    x.update(capi.F_X)
    for _ in range(2):
        c.sgs(1, want_ddq=False)
        x.update(capi.F_X)
    nsw = 10
    sync()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(nsw):
        c.sgs(1, want_ddq=False)
        x.update(capi.F_X)
    e1.record(stream)
    sync()
    ms = maxms(e0.elapsed_time(e1) / nsw)
    nsgs_it, reps = 5, 5
    c.set_field(capi.F_Q, q)
    hp.implicit_iterate(nsgs_it, refresh_jac=False)
    sync()
    e2, e3 = ev(), ev()
    e2.record(stream)
    for _ in range(reps):
        hp.implicit_iterate(nsgs_it, refresh_jac=False)
    e3.record(stream)
    sync()
    ms_it = maxms(e2.elapsed_time(e3) / reps)
    nblocks = int(c.get_crs()[1].size)
    pbs = pass_bytes(c.nedge + c.ngedge, c.nnode + c.gnode, nblocks)
    bytes_sweep = total(pass_bytes(c.nedge, c.nnode, nblocks)["sgs_sweep"])
    bytes_it = total(pbs["gradient"] + pbs["limiter"] + pbs["residual"]) + nsgs_it * bytes_sweep
    nn_g, nb_g = total(c.nnode), total(nblocks)
    ne_g = total(2 * c.nedge + c.ngedge) // 2
    out = {"sweeps_per_s": 1e3 / ms, "ms_per_sweep": ms, "nodes": nn_g, "blocks": nb_g, "block": "5x5", "n_gpus": world,
           "what": f"one sweep = forward + backward over every partition + halo exchange of x ('{args.halo}'); block-Jacobi "
                   "across partitions as in the reference",
           "algorithmic_bytes_per_sweep": bytes_sweep, "GBps": bytes_sweep / (ms * 1e-3) / 1e9,
           "frac_hbm": bytes_sweep / (ms * 1e-3) / 1e9 / (peak * world),
           "implicit_iteration": {"what": f"UpdateBCs + gradient + limiter + residual + {nsgs_it} SGS sweeps + ApplyDQ with the "
                                          "reference's halo exchanges, Jacobian kept",
                                  "ms": ms_it, "Medges_s": ne_g / (ms_it * 1e-3) / 1e6, "algorithmic_bytes": bytes_it,
                                  "GBps": bytes_it / (ms_it * 1e-3) / 1e9,
                                  "frac_hbm": bytes_it / (ms_it * 1e-3) / 1e9 / (peak * world),
                                  "clip_fallbacks": hp.clip_fallbacks}}
    if hasattr(x, "close"):
        x.close()
    c.close()
    return out


def fr_bench_multi(args, rank, world, local_rank, peak, torch, dist, stream):
    """BASELINE configs[4] as named: reacting 5-species air (compressibleEulerFR), implicit, on N partitions (one z-slab of
    n^3 hexes per GPU, colour-sorted numbering, 9x9 block-CRS rows with ghost columns).  One iteration = UpdateBCs,
    gradient, limiter, HLLC residual + finite-rate source, nsgs SGS sweeps, ApplyDQ with the reference's halo
    exchanges (q 21 wide, qgrad 42, limiter 9, x 9 after every sweep); Jacobian kept.  Max over ranks."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_slab_case
    from proteuscfd_b200.parallel import DistributedHotPath, NcclExchange, PObj, PutExchange, TorchGroup
    n = args.fr_n or args.n
    nsgs, reps = 5, 3
    torch.cuda.empty_cache()
    dev = torch.device("cuda", local_rank)
    mesh, params, q, beta = fr_slab_case(n, rank, world, fr_params_from_fixture(), device=f"cuda:{local_rank}")
    c = capi.Context(mesh, params, device=local_rank)
    c.set_stream(stream.cuda_stream)
    c.set_field(capi.F_BETA, beta)
    group = TorchGroup(dist)
    pobj = PObj(rank, world).BuildCommMaps(mesh["gNodeOwner"], mesh["gNodeLocalId"], group)
    x = PutExchange(c, pobj, dist, torch, dev, group) if args.halo == "put" else NcclExchange(c, pobj, dist, torch, dev)
    hitflag = torch.zeros(1, device=dev, dtype=torch.int32)

    def any_rank(hit):   # the (rare) pressure-clip fallback is a global decision: one 4-byte max all-reduce
        hitflag.fill_(int(hit))
        dist.all_reduce(hitflag, op=dist.ReduceOp.MAX)
        return bool(hitflag.item())

    hp = DistributedHotPath(c, x, any_rank=any_rank)
    hp.setup()
    c.set_field(capi.F_Q, q)
    hp.implicit_iterate(nsgs, refresh_jac=True)     # builds A and its LU; warm-up

    def sync():
        dist.barrier()
        torch.cuda.synchronize()

    def maxms(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def total(v):
        t = torch.tensor([v], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    ev = lambda: torch.cuda.Event(enable_timing=True)
    sync()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(reps):
        hp.implicit_iterate(nsgs, refresh_jac=False)
    e1.record(stream)
    sync()
    ms_it = maxms(e0.elapsed_time(e1) / reps)
    finite = bool(np.isfinite(c.get_field(capi.F_Q)).all())
    c.blank_x()
    x.update(capi.F_X)
    nsw = 6
    sync()
    e2, e3 = ev(), ev()
    e2.record(stream)
    for _ in range(nsw):
        c.sgs(1, want_ddq=False)
        x.update(capi.F_X)
    e3.record(stream)
    sync()
    ms_sweep = maxms(e2.elapsed_time(e3) / nsw)
    e4, e5 = ev(), ev()
    e4.record(stream)
    c.timestep(want_min=False)
    c.jacobian()
    c.prepare_sgs()
    e5.record(stream)
    sync()
    ms_jac = maxms(e4.elapsed_time(e5))
    nblocks = int(c.get_crs()[1].size)
    neqn, nterms = c.neqn, c.nterms
    ne, nn, nl = c.nedge + c.ngedge, c.nnode, c.nnode + c.gnode
    bytes_sweep = total(2 * (nblocks * 8 * neqn * neqn + 4 * nblocks + 4 * (nn + 1) + 4 * neqn * nn + 8 * neqn * 3 * nn))
    bytes_head = (8 * ne + nl * (24 + 8 * nterms + 48 + 24 * nterms)
                  + (8 * ne + nl * (8 * neqn + 16 * neqn)) + (8 * ne + nl * (8 * neqn + 24 * neqn + 24 + 16 * neqn) + nl * 8 * neqn)
                  + (8 * ne + nl * (8 * neqn + 24 * neqn + 24 + 8 * neqn) + nl * 8 * neqn)
                  + 40 * ne + nl * (8 * neqn + 24 * neqn + 8 * neqn + 24 + 8) + nn * 8 * neqn)
    bytes_iter = total(bytes_head) + nsgs * bytes_sweep
    nn_g, nb_g = total(nn), total(nblocks)
    ne_g = total(2 * c.nedge + c.ngedge) // 2
    out = {"workload": f"BASELINE configs[4]: reacting 5-species air (compressibleEulerFR), implicit, {world} z-slab partitions of an "
                       f"n x n x {n * world} Kuhn box (n = {n}: {6 * n ** 3 * world} tets, {nn_g} nodes, {ne_g} edges, {nb_g} 9x9 blocks = "
                       f"{nb_g * 648 / 1e9:.1f} GB) on {world} GPUs; halo exchange '{args.halo}'",
           "n_gpus": world, "nsgs": nsgs,
           "iteration_without_refresh_ms": ms_it, "iteration_Medges_s": ne_g / (ms_it * 1e-3) / 1e6,
           "iteration_algorithmic_bytes": bytes_iter, "iteration_GBps": bytes_iter / (ms_it * 1e-3) / 1e9,
           "iteration_frac_hbm": bytes_iter / (ms_it * 1e-3) / 1e9 / (peak * world),
           "sgs_ms_per_sweep": ms_sweep, "sgs_sweeps_per_s": 1e3 / ms_sweep, "sgs_algorithmic_bytes_per_sweep": bytes_sweep,
           "sgs_GBps": bytes_sweep / (ms_sweep * 1e-3) / 1e9, "sgs_frac_hbm": bytes_sweep / (ms_sweep * 1e-3) / 1e9 / (peak * world),
           "jacobian_refresh_ms": ms_jac, "state_finite_after_run": finite, "clip_fallbacks": hp.clip_fallbacks}
    if hasattr(x, "close"):
        x.close()
    c.close()
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
           "state_finite_after_run": finite, "clip_fallbacks": c.clip_fallbacks(), "non_sgs_kernels_ms": ms_explicit_part,
           "kernels_ms": kern}
    c.close()
    return out


if __name__ == "__main__":
    main()
