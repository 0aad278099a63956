// pcfd_internal.cuh -- definitions shared by the translation units of libpcfd_b200.so (pcfd_kernels.cu: perfect-gas
// eqnsets + the C ABI; pcfd_fr.cu: the reacting eqnset).  Not part of the public interface (include/pcfd.h is).
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pcfd.h"
#include "eqnset_compressible.cuh"
#include "eqnset_compressible_cs.cuh"

std::string& pcfd_create_err();   // message of a failed pcfd_create* (pcfd_last_error(NULL))

namespace {

struct DevMesh {
  int nnode, gnode, nbnode, nedge, nbedge, ngedge;
  const int2* en;        // [nedge]   (left, right)
  const double* ea;      // [nedge*4]
  const int2* ben;       // [nbedge+ngedge]
  const double* bea;     // [(nbedge+ngedge)*4]
  const int* bctype;     // [nbedge+ngedge]
  const double* xyz;     // [(nnode+gnode)*3]
  const double* vol;     // [nnode]
  const int* adjp;       // [nnode+1]
  const int2* adj;       // (.x = other node | role<<31 (1 = this node is the RIGHT node), .y = edge id; >= nedge: half-edge)
  const int* bnormal;    // [nbedge] most-normal neighbour of the wall node (NoSlip half-edges, else -1)
  const double* btwall;  // [nbedge] non-dimensional wall temperature of the half-edge's surface (< 0: adiabatic)
  const double* bubar;   // [nbedge] PowerLawU(1, wall distance of the left node, Re) of FarFieldViscous half-edges (else 1)
};

__device__ __forceinline__ bool is_ghost(const DevMesh& m, int n) { return n >= m.nnode && n < m.nnode + m.gnode; }

__device__ __forceinline__ void load_avec(const double* __restrict__ a, int e, double* v) {
  const double2* p = reinterpret_cast<const double2*>(a + (size_t)e * 4);
  const double2 x = __ldg(p), y = __ldg(p + 1);
  v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
}

// gradient.tcc:141-168 ComputeLSQCoefficients
__device__ __forceinline__ void lsq_weights(const double* s, const double* dxbar, double* we) {
  const double r11 = s[0], r12 = s[1], r13 = s[2], s22 = s[3], s23 = s[4], s33 = s[5];
  const double r12_r11 = (r11 == 0.0) ? 0.0 : r12 / r11;
  const double r22 = s22 - r12 * r12_r11;
  const double r23 = s23 - r12_r11 * r13;
  const double r13_r11 = (r11 == 0.0) ? 0.0 : r13 / r11;
  const double r23_r22 = (r22 == 0.0) ? 0.0 : r23 / r22;
  const double r33 = s33 - r13 * r13_r11 - r23 * r23_r22;
  const double dykdx = (dxbar[1] - (r12_r11)*dxbar[0]);
  we[2] = (r33 == 0.0) ? 0.0 : (dxbar[2] - r13_r11 * dxbar[0] - r23_r22 * dykdx) / r33;
  we[1] = (r22 == 0.0) ? 0.0 : (dykdx - r23 * we[2]) / r22;
  we[0] = (r11 == 0.0) ? 0.0 : (dxbar[0] - r12 * we[1] - r13 * we[2]) / r11;
}

__device__ __forceinline__ double limiter_fn(int type, double t) {
  if (type == 1) {   // Barth, limiters.tcc:263-266
    t = eq::maxd(0.0, t);
    t = eq::mind(1.0, t);
    return t;
  }
  return (t * t + 2.0 * t) / (t * t + t + 2.0);   // Venkatakrishnan, :440
}

}  // namespace

// ===================================================================== context
struct pcfd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  int nnode = 0, gnode = 0, nbnode = 0, nedge = 0, nbedge = 0, ngedge = 0;
  int nb = 0, nn = 0, ntot = 0, nblocks = 0;
  pcfd_params prm{};
  int field_jac_type = 0, boundary_jac_type = 0;   // Param::fieldJacType / boundaryJacType: 0 one-sided, 1 central FD
  int grad_type = 0;   // Param::gradType: 0 weighted least squares, 1 Green-Gauss (pcfd_set_gradient_type)
  // system widths: 5 / 10 / 9 for the perfect-gas eqnsets, ns+4 / 3ns+6 / 2ns+4 for the reacting one
  int neqn = PCFD_NEQN, nvars = PCFD_NVARS, nterms = PCFD_NTERMS;
  struct pcfd_fr_state* fr = nullptr;   // reacting eqnset (pcfd_fr.cu); null for the perfect-gas eqnsets
  struct pcfd_comm* comm = nullptr;     // flag-based direct-put exchange between ranks (pcfd_comm.cuh); null: single rank
  struct pcfd_forces* forces = nullptr; // surface forces (pcfd_forces.cuh); null until pcfd_forces_configure
  DevMesh dm{};
  eq::BcParams bp{};
  double* f[PCFD_F_COUNT] = {};
  size_t fsize[PCFD_F_COUNT] = {};
  int2 *en = nullptr, *ben = nullptr, *adj = nullptr;
  double *ea = nullptr, *bea = nullptr, *xyz = nullptr, *vol = nullptr;
  int *bctype = nullptr, *adjp = nullptr;
  // half-edge work lists: nodes owning a Dirichlet-type half-edge are walked sequentially (bnodes),
  // every other half-edge gets its own thread (blist: BC half-edges first, then ghost half-edges)
  int *bnodes = nullptr, *blist = nullptr;
  unsigned char* bfirst = nullptr;
  int nbn = 0, nblist = 0, nblist_bc = 0;
  double* bdiag = nullptr;
  double *flux = nullptr, *bflux = nullptr, *red = nullptr, *redout = nullptr;
  // ComputeTimesteps riding along in the residual pass of pcfd_explicit_iterate: per-edge / per-half-edge terms
  double *eig = nullptr, *beig = nullptr;
  bool eig_fuse = true;            // PCFD_EIG_FUSE=0: separate k_timestep pass
  bool eig_now = false;            // set by pcfd_explicit_iterate around its residual pass
  // viscous terms (compressibleNS): per-edge viscous flux slots, wall-node bookkeeping, VNN time-step limit
  bool viscous = false;
  eq::ViscParams vp{};
  double *vflux = nullptr, *bvflux = nullptr, *btwall = nullptr, *vnn23 = nullptr;
  int *bnormal = nullptr, *wnodes = nullptr, *tbnodes = nullptr;
  // Proteus_FarFieldViscous (bc.tcc:1092-1108): half-edges of that type, their left nodes, the device table of ubar
  std::vector<int> ffv_edges, ffv_left;
  double* bubar = nullptr;
  bool ffv_ready = false;          // PCFD_F_WALLDIST has been set since the table was last built
  int ntbnodes = 0;
  double *tslots = nullptr, *tbslots = nullptr;   // Spalart-Allmaras per-edge / per-half-edge slots
  // SA under the reacting eqnset: {theta, nu} per edge / half-edge and {rho, nu} per local + ghost node (kfr_turb_props)
  double *tprop_e = nullptr, *tprop_b = nullptr, *tprop_n = nullptr;
  unsigned char* wallflag = nullptr;
  int nwall = 0;
  unsigned char* clipflag = nullptr;
  int *tclip[2] = {nullptr, nullptr}, *dflags = nullptr;
  int* hflag = nullptr;            // pinned host copy of the fused clip flag
  cudaEvent_t ev_flag = nullptr;
  bool fused_clip = true;          // PCFD_FUSED_CLIP=0 keeps the separate clip pass in the composite iterations
  // static LSQ weights per node visit (k_lsq_geo) for k_gradient_geo; PCFD_GRAD_GEO=0: k_gradient recomputes them
  double4* geo = nullptr;
  bool geo_valid = false;
  bool use_geo = true;
  bool limiter_per_node = false;   // PCFD_LIMITER_PER_NODE=1: one thread per node (k_limiter_node; slower on B200: 0.79 vs 0.68 ms)
  int grad_threads = 3;            // PCFD_GRAD_THREADS=1: one thread per node (k_gradient_geo) instead of three (k_gradient_geo3)
  // neighbour min / max of the conservative variables taken by k_gradient_geo3 (pass 1 of Limiter::Compute); valid until
  // q changes (set_field, UpdateBCs, updates, a halo of q)
  double* qmm = nullptr;
  bool qmm_valid = false, use_qmm = true;   // PCFD_LIMITER_QMM=0: k_limiter walks the neighbours itself
  long long clip_fallbacks = 0;
  int* dzeroed = nullptr;          // nodes whose NaN / Inf update ApplyDQ zeroed (solutionSpace.tcc:771-796), device counter
  int *ia = nullptr, *ja = nullptr, *iau = nullptr, *pv = nullptr, *posLR = nullptr, *posRL = nullptr, *bpos = nullptr;
  int *rows_f = nullptr, *rows_b = nullptr;
  std::vector<int> lev_f, lev_b;   // level offsets into rows_f / rows_b
  int *dlev_f = nullptr, *dlev_b = nullptr;   // the same on the device (Spalart-Allmaras contexts: k_sgs_scalar_sweeps)
  bool turb_persist = true;        // PCFD_TURB_PERSIST=0: one launch per level for the scalar sweeps
  int turb_grid = 0;
  // halo maps (PObj::BuildCommMaps, parallel.tcc:461-554): what this rank sends to / receives from each peer
  int rank = 0, nranks = 1;
  std::vector<int> send_counts, send_offsets, recv_counts, recv_offsets;
  int* send_list = nullptr;
  int send_total = 0;
  // time integration (pcfd_set_time_integration): Param::dt, useLocalTimeStepping, torder, SolutionSpace::iter
  double time_dt = -1.0;
  int time_local = 1, torder = 1, iter = 1;
  bool have_qold = false;   // PCFD_F_QOLD has been set: TemporalResidual is live
  bool ludiag = false;
  // scratch of pcfd_gmres (Krylov vectors, block-diagonal preconditioner), grown on demand
  double* gm_buf = nullptr;
  int* gm_pv = nullptr;
  size_t gm_cap = 0;
  double* gm_n = nullptr;          // GMRES preconditioner type 4: copy of the matrix with factored diagonal blocks
  int* gm_npv = nullptr;
  size_t gm_ncap = 0;
  double sgs_prev_norm = 0.0;   // xNorm of the last-but-one sweep when the sweeps of a solve are separate calls (multi-rank)
  int sgs_unroll = 4;      // blocks in flight per lane in k_sgs_level (PCFD_SGS_UNROLL overrides, for tuning)
  // bulk-copy (TMA) streaming variant: per level, shared-memory bytes for the matrix part of a tile (0: the rows of
  // the level are not consecutive in memory -> per-lane-load kernel); PCFD_SGS_TILE_WARPS = 0 disables it
  int sgs_tile_warps = 4, sgs_tile_lpr = 16;   // PCFD_SGS_TILE_WARPS / PCFD_SGS_TILE_LPR (lanes per row: 5, 10, 16)
  int sgs_pf_dist = -1;    // PCFD_SGS_PREFETCH_TILES (-1: automatic, 0: off)
  bool sgs_pdl = true;     // PCFD_SGS_PDL=0: plain launches between the levels of a sweep (no programmatic dependent launch)
  int sgs_ring_stages = 0, sgs_ring_ctas_per_sm = 0, ring_cap_blocks = 0, num_sms = 148;
  std::vector<int> tile_cap_f, tile_cap_b, lev_first_f, lev_first_b, lev_step_f, lev_step_b;
  std::vector<void*> allocs;
  std::string err;
  long long launches = 0;
  // optional per-kernel timing with CUDA events on the launch stream (pcfd_profile_*)
  bool prof = false;
  struct ProfRec { const char* name; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  struct ProfAcc { std::string name; double ms = 0; long long n = 0; };
  std::vector<ProfAcc> prof_acc;
};

namespace {

constexpr int RED_BLOCKS = 296;   // 2 x 148 SMs

int fail(pcfd_ctx* c, const std::string& msg) {
  if (c) c->err = msg; else pcfd_create_err() = msg;
  return 1;
}
#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail(c, std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)
#define PROF(name)                                                                                     \
  do {                                                                                                 \
    if (c->prof) prof_begin(c, name);                                                                  \
  } while (0)
#define LAUNCH_CHECK()                                                                                 \
  do {                                                                                                 \
    c->launches++;                                                                                     \
    if (c->prof) prof_end(c);                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                               \
    if (e_ != cudaSuccess) return fail(c, std::string("kernel launch: ") + cudaGetErrorString(e_));    \
  } while (0)

cudaEvent_t prof_event(pcfd_ctx* c) {
  if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(pcfd_ctx* c, const char* name) {
  pcfd_ctx::ProfRec r{name, prof_event(c), prof_event(c)};
  cudaEventRecord(r.a, c->stream);
  c->prof_pending.push_back(r);
}
void prof_end(pcfd_ctx* c) {
  if (!c->prof_pending.empty()) cudaEventRecord(c->prof_pending.back().b, c->stream);
}
void prof_drain(pcfd_ctx* c) {
  cudaStreamSynchronize(c->stream);
  for (auto& r : c->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      pcfd_ctx::ProfAcc* acc = nullptr;
      for (auto& a : c->prof_acc) if (a.name == r.name) acc = &a;
      if (!acc) { c->prof_acc.push_back({r.name, 0.0, 0}); acc = &c->prof_acc.back(); }
      acc->ms += ms;
      acc->n++;
    }
    c->prof_pool.push_back(r.a);
    c->prof_pool.push_back(r.b);
  }
  c->prof_pending.clear();
}

template <class T>
int dev_alloc(pcfd_ctx* c, T** p, size_t n) {
  void* v = nullptr;
  CK(cudaMalloc(&v, std::max<size_t>(n, 1) * sizeof(T)));
  c->allocs.push_back(v);
  *p = static_cast<T*>(v);
  return 0;
}
template <class T>
int dev_upload(pcfd_ctx* c, T** p, const T* host, size_t n) {
  if (dev_alloc(c, p, n)) return 1;
  if (n) CK(cudaMemcpy(*p, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
inline int nblk(long long n, int bs) { return (int)std::max<long long>(1, (n + bs - 1) / bs); }

// doubles per node of an exchangeable field (0: the field has no ghost rows)
inline int field_width(const pcfd_ctx* c, int field) {
  switch (field) {
    case PCFD_F_Q: return c->nvars;
    case PCFD_F_QGRAD: return c->nterms * 3;
    case PCFD_F_LIMITER: case PCFD_F_X: return c->neqn;
    case PCFD_F_LSQ_S: case PCFD_F_LSQ_SW: return 6;
    case PCFD_F_BETA: case PCFD_F_MUT: case PCFD_F_TVAR: case PCFD_F_TURB_X: case PCFD_F_WALLDIST: return 1;
    case PCFD_F_TGRAD: return 3;
    default: return 0;
  }
}

}  // namespace

// ---- the reacting eqnset's side of every phase entry point (pcfd_fr.cu); the C ABI functions in pcfd_kernels.cu
// forward to these when ctx->fr is set
// CRSMatrix::PrepareSGS (crsmatrix.tcc:840-876) -> LU (matrix.h:110-190) of every diagonal block, N lanes per node: lane r
// holds row r of the block in registers (one thread per node kept the N x N block and the dynamically indexed
// permutation in local memory: 11.8 GB of DRAM traffic for 2.2 GB of 9x9 blocks in the round-1 ncu capture).  The
// permutation lives as its inverse, one integer per lane (pos: where in p this row stands); the pivot search walks the
// positions i..N-1 in order with the reference's strict '>' on |a| (first maximum wins, `row` persists when a column is
// all zero), values fetched by shuffle from the lane that stands at that position; the pivot row is broadcast entry by
// entry.  Same operations on every entry as the sequential routine, so the factors and pv are bit-identical.
template <int N>
__global__ void __launch_bounds__(128) k_lu_diag_lanes(int nnode, const int* __restrict__ iau, double* __restrict__ A,
                                                        int* __restrict__ pv) {
  constexpr int RPW = 32 / N;   // nodes per warp
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int grp = lane / N, r = lane - grp * N;
  const int nd = warp * RPW + grp;
  const bool live = grp < RPW && nd < nnode;
  const int base = grp * N;     // first lane of this node's group
  double a[N];
  double* g = nullptr;
  if (live) {
    g = A + (size_t)iau[nd] * N * N + (size_t)r * N;
#pragma unroll
    for (int k = 0; k < N; k++) a[k] = g[k];
  } else {
#pragma unroll
    for (int k = 0; k < N; k++) a[k] = 0.0;
  }
  int pos = r;                  // p[pos] == r
  int row = 0;
  // Every update below is a select, not a branch: the node groups of a warp take different decisions, and the shuffles
  // that follow want all lanes in the same static instruction.
#pragma unroll
  for (int i = 0; i < N; i++) {
    __syncwarp(full);
    // the reference's scan over the positions i..N-1: the value of column i is fetched from the lane that stands at
    // position j (found by ballot); strict '>' on the magnitudes.  |large| is carried next to large: with
    // `fabs(x) > fabs(large)` nvcc 12.9 emitted the second comparison of every step as DSETP.GT |x|, large -- the abs on
    // large dropped -- and a negative first candidate lost against any second one (caught by the cube fixture, cuobjdump -sass)
    const double mine = a[i];
    double large = 0.0, lmag = 0.0;
#pragma unroll
    for (int j = i; j < N; j++) {
      const unsigned at = __ballot_sync(full, live && pos == j);
      const int src = __ffs((at >> base) & ((1u << N) - 1u)) - 1;
      const double x = __shfl_sync(full, mine, base + (src < 0 ? 0 : src));
      const double xmag = fabs(x);
      const bool gt = xmag > lmag;
      large = gt ? x : large;
      lmag = gt ? xmag : lmag;
      row = gt ? j : row;
    }
    // swap p[i] and p[row]
    pos = (pos == i) ? row : ((pos == row) ? i : pos);
    large = 1.0 / large;
    const unsigned atp = __ballot_sync(full, live && pos == i);
    const int piv = __ffs((atp >> base) & ((1u << N) - 1u)) - 1;
    const bool below = pos > i;
    const double scaled = a[i] * large;
    a[i] = below ? scaled : a[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) {
      const double pk = __shfl_sync(full, a[k], base + (piv < 0 ? 0 : piv));
      const double upd = a[k] - a[i] * pk;
      a[k] = below ? upd : a[k];
    }
  }
  if (live) {
#pragma unroll
    for (int k = 0; k < N; k++) g[k] = a[k];
    pv[(size_t)nd * N + pos] = r;
  }
}

int pcfd_internal_create(const pcfd_mesh_desc* mesh, const pcfd_params* params, int device, int neqn, int nvars,
                         int nterms, pcfd_ctx** out);
void pcfd_fr_destroy(pcfd_ctx* c);
int pcfd_fr_update_bcs(pcfd_ctx* c);
int pcfd_fr_gradient(pcfd_ctx* c);
int pcfd_fr_limiter(pcfd_ctx* c);
int pcfd_fr_residual(pcfd_ctx* c, double* sumsq);
int pcfd_fr_limiter_raw(pcfd_ctx* c);
void pcfd_fr_set_time(pcfd_ctx* c, double dt, int use_local);
int pcfd_fr_residual_fused(pcfd_ctx* c, double* sumsq, bool* clip_hit);
int pcfd_fr_timestep(pcfd_ctx* c, double* dtmin);
int pcfd_fr_turb_props(pcfd_ctx* c);
// {p, cp, mu, rho} of the surface node of every BC half-edge (Forces, pcfd_forces.cuh); also GetDensity(Qinf), viscous
int pcfd_fr_surface_props(pcfd_ctx* c, double V, double* props, double* rho_inf, bool* viscous);
int pcfd_fr_explicit_solve(pcfd_ctx* c);
int pcfd_fr_apply_dq(pcfd_ctx* c);
int pcfd_fr_jacobian(pcfd_ctx* c);
int pcfd_fr_prepare_sgs(pcfd_ctx* c);
int pcfd_fr_sgs(pcfd_ctx* c, int nsgs, double* ddq);
