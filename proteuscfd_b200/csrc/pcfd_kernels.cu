// pcfd_kernels.cu -- CUDA kernels (sm_100a, FP64) and the C ABI of libpcfd_b200.so.
//
// Design (DESIGN.md has the long form):
//  * every reduction onto a node is a GATHER by the thread that owns the node, over
//    that node's incident edges in EDGE-INDEX ORDER -- the order in which the
//    reference's sequential Driver/Bdriver loops (ucs/driver.tcc:5-140) scatter into
//    it.  No atomics, no colouring of the edge loop, run-to-run deterministic, and
//    the floating-point summation order is the reference's, so results are
//    bit-identical to the reference, not merely within 1e-12.
//  * expensive per-edge work (Roe flux, the 11-flux finite-difference Jacobian) is
//    evaluated ONCE per edge by an edge-parallel kernel that writes to a private
//    slot (flux[e], or the two off-diagonal blocks of the edge); the node-parallel
//    gather then reads those slots.
//  * cheap per-edge work (LSQ gradient terms, limiter, spectral radius) is
//    recomputed from both ends inside the node-parallel kernel.
//  * SGS rows are level-scheduled from the actual node numbering, which reproduces
//    the reference's sequential Gauss-Seidel exactly for ANY numbering; with a
//    colour-sorted numbering the levels are the colours.
//
// Compile with --fmad=false (see eqnset_compressible.cuh for why).

#include <cooperative_groups.h>

#include "pcfd_internal.cuh"
#include "sgs_tile.cuh"

#define NEQN PCFD_NEQN
#define NVARS PCFD_NVARS
#define NTERMS PCFD_NTERMS
#define NEQN2 (NEQN * NEQN)

#ifndef PCFD_GRAD_MINB
#define PCFD_GRAD_MINB 4
#endif
#ifndef PCFD_GRAD3_MINB
#define PCFD_GRAD3_MINB 5
#endif
#ifndef PCFD_LIM_MINB
#define PCFD_LIM_MINB 4
#endif

namespace {

// gradient terms -> variables (compressible.tcc:1009-1027): {0, 1, 2, 3, 4, 5, 7, 8, 9}, spelled (i < 6) ? i : i + 1 below


__device__ __forceinline__ void load5(const double* __restrict__ p, double* v) {
#pragma unroll
  for (int i = 0; i < 5; i++) v[i] = __ldg(p + i);
}
// q rows are 80 B => 16 B aligned
__device__ __forceinline__ void load_q5(const double* __restrict__ q, int n, double* v) {
  const double2* p = reinterpret_cast<const double2*>(q + (size_t)n * NVARS);
  const double2 a = __ldg(p), b = __ldg(p + 1);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  v[4] = __ldg(q + (size_t)n * NVARS + 4);
}
__device__ __forceinline__ void load_q10(const double* q, int n, double* v) {
  const double2* p = reinterpret_cast<const double2*>(q + (size_t)n * NVARS);
#pragma unroll
  for (int i = 0; i < 5; i++) { const double2 a = p[i]; v[2 * i] = a.x; v[2 * i + 1] = a.y; }
}
__device__ __forceinline__ void store_q10(double* q, int n, const double* v) {
  double2* p = reinterpret_cast<double2*>(q + (size_t)n * NVARS);
#pragma unroll
  for (int i = 0; i < 5; i++) p[i] = make_double2(v[2 * i], v[2 * i + 1]);
}

// ============================================================== LSQ coefficients
// gradient.tcc:115-138 with kernels :381-542 -> Mesh::s, Mesh::sw of the local nodes
__global__ void k_lsq_coeff(DevMesh m, double* __restrict__ s, double* __restrict__ sw) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  const double xn[3] = {m.xyz[3 * n], m.xyz[3 * n + 1], m.xyz[3 * n + 2]};
  double S[6] = {0, 0, 0, 0, 0, 0}, W[6] = {0, 0, 0, 0, 0, 0};
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    const int o = a.x & 0x7fffffff;
    const bool right = a.x < 0;
    if (a.y >= m.nedge && !is_ghost(m, o)) continue;
    const double xo[3] = {m.xyz[3 * o], m.xyz[3 * o + 1], m.xyz[3 * o + 2]};
    double dx[3];   // always x_left - x_right
#pragma unroll
    for (int d = 0; d < 3; d++) dx[d] = right ? (xo[d] - xn[d]) : (xn[d] - xo[d]);
    const double ds2 = 1.0 / (dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
    if (!right) {
      S[0] += dx[0] * dx[0]; S[1] += dx[0] * dx[1]; S[2] += dx[0] * dx[2];
      S[3] += dx[1] * dx[1]; S[4] += dx[1] * dx[2]; S[5] += dx[2] * dx[2];
      W[0] += dx[0] * dx[0] * ds2; W[1] += dx[0] * dx[1] * ds2; W[2] += dx[0] * dx[2] * ds2;
      W[3] += dx[1] * dx[1] * ds2; W[4] += dx[1] * dx[2] * ds2; W[5] += dx[2] * dx[2] * ds2;
    } else {
      const double mx = -dx[0], my = -dx[1], mz = -dx[2];
      S[0] += dx[0] * dx[0]; S[1] += mx * my; S[2] += mx * mz;
      S[3] += dx[1] * dx[1]; S[4] += my * mz; S[5] += dx[2] * dx[2];
      W[0] += dx[0] * dx[0] * ds2; W[1] += (mx * my) * ds2; W[2] += (mx * mz) * ds2;
      W[3] += dx[1] * dx[1] * ds2; W[4] += (my * mz) * ds2; W[5] += dx[2] * dx[2] * ds2;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; k++) { s[6 * (size_t)n + k] = S[k]; sw[6 * (size_t)n + k] = W[k]; }
}


// ====================================================================== gradient
// Gradient::Compute (gradient.tcc:57-112), weighted LSQ kernels :251-378 and the
// symmetry-plane fix :545-565, as one ordered gather per node.
__global__ void __launch_bounds__(128) k_gradient(DevMesh m, const double* __restrict__ q, const double* __restrict__ sw,
                                                   double* __restrict__ qgrad) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double g[NTERMS * 3];
#pragma unroll
  for (int k = 0; k < NTERMS * 3; k++) g[k] = 0.0;
  double qn[NVARS], swn[6];
  load_q10(q, n, qn);
#pragma unroll
  for (int k = 0; k < 6; k++) swn[k] = sw[6 * (size_t)n + k];
  const double xn[3] = {m.xyz[3 * n], m.xyz[3 * n + 1], m.xyz[3 * n + 2]};
  const int kend = m.adjp[n + 1];
  int k = m.adjp[n];
  for (; k < kend; k++) {
    const int2 a = m.adj[k];
    const int o = a.x & 0x7fffffff;
    const bool right = a.x < 0;
    if (a.y >= m.nedge && !is_ghost(m, o)) continue;
    double qo[NVARS];
    load_q10(q, o, qo);
    const double xo[3] = {__ldg(m.xyz + 3 * o), __ldg(m.xyz + 3 * o + 1), __ldg(m.xyz + 3 * o + 2)};
    double dx[3], we[3];   // dx = x_left - x_right
#pragma unroll
    for (int d = 0; d < 3; d++) dx[d] = right ? (xo[d] - xn[d]) : (xn[d] - xo[d]);
    const double dx2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    const double weight = 1.0 / sqrt(dx2);
    dx[0] *= weight; dx[1] *= weight; dx[2] *= weight;
    if (right) { dx[0] = -dx[0]; dx[1] = -dx[1]; dx[2] = -dx[2]; }
    lsq_weights(swn, dx, we);
#pragma unroll
    for (int i = 0; i < NTERMS; i++) {
      const int v = (i < 6) ? i : i + 1;   // c_gradloc
      // dq = weight*(qR - qL)
      const double dq = right ? weight * (qn[v] - qo[v]) : weight * (qo[v] - qn[v]);
#pragma unroll
      for (int j = 0; j < 3; j++) {
        if (right) g[3 * i + j] += +we[j] * dq;
        else g[3 * i + j] += -we[j] * dq;
      }
    }
  }
  // symmetry fix, in half-edge order (the half-edges are the tail of the list)
  for (k = m.adjp[n]; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    if (m.bctype[be] != PCFD_BC_SYMMETRY) continue;
    double av[4];
    load_avec(m.bea, be, av);
#pragma unroll
    for (int i = 0; i < NTERMS; i++) {
      const double dot = g[i * 3] * av[0] + g[i * 3 + 1] * av[1] + g[i * 3 + 2] * av[2];
#pragma unroll
      for (int j = 0; j < 3; j++) g[i * 3 + j] -= dot * av[j];
    }
  }
  double* out = qgrad + (size_t)n * NTERMS * 3;
#pragma unroll
  for (int kk = 0; kk < NTERMS * 3; kk++) out[kk] = g[kk];
}

// ------------------------------------------------------------------ fused gradient + limiter (composite iterations)
// The weighted-LSQ edge weights are pure geometry: for visit k of node n (its k-th incident edge, in edge order)
// weight = 1/|dx| and we[3] = ComputeLSQCoefficients(sw[n], dx/|dx|) (gradient.tcc:141-168, 276-312) never change on a
// static mesh.  k_lsq_geo evaluates them ONCE with the arithmetic of k_gradient and stores {weight, +-we[0..2]} per
// visit (the sign of the scatter, gradient.tcc:306-311, folded in: (-a)*b == -(a*b) exactly), so that the per-iteration
// kernel is left with one subtraction, one product and the ordered sum per term.  BC half-edges (no contribution,
// gradient.tcc:322-378 only acts on parallel edges) get zeros and are skipped through the visit mask.
__global__ void __launch_bounds__(128) k_lsq_geo(DevMesh m, const double* __restrict__ sw, double4* __restrict__ geo) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double swn[6];
#pragma unroll
  for (int k = 0; k < 6; k++) swn[k] = sw[6 * (size_t)n + k];
  const double xn[3] = {m.xyz[3 * n], m.xyz[3 * n + 1], m.xyz[3 * n + 2]};
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    const int o = a.x & 0x7fffffff;
    const bool right = a.x < 0;
    if (a.y >= m.nedge && !is_ghost(m, o)) { geo[k] = make_double4(0.0, 0.0, 0.0, 0.0); continue; }
    const double xo[3] = {__ldg(m.xyz + 3 * o), __ldg(m.xyz + 3 * o + 1), __ldg(m.xyz + 3 * o + 2)};
    double dx[3], we[3];   // dx = x_left - x_right
#pragma unroll
    for (int d = 0; d < 3; d++) dx[d] = right ? (xo[d] - xn[d]) : (xn[d] - xo[d]);
    const double dx2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    const double weight = 1.0 / sqrt(dx2);
    dx[0] *= weight; dx[1] *= weight; dx[2] *= weight;
    if (right) { dx[0] = -dx[0]; dx[1] = -dx[1]; dx[2] = -dx[2]; }
    lsq_weights(swn, dx, we);
    geo[k] = right ? make_double4(weight, we[0], we[1], we[2]) : make_double4(weight, -we[0], -we[1], -we[2]);
  }
}

// Gradient::Compute with the precomputed visit weights: one thread per node, ordered gather like k_gradient, but a
// visit is now 9 x (1 subtraction, 4 products, 3 additions) -- no square root, no divisions, no coordinates.
// (A 16-lanes-per-node form with one lane per visit and the ordered sums through shared memory was measured on B200
// at 2.47 ms against 1.33 ms for k_gradient + k_limiter: 2.3x the instructions for the transposes and shuffles;
// profiles/r2_ncu_explicit.md.)
__global__ void __launch_bounds__(128, PCFD_GRAD_MINB) k_gradient_geo(DevMesh m, const double* __restrict__ q,
                                                                       const double4* __restrict__ geo,
                                                                       double* __restrict__ qgrad) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double g[NTERMS * 3];
#pragma unroll
  for (int k = 0; k < NTERMS * 3; k++) g[k] = 0.0;
  double qn[NVARS];
  load_q10(q, n, qn);
  const int kend = m.adjp[n + 1];
  int k = m.adjp[n];
  for (; k < kend; k++) {
    const int2 a = m.adj[k];
    const int o = a.x & 0x7fffffff;
    const bool right = a.x < 0;
    if (a.y >= m.nedge && !is_ghost(m, o)) continue;
    const double2* gp = reinterpret_cast<const double2*>(geo + k);
    const double2 ga = __ldg(gp), gb = __ldg(gp + 1);
    const double weight = ga.x, we[3] = {ga.y, gb.x, gb.y};   // we carries the scatter sign
    double qo[NVARS];
    load_q10(q, o, qo);
#pragma unroll
    for (int i = 0; i < NTERMS; i++) {
      const int v = (i < 6) ? i : i + 1;   // c_gradloc
      const double dq = right ? weight * (qn[v] - qo[v]) : weight * (qo[v] - qn[v]);
#pragma unroll
      for (int j = 0; j < 3; j++) g[3 * i + j] += we[j] * dq;
    }
  }
  // symmetry fix, in half-edge order (the half-edges are the tail of the list)
  for (k = m.adjp[n]; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    if (m.bctype[be] != PCFD_BC_SYMMETRY) continue;
    double av[4];
    load_avec(m.bea, be, av);
#pragma unroll
    for (int i = 0; i < NTERMS; i++) {
      const double dot = g[i * 3] * av[0] + g[i * 3 + 1] * av[1] + g[i * 3 + 2] * av[2];
#pragma unroll
      for (int j = 0; j < 3; j++) g[i * 3 + j] -= dot * av[j];
    }
  }
  double* out = qgrad + (size_t)n * NTERMS * 3;
#pragma unroll
  for (int kk = 0; kk < NTERMS * 3; kk++) out[kk] = g[kk];
}

// The same gather by THREE threads per node, three gradient terms each: the kernel is bound by the latency of its
// neighbour-row gathers (ncu: 17 % issue slots, 16 stall cycles per instruction on the scoreboard at 25 % occupancy
// with 122 registers), and a third of the accumulators per thread triples the warps in flight.  The threads that own
// the five conservative variables also take the neighbour minimum / maximum of pass 1 of Limiter::Compute
// (Kernel_FindMinMax, limiters.tcc:135-191: from ZERO, over the same visits) while the rows are in registers: qmm[n] =
// {min[5], max[5]}, which saves the limiter kernel -- bound by instruction issue -- its first loop over the neighbours.
__global__ void __launch_bounds__(192, PCFD_GRAD3_MINB) k_gradient_geo3(DevMesh m, const double* __restrict__ q,
                                                                        const double4* __restrict__ geo,
                                                                        double* __restrict__ qgrad, double* __restrict__ qmm) {
  // results leave through shared memory: a thread owns 9 (+ up to 6) scattered doubles, the block's 64 nodes own one
  // contiguous 13.8 kB (+ 5 kB) range, which is written with full 16-byte lanes instead of 8-byte partial sectors
  __shared__ __align__(16) double s_g[192 * 9];
  __shared__ __align__(16) double s_mm[64 * 10];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = tid / 3;
  const int part = tid - 3 * n;              // terms 3*part .. 3*part+2
  const bool active = n < m.nnode;
  if (active) {
    // variables of those terms: {0,1,2}, {3,4,5}, {7,8,9}
    const int v0 = (part == 2) ? 7 : 3 * part;
    double g[9];
#pragma unroll
    for (int k = 0; k < 9; k++) g[k] = 0.0;
    double mn[3] = {0.0, 0.0, 0.0}, mx[3] = {0.0, 0.0, 0.0};
    const double* qr = q + (size_t)n * NVARS + v0;
    const double qn[3] = {qr[0], qr[1], qr[2]};
    const int kend = m.adjp[n + 1];
    int k = m.adjp[n];
    for (; k < kend; k++) {
      const int2 a = m.adj[k];
      const int o = a.x & 0x7fffffff;
      const bool right = a.x < 0;
      if (a.y >= m.nedge && !is_ghost(m, o)) continue;
      const double2* gp = reinterpret_cast<const double2*>(geo + k);
      const double2 ga = __ldg(gp), gb = __ldg(gp + 1);
      const double weight = ga.x, we[3] = {ga.y, gb.x, gb.y};   // we carries the scatter sign
      const double* qp = q + (size_t)o * NVARS + v0;
      const double qo[3] = {__ldg(qp), __ldg(qp + 1), __ldg(qp + 2)};
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double dq = right ? weight * (qn[i] - qo[i]) : weight * (qo[i] - qn[i]);
#pragma unroll
        for (int j = 0; j < 3; j++) g[3 * i + j] += we[j] * dq;
        mx[i] = eq::maxd(mx[i], qo[i]);
        mn[i] = eq::mind(mn[i], qo[i]);
      }
    }
    // symmetry fix, in half-edge order (the half-edges are the tail of the list)
    for (k = m.adjp[n]; k < kend; k++) {
      const int2 a = m.adj[k];
      if (a.y < m.nedge) continue;
      const int be = a.y - m.nedge;
      if (m.bctype[be] != PCFD_BC_SYMMETRY) continue;
      double av[4];
      load_avec(m.bea, be, av);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double dot = g[i * 3] * av[0] + g[i * 3 + 1] * av[1] + g[i * 3 + 2] * av[2];
#pragma unroll
        for (int j = 0; j < 3; j++) g[i * 3 + j] -= dot * av[j];
      }
    }
#pragma unroll
    for (int kk = 0; kk < 9; kk++) s_g[threadIdx.x * 9 + kk] = g[kk];
    if (part < 2) {   // equations 0..2 and 3..4 (variable 5, the temperature, is not limited)
      double* mm = s_mm + (threadIdx.x / 3) * 10;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        if (3 * part + i < 5) { mm[3 * part + i] = mn[i]; mm[5 + 3 * part + i] = mx[i]; }
      }
    }
  }
  __syncthreads();
  const int n0 = blockIdx.x * 64;
  const int nloc = min(64, m.nnode - n0);
  if (nloc <= 0) return;
  {
    double2* dst = reinterpret_cast<double2*>(qgrad + (size_t)n0 * NTERMS * 3);   // 64 * 216 B: 16-byte aligned
    const double2* src = reinterpret_cast<const double2*>(s_g);
    const int cnt = nloc * NTERMS * 3;
    for (int i = threadIdx.x; i < cnt / 2; i += 192) dst[i] = src[i];
    if ((cnt & 1) && threadIdx.x == 0) qgrad[(size_t)n0 * NTERMS * 3 + cnt - 1] = s_g[cnt - 1];
  }
  {
    double2* dst = reinterpret_cast<double2*>(qmm + (size_t)n0 * 10);
    const double2* src = reinterpret_cast<const double2*>(s_mm);
    for (int i = threadIdx.x; i < nloc * 5; i += 192) dst[i] = src[i];
  }
}

// Gradient::Compute with Param::gradType == 1 (gradient.tcc:77-90): Kernel_Green_Gauss_Gradient /
// Bkernel_Green_Gauss_Gradient (:170-248) in the Driver / Bdriver order (interior edges, then ALL half-edges, ghost
// and boundary alike), the division by the dual volume (:83-89) and the symmetry fix (:545-565), as one ordered
// gather per node.  The node's edge list is sorted by edge id with the half-edges at its tail, which is that order.
__global__ void __launch_bounds__(128) k_gradient_gg(DevMesh m, const double* __restrict__ q, double* __restrict__ qgrad) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double g[NTERMS * 3];
#pragma unroll
  for (int k = 0; k < NTERMS * 3; k++) g[k] = 0.0;
  double qn[NVARS];
  load_q10(q, n, qn);
  const int kbeg = m.adjp[n], kend = m.adjp[n + 1];
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    const int o = a.x & 0x7fffffff;
    const bool right = a.x < 0;
    double qo[NVARS], av[4];
    load_q10(q, o, qo);
    if (a.y < m.nedge) load_avec(m.ea, a.y, av);
    else load_avec(m.bea, a.y - m.nedge, av);
    const double area = av[3];
#pragma unroll
    for (int i = 0; i < NTERMS; i++) {
      const int v = (i < 6) ? i : i + 1;   // c_gradloc
      const double faceavg = 0.5 * (qn[v] + qo[v]);
#pragma unroll
      for (int j = 0; j < 3; j++) {
        if (right) g[3 * i + j] += -faceavg * av[j] * area;
        else g[3 * i + j] += faceavg * av[j] * area;
      }
    }
  }
  const double vol = m.vol[n];
#pragma unroll
  for (int kk = 0; kk < NTERMS * 3; kk++) g[kk] /= vol;
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    if (m.bctype[be] != PCFD_BC_SYMMETRY) continue;
    double av[4];
    load_avec(m.bea, be, av);
#pragma unroll
    for (int i = 0; i < NTERMS; i++) {
      const double dot = g[i * 3] * av[0] + g[i * 3 + 1] * av[1] + g[i * 3 + 2] * av[2];
#pragma unroll
      for (int j = 0; j < 3; j++) g[i * 3 + j] -= dot * av[j];
    }
  }
  double* out = qgrad + (size_t)n * NTERMS * 3;
#pragma unroll
  for (int kk = 0; kk < NTERMS * 3; kk++) out[kk] = g[kk];
}

// ======================================================================= limiter

// Limiter::Compute passes 1+2 (limiters.tcc:53-110): neighbour min/max (from ZERO,
// :62-63) and Barth / Venkatakrishnan limiting, both as gathers over the node's edges.
// Writes the UNCLAMPED limiter; pressure clip and clamp follow.  FIVE threads per node, one equation each: the
// limiter of a variable depends on that variable's data only.
// qmm != null: the neighbour minimum / maximum of pass 1 were taken by k_gradient_geo3 from the same q
__global__ void __launch_bounds__(160) k_limiter(DevMesh m, int type, double chi, const double* __restrict__ q,
                                                  const double* __restrict__ qgrad, double* __restrict__ lim,
                                                  const double* __restrict__ qmm) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = tid / 5;
  if (n >= m.nnode + m.gnode) return;
  const int j = tid - n * 5;
  double l = 1.0;
  if (n < m.nnode && (type == 1 || type == 2 || type == 3)) {
    double qmin = 0.0, qmax = 0.0;
    const int k0 = m.adjp[n], k1 = m.adjp[n + 1];
    if (qmm != nullptr) {
      qmin = qmm[(size_t)n * 10 + j];
      qmax = qmm[(size_t)n * 10 + 5 + j];
    } else {
      for (int k = k0; k < k1; k++) {
        const int2 a = m.adj[k];
        const int o = a.x & 0x7fffffff;
        if (a.y >= m.nedge && !is_ghost(m, o)) continue;
        const double qo = __ldg(q + (size_t)o * NVARS + j);
        qmax = eq::maxd(qmax, qo);
        qmin = eq::mind(qmin, qo);
      }
    }
    const double qn = __ldg(q + (size_t)n * NVARS + j);
    const double* gp = qgrad + (size_t)n * NTERMS * 3 + j * 3;
    const double g0 = gp[0], g1 = gp[1], g2 = gp[2];
    const double xn[3] = {m.xyz[3 * n], m.xyz[3 * n + 1], m.xyz[3 * n + 2]};
    for (int k = k0; k < k1; k++) {
      const int2 a = m.adj[k];
      const int o = a.x & 0x7fffffff;
      if (a.y >= m.nedge && !is_ghost(m, o)) continue;
      const double qo = __ldg(q + (size_t)o * NVARS + j);
      const double dQ = qo - qn;
      double dx[3];
#pragma unroll
      for (int d = 0; d < 3; d++) dx[d] = 0.5 * (__ldg(m.xyz + 3 * o + d) - xn[d]);
      // eq::extrapolate with a unit limiter
      const double corr = 0.5 * chi * dQ + (1.0 - chi) * (g0 * dx[0] + g1 * dx[1] + g2 * dx[2]);
      const double QH = qn + corr * 1.0;
      double t = 1.0;
      if (type == 3) {
        // Kernel_VenkatMod / Bkernel_VenkatMod (limiters.tcc:534-735), eps^2 = 6 pi V K^3 with K = 1.  Three spellings of
        // delta+ in the reference: left node of an interior edge (unset when QH == qn: DM = 0 makes the ratio 1 for any
        // finite value, 0 is used), right node, ghost half-edge (measured from the extrapolated value).
        const double DM = QH - qn;
        double DP = 0.0;
        if (a.y >= m.nedge) DP = (QH > qn) ? (qmax - QH) : (qmin - QH);
        else if (a.x < 0) DP = (QH > qn) ? (qmax - qn) : (qmin - qn);
        else if (QH > qn) DP = qmax - qn;
        else if (QH < qn) DP = qmin - qn;
        const double ep2 = (6.0 * 3.141592653589793 * m.vol[n]) * (1.0 * 1.0 * 1.0);
        t = (DP * DP + ep2 + 2.0 * DM * DP) / (DP * DP + 2.0 * DM * DM + DM * DP + ep2);
      } else {
        if (QH > qn) t = (qmax - qn) / (QH - qn);
        else if (QH < qn) t = (qmin - qn) / (QH - qn);
        t = limiter_fn(type, t);
      }
      l = eq::mind(l, t);
    }
  }
  lim[(size_t)n * 5 + j] = l;
}

// The same two passes with ONE thread per node carrying all five equations: adjacency, neighbour coordinates and the
// half edge vector are fetched once per visit instead of once per equation, and the five independent division chains of
// a visit give the FP64 pipe the instruction-level parallelism the five-thread form gets from its threads.
__global__ void __launch_bounds__(128, PCFD_LIM_MINB) k_limiter_node(DevMesh m, int type, double chi, const double* __restrict__ q,
                                                                      const double* __restrict__ qgrad, double* __restrict__ lim) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode + m.gnode) return;
  double l[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
  if (n < m.nnode && (type == 1 || type == 2 || type == 3)) {
    double qmin[5] = {0, 0, 0, 0, 0}, qmax[5] = {0, 0, 0, 0, 0};
    const int k0 = m.adjp[n], k1 = m.adjp[n + 1];
    for (int k = k0; k < k1; k++) {
      const int2 a = m.adj[k];
      const int o = a.x & 0x7fffffff;
      if (a.y >= m.nedge && !is_ghost(m, o)) continue;
      double qo[5];
      load_q5(q, o, qo);
#pragma unroll
      for (int j = 0; j < 5; j++) { qmax[j] = eq::maxd(qmax[j], qo[j]); qmin[j] = eq::mind(qmin[j], qo[j]); }
    }
    double qn[5], g[15];
    load_q5(q, n, qn);
    const double* gp = qgrad + (size_t)n * NTERMS * 3;
#pragma unroll
    for (int j = 0; j < 15; j++) g[j] = gp[j];
    const double xn[3] = {m.xyz[3 * n], m.xyz[3 * n + 1], m.xyz[3 * n + 2]};
    const double ep2 = (type == 3) ? (6.0 * 3.141592653589793 * m.vol[n]) * (1.0 * 1.0 * 1.0) : 0.0;
    for (int k = k0; k < k1; k++) {
      const int2 a = m.adj[k];
      const int o = a.x & 0x7fffffff;
      if (a.y >= m.nedge && !is_ghost(m, o)) continue;
      double qo[5], dx[3];
      load_q5(q, o, qo);
#pragma unroll
      for (int d = 0; d < 3; d++) dx[d] = 0.5 * (__ldg(m.xyz + 3 * o + d) - xn[d]);
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const double dQ = qo[j] - qn[j];
        const double corr = 0.5 * chi * dQ + (1.0 - chi) * (g[3 * j] * dx[0] + g[3 * j + 1] * dx[1] + g[3 * j + 2] * dx[2]);
        const double QH = qn[j] + corr * 1.0;
        double t = 1.0;
        if (type == 3) {   // see k_limiter
          const double DM = QH - qn[j];
          double DP = 0.0;
          if (a.y >= m.nedge) DP = (QH > qn[j]) ? (qmax[j] - QH) : (qmin[j] - QH);
          else if (a.x < 0) DP = (QH > qn[j]) ? (qmax[j] - qn[j]) : (qmin[j] - qn[j]);
          else if (QH > qn[j]) DP = qmax[j] - qn[j];
          else if (QH < qn[j]) DP = qmin[j] - qn[j];
          t = (DP * DP + ep2 + 2.0 * DM * DP) / (DP * DP + 2.0 * DM * DM + DM * DP + ep2);
        } else {
          if (QH > qn[j]) t = (qmax[j] - qn[j]) / (QH - qn[j]);
          else if (QH < qn[j]) t = (qmin[j] - qn[j]) / (QH - qn[j]);
          t = limiter_fn(type, t);
        }
        l[j] = eq::mind(l[j], t);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 5; j++) lim[(size_t)n * 5 + j] = l[j];
}

// Kernel_PressureClip (limiters.tcc:737-815) is sequential in the reference: an edge
// sees the clips of earlier edges.  Here tclip[n] = index of the edge that zeroes
// node n's limiter (INT_MAX: never).  k_clip_edges evaluates every edge under the
// node states implied by tclip (node zeroed at edge e iff tclip[node] < e) and
// records which sides it clips; k_clip_nodes takes the first clipping edge per node.
// The pair is iterated to the (unique) fixed point, which is the sequential result;
// with no clips anywhere that is one pass.
__global__ void __launch_bounds__(128) k_clip_edges(DevMesh m, double chi, double gamma, const double* __restrict__ q,
                                                     const double* __restrict__ qgrad, const double* __restrict__ lim,
                                                     const int* __restrict__ tclip, unsigned char* __restrict__ flag,
                                                     int* __restrict__ any) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const int2 lr = m.en[e];
  const int l = lr.x, r = lr.y;
  double qL[5], qR[5], limL[5], limR[5], dQ[5], dx[3], QL[5], QR[5], Qroe[5], gr[15];
  load_q5(q, l, qL);
  load_q5(q, r, qR);
  const bool zl = tclip[l] < e, zr = tclip[r] < e;
#pragma unroll
  for (int j = 0; j < 5; j++) { limL[j] = zl ? 0.0 : lim[(size_t)l * 5 + j]; limR[j] = zr ? 0.0 : lim[(size_t)r * 5 + j]; }
#pragma unroll
  for (int d = 0; d < 3; d++) dx[d] = 0.5 * (__ldg(m.xyz + 3 * r + d) - __ldg(m.xyz + 3 * l + d));
#pragma unroll
  for (int j = 0; j < 5; j++) dQ[j] = qR[j] - qL[j];
#pragma unroll
  for (int j = 0; j < 15; j++) gr[j] = __ldg(qgrad + (size_t)l * NTERMS * 3 + j);
  eq::extrapolate(chi, QL, qL, dQ, gr, dx, limL);
  bool cl = eq::bad_extrapolation(QL, gamma);
#pragma unroll
  for (int d = 0; d < 3; d++) dx[d] = -dx[d];
#pragma unroll
  for (int j = 0; j < 5; j++) dQ[j] = -dQ[j];
#pragma unroll
  for (int j = 0; j < 15; j++) gr[j] = __ldg(qgrad + (size_t)r * NTERMS * 3 + j);
  eq::extrapolate(chi, QR, qR, dQ, gr, dx, limR);
  bool cr = eq::bad_extrapolation(QR, gamma);
  eq::roe_variables(QL, QR, gamma, Qroe);
  if (eq::bad_extrapolation(Qroe, gamma)) cl = cr = true;
  const unsigned char f = (cl ? 1 : 0) | (cr ? 2 : 0);
  flag[e] = f;
  if (f) *any = 1;
}

__global__ void k_clip_nodes(DevMesh m, const unsigned char* __restrict__ flag, const int* __restrict__ told,
                             int* __restrict__ tnew, int* __restrict__ changed) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  int t = INT_MAX;
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y >= m.nedge) break;
    const unsigned char f = flag[a.y];
    if (f & ((a.x < 0) ? 2 : 1)) { t = a.y; break; }
  }
  tnew[n] = t;
  if (t != told[n]) *changed = 1;
}

// final clip + clamp of negatives (limiters.tcc:118-125)
__global__ void k_limiter_final(int ntot, int nnode, const int* __restrict__ tclip, double* __restrict__ lim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot * 5) return;
  const int n = i / 5;
  double v = lim[i];
  if (tclip != nullptr && n < nnode && tclip[n] != INT_MAX) v = 0.0;
  if (v < 0.0) v = 0.0;
  lim[i] = v;
}

// ====================================================================== residual
// Kernel_Inviscid_Flux (residual.tcc:192-296): MUSCL reconstruction + Roe flux, once per edge.
// DET = true is the fused fast path of the composite iterations: `lim` then holds the RAW limiter of k_limiter (before
// Kernel_PressureClip and the clamp of negatives, limiters.tcc:105-125).  The flux uses the clamped value, and the
// pressure-clip test (k_clip_edges) is evaluated on the way with the raw one -- same reconstruction, same Roe state,
// almost no extra work.  If no edge anywhere raises *any, clip(lim) == lim and this flux is final; otherwise the
// caller discards it and takes the ordered clip path.
#ifndef PCFD_FLUX_MINB
#define PCFD_FLUX_MINB 5   /* measured on B200: 1.33 -> 1.01 ms at 10 M cells (102 registers, 152 B of spills) */
#endif
// EIG: the edge's term of ComputeTimesteps (Kernel_Timestep, timestep.tcc:80-111: max eigenvalue of the averaged state
// times the face area) is evaluated on the way from the two node states already in registers and stored per edge; the
// residual gather then sums it per node in edge order and forms the time step -- the separate k_timestep pass (which
// evaluates every edge twice, once from each end) disappears from the explicit iteration.
template <bool DET, bool EIG>
__global__ void __launch_bounds__(128, PCFD_FLUX_MINB) k_flux_edges(DevMesh m, int sorder, double chi, double gamma,
                                                     const double* __restrict__ q, const double* __restrict__ qgrad,
                                                     const double* __restrict__ lim, double* __restrict__ flux,
                                                     int* __restrict__ any, double* __restrict__ eig) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const int2 lr = m.en[e];
  const int l = lr.x, r = lr.y;
  double av[4], QL[5], QR[5], f[5];
  load_avec(m.ea, e, av);
  load_q5(q, l, QL);
  load_q5(q, r, QR);
  if (EIG) {
    double Q[5];
#pragma unroll
    for (int j = 0; j < 5; j++) Q[j] = 0.5 * (QL[j] + QR[j]);
    eig[e] = eq::max_eigenvalue(Q, av, 0.0, gamma) * av[3];
  }
  bool neg = false;
  if (sorder > 1) {
    double dQ[5], dx[3], gr[15], lmL[5], lmR[5], qL[5], qR[5];
#pragma unroll
    for (int j = 0; j < 5; j++) { qL[j] = QL[j]; qR[j] = QR[j]; dQ[j] = qR[j] - qL[j]; }
#pragma unroll
    for (int d = 0; d < 3; d++) dx[d] = 0.5 * (__ldg(m.xyz + 3 * r + d) - __ldg(m.xyz + 3 * l + d));
    load5(lim + (size_t)l * 5, lmL);
    load5(lim + (size_t)r * 5, lmR);
    if (DET) {
      double cL[5], cR[5];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        neg = neg || (lmL[j] < 0.0) || (lmR[j] < 0.0);
        cL[j] = (lmL[j] < 0.0) ? 0.0 : lmL[j];
        cR[j] = (lmR[j] < 0.0) ? 0.0 : lmR[j];
      }
      if (neg) {   // rare: the clip test sees the unclamped limiter, the flux the clamped one
        double TL[5], TR[5], Troe[5], mdQ[5], mdx[3];
#pragma unroll
        for (int j = 0; j < 15; j++) gr[j] = __ldg(qgrad + (size_t)l * NTERMS * 3 + j);
        eq::extrapolate(chi, TL, qL, dQ, gr, dx, lmL);
#pragma unroll
        for (int j = 0; j < 5; j++) mdQ[j] = -dQ[j];
#pragma unroll
        for (int d = 0; d < 3; d++) mdx[d] = -dx[d];
#pragma unroll
        for (int j = 0; j < 15; j++) gr[j] = __ldg(qgrad + (size_t)r * NTERMS * 3 + j);
        eq::extrapolate(chi, TR, qR, mdQ, gr, mdx, lmR);
        eq::roe_variables(TL, TR, gamma, Troe);
        if (eq::bad_extrapolation(TL, gamma) || eq::bad_extrapolation(TR, gamma) || eq::bad_extrapolation(Troe, gamma)) *any = 1;
      }
#pragma unroll
      for (int j = 0; j < 5; j++) { lmL[j] = cL[j]; lmR[j] = cR[j]; }
    }
#pragma unroll
    for (int j = 0; j < 15; j++) gr[j] = __ldg(qgrad + (size_t)l * NTERMS * 3 + j);
    eq::extrapolate(chi, QL, qL, dQ, gr, dx, lmL);
#pragma unroll
    for (int j = 0; j < 5; j++) dQ[j] = -dQ[j];
#pragma unroll
    for (int d = 0; d < 3; d++) dx[d] = -dx[d];
#pragma unroll
    for (int j = 0; j < 15; j++) gr[j] = __ldg(qgrad + (size_t)r * NTERMS * 3 + j);
    eq::extrapolate(chi, QR, qR, dQ, gr, dx, lmR);
  }
  bool bad = false;
  eq::numerical_flux(QL, QR, av, 0.0, gamma, f, DET ? &bad : nullptr);
  if (DET && bad && !neg && sorder > 1) *any = 1;   // with a clamped component the raw-limiter test above decides
#pragma unroll
  for (int j = 0; j < 5; j++) flux[(size_t)e * 5 + j] = f[j];
}

// Bkernel_Inviscid_Flux (residual.tcc:299-387): boundary and ghost half-edges
__global__ void __launch_bounds__(128) k_flux_bedges(DevMesh m, int sorder, double chi, double gamma,
                                                      const double* __restrict__ q, const double* __restrict__ qgrad,
                                                      const double* __restrict__ lim, double* __restrict__ bflux) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nbedge + m.ngedge) return;
  const int2 lr = m.ben[e];
  const int l = lr.x, r = lr.y;
  double av[4], QL[5], QR[5], f[5];
  load_avec(m.bea, e, av);
  load_q5(q, l, QL);
  load_q5(q, r, QR);
  if (sorder > 1 && is_ghost(m, r)) {
    double dQ[5], dx[3], gr[15], lm[5], qL[5], qR[5];
#pragma unroll
    for (int j = 0; j < 5; j++) { qL[j] = QL[j]; qR[j] = QR[j]; dQ[j] = qR[j] - qL[j]; }
#pragma unroll
    for (int d = 0; d < 3; d++) dx[d] = 0.5 * (__ldg(m.xyz + 3 * r + d) - __ldg(m.xyz + 3 * l + d));
#pragma unroll
    for (int j = 0; j < 15; j++) gr[j] = __ldg(qgrad + (size_t)l * NTERMS * 3 + j);
    load5(lim + (size_t)l * 5, lm);
    eq::extrapolate(chi, QL, qL, dQ, gr, dx, lm);
#pragma unroll
    for (int j = 0; j < 5; j++) dQ[j] = -dQ[j];
#pragma unroll
    for (int d = 0; d < 3; d++) dx[d] = -dx[d];
#pragma unroll
    for (int j = 0; j < 15; j++) gr[j] = __ldg(qgrad + (size_t)r * NTERMS * 3 + j);
    load5(lim + (size_t)r * 5, lm);
    eq::extrapolate(chi, QR, qR, dQ, gr, dx, lm);
  }
  eq::numerical_flux(QL, QR, av, 0.0, gamma, f);
#pragma unroll
  for (int j = 0; j < 5; j++) bflux[(size_t)e * 5 + j] = f[j];
}

// DriverScatter (driver.tcc:274-306) turned into an ordered gather: b[n] accumulates
// +flux (n is the right node) / -flux (left node) in edge order, half-edges last,
// then the (zero) source term of residual.tcc:109-115.
// With VISC the viscous edge fluxes (Kernel_Viscous_Flux / Bkernel_Viscous_Flux, residual.tcc:388-562) follow in
// a second pass over the same list -- the reference runs its viscous Driver/Bdriver after the inviscid pair --
// and Bkernel_BC_Res_Modify (bc.tcc:905-1056 -> ModifyViscousWallResidual, compressible.tcc:1611-1631) zeroes the
// hard-set rows of no-slip wall nodes (wallflag: bit 0 = owns a NoSlip half-edge, bit 1 = an adiabatic one).
// EIG: also sums the per-edge time-step terms (k_flux_edges<., true>, k_eig_bedges) in the same order and writes
// dt = CFL V / sum (+ the Von Neumann limit), exactly what k_timestep computes (timestep.tcc:31-45).
template <bool VISC, bool EIG>
__global__ void __launch_bounds__(128) k_residual_gather(DevMesh m, const double* __restrict__ flux,
                                                          const double* __restrict__ bflux,
                                                          const double* __restrict__ vflux,
                                                          const double* __restrict__ bvflux,
                                                          const unsigned char* __restrict__ wallflag,
                                                          double* __restrict__ b, const double* __restrict__ eig,
                                                          const double* __restrict__ beig, double cfl,
                                                          const double* __restrict__ vnn23, double* __restrict__ dt) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double acc[5] = {0, 0, 0, 0, 0};
  double esum = 0.0;
  const int k0 = m.adjp[n], k1 = m.adjp[n + 1];
  for (int k = k0; k < k1; k++) {
    const int2 a = m.adj[k];
    const bool right = a.x < 0;
    const double* f = (a.y < m.nedge) ? flux + (size_t)a.y * 5 : bflux + (size_t)(a.y - m.nedge) * 5;
    double fv[5];
    load5(f, fv);
#pragma unroll
    for (int j = 0; j < 5; j++) acc[j] += right ? fv[j] : -fv[j];
    if (EIG) esum += (a.y < m.nedge) ? __ldg(eig + a.y) : __ldg(beig + (a.y - m.nedge));
  }
  if (EIG) {
    double d = cfl * (m.vol[n] / esum);
    if (vnn23 != nullptr && n >= 1) d = eq::mind(d, vnn23[n]);
    dt[n] = d;
  }
  if (VISC) {
    for (int k = k0; k < k1; k++) {
      const int2 a = m.adj[k];
      const bool right = a.x < 0;
      const double2* f = reinterpret_cast<const double2*>((a.y < m.nedge) ? vflux + (size_t)a.y * 4
                                                                           : bvflux + (size_t)(a.y - m.nedge) * 4);
      const double2 f01 = __ldg(f), f23 = __ldg(f + 1);
      acc[1] += right ? f01.x : -f01.x;
      acc[2] += right ? f01.y : -f01.y;
      acc[3] += right ? f23.x : -f23.x;
      acc[4] += right ? f23.y : -f23.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 5; j++) acc[j] = acc[j] + 0.0;
  if (VISC) {
    const unsigned char wf = wallflag[n];
    if (wf & 1) { acc[1] = acc[2] = acc[3] = acc[4] = 0.0; }
    if (wf & 2) acc[0] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < 5; j++) b[(size_t)n * 5 + j] = acc[j];
}

// Kernel_Viscous_Flux (residual.tcc:388-466): face gradient = average of the two nodal gradients plus the
// directional correction (dq - g.dx)/|dx|^2 dx, then CompressibleEqnSet::ViscousFlux.  Only the gradient rows the
// flux reads (T, u, v, w = terms 5..8) are formed.  One thread per edge; vflux[e] = flux[1..4] (flux[0] == 0).
__device__ __forceinline__ void face_gradient_Tuvw(const DevMesh& m, int sorder, const double* __restrict__ q,
                                                   const double* __restrict__ qgrad, int l, int r, double* g) {
  const double* gl = qgrad + (size_t)l * NTERMS * 3 + 15;
  const double* gr = qgrad + (size_t)r * NTERMS * 3 + 15;
#pragma unroll
  for (int i = 0; i < 12; i++) g[i] = 0.5 * (__ldg(gl + i) + __ldg(gr + i));
  if (sorder > 1) {
    double dx[3], s2 = 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      dx[d] = (__ldg(m.xyz + 3 * r + d) - __ldg(m.xyz + 3 * l + d));
      s2 += dx[d] * dx[d];
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int loc = (j == 0) ? 5 : 6 + j;   // c_gradloc[5..8] = 5, 7, 8, 9
      const double qdots = dx[0] * g[j * 3] + dx[1] * g[j * 3 + 1] + dx[2] * g[j * 3 + 2];
      const double dq = (__ldg(q + (size_t)r * NVARS + loc) - __ldg(q + (size_t)l * NVARS + loc) - qdots) / s2;
#pragma unroll
      for (int d = 0; d < 3; d++) g[j * 3 + d] += dq * dx[d];
    }
  }
}

__global__ void __launch_bounds__(128) k_vflux_edges(DevMesh m, int sorder, eq::ViscParams vp,
                                                      const double* __restrict__ q, const double* __restrict__ qgrad,
                                                      const double* __restrict__ mut, double* __restrict__ vflux) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const int2 lr = m.en[e];
  const int l = lr.x, r = lr.y;
  double av[4], qL[5], qR[5], Qavg[5], g[12], f[4];
  load_avec(m.ea, e, av);
  load_q5(q, l, qL);
  load_q5(q, r, qR);
#pragma unroll
  for (int i = 0; i < 5; i++) Qavg[i] = (qL[i] + qR[i]) / 2.0;
  const double T = vp.gamma * eq::pressure(Qavg, vp.gamma) / Qavg[0];
  const double tmut = 0.5 * (__ldg(mut + l) + __ldg(mut + r));
  face_gradient_Tuvw(m, sorder, q, qgrad, l, r, g);
  eq::viscous_flux(vp, Qavg, T, g, av, tmut, f);
  double2* out = reinterpret_cast<double2*>(vflux + (size_t)e * 4);
  out[0] = make_double2(f[0], f[1]);
  out[1] = make_double2(f[2], f[3]);
}

// Bkernel_Viscous_Flux (residual.tcc:468-562): ghost half-edges like interior edges, boundary half-edges with
// the wall node's own gradient and mut
__global__ void __launch_bounds__(128) k_vflux_bedges(DevMesh m, int sorder, eq::ViscParams vp,
                                                       const double* __restrict__ q, const double* __restrict__ qgrad,
                                                       const double* __restrict__ mut, double* __restrict__ bvflux) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nbedge + m.ngedge) return;
  const int2 lr = m.ben[e];
  const int l = lr.x, r = lr.y;
  double av[4], qL[5], qR[5], Qavg[5], g[12], f[4];
  load_avec(m.bea, e, av);
  load_q5(q, l, qL);
  load_q5(q, r, qR);
#pragma unroll
  for (int i = 0; i < 5; i++) Qavg[i] = 0.5 * (qL[i] + qR[i]);
  const double T = vp.gamma * eq::pressure(Qavg, vp.gamma) / Qavg[0];
  double tmut;
  if (is_ghost(m, r)) {
    tmut = (__ldg(mut + l) + __ldg(mut + r)) / 2.0;
    face_gradient_Tuvw(m, sorder, q, qgrad, l, r, g);
  } else {
    tmut = __ldg(mut + l);
    const double* gl = qgrad + (size_t)l * NTERMS * 3 + 15;
#pragma unroll
    for (int i = 0; i < 12; i++) g[i] = __ldg(gl + i);
  }
  eq::viscous_flux(vp, Qavg, T, g, av, tmut, f);
  double2* out = reinterpret_cast<double2*>(bvflux + (size_t)e * 4);
  out[0] = make_double2(f[0], f[1]);
  out[1] = make_double2(f[2], f[3]);
}

// ====================================================================== timestep
// ComputeTimesteps (timestep.tcc:7-49) + Kernel_Timestep/Bkernel_Timestep (:80-143)
// vnn23 (may be null) = VNN * pow(vol, 2/3), the Von Neumann limit of timestep.tcc:37-41, formed once on the host
// with the same libm pow as the reference; the reference applies it from node 1 on.
__global__ void __launch_bounds__(128) k_timestep(DevMesh m, double gamma, double cfl, const double* __restrict__ q,
                                                   const double* __restrict__ vnn23, double* __restrict__ dt) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double qn[5];
  load_q5(q, n, qn);
  double acc = 0.0;
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    const int o = a.x & 0x7fffffff;
    double qo[5], av[4], Q[5];
    load_q5(q, o, qo);
    if (a.y < m.nedge) load_avec(m.ea, a.y, av);
    else load_avec(m.bea, a.y - m.nedge, av);
    // 0.5*(qL + qR): addition commutes exactly, so the role does not matter
#pragma unroll
    for (int j = 0; j < 5; j++) Q[j] = (a.x < 0) ? 0.5 * (qo[j] + qn[j]) : 0.5 * (qn[j] + qo[j]);
    const double maxeig = eq::max_eigenvalue(Q, av, 0.0, gamma);
    acc += maxeig * av[3];
  }
  double d = cfl * (m.vol[n] / acc);
  if (vnn23 != nullptr && n >= 1) d = eq::mind(d, vnn23[n]);
  dt[n] = d;
}

// Bkernel_Timestep (timestep.tcc:114-143) per half-edge, for the fused form: launched BEFORE UpdateBCs, where
// ComputeTimesteps sits in the iteration, so that it sees the phantom / ghost states k_timestep would see
__global__ void __launch_bounds__(128) k_eig_bedges(DevMesh m, double gamma, const double* __restrict__ q,
                                                     double* __restrict__ beig) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nbedge + m.ngedge) return;
  const int2 lr = m.ben[e];
  double av[4], qL[5], qR[5], Q[5];
  load_avec(m.bea, e, av);
  load_q5(q, lr.x, qL);
  load_q5(q, lr.y, qR);
#pragma unroll
  for (int j = 0; j < 5; j++) Q[j] = 0.5 * (qL[j] + qR[j]);
  beig[e] = eq::max_eigenvalue(Q, av, 0.0, gamma) * av[3];
}

// deterministic min / sum-of-squares reductions (fixed grid, fixed tree)
template <int BLOCK>
__global__ void k_min_partial(const double* __restrict__ v, int n, double* __restrict__ part) {
  __shared__ double sh[BLOCK];
  double mval = INFINITY;
  for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += gridDim.x * BLOCK) mval = fmin(mval, v[i]);
  sh[threadIdx.x] = mval;
  __syncthreads();
  for (int s = BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
template <int BLOCK>
__global__ void k_min_final(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double sh[BLOCK];
  double mval = INFINITY;
  for (int i = threadIdx.x; i < n; i += BLOCK) mval = fmin(mval, part[i]);
  sh[threadIdx.x] = mval;
  __syncthreads();
  for (int s = BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}
// sum of squares of v[i*stride + c] for c < stride, plus the total: part[block*(stride+1) + {0..stride}]
template <int BLOCK, int STRIDE>
__global__ void k_sumsq_partial(const double* __restrict__ v, int nrows, double* __restrict__ part) {
  __shared__ double sh[BLOCK];
  double acc[STRIDE];
#pragma unroll
  for (int c = 0; c < STRIDE; c++) acc[c] = 0.0;
  for (int i = blockIdx.x * BLOCK + threadIdx.x; i < nrows; i += gridDim.x * BLOCK) {
#pragma unroll
    for (int c = 0; c < STRIDE; c++) { const double t = v[(size_t)i * STRIDE + c]; acc[c] += t * t; }
  }
  for (int c = 0; c < STRIDE; c++) {
    sh[threadIdx.x] = acc[c];
    __syncthreads();
    for (int s = BLOCK / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * STRIDE + c] = sh[0];
    __syncthreads();
  }
}
template <int BLOCK, int STRIDE>
__global__ void k_sumsq_final(const double* __restrict__ part, int nparts, double* __restrict__ out) {
  __shared__ double sh[BLOCK];
  double total = 0.0;
  for (int c = 0; c < STRIDE; c++) {
    double a = 0.0;
    for (int i = threadIdx.x; i < nparts; i += BLOCK) a += part[(size_t)i * STRIDE + c];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = BLOCK / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) { out[1 + c] = sh[0]; total += sh[0]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = total;
}

// ======================================================================== update
// ExplicitSolve (solve.tcc:71-98) and the ApplyDQ loop (solutionSpace.tcc:802-804)
__global__ void k_explicit(int nnode, double gamma, const double* __restrict__ b, const double* __restrict__ dt,
                           const double* __restrict__ vol, double* __restrict__ x, double* __restrict__ q) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnode) return;
  double dq[5], Q[NVARS];
  const double d = dt[n], v = vol[n];
#pragma unroll
  for (int j = 0; j < 5; j++) { dq[j] = b[(size_t)n * 5 + j] * d / v; x[(size_t)n * 5 + j] = dq[j]; }
  load_q10(q, n, Q);
  eq::apply_dq(dq, Q, gamma);
  store_q10(q, n, Q);
}
// NewtonIterate first zeroes the whole update of a node with a NaN / Inf component, in crs->x as well
// (solutionSpace.tcc:771-796); zeroed nodes are counted (pcfd_zeroed_updates)
__global__ void k_apply_dq(int nnode, double gamma, double* __restrict__ x, double* __restrict__ q, int* __restrict__ zeroed) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnode) return;
  double dq[5], Q[NVARS];
  load5(x + (size_t)n * 5, dq);
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 5; j++) bad = bad || !isfinite(dq[j]);
  if (bad) {
#pragma unroll
    for (int j = 0; j < 5; j++) { dq[j] = 0.0; x[(size_t)n * 5 + j] = 0.0; }
    atomicAdd(zeroed, 1);
  }
  load_q10(q, n, Q);
  eq::apply_dq(dq, Q, gamma);
  store_q10(q, n, Q);
}

// ============================================================ boundary conditions
// UpdateBCs (bc.tcc:1399-1457) -> BC_Kernel (:723-745).  One thread per local node
// that owns half-edges, walking them in half-edge order: two half-edges interact
// only through their shared left node, so this is the reference's sequence.
__global__ void __launch_bounds__(128) k_update_bcs(DevMesh m, eq::BcParams bp, const int* __restrict__ bnodes, int nb,
                                                     double* __restrict__ q) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb) return;
  const int n = bnodes[t];
  double QL[NVARS];
  load_q10(q, n, QL);
  bool touched = false;
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    const int type = m.bctype[be];
    if (type == PCFD_BC_PARALLEL) continue;
    const int r = a.x & 0x7fffffff;
    double QR[NVARS], av[4], nQ[NVARS];
    load_q10(q, r, QR);
    load_avec(m.bea, be, av);
    double tw = 0.0;
    if (type == PCFD_BC_NOSLIP) { load_q10(q, m.bnormal[be], nQ); tw = m.btwall[be]; }
    eq::boundary_variables(bp, QL, QR, av, type, nQ, tw, (type == PCFD_BC_FARFIELD_VISCOUS) ? m.bubar[be] : 1.0);
    store_q10(q, r, QR);
    touched = true;
  }
  if (touched) store_q10(q, n, QL);
}

// The same update with one thread per half-edge, for left nodes that own no Dirichlet-type
// half-edge (those rewrite QL itself and stay on the sequential kernel above).  The only
// coupling between two half-edges of a node is then ComputeAuxiliaryVariables(QL) at the end
// of each call (bc.tcc:1392-1396): the first half-edge sees the stored aux values, every later
// one sees them recomputed from QL[0..4] -- which a thread reproduces locally.
__global__ void __launch_bounds__(128) k_update_bcs_edges(DevMesh m, eq::BcParams bp, const int* __restrict__ list, int n,
                                                           const unsigned char* __restrict__ bfirst, double* q) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int be = list[t];
  const int2 lr = m.ben[be];
  const bool first = bfirst[be] != 0;
  double QL[NVARS], QR[NVARS], av[4];
  load_q10(q, lr.x, QL);
  load_q10(q, lr.y, QR);
  load_avec(m.bea, be, av);
  if (!first) eq::aux(QL, bp.gamma);
  const int type = m.bctype[be];
  eq::boundary_variables(bp, QL, QR, av, type, nullptr, 0.0, (type == PCFD_BC_FARFIELD_VISCOUS) ? m.bubar[be] : 1.0);
  store_q10(q, lr.y, QR);
  if (first) {   // only aux of QL can have changed
    double2* pq = reinterpret_cast<double2*>(q + (size_t)lr.x * NVARS);
    q[(size_t)lr.x * NVARS + 5] = QL[5];
    pq[3] = make_double2(QL[6], QL[7]);
    pq[4] = make_double2(QL[8], QL[9]);
  }
}

// ====================================================================== Jacobian
// Kernel_NumJac (jacobian.tcc:254-304): one-sided finite differences, h = 1e-8, of the
// FIRST-ORDER flux; writes A(l,r) = dF/dqR and A(r,l) = -dF/dqL into their slots.
template <int MINB = 2>
__global__ void __launch_bounds__(128, MINB) k_jac_edges(DevMesh m, double gamma, const double* __restrict__ q,
                                                    const int* __restrict__ posLR, const int* __restrict__ posRL,
                                                    double* __restrict__ A) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const double h = 1.0e-8;
  const int2 lr = m.en[e];
  double av[4], QL[5], QR[5], fS[5];
  load_avec(m.ea, e, av);
  load_q5(q, lr.x, QL);
  load_q5(q, lr.y, QR);
  eq::numerical_flux(QL, QR, av, 0.0, gamma, fS);
  double* pR = A + (size_t)posLR[e] * NEQN2;   // row l, column r
  double* pL = A + (size_t)posRL[e] * NEQN2;   // row r, column l
#pragma unroll 1
  for (int i = 0; i < 5; i++) {
    double QP[5], fL[5], fR[5];
#pragma unroll
    for (int j = 0; j < 5; j++) QP[j] = QL[j];
    QP[i] += h;
    eq::numerical_flux(QP, QR, av, 0.0, gamma, fL);
#pragma unroll
    for (int j = 0; j < 5; j++) QP[j] = QR[j];
    QP[i] += h;
    eq::numerical_flux(QL, QP, av, 0.0, gamma, fR);
#pragma unroll
    for (int j = 0; j < 5; j++) {
      pL[j * 5 + i] = 0.0 + (fS[j] - fL[j]) / h;   // "+=" onto the blanked matrix
      pR[j * 5 + i] = 0.0 + (fR[j] - fS[j]) / h;
    }
  }
}

// Kernel_NumJac_Complex (jacobian.tcc:370-433), Param::fieldJacType == 2: the first-order Roe flux on a complex state
// (eqnset_compressible_cs.cuh), one conservative variable perturbed by i * 1e-11; A(r,l) = -imag(F(qL + ih)) / h,
// A(l,r) = imag(F(qR + ih)) / h.  (The reference also recomputes the auxiliary variables of the perturbed state; the
// Roe flux does not read them.)
__global__ void __launch_bounds__(128) k_jac_edges_complex(DevMesh m, double gamma, const double* __restrict__ q,
                                                            const int* __restrict__ posLR, const int* __restrict__ posRL,
                                                            double* __restrict__ A) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const double h = 1.0e-11;
  const int2 lr = m.en[e];
  double av[4], qL[5], qR[5];
  load_avec(m.ea, e, av);
  load_q5(q, lr.x, qL);
  load_q5(q, lr.y, qR);
  eqcs::cplx QL[5], QR[5];
#pragma unroll
  for (int j = 0; j < 5; j++) { QL[j] = eqcs::cplx(qL[j]); QR[j] = eqcs::cplx(qR[j]); }
  double* pR = A + (size_t)posLR[e] * NEQN2;   // row l, column r
  double* pL = A + (size_t)posRL[e] * NEQN2;   // row r, column l
#pragma unroll 1
  for (int i = 0; i < 5; i++) {
    eqcs::cplx QP[5], fL[5], fR[5];
#pragma unroll
    for (int j = 0; j < 5; j++) QP[j] = QL[j];
    QP[i].im += h;
    eqcs::roe_flux(QP, QR, av, 0.0, gamma, fL);
#pragma unroll
    for (int j = 0; j < 5; j++) QP[j] = QR[j];
    QP[i].im += h;
    eqcs::roe_flux(QL, QP, av, 0.0, gamma, fR);
#pragma unroll
    for (int j = 0; j < 5; j++) {
      // EqnSet::NumericalFlux kneecaps a flux entry whose REAL part is NaN (eqnset.tcc:73-88)
      const double iL = isnan(fL[j].re) ? 0.0 : fL[j].im, iR = isnan(fR[j].re) ? 0.0 : fR[j].im;
      pL[j * 5 + i] = 0.0 + (-iL / h);   // "+=" onto the blanked matrix
      pR[j * 5 + i] = 0.0 + (iR / h);
    }
  }
}

// Kernel_NumJac_Centered (jacobian.tcc:306-366), Param::fieldJacType == 1: central differences, h = 1e-8, of the
// first-order flux; A(r,l) = (F(qL-h) - F(qL+h))/2h, A(l,r) = (F(qR+h) - F(qR-h))/2h.
__global__ void __launch_bounds__(128) k_jac_edges_central(DevMesh m, double gamma, const double* __restrict__ q,
                                                            const int* __restrict__ posLR, const int* __restrict__ posRL,
                                                            double* __restrict__ A) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const double h = 1.0e-8;
  const int2 lr = m.en[e];
  double av[4], QL[5], QR[5];
  load_avec(m.ea, e, av);
  load_q5(q, lr.x, QL);
  load_q5(q, lr.y, QR);
  double* pR = A + (size_t)posLR[e] * NEQN2;   // row l, column r
  double* pL = A + (size_t)posRL[e] * NEQN2;   // row r, column l
#pragma unroll 1
  for (int i = 0; i < 5; i++) {
    double QP[5], fL[5], fR[5], fLd[5], fRd[5];
#pragma unroll
    for (int j = 0; j < 5; j++) QP[j] = QL[j];
    QP[i] += h;
    eq::numerical_flux(QP, QR, av, 0.0, gamma, fL);
#pragma unroll
    for (int j = 0; j < 5; j++) QP[j] = QR[j];
    QP[i] += h;
    eq::numerical_flux(QL, QP, av, 0.0, gamma, fR);
#pragma unroll
    for (int j = 0; j < 5; j++) QP[j] = QL[j];
    QP[i] -= h;
    eq::numerical_flux(QP, QR, av, 0.0, gamma, fLd);
#pragma unroll
    for (int j = 0; j < 5; j++) QP[j] = QR[j];
    QP[i] -= h;
    eq::numerical_flux(QL, QP, av, 0.0, gamma, fRd);
#pragma unroll
    for (int j = 0; j < 5; j++) {
      pL[j * 5 + i] = 0.0 + (fLd[j] - fL[j]) / (2.0 * h);   // "+=" onto the blanked matrix
      pR[j * 5 + i] = 0.0 + (fR[j] - fRd[j]) / (2.0 * h);
    }
  }
}

// Kernel_Viscous_Jac (jacobian.tcc:768-800): analytic viscous blocks added onto the two off-diagonal blocks of
// the edge.  A separate pass AFTER the boundary Jacobian kernels, as in the reference (jacobian.tcc:183-193):
// Bkernel_NumJac updates the wall-node and phantom states this kernel reads.
__global__ void __launch_bounds__(128) k_vjac_edges(DevMesh m, eq::ViscParams vp, const double* __restrict__ q,
                                                     const double* __restrict__ mut, const int* __restrict__ posLR,
                                                     const int* __restrict__ posRL, double* A) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const int2 lr = m.en[e];
  double av[4], qL[NVARS], qR[NVARS], dx[3], s2 = 0.0;
  load_avec(m.ea, e, av);
  load_q10(q, lr.x, qL);
  load_q10(q, lr.y, qR);
#pragma unroll
  for (int d = 0; d < 3; d++) {
    dx[d] = (__ldg(m.xyz + 3 * lr.y + d) - __ldg(m.xyz + 3 * lr.x + d));
    s2 += dx[d] * dx[d];
  }
  const double tmut = (__ldg(mut + lr.x) + __ldg(mut + lr.y)) / 2.0;
  eq::viscous_jacobian(vp, qL, qR, dx, s2, av, tmut, A + (size_t)posRL[e] * NEQN2, A + (size_t)posLR[e] * NEQN2);
}

// Bkernel_BC_Jac_Modify (jacobian.tcc:247-249, bc.tcc:747-903) -> ModifyViscousWallJacobian
// (compressible.tcc:1576-1609) with CRSMatrix::BlankSubRow (crsmatrix.tcc:524-541).  One thread per wall node,
// its NoSlip half-edges in half-edge order (the reference re-applies the modification per half-edge).
__global__ void k_jac_wall(DevMesh m, double gamma, const int* __restrict__ wnodes, int nw, const int* __restrict__ ia,
                           const int* __restrict__ ja, const int* __restrict__ iau, double* A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nw) return;
  const int n = wnodes[t];
  const int r0 = ia[n], r1 = ia[n + 1];
  double* diag = A + (size_t)iau[n] * NEQN2;
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    if (m.bctype[be] != PCFD_BC_NOSLIP) continue;
    const double Twall = m.btwall[be];
    const bool adiabatic = Twall < 0.0;
    for (int sub = 0; sub < NEQN; sub++) {
      if (sub == 0 && !adiabatic) continue;
      for (int kk = r0; kk < r1; kk++)
        for (int j = 0; j < NEQN; j++) A[(size_t)kk * NEQN2 + sub * NEQN + j] = 0.0;
      diag[sub * NEQN + sub] = 1.0;
    }
    if (adiabatic) {
      const int nn_ = m.bnormal[be];
      for (int kk = r0; kk < r1; kk++)
        if (ja[kk] == nn_) {
          A[(size_t)kk * NEQN2 + 0] = -1.0;
          A[(size_t)kk * NEQN2 + 4 * NEQN + 4] = -1.0;
          break;
        }
    } else {
      const double v2 = 0.0;   // static wall
      diag[4 * NEQN + 0] = -(Twall / (gamma * (gamma - 1.0)) + 0.5 * v2);
    }
  }
}

// Bkernel_NumJac (jacobian.tcc:459-544), boundaryJacEval == 0, for ONE half-edge: updates the
// phantom state like the reference does, writes dF/dqL to bdiag[be] (summed into the node's
// diagonal block, in half-edge order, by k_jac_diag) and, for ghost half-edges, dF/dqR to A(l,ghost).
__device__ __forceinline__ void jac_half_edge(const DevMesh& m, const eq::BcParams& bp, int be, double* QL, double* q,
                                              const int* __restrict__ bpos, double* __restrict__ bdiag,
                                              double* __restrict__ A) {
  const double h = 1.0e-8;
  const double gamma = bp.gamma;
  const int type = m.bctype[be];
  const int r = m.ben[be].y;
  const bool ghost = is_ghost(m, r);
  double QR[NVARS], av[4], fS[5], nQ[NVARS];
  double tw = 0.0;
  load_q10(q, r, QR);
  load_avec(m.bea, be, av);
  if (type == PCFD_BC_NOSLIP) { load_q10(q, m.bnormal[be], nQ); tw = m.btwall[be]; }
  // viscous far field: ONE copy of the free stream for this half-edge, scaled in place by every evaluation below
  const bool ffv = type == PCFD_BC_FARFIELD_VISCOUS;
  const double ubar = ffv ? m.bubar[be] : 1.0;
  double Qref[NVARS];
  if (ffv) {
#pragma unroll
    for (int j = 0; j < NVARS; j++) Qref[j] = bp.qinf[j];
  }
  eq::boundary_variables(bp, QL, QR, av, type, nQ, tw, ubar, ffv ? Qref : nullptr);
  if (type != PCFD_BC_PARALLEL) store_q10(q, r, QR);
  eq::numerical_flux(QL, QR, av, 0.0, gamma, fS);
  double* pR = ghost ? A + (size_t)bpos[be] * NEQN2 : nullptr;
  double* bd = bdiag + (size_t)be * NEQN2;
#pragma unroll 1
  for (int i = 0; i < 5; i++) {
    double QPL[NVARS], QPR[NVARS], fL[5], fR[5];
#pragma unroll
    for (int j = 0; j < NVARS; j++) { QPL[j] = QL[j]; QPR[j] = QR[j]; }
    QPL[i] += h;
    QPR[i] += h;
    eq::aux(QPL, gamma);
    eq::aux(QPR, gamma);
    if (ghost) {
      eq::numerical_flux(QL, QPR, av, 0.0, gamma, fR);
      eq::numerical_flux(QPL, QR, av, 0.0, gamma, fL);
#pragma unroll
      for (int j = 0; j < 5; j++) pR[j * 5 + i] = 0.0 + (fR[j] - fS[j]) / h;
    } else {
#pragma unroll
      for (int j = 0; j < NVARS; j++) QPR[j] = QR[j];
      eq::aux(QPR, gamma);
      eq::boundary_variables(bp, QPL, QPR, av, type, nQ, tw, ubar, ffv ? Qref : nullptr);
      eq::numerical_flux(QPL, QPR, av, 0.0, gamma, fL);
    }
#pragma unroll
    for (int j = 0; j < 5; j++) bd[j * 5 + i] = (fL[j] - fS[j]) / h;
  }
}

// Bkernel_NumJac_Centered (jacobian.tcc:546-640), Param::boundaryJacType == 1, boundaryJacEval == 0, for ONE half-edge.
// The BC is re-evaluated for the +h state only: the -h branch tests boundaryJacEval without the negation (:604), so it
// takes the flux of the perturbed left state against the unperturbed phantom state.
__device__ __forceinline__ void jac_half_edge_central(const DevMesh& m, const eq::BcParams& bp, int be, double* QL, double* q,
                                                      const int* __restrict__ bpos, double* __restrict__ bdiag,
                                                      double* __restrict__ A) {
  const double h = 1.0e-8;
  const double gamma = bp.gamma;
  const int type = m.bctype[be];
  const int r = m.ben[be].y;
  const bool ghost = is_ghost(m, r);
  double QR[NVARS], av[4], nQ[NVARS];
  double tw = 0.0;
  load_q10(q, r, QR);
  load_avec(m.bea, be, av);
  if (type == PCFD_BC_NOSLIP) { load_q10(q, m.bnormal[be], nQ); tw = m.btwall[be]; }
  // viscous far field: ONE copy of the free stream for this half-edge, scaled in place by every evaluation below
  const bool ffv = type == PCFD_BC_FARFIELD_VISCOUS;
  const double ubar = ffv ? m.bubar[be] : 1.0;
  double Qref[NVARS];
  if (ffv) {
#pragma unroll
    for (int j = 0; j < NVARS; j++) Qref[j] = bp.qinf[j];
  }
  eq::boundary_variables(bp, QL, QR, av, type, nQ, tw, ubar, ffv ? Qref : nullptr);
  if (type != PCFD_BC_PARALLEL) store_q10(q, r, QR);
  double* pR = ghost ? A + (size_t)bpos[be] * NEQN2 : nullptr;
  double* bd = bdiag + (size_t)be * NEQN2;
#pragma unroll 1
  for (int i = 0; i < 5; i++) {
    double QPL[NVARS], QPR[NVARS], fL[5], fR[5], fLd[5], fRd[5];
#pragma unroll
    for (int j = 0; j < NVARS; j++) { QPL[j] = QL[j]; QPR[j] = QR[j]; }
    QPL[i] += h;
    QPR[i] += h;
    eq::aux(QPL, gamma);
    eq::aux(QPR, gamma);
    eq::numerical_flux(QL, QPR, av, 0.0, gamma, fR);
    if (ghost) {
      eq::numerical_flux(QPL, QR, av, 0.0, gamma, fL);
    } else {
#pragma unroll
      for (int j = 0; j < NVARS; j++) QPR[j] = QR[j];
      eq::aux(QPR, gamma);
      eq::boundary_variables(bp, QPL, QPR, av, type, nQ, tw, ubar, ffv ? Qref : nullptr);
      eq::numerical_flux(QPL, QPR, av, 0.0, gamma, fL);
    }
#pragma unroll
    for (int j = 0; j < NVARS; j++) { QPL[j] = QL[j]; QPR[j] = QR[j]; }
    QPL[i] -= h;
    QPR[i] -= h;
    eq::aux(QPL, gamma);
    eq::aux(QPR, gamma);
    eq::numerical_flux(QL, QPR, av, 0.0, gamma, fRd);
    eq::numerical_flux(QPL, QR, av, 0.0, gamma, fLd);
    if (ghost) {
#pragma unroll
      for (int j = 0; j < 5; j++) pR[j * 5 + i] = 0.0 + (fR[j] - fRd[j]) / (2.0 * h);
    }
#pragma unroll
    for (int j = 0; j < 5; j++) bd[j * 5 + i] = (fL[j] - fLd[j]) / (2.0 * h);
  }
}

__global__ void __launch_bounds__(64) k_jac_bnodes_central(DevMesh m, eq::BcParams bp, const int* __restrict__ bnodes, int nb,
                                                            double* q, const int* __restrict__ bpos, double* __restrict__ bdiag,
                                                            double* __restrict__ A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb) return;
  const int n = bnodes[t];
  double QL[NVARS];
  load_q10(q, n, QL);
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    jac_half_edge_central(m, bp, a.y - m.nedge, QL, q, bpos, bdiag, A);
  }
  store_q10(q, n, QL);
}

__global__ void __launch_bounds__(64) k_jac_bedges_central(DevMesh m, eq::BcParams bp, const int* __restrict__ list, int n,
                                                            const unsigned char* __restrict__ bfirst, double* q,
                                                            const int* __restrict__ bpos, double* __restrict__ bdiag,
                                                            double* __restrict__ A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int be = list[t];
  const int l = m.ben[be].x;
  const int type = m.bctype[be];
  const bool first = bfirst[be] != 0;
  double QL[NVARS];
  load_q10(q, l, QL);
  if (!first && type != PCFD_BC_PARALLEL) eq::aux(QL, bp.gamma);
  jac_half_edge_central(m, bp, be, QL, q, bpos, bdiag, A);
  if (first && type != PCFD_BC_PARALLEL) {
    double2* pq = reinterpret_cast<double2*>(q + (size_t)l * NVARS);
    q[(size_t)l * NVARS + 5] = QL[5];
    pq[3] = make_double2(QL[6], QL[7]);
    pq[4] = make_double2(QL[8], QL[9]);
  }
}

// nodes that own a Dirichlet-type half-edge: sequential over the node's half-edges
__global__ void __launch_bounds__(64) k_jac_bnodes(DevMesh m, eq::BcParams bp, const int* __restrict__ bnodes, int nb,
                                                    double* q, const int* __restrict__ bpos, double* __restrict__ bdiag,
                                                    double* __restrict__ A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb) return;
  const int n = bnodes[t];
  double QL[NVARS];
  load_q10(q, n, QL);
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    jac_half_edge(m, bp, a.y - m.nedge, QL, q, bpos, bdiag, A);
  }
  store_q10(q, n, QL);
}

// all other half-edges, one thread each (see k_update_bcs_edges for why this is the same sequence)
template <int MINB = 4>
__global__ void __launch_bounds__(64, MINB) k_jac_bedges(DevMesh m, eq::BcParams bp, const int* __restrict__ list, int n,
                                                    const unsigned char* __restrict__ bfirst, double* q,
                                                    const int* __restrict__ bpos, double* __restrict__ bdiag,
                                                    double* __restrict__ A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int be = list[t];
  const int l = m.ben[be].x;
  const int type = m.bctype[be];
  const bool first = bfirst[be] != 0;
  double QL[NVARS];
  load_q10(q, l, QL);
  if (!first && type != PCFD_BC_PARALLEL) eq::aux(QL, bp.gamma);
  jac_half_edge(m, bp, be, QL, q, bpos, bdiag, A);
  if (first && type != PCFD_BC_PARALLEL) {
    double2* pq = reinterpret_cast<double2*>(q + (size_t)l * NVARS);
    q[(size_t)l * NVARS + 5] = QL[5];
    pq[3] = make_double2(QL[6], QL[7]);
    pq[4] = make_double2(QL[8], QL[9]);
  }
}

// TemporalResidual (residual.tcc:125-179, no GCL): b -= cnp1 V/dt (Q - Q^n) and b -= cnm1 V/dt (Q^n - Q^{n-1}) on the
// conservative variables.  It runs between the spatial residual and Bkernel_BC_Res_Modify (residual.tcc:20-43), whose
// hard-set wall rows (already applied by the gather) are therefore re-applied here.
__global__ void __launch_bounds__(128) k_temporal_residual(int nnode, double cnp1, double cnm1, double tdt,
                                                            const double* __restrict__ vol, const double* __restrict__ q,
                                                            const double* __restrict__ qold, const double* __restrict__ qoldm1,
                                                            const unsigned char* __restrict__ wallflag, double* __restrict__ b) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnode) return;
  const double dt = cnp1 * vol[n] / tdt;
  const double dtm1 = cnm1 * vol[n] / tdt;
  double r[NEQN];
#pragma unroll
  for (int j = 0; j < NEQN; j++) {
    const double qo = qold[(size_t)n * NVARS + j];
    const double dq = q[(size_t)n * NVARS + j] - qo;
    const double dqm1 = qo - qoldm1[(size_t)n * NVARS + j];
    double v = b[(size_t)n * NEQN + j];
    v -= dt * dq;
    v -= dtm1 * dqm1;
    r[j] = v;
  }
  if (wallflag) {
    const unsigned char wf = wallflag[n];
    if (wf & 1) { r[1] = r[2] = r[3] = r[4] = 0.0; }
    if (wf & 2) r[0] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < NEQN; j++) b[(size_t)n * NEQN + j] = r[j];
}

// Kernel_Diag_NumJac (jacobian.tcc:434-456) + ContributeTemporalTerms (:214-250 ->
// eqnset.tcc:195-208): diag(n) -= A(other,n) in edge order, then
// += vol/dt on its diagonal.
__global__ void __launch_bounds__(128) k_jac_diag(DevMesh m, const int* __restrict__ iau, const int* __restrict__ posLR,
                                                   const int* __restrict__ posRL, const double* __restrict__ dt,
                                                   const double* __restrict__ bdiag, double cnp1, double tdt, double* A) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double* dg = A + (size_t)iau[n] * NEQN2;
  double d[NEQN2];
#pragma unroll
  for (int k = 0; k < NEQN2; k++) d[k] = 0.0;
  const int kbeg = m.adjp[n], kend = m.adjp[n + 1];
  // Bdriver runs before the diagonal pass: half-edge terms first, in half-edge order
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const double* src = bdiag + (size_t)(a.y - m.nedge) * NEQN2;
#pragma unroll
    for (int kk = 0; kk < NEQN2; kk++) d[kk] += __ldg(src + kk);
  }
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y >= m.nedge) break;
    // this node is the right node: the other row's block for column n is A(l,r) at posLR
    const int pos = (a.x < 0) ? posLR[a.y] : posRL[a.y];
    const double* src = A + (size_t)pos * NEQN2;
#pragma unroll
    for (int kk = 0; kk < NEQN2; kk++) d[kk] += -src[kk];
  }
  // tdt > 0: unsteady with local (pseudo) time stepping: cnp1 V/dt + V/dtau; else cnp1 V/dtau (eqnset.tcc:199-206)
  const double tt = (tdt > 0.0) ? cnp1 * m.vol[n] / tdt + m.vol[n] / dt[n] : cnp1 * m.vol[n] / dt[n];
#pragma unroll
  for (int kk = 0; kk < NEQN; kk++) d[kk * NEQN + kk] += tt;
#pragma unroll
  for (int k = 0; k < NEQN2; k++) dg[k] = d[k];
}

// PrepareSGS: k_lu_diag_lanes<NEQN> (pcfd_internal.cuh)

// ========================================================================== SGS
// One level of CRS::SGS (crs.tcc:90-145).  NEQN lanes cooperate on one row: lane i
// owns block-row i, accumulates rhs[i] -= (M_k x_k)[i] block after block in ja order
// (MatVecMult, matrix.h:63-74), the lanes exchange rhs by shuffle and each runs the
// permuted LuSolve (matrix.h:237-264) redundantly; lane i stores x[i].
template <int U>
__global__ void __launch_bounds__(128) k_sgs_level(const int* __restrict__ rows, int nrows, const int* __restrict__ ia,
                                                    const int* __restrict__ ja, const int* __restrict__ iau,
                                                    const double* __restrict__ A, const int* __restrict__ pv,
                                                    const double* __restrict__ b, double* x) {
  constexpr int RPW = 32 / NEQN;   // rows per warp
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int grp = lane / NEQN;
  const int i = lane - grp * NEQN;
  const int slot = warp * RPW + grp;
  const bool active = (grp < RPW) && (slot < nrows);
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const int row = rows[slot];
  const int k0 = __ldg(ia + row), k1 = __ldg(ia + row + 1);
  int p[NEQN];
#pragma unroll
  for (int j = 0; j < NEQN; j++) p[j] = __ldg(pv + (size_t)row * NEQN + j);
  double rhs = __ldg(b + (size_t)row * NEQN + i);
  int k = k0 + 1;
  // U blocks at a time: all the (streaming, read-once) matrix loads and the x gathers of the
  // group are issued before the first use, the arithmetic stays in ja order
  for (; k + U <= k1; k += U) {
    int col[U];
    double a[U][NEQN], xv[U][NEQN];
#pragma unroll
    for (int u = 0; u < U; u++) col[u] = __ldg(ja + k + u);
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < NEQN; j++) a[u][j] = __ldcs(A + (size_t)(k + u) * NEQN2 + i * NEQN + j);
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < NEQN; j++) xv[u][j] = x[(size_t)col[u] * NEQN + j];
#pragma unroll
    for (int u = 0; u < U; u++) {
      double v = a[u][0] * xv[u][0];
#pragma unroll
      for (int j = 1; j < NEQN; j++) v += a[u][j] * xv[u][j];
      rhs -= v;
    }
  }
  for (; k < k1; k++) {
    const double* a = A + (size_t)k * NEQN2 + i * NEQN;
    const double* xv = x + (size_t)__ldg(ja + k) * NEQN;
    double v = __ldcs(a) * xv[0];
#pragma unroll
    for (int j = 1; j < NEQN; j++) v += __ldcs(a + j) * xv[j];
    rhs -= v;
  }
  // gather the full right-hand side of this row into every lane of the group
  double bb[NEQN], xx[NEQN];
#pragma unroll
  for (int j = 0; j < NEQN; j++) bb[j] = __shfl_sync(mask, rhs, grp * NEQN + j);
  const double* d = A + (size_t)__ldg(iau + row) * NEQN2;
  // forward: x_r = b[p_r] - sum_{j<r} a[p_r][j] x_j
#pragma unroll
  for (int r = 0; r < NEQN; r++) {
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < r; j++) sum += d[p[r] * NEQN + j] * xx[j];
    double bp = bb[0];
#pragma unroll
    for (int j = 1; j < NEQN; j++) bp = (p[r] == j) ? bb[j] : bp;
    xx[r] = bp - sum;
  }
  // backward: b_r = (x_r - sum_{j>r, descending} a[p_r][j] b_j) / a[p_r][r]
#pragma unroll
  for (int r = NEQN - 1; r >= 0; r--) {
    double sum = 0.0;
#pragma unroll
    for (int j = NEQN - 1; j > r; j--) sum += d[p[r] * NEQN + j] * bb[j];
    bb[r] = (xx[r] - sum) / d[p[r] * NEQN + r];
  }
  double out = bb[0];
#pragma unroll
  for (int j = 1; j < NEQN; j++) out = (i == j) ? bb[j] : out;
  x[(size_t)row * NEQN + i] = out;
}

// ---- persistent variant: one warp per CTA walks tiles blockIdx.x, +gridDim.x, ... of the level through an
// S-stage ring of shared-memory buffers.  Lane 0 is the producer: it keeps S-1 bulk copies (cp.async.bulk, one
// mbarrier per stage, complete_tx bytes) in flight ahead of the tile being computed, so the shared memory of every
// resident CTA is in flight nearly all the time and the ia -> copy dependency is off the critical path; the x / ja /
// b / pv fetches of tile k+1 are issued before tile k is computed.  Arithmetic and its order are those of k_sgs_tile.
template <int LPR>
struct SgsRowState {
  static constexpr int R = (16 + LPR - 1) / LPR;
  int row, k0, nb, kb0;
  int p[NEQN];
  double rhs;
  double xr[R][NEQN];
  bool active;
};

template <int LPR, int S>
__global__ void __launch_bounds__(32) k_sgs_ring(int row0, int step, int nrows, const int* __restrict__ ia,
                                                  const int* __restrict__ ja, const double* __restrict__ A,
                                                  const int* __restrict__ pv, const double* __restrict__ b, double* x,
                                                  int stage_bytes) {
  constexpr int RPW = 32 / LPR;
  constexpr int R = SgsRowState<LPR>::R;
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);
  unsigned char* stage0 = smem + ((S * 8 + 15) & ~15);
  const int lane = threadIdx.x;
  const int grp = lane / LPR;
  const int t = lane - grp * LPR;
  const unsigned gmask = (LPR == 16 ? 0xffffu : ((1u << LPR) - 1u)) << ((grp * LPR) & 31);
  const int ntiles = (nrows + RPW - 1) / RPW;
  const int G = gridDim.x;
  const int K = (ntiles - (int)blockIdx.x + G - 1) / G;   // tiles of this CTA
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < S; st++) mbar_init(bars + st, 1);
  }
  __syncwarp();
  const unsigned long long pol = policy_evict_first();

  auto issue = [&](int k) {   // producer (lane 0): bulk copy of tile k into stage k % S
    const int s0 = ((int)blockIdx.x + k * G) * RPW;
    const int s1 = min(s0 + RPW, nrows);
    const int rfirst = row0 + step * s0, rlast = row0 + step * (s1 - 1);
    const int kb0 = __ldg(ia + min(rfirst, rlast)), kb1 = __ldg(ia + max(rfirst, rlast) + 1);
    const unsigned offA = (kb0 & 1) * 8u;
    const unsigned bytes = ((unsigned)(kb1 - kb0) * (NEQN2 * 8u) + offA + 15u) & ~15u;
    const int st = k % S;
    mbar_expect_tx(bars + st, bytes);
    bulk_g2s_hint(stage0 + (size_t)st * stage_bytes, reinterpret_cast<const unsigned char*>(A) + (size_t)kb0 * (NEQN2 * 8) - offA,
                  bytes, bars + st, pol);
  };
  auto fetch = [&](int k, SgsRowState<LPR>& rs) {   // every lane: row metadata + x gather of tile k
    const int s0 = ((int)blockIdx.x + k * G) * RPW;
    const int s1 = min(s0 + RPW, nrows);
    const int rfirst = row0 + step * s0, rlast = row0 + step * (s1 - 1);
    rs.kb0 = __ldg(ia + min(rfirst, rlast));
    const int slot = s0 + grp;
    rs.active = (grp < RPW) && (slot < s1);
    rs.row = 0; rs.k0 = 0; rs.nb = 0; rs.rhs = 0.0;
    if (rs.active) {
      rs.row = row0 + step * slot;
      rs.k0 = __ldg(ia + rs.row);
      rs.nb = __ldg(ia + rs.row + 1) - rs.k0 - 1;
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int u = r * LPR + t;
        if (u < rs.nb) {
          const double* xc = x + (size_t)__ldg(ja + rs.k0 + 1 + u) * NEQN;
#pragma unroll
          for (int j = 0; j < NEQN; j++) rs.xr[r][j] = xc[j];
        }
      }
#pragma unroll
      for (int j = 0; j < NEQN; j++) rs.p[j] = __ldg(pv + (size_t)rs.row * NEQN + j);
      if (t < NEQN) rs.rhs = __ldg(b + (size_t)rs.row * NEQN + t);
    }
  };

  if (lane == 0) {
    for (int k = 0; k < S - 1 && k < K; k++) issue(k);
  }
  SgsRowState<LPR> cur, nxt;
  if (K > 0) fetch(0, cur);
  for (int k = 0; k < K; k++) {
    if (lane == 0 && k + S - 1 < K) issue(k + S - 1);   // stage (k-1) % S was released at the end of iteration k-1
    if (k + 1 < K) fetch(k + 1, nxt);
    const int st = k % S;
    mbar_wait(bars + st, (unsigned)((k / S) & 1));
    if (cur.active) {
      const unsigned offA = (cur.kb0 & 1) * 8u;
      double* tA = reinterpret_cast<double*>(stage0 + (size_t)st * stage_bytes + offA) - (size_t)cur.kb0 * NEQN2;
      const int k0 = cur.k0, nb = cur.nb;
      for (int base = 0;;) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int u = base + r * LPR + t;
          if (u < nb) {
            double* m = tA + (size_t)(k0 + 1 + u) * NEQN2;
            double v[NEQN];
#pragma unroll
            for (int ii = 0; ii < NEQN; ii++) {
              double acc = m[ii * NEQN] * cur.xr[r][0];
#pragma unroll
              for (int j = 1; j < NEQN; j++) acc += m[ii * NEQN + j] * cur.xr[r][j];
              v[ii] = acc;
            }
#pragma unroll
            for (int ii = 0; ii < NEQN; ii++) m[ii] = v[ii];
          }
        }
        base += R * LPR;
        if (base >= nb) break;
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int u = base + r * LPR + t;
          if (u < nb) {
            const double* xc = x + (size_t)__ldg(ja + k0 + 1 + u) * NEQN;
#pragma unroll
            for (int j = 0; j < NEQN; j++) cur.xr[r][j] = xc[j];
          }
        }
      }
      __syncwarp(gmask);
      double rhs = cur.rhs;
      if (t < NEQN) {
        const double* vv = tA + (size_t)(k0 + 1) * NEQN2 + t;
        for (int u = 0; u < nb; u++) rhs -= vv[(size_t)u * NEQN2];
      }
      double bb[NEQN], xx[NEQN];
#pragma unroll
      for (int j = 0; j < NEQN; j++) bb[j] = __shfl_sync(gmask, rhs, grp * LPR + j);
      if (t < NEQN) {
        const double* d = tA + (size_t)k0 * NEQN2;
#pragma unroll
        for (int r = 0; r < NEQN; r++) {
          double sum = 0.0;
#pragma unroll
          for (int j = 0; j < r; j++) sum += d[cur.p[r] * NEQN + j] * xx[j];
          double bp = bb[0];
#pragma unroll
          for (int j = 1; j < NEQN; j++) bp = (cur.p[r] == j) ? bb[j] : bp;
          xx[r] = bp - sum;
        }
#pragma unroll
        for (int r = NEQN - 1; r >= 0; r--) {
          double sum = 0.0;
#pragma unroll
          for (int j = NEQN - 1; j > r; j--) sum += d[cur.p[r] * NEQN + j] * bb[j];
          bb[r] = (xx[r] - sum) / d[cur.p[r] * NEQN + r];
        }
        double out = bb[0];
#pragma unroll
        for (int j = 1; j < NEQN; j++) out = (t == j) ? bb[j] : out;
        x[(size_t)cur.row * NEQN + t] = out;
      }
    }
    // release the stage: generic-proxy accesses (reads and the parked v_u writes) before the next async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    cur = nxt;
  }
}

// PObj::UpdateGeneralVectors pack loop (parallel.tcc:829-846) with persistent send lists: row j of the
// send buffer is node list[j] of field v (width n doubles).  dst may be local staging memory or, for the
// direct-put exchange, a peer GPU's ghost segment mapped through CUDA IPC.
__global__ void k_halo_pack(const int* __restrict__ list, int count, int n, const double* __restrict__ v,
                            double* __restrict__ dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)count * n) return;
  const int j = (int)(t / n), cidx = (int)(t - (long long)j * n);
  dst[t] = v[(size_t)list[j] * n + cidx];
}

// ====================================================== Spalart-Allmaras turbulence model
// TurbulenceModel::Compute (turb.tcc:163-339) with Spalart (spalart.tcc:141-361): a segregated SCALAR system on the
// flow's CRS pattern (one double per block), first-order convection (turbulenceSpatialOrder = 1).  Same pattern as
// the flow: per-edge quantities once per edge into private slots, then ordered gathers per node.
namespace sa {
constexpr double kTinf = 1.341946;   // spalart.tcc:79
constexpr double sigma = 2.0 / 3.0, cb1 = 0.1355, cb2 = 0.622, kappa = 0.41, cw2 = 0.3, cw3 = 2.0, cv1 = 7.1;
constexpr double ct3 = 1.2, ct4 = 0.5;

// spalart.tcc:303-340 Diffusive
__device__ __forceinline__ void diffusive(double Re, double nu, const double* tg, double nutL, double nutR, const double* av,
                                          double dgrad, double* resL, double* resR, double* jacL, double* jacR) {
  const double nut = 0.5 * (nutL + nutR);
  const double gdot = tg[0] * av[0] + tg[1] * av[1] + tg[2] * av[2];
  const double area = av[3];
  const double Reinv = 1.0 / Re;
  const double c1 = (1.0 + cb2) * nut + nu;
  double rl = c1 * gdot - cb2 * nutL * gdot;
  rl *= 1.0 / sigma * area * Reinv;
  double rr = c1 * gdot - cb2 * nutR * gdot;
  rr *= 1.0 / sigma * area * Reinv;
  double jl = c1 * dgrad - cb2 * nutL * dgrad;
  jl *= 1.0 / sigma * area * Reinv;
  double jr = c1 * dgrad - cb2 * nutR * dgrad;
  jr *= 1.0 / sigma * area * Reinv;
  *resL = rl; *resR = rr; *jacL = jl; *jacR = jr;
}

// spalart.tcc:206-300 Source.  exp / pow are CUDA libm (<= 2 ulp), the reference's are glibc: parity to rounding.
__device__ __forceinline__ void source(double Re, double nu, double d, const double* vgrad, double nut, double vol,
                                       double* res, double* jac) {
  const double cw1 = cb1 / (kappa * kappa) + (1.0 + cb2) / sigma;
  const double cw36 = cw3 * cw3 * cw3 * cw3 * cw3 * cw3;
  const double cv13 = cv1 * cv1 * cv1;
  double chi = nut / nu;
  if (nu == 0.0) chi = 0.0;
  const double chi2 = chi * chi;
  const double chi3 = chi2 * chi;
  const double ft2 = ct3 * exp(-ct4 * chi2);
  const double d2 = d * d;
  const double uy = vgrad[1], uz = vgrad[2], vx = vgrad[3], vz = vgrad[5], wx = vgrad[6], wy = vgrad[7];
  const double o0 = wy - vz, o1 = uz - wx, o2 = vx - uy;
  const double Reinv = 1.0 / Re;
  const double fv1 = chi3 / (chi3 + cv13);
  const double fv2 = 1.0 - chi / (1.0 + chi * fv1);
  const double magw = sqrt(o0 * o0 + o1 * o1 + o2 * o2);
  double sv = magw + (nut / (kappa * kappa * d2)) * fv2 * Reinv;
  sv = eq::maxd(eq::maxd(sv, 0.3 * magw), 1.0e-12);
  const double r = eq::mind(10.0, nut / (sv * kappa * kappa * d2) * Reinv);
  double r6 = r * r * r;
  r6 = r6 * r6;
  const double g = r + cw2 * (r6 - r);
  double g6 = g * g * g;
  g6 = g6 * g6;
  const double fw = g * pow(((1.0 + cw36) / (g6 + cw36)), 1.0 / 6.0);
  const double pi = cb1 * (1.0 - ft2) * sv;
  const double prod = pi * nut;
  const double di = (cw1 * fw - cb1 * ft2 / (kappa * kappa)) * Reinv * (nut / (d2));
  const double dest = di * nut;
  const double dsdnut = fv2 * Reinv / (kappa * kappa * d2);
  const double dPdnut = pi * nut * dsdnut / sv;
  const double dDdnut = di;
  *res = (prod - dest) * vol;
  *jac = (eq::maxd(0.0, -(pi - di)) + eq::maxd(0.0, -(dPdnut - dDdnut))) * vol;
}

// spalart.tcc:343-359 ComputeEddyViscosity
__device__ __forceinline__ double eddy_viscosity(double rho, double nu, double nut) {
  const double cv13 = cv1 * cv1 * cv1;
  const double chi = nut / nu;
  if (nut <= 0.0) return 0.0;
  const double chi3 = chi * chi * chi;
  const double fv1 = chi3 / (chi3 + cv13);
  return rho * nut * fv1;
}
}  // namespace sa

// Spalart::BC_Kernel (spalart.tcc:141-170): one thread per node owning BC half-edges, in half-edge order (a
// no-slip half-edge zeroes the node's own value, which a later symmetry half-edge of the same node copies)
__global__ void k_turb_bcs(DevMesh m, const int* __restrict__ nodes, int nn_, double* tvar) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nn_) return;
  const int n = nodes[t];
  double tl = tvar[n];
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int type = m.bctype[a.y - m.nedge];
    const int r = a.x & 0x7fffffff;
    if (type == PCFD_BC_PARALLEL) continue;
    if (type == PCFD_BC_NOSLIP) { tvar[r] = 0.0; tl = 0.0; }
    else if (type == PCFD_BC_SYMMETRY || type == PCFD_BC_IMPERMEABLE_WALL) tvar[r] = tl;
    else tvar[r] = sa::kTinf;
  }
  tvar[n] = tl;
}

// unweighted LSQ gradient of the turbulence variable (turb.tcc:186-190 -> gradient.tcc:251-378 with Mesh::s,
// weight = 1) + symmetry fix (:545-565)
__global__ void __launch_bounds__(128) k_turb_gradient(DevMesh m, const double* __restrict__ tvar, const double* __restrict__ s,
                                                        double* __restrict__ tgrad) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double g[3] = {0.0, 0.0, 0.0}, sn[6];
#pragma unroll
  for (int k = 0; k < 6; k++) sn[k] = s[6 * (size_t)n + k];
  const double tn = tvar[n];
  const double xn[3] = {m.xyz[3 * n], m.xyz[3 * n + 1], m.xyz[3 * n + 2]};
  const int kbeg = m.adjp[n], kend = m.adjp[n + 1];
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    const int o = a.x & 0x7fffffff;
    const bool right = a.x < 0;
    if (a.y >= m.nedge && !is_ghost(m, o)) continue;
    const double to = __ldg(tvar + o);
    double dx[3], we[3];   // x_left - x_right, negated for the right node
#pragma unroll
    for (int d = 0; d < 3; d++) dx[d] = right ? (__ldg(m.xyz + 3 * o + d) - xn[d]) : (xn[d] - __ldg(m.xyz + 3 * o + d));
    if (right) { dx[0] = -dx[0]; dx[1] = -dx[1]; dx[2] = -dx[2]; }
    lsq_weights(sn, dx, we);
    const double dq = right ? 1.0 * (tn - to) : 1.0 * (to - tn);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (right) g[j] += +we[j] * dq;
      else g[j] += -we[j] * dq;
    }
  }
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    if (m.bctype[be] != PCFD_BC_SYMMETRY) continue;
    double av[4];
    load_avec(m.bea, be, av);
    const double dot = g[0] * av[0] + g[1] * av[1] + g[2] * av[2];
#pragma unroll
    for (int j = 0; j < 3; j++) g[j] -= dot * av[j];
  }
#pragma unroll
  for (int j = 0; j < 3; j++) tgrad[(size_t)n * 3 + j] = g[j];
}

// Kernel_Convective (turb.tcc:342-449) + Kernel_Diffusive (:563-650) for one interior edge.  slots[e] =
// {convective flux, diffusive resL, diffusive resR}; the two off-diagonal entries are final after this kernel.
// What the model asks of the eqnset (GetTheta, ComputeViscosity / GetDensity of the averaged state, GetRe, the place of the
// velocity gradient): evaluated here for the perfect gas (PROPS == false), read from the tables kfr_turb_props wrote
// for the reacting eqnset (PROPS == true).
struct TurbGas {
  double Re;                  // EqnSet::GetRe(): Re / Mach for the perfect gas (compressible.tcc), Param::Re otherwise (eqnset.h:137)
  int gstride, goff;          // qgrad row width and GetVelocityGradLocation()*3
  const double *pe, *pb, *pn; // {theta, nu} per edge, per half-edge; {rho, nu} per node (PROPS only)
};
template <bool PROPS>
__global__ void __launch_bounds__(128) k_turb_edges(DevMesh m, eq::ViscParams vp, TurbGas tg_, const double* __restrict__ q,
                                                     const double* __restrict__ tvar, const double* __restrict__ tgrad,
                                                     const int* __restrict__ posLR, const int* __restrict__ posRL,
                                                     double* __restrict__ slots, double* __restrict__ A) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const int2 lr = m.en[e];
  const int l = lr.x, r = lr.y;
  double av[4], theta, nu;
  load_avec(m.ea, e, av);
  if constexpr (PROPS) {
    theta = __ldg(tg_.pe + 2 * (size_t)e);
    nu = __ldg(tg_.pe + 2 * (size_t)e + 1);
  } else {
    double qL[5], qR[5], qa[5];
    load_q5(q, l, qL);
    load_q5(q, r, qR);
#pragma unroll
    for (int i = 0; i < 5; i++) qa[i] = 0.5 * (qL[i] + qR[i]);
    theta = eq::theta(qa, av, 0.0);
    const double T = vp.gamma * eq::pressure(qa, vp.gamma) / qa[0];
    nu = eq::viscosity(vp, T) / qa[0];
  }
  const double tL = __ldg(tvar + l), tR = __ldg(tvar + r);
  const double ta = theta * av[3];
  double aRL = 0.0, aLR = 0.0, conv;   // A(r,l), A(l,r)
  if (theta > 0.0) { aRL -= ta; conv = ta * tL; }
  else { aLR += ta; conv = ta * tR; }
  double de[3], ds2 = 0.0, tg[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    de[d] = __ldg(m.xyz + 3 * r + d) - __ldg(m.xyz + 3 * l + d);
    ds2 += de[d] * de[d];
  }
  const double dsum = de[0] * av[0] + de[1] * av[1] + de[2] * av[2];
  const double dgrad = dsum / ds2;
#pragma unroll
  for (int d = 0; d < 3; d++) tg[d] = 0.5 * (__ldg(tgrad + (size_t)l * 3 + d) + __ldg(tgrad + (size_t)r * 3 + d));
  double resL, resR, jacL, jacR;
  sa::diffusive(tg_.Re, nu, tg, tL, tR, av, dgrad, &resL, &resR, &jacL, &jacR);
  aRL -= jacL;
  aLR -= jacR;
  A[posRL[e]] = aRL;
  A[posLR[e]] = aLR;
  slots[(size_t)e * 3] = conv;
  slots[(size_t)e * 3 + 1] = resL;
  slots[(size_t)e * 3 + 2] = resR;
}

// Bkernel_Convective (turb.tcc:451-561) + Bkernel_Diffusive (:653-755) for one half-edge.  bslots[be] =
// {convective flux, diffusive resL, convective diagonal term, diffusive diagonal term}
template <bool PROPS>
__global__ void __launch_bounds__(128) k_turb_bedges(DevMesh m, eq::ViscParams vp, TurbGas tg_, const double* __restrict__ q,
                                                      const double* __restrict__ tvar, const double* __restrict__ tgrad,
                                                      const int* __restrict__ bpos, double* __restrict__ bslots,
                                                      double* __restrict__ A) {
  const int be = blockIdx.x * blockDim.x + threadIdx.x;
  if (be >= m.nbedge + m.ngedge) return;
  const int2 lr = m.ben[be];
  const int l = lr.x, r = lr.y;
  const bool ghost = is_ghost(m, r);
  double av[4], theta, nu;
  load_avec(m.bea, be, av);
  if constexpr (PROPS) {
    theta = __ldg(tg_.pb + 2 * (size_t)be);
    nu = __ldg(tg_.pb + 2 * (size_t)be + 1);
  } else {
    double qL[5], qR[5], qa[5];
    load_q5(q, l, qL);
    load_q5(q, r, qR);
#pragma unroll
    for (int i = 0; i < 5; i++) qa[i] = 0.5 * (qL[i] + qR[i]);
    theta = eq::theta(qa, av, 0.0);
    const double T = vp.gamma * eq::pressure(qa, vp.gamma) / qa[0];
    nu = eq::viscosity(vp, T) / qa[0];
  }
  const double tL = tvar[l], tR = tvar[r];
  const double ta = theta * av[3];
  double conv, dconv = 0.0, aLR = 0.0;
  if (theta > 0.0) { dconv = ta; conv = ta * tL; }
  else { if (ghost) aLR += ta; conv = ta * tR; }
  double tg[3], dgrad = 0.0;   // the reference leaves dgrad unset on physical boundaries (turb.tcc:669, 731)
  if (ghost) {
    double de[3], ds2 = 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      de[d] = (m.xyz[3 * r + d] - m.xyz[3 * l + d]);
      ds2 += de[d] * de[d];
    }
    const double dsum = de[0] * av[0] + de[1] * av[1] + de[2] * av[2];
    dgrad = dsum / ds2;
#pragma unroll
    for (int d = 0; d < 3; d++) tg[d] = 0.5 * (tgrad[(size_t)l * 3 + d] + tgrad[(size_t)r * 3 + d]);
    const double qdots = de[0] * tg[0] + de[1] * tg[1] + de[2] * tg[2];
    const double dq = (tR - tL - qdots) / ds2;
#pragma unroll
    for (int d = 0; d < 3; d++) tg[d] += dq * de[d];
  } else {
#pragma unroll
    for (int d = 0; d < 3; d++) tg[d] = tgrad[(size_t)l * 3 + d];
  }
  double resL, resR, jacL, jacR;
  sa::diffusive(tg_.Re, nu, tg, tL, tR, av, dgrad, &resL, &resR, &jacL, &jacR);
  double ddiff = 0.0;
  if (ghost) { aLR -= jacR; A[bpos[be]] = aLR; }
  else ddiff = jacL;
  bslots[(size_t)be * 4] = conv;
  bslots[(size_t)be * 4 + 1] = resL;
  bslots[(size_t)be * 4 + 2] = dconv;
  bslots[(size_t)be * 4 + 3] = ddiff;
}

// per node, in the reference's order: convective edges, convective half-edges, diffusive edges, diffusive half-edges,
// source (turb.tcc:207-233), Kernel_Diag_NumJac (:241-243), temporal term (:246-252)
template <bool PROPS>
__global__ void __launch_bounds__(128) k_turb_node(DevMesh m, eq::ViscParams vp, TurbGas tg_, const double* __restrict__ q,
                                                    const double* __restrict__ qgrad, const double* __restrict__ tvar,
                                                    const double* __restrict__ dist, const double* __restrict__ dt,
                                                    const double* __restrict__ slots, const double* __restrict__ bslots,
                                                    const int* __restrict__ iau, const int* __restrict__ posLR,
                                                    const int* __restrict__ posRL, double* __restrict__ b, double* A) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  const int k0 = m.adjp[n], k1 = m.adjp[n + 1];
  double res = 0.0, diag = 0.0;
  for (int k = k0; k < k1; k++) {   // Convective: Driver, then Bdriver
    const int2 a = m.adj[k];
    if (a.y < m.nedge) {
      const double f = __ldg(slots + (size_t)a.y * 3);
      res += (a.x < 0) ? f : -f;
    } else {
      const double* bs = bslots + (size_t)(a.y - m.nedge) * 4;
      res += -__ldg(bs);
      diag += __ldg(bs + 2);
    }
  }
  for (int k = k0; k < k1; k++) {   // DiffusiveDriver
    const int2 a = m.adj[k];
    if (a.y < m.nedge) {
      if (a.x < 0) res -= __ldg(slots + (size_t)a.y * 3 + 2);
      else res += __ldg(slots + (size_t)a.y * 3 + 1);
    } else {
      const double* bs = bslots + (size_t)(a.y - m.nedge) * 4;
      res += __ldg(bs + 1);
      diag += __ldg(bs + 3);
    }
  }
  const double d = dist[n];
  if (!(d < 1.0e-16)) {
    double nu;
    if constexpr (PROPS) {
      nu = __ldg(tg_.pn + 2 * (size_t)n + 1);
    } else {
      double Q[NVARS];
      load_q10(q, n, Q);
      nu = eq::viscosity(vp, Q[5]) / Q[0];
    }
    double vg[9], tres, tjac;
#pragma unroll
    for (int i = 0; i < 9; i++) vg[i] = __ldg(qgrad + (size_t)n * tg_.gstride + tg_.goff + i);   // GetVelocityGradLocation()*3
    sa::source(tg_.Re, nu, d, vg, tvar[n], m.vol[n], &tres, &tjac);
    res += tres;
    diag += tjac;
  }
  for (int k = k0; k < k1; k++) {   // Kernel_Diag_NumJac
    const int2 a = m.adj[k];
    if (a.y >= m.nedge) break;
    const int pos = (a.x < 0) ? posLR[a.y] : posRL[a.y];
    diag += -A[pos];
  }
  diag += 1.0 * m.vol[n] / dt[n];
  b[n] = res;
  A[iau[n]] = diag;
}

// Spalart::BC_Jac_Kernel (spalart.tcc:173-203) on no-slip nodes, then CRSMatrix::PrepareSGS for neqn == 1
// (crsmatrix.tcc:852-858: the diagonal is replaced by its inverse)
__global__ void k_turb_wall(const int* __restrict__ wnodes, int nw, const int* __restrict__ ia, const int* __restrict__ iau,
                            double* b, double* x, double* A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nw) return;
  const int n = wnodes[t];
  b[n] = 0.0;
  x[n] = 0.0;
  for (int k = ia[n]; k < ia[n + 1]; k++) A[k] = 0.0;
  A[iau[n]] = 1.0;
}
// ... and the clip of NaN / Inf entries of the right-hand side (turb.tcc:259-276: done when the residual norm is not
// finite, i.e. when such an entry exists; the norm itself is taken before, as there)
__global__ void k_turb_invdiag(int nnode, const int* __restrict__ iau, double* A, double* b) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnode) return;
  A[iau[n]] = 1.0 / A[iau[n]];
  if (!isfinite(b[n])) b[n] = 0.0;
}

// one level of CRS::SGS for neqn == 1 (crs.tcc:109-113): x = Dinv * (b - sum A_k x_k), ja order
__global__ void __launch_bounds__(128) k_sgs_scalar_level(const int* __restrict__ rows, int nrows, const int* __restrict__ ia,
                                                           const int* __restrict__ ja, const double* __restrict__ A,
                                                           const double* __restrict__ b, double* x) {
  // consecutive levels are chained by programmatic dependent launch (as the block sweeps, sgs_tile.cuh): the next
  // level may start its prologue (row, extents, right-hand side, inverse diagonal: nothing a level writes) while this
  // one drains; x -- the only thing a level inherits -- is read after griddepcontrol.wait
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  int row = 0, k0 = 0, k1 = 0;
  double rhs = 0.0, dinv = 0.0;
  if (t < nrows) {
    row = rows[t];
    k0 = ia[row]; k1 = ia[row + 1];
    rhs = b[row];
    dinv = A[k0];
  }
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (t >= nrows) return;
  for (int k = k0 + 1; k < k1; k++) {
    const double vout = __ldcs(A + k) * x[__ldg(ja + k)];
    rhs -= vout;
  }
  x[row] = dinv * rhs;
}

// The same sweeps as ONE cooperative launch: a level of the scalar system is ~0.1 M rows of ~15 entries -- 3 us of data
// behind ~20 us of launch latency and ramp -- and TurbulenceModel::Compute runs 2 x levels x nSgs of them (80 at 10 M
// cells).  Every block walks the levels of `nsweeps` forward + backward sweeps with a grid barrier in between; x is read
// through L2 (ld.global.cg): a value another SM wrote one level ago must not come from this SM's L1.  Row arithmetic and
// order are those of k_sgs_scalar_level.
__global__ void __launch_bounds__(256) k_sgs_scalar_sweeps(const int* __restrict__ rows_f, const int* __restrict__ rows_b,
                                                            const int* __restrict__ lev_f, int nlev_f,
                                                            const int* __restrict__ lev_b, int nlev_b, int nsweeps,
                                                            const int* __restrict__ ia, const int* __restrict__ ja,
                                                            const double* __restrict__ A, const double* __restrict__ b,
                                                            double* x) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int s = 0; s < nsweeps; s++) {
    for (int dir = 0; dir < 2; dir++) {
      const int* rows = dir ? rows_b : rows_f;
      const int* lev = dir ? lev_b : lev_f;
      const int nlev = dir ? nlev_b : nlev_f;
      for (int l = 0; l < nlev; l++) {
        const int r0 = lev[l], r1 = lev[l + 1];
        for (int t = r0 + tid; t < r1; t += nth) {
          const int row = rows[t];
          const int k0 = ia[row], k1 = ia[row + 1];
          double rhs = b[row];
          const double dinv = A[k0];
          for (int k = k0 + 1; k < k1; k++) {
            const double vout = __ldcs(A + k) * __ldcg(x + __ldg(ja + k));
            rhs -= vout;
          }
          x[row] = dinv * rhs;
        }
        grid.sync();
      }
    }
  }
}

// tvar += x with the clip at zero (turb.tcc:306-320), then mut = rho nu~ fv1 for local and ghost nodes (:324-336)
__global__ void k_turb_update(int nnode, double* __restrict__ x, double* tvar) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnode) return;
  if (!isfinite(x[n])) x[n] = 0.0;   // turb.tcc:283-301: NaN / Inf entries of the update are clipped to zero
  double t = tvar[n] + x[n];
  if (t < 0.0) t = 0.0;
  tvar[n] = t;
}
template <bool PROPS>
__global__ void k_turb_mut(int nn_, eq::ViscParams vp, TurbGas tg_, const double* __restrict__ q, const double* __restrict__ tvar,
                           double* __restrict__ mut) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn_) return;
  double rho, nu;
  if constexpr (PROPS) {
    rho = tg_.pn[2 * (size_t)n];
    nu = tg_.pn[2 * (size_t)n + 1];
  } else {
    rho = q[(size_t)n * NVARS];
    nu = eq::viscosity(vp, q[(size_t)n * NVARS + 5]) / rho;
  }
  mut[n] = sa::eddy_viscosity(rho, nu, tvar[n]);
}

__global__ void k_fill_int(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// p[i] = *src if src is given, else v
__global__ void k_fill_double(double* p, int n, double v, const double* __restrict__ src) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = src ? src[0] : v;
}

}  // namespace

std::string& pcfd_create_err() {
  static std::string e;
  return e;
}

namespace {

// level schedule of the sequential sweep: level[i] = 1 + max(level[j]) over the
// columns j of row i that the sweep has already updated (j < i forward, j > i backward)
void build_levels(int n, const int* ia, const int* ja, bool forward, std::vector<int>& rows, std::vector<int>& off) {
  std::vector<int> lev(n, 0);
  int nlev = 0;
  for (int kk = 0; kk < n; kk++) {
    const int i = forward ? kk : n - 1 - kk;
    int l = 0;
    for (int k = ia[i] + 1; k < ia[i + 1]; k++) {
      const int j = ja[k];
      if (j >= n) continue;
      if (forward ? (j < i) : (j > i)) l = std::max(l, lev[j] + 1);
    }
    lev[i] = l;
    nlev = std::max(nlev, l + 1);
  }
  off.assign(nlev + 1, 0);
  for (int i = 0; i < n; i++) off[lev[i] + 1]++;
  for (int l = 0; l < nlev; l++) off[l + 1] += off[l];
  rows.resize(n);
  std::vector<int> cur(off.begin(), off.end() - 1);
  // rows inside a level in sweep order (ascending forward, descending backward)
  for (int kk = 0; kk < n; kk++) {
    const int i = forward ? kk : n - 1 - kk;
    rows[cur[lev[i]]++] = i;
  }
}

// Alternative schedule for numberings that are already colour-sorted: cut 0..n-1 into maximal CONTIGUOUS index
// ranges that contain no two coupled rows.  Sweeping the ranges in order (rows of a range in parallel) is the
// sequential Gauss-Seidel sweep exactly, forward and -- with the ranges reversed -- backward, and every range is one
// contiguous slab of A, which is what the bulk-copy kernel streams.  Returns the range offsets.
std::vector<int> contiguous_ranges(int n, const int* ia, const int* ja) {
  std::vector<int> dep(n, -1), off;
  int start = 0;
  off.push_back(0);
  for (int i = 0; i < n; i++) {
    bool conflict = dep[i] >= start;
    for (int k = ia[i] + 1; k < ia[i + 1]; k++) {
      const int j = ja[k];
      if (j >= n) continue;
      if (j < i) conflict = conflict || j >= start;
      else dep[j] = std::max(dep[j], i);
    }
    if (conflict) { off.push_back(i); start = i; }
  }
  off.push_back(n);
  return off;
}

}  // namespace

#include "pcfd_comm.cuh"
#include "pcfd_gmres.cuh"
#include "pcfd_forces.cuh"

// ======================================================================= C ABI
extern "C" {

int pcfd_abi_version(void) { return PCFD_ABI_VERSION; }

const char* pcfd_last_error(const pcfd_ctx* ctx) { return ctx ? ctx->err.c_str() : pcfd_create_err().c_str(); }

int pcfd_destroy(pcfd_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->fr) pcfd_fr_destroy(c);
  forces_free(c);
  if (c->comm) {
    pcfd_comm_disconnect(c);
    if (c->comm->hgout) cudaFreeHost(c->comm->hgout);
    if (c->comm->ev) cudaEventDestroy(c->comm->ev);
    delete c->comm;
    c->comm = nullptr;
  }
  for (void* p : c->allocs) cudaFree(p);
  if (c->gm_buf) cudaFree(c->gm_buf);
  if (c->gm_pv) cudaFree(c->gm_pv);
  if (c->gm_n) cudaFree(c->gm_n);
  if (c->gm_npv) cudaFree(c->gm_npv);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->hflag) cudaFreeHost(c->hflag);
  if (c->ev_flag) cudaEventDestroy(c->ev_flag);
  delete c;
  return 0;
}

// shared by pcfd_create (perfect gas: 5 / 10 / 9) and pcfd_create_fr (reacting: ns+4 / 3ns+6 / 2ns+4, pcfd_fr.cu)
static int create_impl(const pcfd_mesh_desc* mesh, const pcfd_params* params, int device, int neqn, int nvars, int nterms,
                       pcfd_ctx** out) {
  pcfd_ctx* c = nullptr;
  if (!mesh || !params || !out) return fail(c, "pcfd_create: null argument");
  *out = nullptr;
  // param.tcc:401-404: the NS eqnset ids switch the viscous terms on
  const bool viscous = params->eqnset == PCFD_EQNSET_COMPRESSIBLE_NS;
  if (viscous && !(params->Re > 0.0 && params->Pr > 0.0 && params->PrT > 0.0 && params->tref > 0.0 && params->mach > 0.0))
    return fail(c, "pcfd_create: compressibleNS needs positive Re, Pr, PrT, tref and mach");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(c, "pcfd_create: no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(c, "pcfd_create: bad device ordinal");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(c, "pcfd_create: cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(c, std::string("pcfd_create: built for sm_100a only, device is ") + prop.name);

  c = new pcfd_ctx();
  c->neqn = neqn; c->nvars = nvars; c->nterms = nterms;
  struct Guard { pcfd_ctx* c; bool ok = false; ~Guard() { if (!ok) { pcfd_create_err() = c->err; pcfd_destroy(c); } } } guard{c};
  c->device = device;
  if (const char* e = getenv("PCFD_SGS_UNROLL")) c->sgs_unroll = atoi(e);
  if (const char* e = getenv("PCFD_SGS_TILE_WARPS")) c->sgs_tile_warps = atoi(e);
  if (c->sgs_tile_warps != 0 && c->sgs_tile_warps != 1 && c->sgs_tile_warps != 2 && c->sgs_tile_warps != 4) c->sgs_tile_warps = 4;
  if (const char* e = getenv("PCFD_SGS_PREFETCH_TILES")) c->sgs_pf_dist = atoi(e);
  if (const char* e = getenv("PCFD_SGS_PDL")) c->sgs_pdl = atoi(e) != 0;
  if (const char* e = getenv("PCFD_SGS_RING_STAGES")) c->sgs_ring_stages = atoi(e);
  if (const char* e = getenv("PCFD_SGS_RING_CTAS")) c->sgs_ring_ctas_per_sm = atoi(e);
  if (c->sgs_ring_stages > 0) c->sgs_tile_warps = 1;   // ring tiles are one warp's rows
  c->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("PCFD_SGS_TILE_LPR")) c->sgs_tile_lpr = atoi(e);
  if (c->sgs_tile_lpr != 5 && c->sgs_tile_lpr != 10 && c->sgs_tile_lpr != 16) c->sgs_tile_lpr = 16;
  if (neqn != NEQN) {   // wider blocks: the ring variant is 5x5 only; a row needs >= neqn lanes; fewer rows per tile
    c->sgs_ring_stages = 0;
    c->sgs_tile_lpr = 16;
    if (neqn > 16) c->sgs_tile_warps = 0;
    else if (!getenv("PCFD_SGS_TILE_WARPS")) c->sgs_tile_warps = 2;
  }
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  c->prm = *params;
  c->nnode = mesh->nnode; c->gnode = mesh->gnode; c->nbnode = mesh->nbnode;
  c->nedge = mesh->nedge; c->nbedge = mesh->nbedge; c->ngedge = mesh->ngedge;
  c->nb = c->nbedge + c->ngedge;
  c->nn = c->nnode + c->gnode;
  c->ntot = c->nn + c->nbnode;
  const int nnode = c->nnode, nedge = c->nedge, nb = c->nb;
  if (nnode <= 0) return fail(c, "pcfd_create: empty mesh");

  // ---- validate connectivity
  for (int e = 0; e < nedge; e++) {
    const int l = mesh->edges_n[2 * e], r = mesh->edges_n[2 * e + 1];
    if (l < 0 || r < 0 || l >= nnode || r >= nnode) return fail(c, "pcfd_create: interior edge references a non-local node");
  }
  for (int e = 0; e < nb; e++) {
    const int l = mesh->bedges_n[2 * e], r = mesh->bedges_n[2 * e + 1];
    if (l < 0 || l >= nnode || r < nnode || r >= c->ntot) return fail(c, "pcfd_create: bad half-edge nodes");
  }

  // ---- node -> edge gather lists, in edge order (interior edges, then half-edges)
  std::vector<int> adjp(nnode + 1, 0);
  for (int e = 0; e < nedge; e++) { adjp[mesh->edges_n[2 * e] + 1]++; adjp[mesh->edges_n[2 * e + 1] + 1]++; }
  for (int e = 0; e < nb; e++) adjp[mesh->bedges_n[2 * e] + 1]++;
  for (int i = 0; i < nnode; i++) adjp[i + 1] += adjp[i];
  std::vector<int2> adj(adjp[nnode]);
  {
    std::vector<int> cur(adjp.begin(), adjp.end() - 1);
    for (int e = 0; e < nedge; e++) {
      const int l = mesh->edges_n[2 * e], r = mesh->edges_n[2 * e + 1];
      adj[cur[l]++] = make_int2(r, e);
      adj[cur[r]++] = make_int2(l | (int)0x80000000, e);
    }
    for (int e = 0; e < nb; e++) {
      const int l = mesh->bedges_n[2 * e], r = mesh->bedges_n[2 * e + 1];
      adj[cur[l]++] = make_int2(r, nedge + e);
    }
  }
  std::vector<int> bnodes, blist;
  std::vector<unsigned char> bfirst(std::max(nb, 1), 0);
  {
    std::vector<char> seq(nnode, 0), seen(nnode, 0);
    for (int e = 0; e < nb; e++) {
      const int t = mesh->bedges_bctype[e];
      if (t == PCFD_BC_DIRICHLET || t == PCFD_BC_SONIC_INFLOW || t == PCFD_BC_NOSLIP) seq[mesh->bedges_n[2 * e]] = 1;
    }
    for (int i = 0; i < nnode; i++) if (seq[i]) bnodes.push_back(i);
    for (int e = 0; e < nb; e++) {      // BC half-edges (everything that is not a parallel boundary)
      const int l = mesh->bedges_n[2 * e];
      if (mesh->bedges_bctype[e] == PCFD_BC_PARALLEL) continue;
      if (!seen[l]) { seen[l] = 1; bfirst[e] = 1; }
      if (!seq[l]) blist.push_back(e);
    }
    c->nblist_bc = (int)blist.size();
    for (int e = 0; e < nb; e++) {      // ghost half-edges
      if (mesh->bedges_bctype[e] != PCFD_BC_PARALLEL) continue;
      if (!seq[mesh->bedges_n[2 * e]]) blist.push_back(e);
    }
    c->nblist = (int)blist.size();
  }
  c->nbn = (int)bnodes.size();

  // ---- no-slip walls: the "most normal node off the wall" search of bc.tcc:1182-1206 is pure geometry, so it is
  // done once here (same arithmetic); wall temperature per half-edge; per-node flags for the residual hook
  std::vector<int> bnormal(std::max(c->nbedge, 1), -1), wnodes;
  std::vector<double> btwall(std::max(c->nbedge, 1), 1.0 / (params->tref > 0.0 ? params->tref : 1.0));
  std::vector<unsigned char> wallflag(nnode, 0);
  for (int e = 0; e < c->nbedge; e++) {
    if (mesh->bedges_twall) btwall[e] = mesh->bedges_twall[e];
    if (mesh->bedges_bctype[e] != PCFD_BC_NOSLIP) continue;
    const int left = mesh->bedges_n[2 * e];
    const double* av = mesh->bedges_a + 4 * (size_t)e;
    const double* wx = mesh->xyz + 3 * (size_t)left;
    double dotmax = 0.0;
    int nn_ = -1;
    for (int p = mesh->ipsp[left]; p < mesh->ipsp[left + 1]; p++) {
      const int pt = mesh->psp[p];
      const double* px = mesh->xyz + 3 * (size_t)pt;
      double dx[3] = {px[0] - wx[0], px[1] - wx[1], px[2] - wx[2]};
      const double mag = std::sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
      double dot = 0.0;
      for (int i = 0; i < 3; i++) { dx[i] = dx[i] / mag; dot -= dx[i] * av[i]; }
      if (dot >= dotmax) { nn_ = pt; dotmax = dot; }
    }
    if (nn_ < 0) return fail(c, "pcfd_create: no-slip wall node without a neighbour off the wall");
    bnormal[e] = nn_;
    if (!wallflag[left]) wnodes.push_back(left);
    wallflag[left] |= 1;
    if (btwall[e] < 0.0) wallflag[left] |= 2;
  }
  std::vector<double> bubar(std::max(c->nbedge, 1), 1.0);
  for (int e = 0; e < c->nbedge; e++) {
    if (mesh->bedges_bctype[e] != PCFD_BC_FARFIELD_VISCOUS) continue;
    if (!viscous && params->eqnset != PCFD_EQNSET_COMPRESSIBLE_NS_FR)
      return fail(c, "pcfd_create: the viscous far-field BC needs a viscous eqnset (Re and the wall distance)");
    c->ffv_edges.push_back(e);
    c->ffv_left.push_back(mesh->bedges_n[2 * e]);
  }
  std::sort(wnodes.begin(), wnodes.end());
  for (int e = 0; e < c->nbedge; e++) {
    // an adiabatic wall copies rho and rho*E from the normal node (compressible.tcc:1548-1554): in the
    // reference's sequential BC loop that is order-dependent only if the normal node is itself hard-set
    if (bnormal[e] >= 0 && btwall[e] < 0.0 && bnormal[e] < nnode && wallflag[bnormal[e]])
      return fail(c, "pcfd_create: adiabatic no-slip wall whose most-normal neighbour is itself a wall node is not supported");
  }
  c->nwall = (int)wnodes.size();
  std::vector<int> tbnodes;   // nodes owning at least one BC (non-parallel) half-edge: Spalart::BC_Kernel walks them
  {
    std::vector<char> has(nnode, 0);
    for (int e = 0; e < nb; e++)
      if (mesh->bedges_bctype[e] != PCFD_BC_PARALLEL) has[mesh->bedges_n[2 * e]] = 1;
    for (int i = 0; i < nnode; i++) if (has[i]) tbnodes.push_back(i);
  }
  c->ntbnodes = (int)tbnodes.size();
  if (params->turb_model != 0 && params->turb_model != 1) return fail(c, "pcfd_create: unknown turbulence model");
  if (params->turb_model == 1 && !viscous && params->eqnset != PCFD_EQNSET_COMPRESSIBLE_NS_FR)
    return fail(c, "pcfd_create: Spalart-Allmaras needs a viscous eqnset (compressibleNS or compressibleNSFR)");
  c->viscous = viscous;
  c->vp = eq::ViscParams{params->gamma, params->Re, params->Pr, params->PrT, params->tref, params->mach};
  std::vector<double> vnn23;
  if (params->enable_vnn) {   // timestep.tcc:37-41
    vnn23.resize(nnode);
    for (int i = 0; i < nnode; i++) vnn23[i] = params->vnn * std::pow(mesh->vol[i], 2.0 / 3.0);
  }

  // ---- block-CRS pattern: CRSMatrix::Init (crsmatrix.tcc:48-97), diagonal first, then psp order
  std::vector<int> ia(nnode + 1), iau(nnode);
  ia[0] = 0;
  for (int i = 0; i < nnode; i++) ia[i + 1] = ia[i] + (mesh->ipsp[i + 1] - mesh->ipsp[i]) + 1;
  c->nblocks = ia[nnode];
  std::vector<int> ja(c->nblocks);
  for (int i = 0; i < nnode; i++) {
    int k = ia[i];
    iau[i] = k;
    ja[k++] = i;
    for (int p = mesh->ipsp[i]; p < mesh->ipsp[i + 1]; p++) ja[k++] = mesh->psp[p];
  }
  auto find = [&](int row, int col) {   // CRSMatrix::GetPointer (crsmatrix.tcc:740-781), once at setup
    for (int k = ia[row]; k < ia[row + 1]; k++) if (ja[k] == col) return k;
    return -1;
  };
  std::vector<int> posLR(nedge), posRL(nedge), bpos(nb, -1);
  for (int e = 0; e < nedge; e++) {
    const int l = mesh->edges_n[2 * e], r = mesh->edges_n[2 * e + 1];
    posLR[e] = find(l, r);
    posRL[e] = find(r, l);
    if (posLR[e] < 0 || posRL[e] < 0) return fail(c, "pcfd_create: edge without a matching psp entry");
  }
  for (int e = c->nbedge; e < nb; e++) {
    bpos[e] = find(mesh->bedges_n[2 * e], mesh->bedges_n[2 * e + 1]);
    if (bpos[e] < 0) return fail(c, "pcfd_create: ghost half-edge without a matching psp entry");
  }
  std::vector<int> rows_f, rows_b;
  {
    const std::vector<int> rng = contiguous_ranges(nnode, ia.data(), ja.data());
    const int nr = (int)rng.size() - 1;
    if (nr <= 64 && nnode / nr >= 1024 && !getenv("PCFD_SGS_LEVEL_SCHEDULE")) {
      // colour-sorted numbering: the ranges are the colours
      c->lev_f = rng;
      rows_f.resize(nnode);
      rows_b.resize(nnode);
      for (int i = 0; i < nnode; i++) { rows_f[i] = i; rows_b[i] = nnode - 1 - i; }
      c->lev_b.assign(nr + 1, 0);
      for (int l = 0; l <= nr; l++) c->lev_b[l] = nnode - rng[nr - l];
    } else {
      build_levels(nnode, ia.data(), ja.data(), true, rows_f, c->lev_f);
      build_levels(nnode, ia.data(), ja.data(), false, rows_b, c->lev_b);
    }
  }
  // per level: can a tile of rows be fetched as one contiguous byte range (rows consecutive in memory), and how
  // many blocks does the largest tile hold
  auto tile_caps = [&](const std::vector<int>& rows, const std::vector<int>& off, std::vector<int>& cap,
                       std::vector<int>& first, std::vector<int>& stepv) {
    cap.assign(off.size() - 1, 0);
    first.assign(off.size() - 1, 0);
    stepv.assign(off.size() - 1, 1);
    for (size_t l = 0; l + 1 < off.size(); l++) {
      first[l] = rows[off[l]];
      stepv[l] = (off[l + 1] - off[l] > 1 && rows[off[l] + 1] < rows[off[l]]) ? -1 : 1;
    }
    if (c->sgs_tile_warps == 0) return;
    const int RT = c->sgs_tile_warps * (32 / c->sgs_tile_lpr);
    for (size_t l = 0; l + 1 < off.size(); l++) {
      bool consecutive = true;
      for (int s = off[l]; s + 1 < off[l + 1] && consecutive; s++) consecutive = rows[s + 1] - rows[s] == stepv[l];
      if (!consecutive) continue;
      int mx = 0;
      for (int s0 = off[l]; s0 < off[l + 1]; s0 += RT) {
        const int s1 = std::min(s0 + RT, off[l + 1]);
        const int rlo = std::min(rows[s0], rows[s1 - 1]), rhi = std::max(rows[s0], rows[s1 - 1]);
        if (rhi - rlo + 1 != s1 - s0) { mx = 0; consecutive = false; break; }
        mx = std::max(mx, ia[rhi + 1] - ia[rlo]);
      }
      if (consecutive && (size_t)mx * ((size_t)neqn * neqn * 8 + 4) + 64 <= 100 * 1024) cap[l] = mx;
    }
  };
  tile_caps(rows_f, c->lev_f, c->tile_cap_f, c->lev_first_f, c->lev_step_f);
  tile_caps(rows_b, c->lev_b, c->tile_cap_b, c->lev_first_b, c->lev_step_b);
  for (int v : c->tile_cap_f) c->ring_cap_blocks = std::max(c->ring_cap_blocks, v);
  for (int v : c->tile_cap_b) c->ring_cap_blocks = std::max(c->ring_cap_blocks, v);

  // ---- upload
  if (dev_upload(c, &c->en, reinterpret_cast<const int2*>(mesh->edges_n), (size_t)nedge)) return 1;
  if (dev_upload(c, &c->ea, mesh->edges_a, (size_t)nedge * 4)) return 1;
  if (dev_upload(c, &c->ben, reinterpret_cast<const int2*>(mesh->bedges_n), (size_t)nb)) return 1;
  if (dev_upload(c, &c->bea, mesh->bedges_a, (size_t)nb * 4)) return 1;
  if (dev_upload(c, &c->bctype, mesh->bedges_bctype, (size_t)nb)) return 1;
  if (dev_upload(c, &c->xyz, mesh->xyz, (size_t)c->nn * 3)) return 1;
  if (dev_upload(c, &c->vol, mesh->vol, (size_t)nnode)) return 1;
  if (dev_upload(c, &c->adjp, adjp.data(), adjp.size())) return 1;
  if (dev_upload(c, &c->adj, adj.data(), adj.size())) return 1;
  if (dev_upload(c, &c->bnodes, bnodes.data(), bnodes.size())) return 1;
  if (dev_upload(c, &c->blist, blist.data(), blist.size())) return 1;
  if (dev_upload(c, &c->bfirst, bfirst.data(), bfirst.size())) return 1;
  if (dev_upload(c, &c->bnormal, bnormal.data(), bnormal.size())) return 1;
  if (dev_upload(c, &c->btwall, btwall.data(), btwall.size())) return 1;
  if (dev_upload(c, &c->bubar, bubar.data(), bubar.size())) return 1;
  if (dev_upload(c, &c->wallflag, wallflag.data(), wallflag.size())) return 1;
  if (dev_upload(c, &c->wnodes, wnodes.data(), wnodes.size())) return 1;
  if (dev_upload(c, &c->tbnodes, tbnodes.data(), tbnodes.size())) return 1;
  if (!vnn23.empty() && dev_upload(c, &c->vnn23, vnn23.data(), vnn23.size())) return 1;
  if (dev_upload(c, &c->ia, ia.data(), ia.size())) return 1;
  {   // 16 bytes of slack behind ja (and A): the bulk copies of k_sgs_tile round their byte ranges to 16
    std::vector<int> japad(ja);
    japad.resize(ja.size() + 4, 0);
    if (dev_upload(c, &c->ja, japad.data(), japad.size())) return 1;
  }
  if (dev_upload(c, &c->iau, iau.data(), iau.size())) return 1;
  if (dev_upload(c, &c->posLR, posLR.data(), posLR.size())) return 1;
  if (dev_upload(c, &c->posRL, posRL.data(), posRL.size())) return 1;
  if (dev_upload(c, &c->bpos, bpos.data(), bpos.size())) return 1;
  if (dev_upload(c, &c->rows_f, rows_f.data(), rows_f.size())) return 1;
  if (dev_upload(c, &c->rows_b, rows_b.data(), rows_b.size())) return 1;
  if (dev_alloc(c, &c->pv, (size_t)nnode * neqn + 8)) return 1;

  c->fsize[PCFD_F_Q] = (size_t)c->ntot * nvars;
  c->fsize[PCFD_F_QGRAD] = (size_t)c->nn * nterms * 3;
  c->fsize[PCFD_F_LIMITER] = (size_t)c->nn * neqn;
  c->fsize[PCFD_F_B] = (size_t)nnode * neqn;
  c->fsize[PCFD_F_X] = (size_t)c->nn * neqn;
  c->fsize[PCFD_F_TIMESTEP] = (size_t)nnode;
  c->fsize[PCFD_F_BETA] = (size_t)c->ntot;
  c->fsize[PCFD_F_LSQ_S] = (size_t)c->nn * 6;
  c->fsize[PCFD_F_LSQ_SW] = (size_t)c->nn * 6;
  c->fsize[PCFD_F_MUT] = (size_t)c->ntot;
  const bool sa_on = params->turb_model == 1;
  c->fsize[PCFD_F_TVAR] = sa_on ? (size_t)c->ntot : 0;
  c->fsize[PCFD_F_TGRAD] = sa_on ? (size_t)c->nn * 3 : 0;
  c->fsize[PCFD_F_WALLDIST] = (sa_on || !c->ffv_edges.empty()) ? (size_t)c->nn : 0;
  c->fsize[PCFD_F_TURB_B] = sa_on ? (size_t)nnode : 0;
  c->fsize[PCFD_F_TURB_X] = sa_on ? (size_t)c->nn : 0;
  c->fsize[PCFD_F_TURB_A] = sa_on ? (size_t)c->nblocks : 0;
  c->fsize[PCFD_F_A] = 0;   // allocated on first use (implicit runs only)
  for (int k = 0; k < PCFD_F_COUNT; k++) {
    if (k == PCFD_F_A || k == PCFD_F_QOLD || k == PCFD_F_QOLDM1) continue;   // allocated on first use
    if (dev_alloc(c, &c->f[k], c->fsize[k] + 4)) return 1;   // slack: 16-byte rounded bulk prefetches
    CK(cudaMemset(c->f[k], 0, std::max<size_t>(c->fsize[k], 1) * sizeof(double)));
  }
  if (dev_alloc(c, &c->flux, (size_t)nedge * neqn)) return 1;
  if (dev_alloc(c, &c->bflux, (size_t)nb * neqn)) return 1;
  if (neqn == NEQN) {
    if (dev_alloc(c, &c->eig, (size_t)nedge)) return 1;
    if (dev_alloc(c, &c->beig, (size_t)nb)) return 1;
  }
  if (const char* e = getenv("PCFD_EIG_FUSE")) c->eig_fuse = atoi(e) != 0;
  if (sa_on) {
    if (dev_upload(c, &c->dlev_f, c->lev_f.data(), c->lev_f.size())) return 1;
    if (dev_upload(c, &c->dlev_b, c->lev_b.data(), c->lev_b.size())) return 1;
    if (const char* e = getenv("PCFD_TURB_PERSIST")) c->turb_persist = atoi(e) != 0;
    if (dev_alloc(c, &c->tslots, (size_t)nedge * 3)) return 1;
    if (dev_alloc(c, &c->tbslots, (size_t)nb * 4)) return 1;
    if (neqn != NEQN) {   // reacting eqnset: tables of kfr_turb_props
      if (dev_alloc(c, &c->tprop_e, (size_t)nedge * 2)) return 1;
      if (dev_alloc(c, &c->tprop_b, (size_t)nb * 2)) return 1;
      if (dev_alloc(c, &c->tprop_n, (size_t)c->nn * 2)) return 1;
    }
  }
  if (viscous) {
    if (dev_alloc(c, &c->vflux, (size_t)nedge * 4)) return 1;
    if (dev_alloc(c, &c->bvflux, (size_t)nb * 4)) return 1;
  }
  if (dev_alloc(c, &c->dzeroed, 4)) return 1;
  CK(cudaMemset(c->dzeroed, 0, 4 * sizeof(int)));
  if (dev_alloc(c, &c->red, (size_t)RED_BLOCKS * 8)) return 1;
  if (dev_alloc(c, &c->redout, 16)) return 1;
  if (dev_alloc(c, &c->clipflag, (size_t)nedge)) return 1;
  if (dev_alloc(c, &c->tclip[0], (size_t)nnode)) return 1;
  if (dev_alloc(c, &c->tclip[1], (size_t)nnode)) return 1;
  if (dev_alloc(c, &c->dflags, 4)) return 1;
  CK(cudaMallocHost(reinterpret_cast<void**>(&c->hflag), 64));   // clip flag (int) + small pinned scratch
  CK(cudaEventCreateWithFlags(&c->ev_flag, cudaEventDisableTiming));
  if (const char* e = getenv("PCFD_FUSED_CLIP")) c->fused_clip = atoi(e) != 0;
  if (const char* e = getenv("PCFD_GRAD_GEO")) c->use_geo = atoi(e) != 0;
  if (const char* e = getenv("PCFD_GRAD_THREADS")) c->grad_threads = atoi(e) == 1 ? 1 : 3;
  if (const char* e = getenv("PCFD_LIMITER_QMM")) c->use_qmm = atoi(e) != 0;
  if (neqn == NEQN && dev_alloc(c, &c->qmm, (size_t)nnode * 10)) return 1;
  if (const char* e = getenv("PCFD_LIMITER_PER_NODE")) c->limiter_per_node = atoi(e) != 0;
  if (neqn == NEQN && dev_alloc(c, &c->geo, adj.size())) return 1;

  c->dm = DevMesh{c->nnode, c->gnode, c->nbnode, c->nedge, c->nbedge, c->ngedge, c->en, c->ea, c->ben, c->bea,
                  c->bctype, c->xyz, c->vol, c->adjp, c->adj, c->bnormal, c->btwall, c->bubar};
  c->bp.gamma = params->gamma;
  c->bp.no_cvbc = params->no_cvbc;
  for (int i = 0; i < NVARS; i++) c->bp.qinf[i] = params->qinf[i];
  CK(cudaDeviceSynchronize());
  guard.ok = true;
  *out = c;
  return 0;
}

int pcfd_create(const pcfd_mesh_desc* mesh, const pcfd_params* params, int device, pcfd_ctx** out) {
  if (params && params->eqnset != PCFD_EQNSET_COMPRESSIBLE_EULER && params->eqnset != PCFD_EQNSET_COMPRESSIBLE_NS) {
    if (out) *out = nullptr;
    return fail(nullptr, "pcfd_create: unsupported eqnset id (the reacting eqnset is created with pcfd_create_fr)");
  }
  return create_impl(mesh, params, device, NEQN, NVARS, NTERMS, out);
}

int pcfd_set_stream(pcfd_ctx* c, void* s) {
  if (!c) return 1;
  c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
  return 0;
}
int pcfd_synchronize(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (comm_on(c)) return comm_check_err(c);
  return 0;
}
int pcfd_set_cfl(pcfd_ctx* c, double cfl) {
  if (!c) return 1;
  c->prm.cfl = cfl;
  return 0;
}
long long pcfd_launch_count(const pcfd_ctx* c) { return c ? c->launches : 0; }

long long pcfd_zeroed_updates(pcfd_ctx* c) {
  if (!c) return -1;
  int h = 0;
  if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess ||
      cudaMemcpy(&h, c->dzeroed, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  return h;
}

int pcfd_profile_enable(pcfd_ctx* c, int on) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  if (!on && c->prof) prof_drain(c);
  c->prof = on != 0;
  return 0;
}
int pcfd_profile_reset(pcfd_ctx* c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  prof_drain(c);
  c->prof_acc.clear();
  return 0;
}
int pcfd_profile_count(pcfd_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  prof_drain(c);
  return (int)c->prof_acc.size();
}
int pcfd_profile_get(pcfd_ctx* c, int i, const char** name, double* total_ms, long long* launches) {
  if (!c || i < 0 || i >= (int)c->prof_acc.size()) return 1;
  if (name) *name = c->prof_acc[i].name.c_str();
  if (total_ms) *total_ms = c->prof_acc[i].ms;
  if (launches) *launches = c->prof_acc[i].n;
  return 0;
}

// qold / qoldm1: unsteady runs only, allocated on first use
static int ensure_time_fields(pcfd_ctx* c) {
  if (c->f[PCFD_F_QOLD]) return 0;
  for (int k : {PCFD_F_QOLD, PCFD_F_QOLDM1}) {
    c->fsize[k] = (size_t)c->nnode * c->nvars;
    if (dev_alloc(c, &c->f[k], c->fsize[k] + 4)) return 1;
    CK(cudaMemsetAsync(c->f[k], 0, c->fsize[k] * sizeof(double), c->stream));
  }
  return 0;
}
static bool is_time_field(int field) { return field == PCFD_F_QOLD || field == PCFD_F_QOLDM1; }

static int ensure_matrix(pcfd_ctx* c) {
  if (c->f[PCFD_F_A]) return 0;
  c->fsize[PCFD_F_A] = (size_t)c->nblocks * c->neqn * c->neqn;
  if (dev_alloc(c, &c->f[PCFD_F_A], c->fsize[PCFD_F_A] + 2)) return 1;
  if (dev_alloc(c, &c->bdiag, (size_t)c->nb * c->neqn * c->neqn)) return 1;
  CK(cudaMemsetAsync(c->f[PCFD_F_A], 0, c->fsize[PCFD_F_A] * sizeof(double), c->stream));
  return 0;
}

size_t pcfd_field_size(const pcfd_ctx* c, int field) {
  if (!c || field < 0 || field >= PCFD_F_COUNT) return 0;
  if (field == PCFD_F_A) return (size_t)c->nblocks * c->neqn * c->neqn;
  if (is_time_field(field)) return (size_t)c->nnode * c->nvars;
  return c->fsize[field];
}
int pcfd_set_field(pcfd_ctx* c, int field, const double* host, size_t n) {
  if (!c) return 1;
  if (field < 0 || field >= PCFD_F_COUNT || !host) return fail(c, "pcfd_set_field: bad argument");
  CK(cudaSetDevice(c->device));
  if (field == PCFD_F_A) { if (ensure_matrix(c)) return 1; c->ludiag = false; }
  if (field == PCFD_F_LSQ_SW) c->geo_valid = false;
  if (field == PCFD_F_Q) { c->qmm_valid = false; if (c->comm) c->comm->ghost_q_fresh = false; }
  if (is_time_field(field)) {
    if (ensure_time_fields(c)) return 1;
    if (field == PCFD_F_QOLD) c->have_qold = true;
  }
  if (n != c->fsize[field]) return fail(c, "pcfd_set_field: size mismatch");
  CK(cudaMemcpyAsync(c->f[field], host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (field == PCFD_F_WALLDIST && !c->ffv_edges.empty()) {
    // PowerLawU(1, d, Re) (powerLaw.h:11-29) per FarFieldViscous half-edge, with the C library's pow like the reference
    std::vector<double> ub(c->ffv_edges.size());
    for (size_t k = 0; k < ub.size(); k++) {
      const double deltaTurb = 0.382 * 10.0 / (std::pow(c->prm.Re, 0.2));
      const double u = 1.0 * std::pow(host[c->ffv_left[k]] / deltaTurb, 1.0 / 7.0);
      ub[k] = (u < 1.0) ? u : 1.0;
    }
    for (size_t k = 0; k < ub.size(); k++)
      CK(cudaMemcpy(c->bubar + c->ffv_edges[k], &ub[k], sizeof(double), cudaMemcpyHostToDevice));
    c->ffv_ready = true;
  }
  return 0;
}
int pcfd_get_field(pcfd_ctx* c, int field, double* host, size_t n) {
  if (!c) return 1;
  if (field < 0 || field >= PCFD_F_COUNT || !host) return fail(c, "pcfd_get_field: bad argument");
  CK(cudaSetDevice(c->device));
  if (field == PCFD_F_A && ensure_matrix(c)) return 1;
  if (is_time_field(field) && ensure_time_fields(c)) return 1;
  if (n != c->fsize[field]) return fail(c, "pcfd_get_field: size mismatch");
  CK(cudaMemcpyAsync(host, c->f[field], n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
void* pcfd_field_device_ptr(pcfd_ctx* c, int field) {
  if (!c || field < 0 || field >= PCFD_F_COUNT) return nullptr;
  if (field == PCFD_F_A) { cudaSetDevice(c->device); if (ensure_matrix(c)) return nullptr; }
  if (is_time_field(field)) { cudaSetDevice(c->device); if (ensure_time_fields(c)) return nullptr; }
  return c->f[field];
}
int pcfd_crs_sizes(const pcfd_ctx* c, int* nrows, int* nblocks) {
  if (!c) return 1;
  if (nrows) *nrows = c->nnode;
  if (nblocks) *nblocks = c->nblocks;
  return 0;
}
int pcfd_get_crs(pcfd_ctx* c, int* ia, int* ja, int* iau, int* pv) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (ia) CK(cudaMemcpy(ia, c->ia, (size_t)(c->nnode + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  if (ja) CK(cudaMemcpy(ja, c->ja, (size_t)c->nblocks * sizeof(int), cudaMemcpyDeviceToHost));
  if (iau) CK(cudaMemcpy(iau, c->iau, (size_t)c->nnode * sizeof(int), cudaMemcpyDeviceToHost));
  if (pv) CK(cudaMemcpy(pv, c->pv, (size_t)c->nnode * c->neqn * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int pcfd_lsq_coefficients(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  PROF("k_lsq_coeff");
  k_lsq_coeff<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->f[PCFD_F_LSQ_S], c->f[PCFD_F_LSQ_SW]);
  LAUNCH_CHECK();
  c->geo_valid = false;
  if (comm_on(c)) {   // gradient.tcc:131-134: halos of s and sw
    if (comm_update(c, PCFD_F_LSQ_S)) return 1;
    return comm_update(c, PCFD_F_LSQ_SW);
  }
  return 0;
}

int pcfd_update_bcs(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  c->qmm_valid = false;
  if (c->comm && !c->comm->no_hardset) c->comm->ghost_q_fresh = false;   // hard-set BC nodes: owned rows change
  if (!c->ffv_edges.empty() && !c->ffv_ready)
    return fail(c, "pcfd_update_bcs: the viscous far-field BC needs field PCFD_F_WALLDIST (pcfd_set_field) first");
  if (c->fr) return pcfd_fr_update_bcs(c);
  if (c->nbn) {
    PROF("k_update_bcs");
    k_update_bcs<<<nblk(c->nbn, 128), 128, 0, c->stream>>>(c->dm, c->bp, c->bnodes, c->nbn, c->f[PCFD_F_Q]);
    LAUNCH_CHECK();
  }
  if (c->nblist_bc) {
    PROF("k_update_bcs_edges");
    k_update_bcs_edges<<<nblk(c->nblist_bc, 128), 128, 0, c->stream>>>(c->dm, c->bp, c->blist, c->nblist_bc, c->bfirst,
                                                                       c->f[PCFD_F_Q]);
    LAUNCH_CHECK();
  }
  return 0;
}

static int run_gradient_geo(pcfd_ctx* c);

int pcfd_gradient(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (c->fr) return pcfd_fr_gradient(c);
  if (c->grad_type == 1) {
    PROF("k_gradient_gg");
    k_gradient_gg<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD]);
    LAUNCH_CHECK();
    return 0;
  }
  if (c->use_geo && c->geo) return run_gradient_geo(c);
  PROF("k_gradient");
  k_gradient<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->f[PCFD_F_Q], c->f[PCFD_F_LSQ_SW], c->f[PCFD_F_QGRAD]);
  LAUNCH_CHECK();
  return 0;
}

int pcfd_set_jacobian_type(pcfd_ctx* c, int field_type, int boundary_type) {
  if (!c) return 1;
  if (field_type < 0 || field_type > 2 || (boundary_type != 0 && boundary_type != 1) || (field_type == 2 && c->fr))
    return fail(c, "pcfd_set_jacobian_type: field 0 (one-sided differences), 1 (central differences) or 2 (complex step: "
                   "perfect-gas eqnsets only); boundary 0 or 1 -- the complex-step boundary Jacobian "
                   "(Bkernel_NumJac_Complex, jacobian.tcc:170-172) is not built");
  c->field_jac_type = field_type;
  c->boundary_jac_type = boundary_type;
  return 0;
}

int pcfd_set_gradient_type(pcfd_ctx* c, int type) {
  if (!c) return 1;
  if (type != 0 && type != 1) return fail(c, "pcfd_set_gradient_type: 0 (weighted least squares) or 1 (Green-Gauss), gradient.tcc:68-90");
  c->grad_type = type;
  return 0;
}

static int run_limiter_final(pcfd_ctx* c, const int* tclip);
static int run_limiter_raw(pcfd_ctx* c);
static int run_flux(pcfd_ctx* c, bool fused = false, bool* clip_hit = nullptr);
static int run_sumsq(pcfd_ctx* c, const double* v, int nrows, double* host_out);

int pcfd_limiter(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (c->fr) return pcfd_fr_limiter(c);
  const int type = c->prm.limiter;
  double* lim = c->f[PCFD_F_LIMITER];
  if (run_limiter_raw(c)) return 1;
  if (type == 0) return 0;
  // pressure clip: iterate (edges -> flags, nodes -> first clipping edge) to the fixed point
  int cur = 0;
  bool clipped = false;
  PROF("k_fill_int");
  k_fill_int<<<nblk(c->nnode, 256), 256, 0, c->stream>>>(c->tclip[0], c->nnode, INT_MAX);
  LAUNCH_CHECK();
  for (int it = 0; it < c->nedge + 2; it++) {
    int hflags[2] = {0, 0};
    CK(cudaMemsetAsync(c->dflags, 0, 2 * sizeof(int), c->stream));
    PROF("k_clip_edges");
    k_clip_edges<<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, c->prm.chi, c->prm.gamma, c->f[PCFD_F_Q],
                                                             c->f[PCFD_F_QGRAD], lim, c->tclip[cur], c->clipflag,
                                                             c->dflags);
    LAUNCH_CHECK();
    if (it == 0) {
      CK(cudaMemcpyAsync(hflags, c->dflags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      if (!hflags[0]) break;   // no edge clips anything: the common case
      clipped = true;
    }
    PROF("k_clip_nodes");
    k_clip_nodes<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->clipflag, c->tclip[cur], c->tclip[cur ^ 1],
                                                             c->dflags + 1);
    LAUNCH_CHECK();
    cur ^= 1;
    CK(cudaMemcpyAsync(hflags, c->dflags, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (!hflags[1]) break;     // tclip reproduced itself: fixed point
  }
  return run_limiter_final(c, clipped ? c->tclip[cur] : nullptr);
}

// Gradient -> Limiter -> ComputeResiduals of the composite iterations.  With a limiter on, the pressure-clip test
// rides along in the edge-flux kernel (k_flux_edges<true>); only when some edge actually clips (rare: the limiter
// exists to prevent exactly that) are the ordered clip passes and the flux redone.
// Gradient::Compute with the static LSQ visit weights (k_lsq_geo), (re)built when the coefficients changed
static int run_gradient_geo(pcfd_ctx* c) {
  if (!c->geo_valid) {
    PROF("k_lsq_geo");
    k_lsq_geo<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->f[PCFD_F_LSQ_SW], c->geo);
    LAUNCH_CHECK();
    c->geo_valid = true;
  }
  PROF("k_gradient");
  if (c->grad_threads == 3) {
    k_gradient_geo3<<<nblk((long long)c->nnode * 3, 192), 192, 0, c->stream>>>(c->dm, c->f[PCFD_F_Q], c->geo, c->f[PCFD_F_QGRAD],
                                                                              c->qmm);
    c->qmm_valid = true;
  } else {
    k_gradient_geo<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->f[PCFD_F_Q], c->geo, c->f[PCFD_F_QGRAD]);
  }
  LAUNCH_CHECK();
  return 0;
}

// passes 1 + 2 of Limiter::Compute (raw limiter): five threads per node (one per equation) or one thread per node
static int run_limiter_raw(pcfd_ctx* c) {
  PROF("k_limiter");
  if (c->limiter_per_node)
    k_limiter_node<<<nblk(c->nn, 128), 128, 0, c->stream>>>(c->dm, c->prm.limiter, c->prm.chi, c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD],
                                                            c->f[PCFD_F_LIMITER]);
  else
    k_limiter<<<nblk((long long)c->nn * 5, 160), 160, 0, c->stream>>>(c->dm, c->prm.limiter, c->prm.chi, c->f[PCFD_F_Q],
                                                                     c->f[PCFD_F_QGRAD], c->f[PCFD_F_LIMITER],
                                                                     (c->qmm_valid && c->use_qmm) ? c->qmm : nullptr);
  LAUNCH_CHECK();
  return 0;
}

static int gradient_limiter_residual(pcfd_ctx* c, double* sumsq) {
  const bool dist = comm_on(c);
  if (c->fr) {
    if (c->prm.sorder > 1) {
      if (pcfd_gradient(c)) return 1;
      if (dist && comm_update(c, PCFD_F_QGRAD)) return 1;          // gradient.tcc:98
      if (c->prm.limiter != 0 && c->fused_clip) {
        bool hit = false;
        if (pcfd_fr_limiter_raw(c)) return 1;
        if (dist && comm_update(c, PCFD_F_LIMITER)) return 1;      // limiters.tcc:128 (raw values; both sides clamp)
        if (pcfd_fr_residual_fused(c, sumsq, &hit)) return 1;
        if (dist) {   // one decision for all ranks
          double mine = hit ? 1.0 : 0.0, all[COMM_MAXR];
          if (pcfd_comm_allgather(c, &mine, 1, all)) return 1;
          for (int r = 0; r < c->nranks; r++) hit = hit || all[r] != 0.0;
        }
        if (!hit) return 0;
        c->clip_fallbacks++;
      }
      if (pcfd_limiter(c)) return 1;
      if (dist && comm_update(c, PCFD_F_LIMITER)) return 1;
    }
    return pcfd_residual(c, sumsq);
  }
  if (c->prm.sorder > 1) {
    const int type = c->prm.limiter;
    if (pcfd_gradient(c)) return 1;
    if (type != 0 && c->fused_clip) {
      if (run_limiter_raw(c)) return 1;
      // gradient.tcc:98 and limiters.tcc:128 (raw values; both sides clamp) in ONE put: the limiter of the owned nodes
      // does not read ghost gradients, so both halos can leave together; waited for inside run_flux
      if (dist && comm_post(c, PCFD_F_QGRAD, PCFD_F_LIMITER)) return 1;
      bool hit = false;
      if (run_flux(c, true, &hit)) return 1;
      if (!hit) {
        if (sumsq) return run_sumsq(c, c->f[PCFD_F_B], c->nnode, sumsq);
        return 0;
      }
      c->clip_fallbacks++;
    } else if (dist && comm_update(c, PCFD_F_QGRAD)) {              // gradient.tcc:98
      return 1;
    }
    if (pcfd_limiter(c)) return 1;
    if (dist && comm_update(c, PCFD_F_LIMITER)) return 1;
  }
  return pcfd_residual(c, sumsq);
}

// The two halves of the fused limiter / residual path as separate entry points, so that a multi-rank host can put the
// limiter halo exchange between them (the clamp of negatives is element-wise, so exchanging the raw limiter and clamping
// on both sides equals exchanging the clamped one).
int pcfd_limiter_raw(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (c->fr) return pcfd_fr_limiter_raw(c);
  return run_limiter_raw(c);
}

long long pcfd_clip_fallbacks(const pcfd_ctx* c) { return c ? c->clip_fallbacks : -1; }

int pcfd_set_time_integration(pcfd_ctx* c, double dt, int use_local_time_stepping, int torder, int iter) {
  if (!c) return 1;
  if (torder != 1 && torder != 2) return fail(c, "pcfd_set_time_integration: torder must be 1 or 2 (param.tcc:167)");
  if (dt == 0.0) return fail(c, "pcfd_set_time_integration: dt must be non-zero (negative: steady)");
  c->time_dt = dt;
  c->time_local = use_local_time_stepping ? 1 : 0;
  c->torder = torder;
  c->iter = iter;
  if (c->fr) pcfd_fr_set_time(c, dt, c->time_local);
  return 0;
}

int pcfd_residual_fused(pcfd_ctx* c, double* sumsq, int* clip_hit) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (!clip_hit) return fail(c, "pcfd_residual_fused: clip_hit must not be NULL");
  if (c->prm.sorder < 2 || c->prm.limiter == 0) {   // nothing to clip: the plain residual
    *clip_hit = 0;
    return pcfd_residual(c, sumsq);
  }
  bool hit = false;
  if (c->fr) {
    if (pcfd_fr_residual_fused(c, sumsq, &hit)) return 1;
    *clip_hit = hit ? 1 : 0;
    if (hit) c->clip_fallbacks++;
    return 0;
  }
  if (run_flux(c, true, &hit)) return 1;
  *clip_hit = hit ? 1 : 0;
  if (hit) c->clip_fallbacks++;
  if (sumsq && !hit) return run_sumsq(c, c->f[PCFD_F_B], c->nnode, sumsq);
  return 0;
}

static int run_limiter_final(pcfd_ctx* c, const int* tclip) {
  PROF("k_limiter_final");
  k_limiter_final<<<nblk((long long)c->nn * 5, 256), 256, 0, c->stream>>>(c->nn, c->nnode, tclip, c->f[PCFD_F_LIMITER]);
  LAUNCH_CHECK();
  return 0;
}

// fused = true: lim holds the raw limiter; the edge kernel clamps on the fly and raises dflags[2] if the pressure
// clip would act anywhere, k_limiter_final then clamps in place, and *clip_hit reports the flag (one event wait that
// overlaps with the kernels queued behind it)
// BDF coefficients (residual.tcc:139-149, jacobian.tcc:228-230)
static double time_cnp1(const pcfd_ctx* c) { return (c->iter > 1 && c->torder == 2) ? 1.5 : 1.0; }
static double time_cnm1(const pcfd_ctx* c) { return (c->iter > 1 && c->torder == 2) ? -0.5 : 0.0; }

// TemporalResidual, once the host has provided q^n (ComputeResiduals, residual.tcc:20-24)
static int run_temporal(pcfd_ctx* c) {
  if (!c->torder || !c->have_qold) return 0;
  PROF("k_temporal_residual");
  k_temporal_residual<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->nnode, time_cnp1(c), time_cnm1(c), c->time_dt, c->vol,
                                                                  c->f[PCFD_F_Q], c->f[PCFD_F_QOLD], c->f[PCFD_F_QOLDM1],
                                                                  c->viscous ? c->wallflag : nullptr, c->f[PCFD_F_B]);
  LAUNCH_CHECK();
  return 0;
}

// Across ranks (comm connected, fused): the caller has POSTED the halos of qgrad and of the raw limiter; the interior
// edge kernel -- ghost-independent -- runs while they are in flight, every rank's clip flag goes round through the
// flag page (no collective), and the waits come just before the first kernel that reads ghost rows.  *clip_hit is
// the GLOBAL decision (limiters.tcc:112-117 runs on every rank or on none).
// The halo of q after UpdateBCs (solutionSpace.tcc:665).  Elided when it would deliver, bit for bit, the rows the last
// halo of q delivered: nothing has written owned rows of q since (no explicit / implicit update, no pcfd_set_field,
// and no rank owns hard-set BC nodes, so UpdateBCs only wrote phantom rows).
static int comm_q_after_bcs(pcfd_ctx* c) {
  pcfd_comm* m = c->comm;
  if (m->elide && m->ghost_q_fresh && m->no_hardset) { m->elided++; return 0; }
  return comm_update(c, PCFD_F_Q);
}

// the host reads the clip flag(s) while the kernels queued behind the copy keep the GPU busy
static int flux_clip_decision(pcfd_ctx* c, bool dist, bool* clip_hit) {
  if (!dist) {
    CK(cudaEventSynchronize(c->ev_flag));
    *clip_hit = *c->hflag != 0;
    return 0;
  }
  CK(cudaEventSynchronize(c->comm->ev));
  bool hit = false;
  for (int r = 0; r < c->nranks; r++) hit = hit || c->comm->hgout[r * COMM_GW] != 0.0;
  *clip_hit = hit;
  return 0;
}

static int run_flux(pcfd_ctx* c, bool fused, bool* clip_hit) {
  const bool dist = fused && comm_on(c);
  const bool eig = c->eig_now;   // pcfd_explicit_iterate: the time step rides along (k_eig_bedges has run)
  if (fused) CK(cudaMemsetAsync(c->dflags + 2, 0, sizeof(int), c->stream));
  if (c->nedge) {
    PROF("k_flux_edges");
#define PCFD_FLUX_LAUNCH(DD, EE, ANY)                                                                                   \
  k_flux_edges<DD, EE><<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, c->prm.sorder, c->prm.chi, c->prm.gamma,       \
                                                                  c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD],                   \
                                                                  c->f[PCFD_F_LIMITER], c->flux, ANY, c->eig)
    if (fused) { if (eig) PCFD_FLUX_LAUNCH(true, true, c->dflags + 2); else PCFD_FLUX_LAUNCH(true, false, c->dflags + 2); }
    else { if (eig) PCFD_FLUX_LAUNCH(false, true, nullptr); else PCFD_FLUX_LAUNCH(false, false, nullptr); }
#undef PCFD_FLUX_LAUNCH
    LAUNCH_CHECK();
  }
  if (dist) {
    if (comm_gather_enqueue(c, nullptr, c->dflags + 2, 1)) return 1;
    if (comm_wait(c, PCFD_F_QGRAD)) return 1;
    if (c->comm->field_epoch[PCFD_F_LIMITER] != c->comm->field_epoch[PCFD_F_QGRAD] && comm_wait(c, PCFD_F_LIMITER)) return 1;
    if (run_limiter_final(c, nullptr)) return 1;
  } else if (fused) {
    CK(cudaMemcpyAsync(c->hflag, c->dflags + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev_flag, c->stream));
    if (run_limiter_final(c, nullptr)) return 1;
  }
  if (c->nb) {
    PROF("k_flux_bedges");
    k_flux_bedges<<<nblk(c->nb, 128), 128, 0, c->stream>>>(c->dm, c->prm.sorder, c->prm.chi, c->prm.gamma,
                                                           c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD], c->f[PCFD_F_LIMITER],
                                                           c->bflux);
    LAUNCH_CHECK();
  }
  if (c->viscous) {
    if (c->nedge) {
      PROF("k_vflux_edges");
      k_vflux_edges<<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, c->prm.sorder, c->vp, c->f[PCFD_F_Q],
                                                                c->f[PCFD_F_QGRAD], c->f[PCFD_F_MUT], c->vflux);
      LAUNCH_CHECK();
    }
    if (c->nb) {
      PROF("k_vflux_bedges");
      k_vflux_bedges<<<nblk(c->nb, 128), 128, 0, c->stream>>>(c->dm, c->prm.sorder, c->vp, c->f[PCFD_F_Q],
                                                              c->f[PCFD_F_QGRAD], c->f[PCFD_F_MUT], c->bvflux);
      LAUNCH_CHECK();
    }
    PROF("k_residual_gather");
    if (eig)
      k_residual_gather<true, true><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->flux, c->bflux, c->vflux, c->bvflux,
                                                                                c->wallflag, c->f[PCFD_F_B], c->eig, c->beig,
                                                                                c->prm.cfl, c->vnn23, c->f[PCFD_F_TIMESTEP]);
    else
      k_residual_gather<true, false><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->flux, c->bflux, c->vflux, c->bvflux,
                                                                                 c->wallflag, c->f[PCFD_F_B], nullptr, nullptr,
                                                                                 0.0, nullptr, nullptr);
    LAUNCH_CHECK();
    if (run_temporal(c)) return 1;
    if (fused) return flux_clip_decision(c, dist, clip_hit);
    return 0;
  }
  PROF("k_residual_gather");
  if (eig)
    k_residual_gather<false, true><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->flux, c->bflux, nullptr, nullptr, nullptr,
                                                                               c->f[PCFD_F_B], c->eig, c->beig, c->prm.cfl,
                                                                               c->vnn23, c->f[PCFD_F_TIMESTEP]);
  else
    k_residual_gather<false, false><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->flux, c->bflux, nullptr, nullptr, nullptr,
                                                                                c->f[PCFD_F_B], nullptr, nullptr, 0.0, nullptr,
                                                                                nullptr);
  LAUNCH_CHECK();
  if (run_temporal(c)) return 1;
  if (fused) return flux_clip_decision(c, dist, clip_hit);
  return 0;
}

static int run_sumsq(pcfd_ctx* c, const double* v, int nrows, double* host_out) {
  PROF("k_sumsq_partial");
  k_sumsq_partial<256, NEQN><<<RED_BLOCKS, 256, 0, c->stream>>>(v, nrows, c->red);
  LAUNCH_CHECK();
  PROF("k_sumsq_final");
  k_sumsq_final<256, NEQN><<<1, 256, 0, c->stream>>>(c->red, RED_BLOCKS, c->redout);
  LAUNCH_CHECK();
  if (host_out) {
    CK(cudaMemcpyAsync(host_out, c->redout, (1 + NEQN) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int pcfd_residual(pcfd_ctx* c, double* sumsq) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (c->fr) return pcfd_fr_residual(c, sumsq);
  if (run_flux(c)) return 1;
  if (sumsq) return run_sumsq(c, c->f[PCFD_F_B], c->nnode, sumsq);
  return 0;
}

// ComputeTimesteps without local time stepping (timestep.tcc:47-74): Param::dt in every cell when it is positive,
// otherwise the smallest CFL (and VNN) limited step of THIS rank's cells in every cell (the reference takes no
// minimum across ranks here; the all-reduce of solutionSpace.tcc:712-714 only feeds the residual file).
static int timestep_global(pcfd_ctx* c, double* dtmin) {
  double* dt = c->f[PCFD_F_TIMESTEP];
  if (c->time_dt > 0.0) {
    PROF("k_fill_double");
    k_fill_double<<<nblk(c->nnode, 256), 256, 0, c->stream>>>(dt, c->nnode, c->time_dt, nullptr);
    LAUNCH_CHECK();
    if (dtmin) *dtmin = c->time_dt;
    return 0;
  }
  PROF("k_min_partial");
  k_min_partial<256><<<RED_BLOCKS, 256, 0, c->stream>>>(dt, c->nnode, c->red);
  LAUNCH_CHECK();
  PROF("k_min_final");
  k_min_final<256><<<1, 256, 0, c->stream>>>(c->red, RED_BLOCKS, c->redout + 8);
  LAUNCH_CHECK();
  PROF("k_fill_double");
  k_fill_double<<<nblk(c->nnode, 256), 256, 0, c->stream>>>(dt, c->nnode, 0.0, c->redout + 8);
  LAUNCH_CHECK();
  if (dtmin) {
    CK(cudaMemcpyAsync(dtmin, c->redout + 8, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int pcfd_timestep(pcfd_ctx* c, double* dtmin) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  const bool global_dt = !c->time_local;
  if (global_dt && c->time_dt > 0.0) return timestep_global(c, dtmin);   // no eigenvalue pass at all (timestep.tcc:50-54)
  if (c->fr) {
    if (pcfd_fr_timestep(c, global_dt ? nullptr : dtmin)) return 1;
    return global_dt ? timestep_global(c, dtmin) : 0;
  }
  PROF("k_timestep");
  k_timestep<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->prm.gamma, c->prm.cfl, c->f[PCFD_F_Q], c->vnn23,
                                                         c->f[PCFD_F_TIMESTEP]);
  LAUNCH_CHECK();
  if (global_dt) return timestep_global(c, dtmin);
  if (dtmin) {
    PROF("k_min_partial");
    k_min_partial<256><<<RED_BLOCKS, 256, 0, c->stream>>>(c->f[PCFD_F_TIMESTEP], c->nnode, c->red);
    LAUNCH_CHECK();
    PROF("k_min_final");
    k_min_final<256><<<1, 256, 0, c->stream>>>(c->red, RED_BLOCKS, c->redout + 8);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(dtmin, c->redout + 8, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int pcfd_explicit_solve(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  c->qmm_valid = false;
  if (c->comm) c->comm->ghost_q_fresh = false;
  if (c->fr) return pcfd_fr_explicit_solve(c);
  PROF("k_explicit");
  k_explicit<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->nnode, c->prm.gamma, c->f[PCFD_F_B], c->f[PCFD_F_TIMESTEP],
                                                         c->vol, c->f[PCFD_F_X], c->f[PCFD_F_Q]);
  LAUNCH_CHECK();
  return 0;
}

int pcfd_apply_dq(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  c->qmm_valid = false;
  if (c->comm) c->comm->ghost_q_fresh = false;
  if (c->fr) return pcfd_fr_apply_dq(c);
  PROF("k_apply_dq");
  k_apply_dq<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->nnode, c->prm.gamma, c->f[PCFD_F_X], c->f[PCFD_F_Q], c->dzeroed);
  LAUNCH_CHECK();
  return 0;
}

int pcfd_jacobian(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (ensure_matrix(c)) return 1;
  if (!c->ffv_edges.empty() && !c->ffv_ready)
    return fail(c, "pcfd_jacobian: the viscous far-field BC needs field PCFD_F_WALLDIST (pcfd_set_field) first");
  if (c->comm && !c->comm->no_hardset) c->comm->ghost_q_fresh = false;
  if (c->fr) return pcfd_fr_jacobian(c);
  double* A = c->f[PCFD_F_A];
  // CRSMatrix::Blank: nothing to do.  Every block of A is written in full below -- interior off-diagonals by k_jac_edges,
  // A(l, ghost) by the ghost half-edges of k_jac_bedges / k_jac_bnodes, diagonals by k_jac_diag -- before anything is
  // added to it; a memset of the whole matrix (5 GB at 10 M cells: 0.9 ms; 16 GB for 9x9 blocks: 2.7 ms) bought nothing
  // (tests/test_gpu_variants.py poisons the matrix with NaN before the refresh)
  c->ludiag = false;
  if (c->nedge) {
    if (c->field_jac_type == 2) {
      PROF("k_jac_edges_complex");
      k_jac_edges_complex<<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, c->prm.gamma, c->f[PCFD_F_Q], c->posLR, c->posRL, A);
    } else if (c->field_jac_type == 1) {
      PROF("k_jac_edges_central");
      k_jac_edges_central<<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, c->prm.gamma, c->f[PCFD_F_Q], c->posLR, c->posRL, A);
    } else {
      PROF("k_jac_edges");
      // register cap, measured at 10 M cells (tools/time_pgjac.py with PCFD_JAC_MINB): 224 registers / 8 warps per SM 13.30 ms, 168: 12.47,
      // 128 (16 warps): 11.44, 96: 13.44 -- the kernel waits on FP64 latency, not on the pipe
      static const int minb = getenv("PCFD_JAC_MINB") ? atoi(getenv("PCFD_JAC_MINB")) : 4;
      const dim3 g(nblk(c->nedge, 128));
      if (minb == 4) k_jac_edges<4><<<g, 128, 0, c->stream>>>(c->dm, c->prm.gamma, c->f[PCFD_F_Q], c->posLR, c->posRL, A);
      else k_jac_edges<2><<<g, 128, 0, c->stream>>>(c->dm, c->prm.gamma, c->f[PCFD_F_Q], c->posLR, c->posRL, A);
    }
    LAUNCH_CHECK();
  }
  if (c->nbn) {
    if (c->boundary_jac_type == 1) {
      PROF("k_jac_bnodes_central");
      k_jac_bnodes_central<<<nblk(c->nbn, 64), 64, 0, c->stream>>>(c->dm, c->bp, c->bnodes, c->nbn, c->f[PCFD_F_Q], c->bpos,
                                                                   c->bdiag, A);
    } else {
      PROF("k_jac_bnodes");
      // (a 128-register cap, which pays for k_jac_bedges, measured slower here: 1.53 against 1.43 ms)
      k_jac_bnodes<<<nblk(c->nbn, 64), 64, 0, c->stream>>>(c->dm, c->bp, c->bnodes, c->nbn, c->f[PCFD_F_Q], c->bpos, c->bdiag, A);
    }
    LAUNCH_CHECK();
  }
  if (c->nblist) {
    if (c->boundary_jac_type == 1) {
      PROF("k_jac_bedges_central");
      k_jac_bedges_central<<<nblk(c->nblist, 64), 64, 0, c->stream>>>(c->dm, c->bp, c->blist, c->nblist, c->bfirst,
                                                                      c->f[PCFD_F_Q], c->bpos, c->bdiag, A);
    } else {
      PROF("k_jac_bedges");
      // 128 registers / 16 warps per SM: 0.94 ms at 10 M cells against 1.22 ms at 255 registers (FP64 latency)
      k_jac_bedges<8><<<nblk(c->nblist, 64), 64, 0, c->stream>>>(c->dm, c->bp, c->blist, c->nblist, c->bfirst, c->f[PCFD_F_Q],
                                                                 c->bpos, c->bdiag, A);
    }
    LAUNCH_CHECK();
  }
  if (c->viscous && c->nedge) {
    PROF("k_vjac_edges");
    k_vjac_edges<<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, c->vp, c->f[PCFD_F_Q], c->f[PCFD_F_MUT], c->posLR,
                                                             c->posRL, A);
    LAUNCH_CHECK();
  }
  PROF("k_jac_diag");
  k_jac_diag<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->iau, c->posLR, c->posRL, c->f[PCFD_F_TIMESTEP],
                                                         c->bdiag, time_cnp1(c), (c->time_local && c->time_dt > 0.0) ? c->time_dt : -1.0,
                                                         A);
  LAUNCH_CHECK();
  if (c->nwall) {
    PROF("k_jac_wall");
    k_jac_wall<<<nblk(c->nwall, 64), 64, 0, c->stream>>>(c->dm, c->prm.gamma, c->wnodes, c->nwall, c->ia, c->ja, c->iau, A);
    LAUNCH_CHECK();
  }
  return 0;
}

int pcfd_prepare_sgs(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (ensure_matrix(c)) return 1;
  if (c->fr) return pcfd_fr_prepare_sgs(c);
  if (c->ludiag) return 0;   // CRSMatrix::ludiag (crsmatrix.tcc:844)
  PROF("k_lu_diag");
  k_lu_diag_lanes<NEQN><<<nblk((long long)((c->nnode + 5) / 6) * 32, 128), 128, 0, c->stream>>>(c->nnode, c->iau, c->f[PCFD_F_A], c->pv);
  LAUNCH_CHECK();
  c->ludiag = true;
  return 0;
}

int pcfd_blank_x(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  CK(cudaMemsetAsync(c->f[PCFD_F_X], 0, c->fsize[PCFD_F_X] * sizeof(double), c->stream));
  return 0;
}

int pcfd_sgs(pcfd_ctx* c, int nsgs, double* ddq) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (ensure_matrix(c)) return 1;
  if (c->fr) return pcfd_fr_sgs(c, nsgs, ddq);
  constexpr int RPW = 32 / NEQN;
  const double* A = c->f[PCFD_F_A];
  double* x = c->f[PCFD_F_X];
  bool prev_tile = false;   // the previous launch of this call was a k_sgs_tile level
  for (int s = 0; s < nsgs; s++) {
    for (int dir = 0; dir < 2; dir++) {
      const std::vector<int>& off = dir ? c->lev_b : c->lev_f;
      const int* rows = dir ? c->rows_b : c->rows_f;
      for (size_t l = 0; l + 1 < off.size(); l++) {
        const int nr = off[l + 1] - off[l];
        const int warps = (nr + RPW - 1) / RPW;
        const int cap = (dir ? c->tile_cap_b : c->tile_cap_f)[l];
        if (cap > 0) {
          // bulk-copy streaming variant: one CTA per tile of tile_warps*6 consecutive rows
          if (c->sgs_ring_stages > 0) {
            const int LPRr = c->sgs_tile_lpr, St = c->sgs_ring_stages;
            const int RPWr = 32 / LPRr;
            const int stage_bytes = (c->ring_cap_blocks * NEQN2 * 8 + 8 + 15) & ~15;
            const size_t shm = ((St * 8 + 15) & ~15) + (size_t)St * stage_bytes;
            const int tiles = (nr + RPWr - 1) / RPWr;
            int per_sm = (int)std::min<size_t>(32, (size_t)(220 * 1024) / (shm + 1024));
            if (c->sgs_ring_ctas_per_sm > 0) per_sm = std::min(per_sm, c->sgs_ring_ctas_per_sm);
            const int grid = std::max(1, std::min(tiles, c->num_sms * std::max(per_sm, 1)));
            const int row0 = dir ? c->lev_first_b[l] : c->lev_first_f[l];
            const int step = dir ? c->lev_step_b[l] : c->lev_step_f[l];
            PROF("k_sgs_ring");
#define PCFD_RING_LAUNCH(LL, SS)                                                                                     \
  do {                                                                                                               \
    static size_t set_##LL##_##SS = 0;                                                                               \
    if (shm > set_##LL##_##SS) {                                                                                     \
      CK(cudaFuncSetAttribute(k_sgs_ring<LL, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));           \
      set_##LL##_##SS = shm;                                                                                         \
    }                                                                                                                \
    k_sgs_ring<LL, SS><<<grid, 32, shm, c->stream>>>(row0, step, nr, c->ia, c->ja, A, c->pv, c->f[PCFD_F_B], x,      \
                                                     stage_bytes);                                                   \
  } while (0)
#define PCFD_RING_LPR(SS)                                                                                            \
  do {                                                                                                               \
    if (LPRr == 5) PCFD_RING_LAUNCH(5, SS); else if (LPRr == 10) PCFD_RING_LAUNCH(10, SS); else PCFD_RING_LAUNCH(16, SS); \
  } while (0)
            if (St == 2) PCFD_RING_LPR(2); else if (St == 3) PCFD_RING_LPR(3); else if (St == 6) PCFD_RING_LPR(6); else PCFD_RING_LPR(4);
#undef PCFD_RING_LPR
#undef PCFD_RING_LAUNCH
            LAUNCH_CHECK();
            continue;
          }
          const int W = c->sgs_tile_warps, LPR = c->sgs_tile_lpr;
          const int RT = W * (32 / LPR);
          const int capA = (cap * NEQN2 * 8 + 8 + 15) & ~15;
          const size_t shm = 16 + (size_t)capA;
          const int tiles = (nr + RT - 1) / RT;
          // L2 prefetch distance in tiles: ~24 MB of matrix ahead of the compute front (measured best on B200; much
          // further and the prefetched lines are evicted before use)
          const int pf = c->sgs_pf_dist >= 0 ? c->sgs_pf_dist : (int)((size_t)(24 << 20) / shm);
          const int row0 = dir ? c->lev_first_b[l] : c->lev_first_f[l];
          const int step = dir ? c->lev_step_b[l] : c->lev_step_f[l];
          // programmatic dependent launch between consecutive levels of this call (never across other kernels)
          const bool chain = c->sgs_pdl && !c->prof && prev_tile;
          PROF("k_sgs_tile");
#define PCFD_TILE_LAUNCH(WW, LL)                                                                                      \
  do {                                                                                                                \
    static size_t set_##WW##_##LL = 0;                                                                                \
    if (shm > set_##WW##_##LL) {                                                                                      \
      CK(cudaFuncSetAttribute(k_sgs_tile_t<NEQN, WW, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));            \
      set_##WW##_##LL = shm;                                                                                          \
    }                                                                                                                 \
    CK(launch_maybe_pdl(k_sgs_tile_t<NEQN, WW, LL>, tiles, WW * 32, shm, c->stream, chain, row0, step, nr, c->ia, c->ja, A, \
                        c->pv, c->f[PCFD_F_B], x, pf));                                                               \
  } while (0)
          if (LPR == 5) { if (W == 1) PCFD_TILE_LAUNCH(1, 5); else if (W == 4) PCFD_TILE_LAUNCH(4, 5); else PCFD_TILE_LAUNCH(2, 5); }
          else if (LPR == 10) { if (W == 1) PCFD_TILE_LAUNCH(1, 10); else if (W == 4) PCFD_TILE_LAUNCH(4, 10); else PCFD_TILE_LAUNCH(2, 10); }
          else { if (W == 1) PCFD_TILE_LAUNCH(1, 16); else if (W == 4) PCFD_TILE_LAUNCH(4, 16); else PCFD_TILE_LAUNCH(2, 16); }
#undef PCFD_TILE_LAUNCH
          LAUNCH_CHECK();
          prev_tile = true;
          continue;
        }
        prev_tile = false;
        PROF("k_sgs_level");
        const int nb_ = nblk((long long)warps * 32, 128);
        switch (c->sgs_unroll) {
          case 1: k_sgs_level<1><<<nb_, 128, 0, c->stream>>>(rows + off[l], nr, c->ia, c->ja, c->iau, A, c->pv, c->f[PCFD_F_B], x); break;
          case 2: k_sgs_level<2><<<nb_, 128, 0, c->stream>>>(rows + off[l], nr, c->ia, c->ja, c->iau, A, c->pv, c->f[PCFD_F_B], x); break;
          case 7: k_sgs_level<7><<<nb_, 128, 0, c->stream>>>(rows + off[l], nr, c->ia, c->ja, c->iau, A, c->pv, c->f[PCFD_F_B], x); break;
          default: k_sgs_level<4><<<nb_, 128, 0, c->stream>>>(rows + off[l], nr, c->ia, c->ja, c->iau, A, c->pv, c->f[PCFD_F_B], x); break;
        }
        LAUNCH_CHECK();
      }
    }
    // xNorm of the last two sweeps only (crs.tcc:149-172 uses nothing else)
    if (ddq && s >= nsgs - 2) {
      prev_tile = false;
      PROF("k_sumsq_partial");
      k_sumsq_partial<256, NEQN><<<RED_BLOCKS, 256, 0, c->stream>>>(x, c->nnode, c->red);
      LAUNCH_CHECK();
      PROF("k_sumsq_final");
      k_sumsq_final<256, NEQN><<<1, 256, 0, c->stream>>>(c->red, RED_BLOCKS, c->redout + ((s == nsgs - 1) ? 0 : 8));
      LAUNCH_CHECK();
    }
  }
  if (ddq) {
    double h[16] = {0};
    CK(cudaMemcpyAsync(h, c->redout, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const double N = (double)c->nnode * NEQN;
    const double xNorm = (nsgs >= 1) ? sqrt(h[0]) / N : 0.0;
    const double xOld = (nsgs >= 2) ? sqrt(h[8]) / N : 0.0;
    *ddq = fabs(xOld - xNorm);
  }
  return 0;
}

int pcfd_halo_configure(pcfd_ctx* c, int rank, int nranks, const int* send_counts, const int* send_list,
                        const int* recv_counts) {
  if (!c) return 1;
  if (nranks < 1 || rank < 0 || rank >= nranks || !send_counts || !recv_counts) return fail(c, "pcfd_halo_configure: bad argument");
  CK(cudaSetDevice(c->device));
  c->rank = rank;
  c->nranks = nranks;
  c->send_counts.assign(send_counts, send_counts + nranks);
  c->recv_counts.assign(recv_counts, recv_counts + nranks);
  c->send_offsets.assign(nranks + 1, 0);
  c->recv_offsets.assign(nranks + 1, 0);
  for (int p = 0; p < nranks; p++) {
    if (send_counts[p] < 0 || recv_counts[p] < 0) return fail(c, "pcfd_halo_configure: negative count");
    c->send_offsets[p + 1] = c->send_offsets[p] + send_counts[p];
    c->recv_offsets[p + 1] = c->recv_offsets[p] + recv_counts[p];
  }
  if (c->recv_offsets[nranks] != c->gnode) return fail(c, "pcfd_halo_configure: receive counts do not add up to gnode");
  c->send_total = c->send_offsets[nranks];
  for (int j = 0; j < c->send_total; j++)
    if (!send_list || send_list[j] < 0 || send_list[j] >= c->nnode) return fail(c, "pcfd_halo_configure: send list entry is not a local node");
  if (dev_upload(c, &c->send_list, send_list, (size_t)c->send_total)) return 1;
  return 0;
}

int pcfd_halo_width(const pcfd_ctx* c, int field) { return c ? field_width(c, field) : 0; }
int pcfd_halo_send_total(const pcfd_ctx* c) { return c ? c->send_total : 0; }

/* pack the rows this rank owes peer `peer` (peer < 0: all peers, in rank order) into dst (device memory,
   count*width doubles) */
int pcfd_halo_pack(pcfd_ctx* c, int field, int peer, void* dst) {
  if (!c) return 1;
  const int n = field_width(c, field);
  if (n == 0 || !dst || peer >= c->nranks) return fail(c, "pcfd_halo_pack: bad argument");
  if (c->send_offsets.empty()) return fail(c, "pcfd_halo_pack: pcfd_halo_configure has not been called");
  CK(cudaSetDevice(c->device));
  const int off = peer < 0 ? 0 : c->send_offsets[peer];
  const int cnt = peer < 0 ? c->send_total : c->send_counts[peer];
  if (cnt == 0) return 0;
  PROF("k_halo_pack");
  k_halo_pack<<<nblk((long long)cnt * n, 256), 256, 0, c->stream>>>(c->send_list + off, cnt, n, c->f[field],
                                                                   static_cast<double*>(dst));
  LAUNCH_CHECK();
  return 0;
}

/* device address where the rows received from `peer` (peer < 0: from all peers) land: the ghost segment of the
   field is contiguous per owner (decomp.cpp:245-273), so receives go straight into it, no unpack */
void* pcfd_halo_recv_ptr(pcfd_ctx* c, int field, int peer) {
  if (!c) return nullptr;
  const int n = field_width(c, field);
  if (n == 0 || peer >= c->nranks || c->recv_offsets.empty()) return nullptr;
  const int off = peer < 0 ? 0 : c->recv_offsets[peer];
  return c->f[field] + ((size_t)c->nnode + off) * n;
}

/* CUDA IPC: let a peer process map this context's field so that its k_halo_pack can store straight into our
   ghost segment (one process per GPU; NVLink P2P).  handle is CUDA_IPC_HANDLE_SIZE (64) bytes. */
int pcfd_ipc_export(pcfd_ctx* c, int field, void* handle) {
  if (!c) return 1;
  if (field < 0 || field >= PCFD_F_COUNT || !handle || !c->f[field]) return fail(c, "pcfd_ipc_export: bad argument");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->f[field]));
  memcpy(handle, &h, sizeof(h));
  return 0;
}
int pcfd_ipc_open(pcfd_ctx* c, const void* handle, void** devptr) {
  if (!c) return 1;
  if (!handle || !devptr) return fail(c, "pcfd_ipc_open: bad argument");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  CK(cudaIpcOpenMemHandle(devptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int pcfd_ipc_close(pcfd_ctx* c, void* devptr) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  CK(cudaIpcCloseMemHandle(devptr));
  return 0;
}

// nsweeps symmetric Gauss-Seidel sweeps of the scalar (turbulence) system: one cooperative launch (k_sgs_scalar_sweeps)
// or, with PCFD_TURB_PERSIST=0 / while profiling per kernel, one launch per level
static int turb_sweeps(pcfd_ctx* c, int nsweeps, const double* tA, const double* tb, double* tx) {
  if (c->turb_persist && !c->prof) {
    if (c->turb_grid == 0) {
      int per_sm = 0, sms = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sgs_scalar_sweeps, 256, 0));
      CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
      c->turb_grid = std::max(1, std::min(per_sm, 4)) * sms;
    }
    int nlev_f = (int)c->lev_f.size() - 1, nlev_b = (int)c->lev_b.size() - 1;
    const int *rows_f = c->rows_f, *rows_b = c->rows_b, *lev_f = c->dlev_f, *lev_b = c->dlev_b, *ia = c->ia, *ja = c->ja;
    void* args[] = {&rows_f, &rows_b, &lev_f, &nlev_f, &lev_b, &nlev_b, &nsweeps, &ia, &ja, &tA, &tb, &tx};
    PROF("k_sgs_scalar_sweeps");
    CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_sgs_scalar_sweeps), dim3(c->turb_grid), dim3(256), args, 0, c->stream));
    LAUNCH_CHECK();
    return 0;
  }
  bool chained = false;   // the first level follows an ordinary kernel: plain launch
  for (int s = 0; s < nsweeps; s++) {
    for (int dir = 0; dir < 2; dir++) {
      const std::vector<int>& off = dir ? c->lev_b : c->lev_f;
      const int* rows = dir ? c->rows_b : c->rows_f;
      for (size_t l = 0; l + 1 < off.size(); l++) {
        const int nr = off[l + 1] - off[l];
        PROF("k_sgs_scalar_level");
        CK(launch_maybe_pdl(k_sgs_scalar_level, nblk(nr, 128), 128, 0, c->stream, c->sgs_pdl && !c->prof && chained,
                            rows + off[l], nr, c->ia, c->ja, tA, tb, tx));
        chained = true;
        LAUNCH_CHECK();
      }
    }
  }
  return 0;
}

// the Spalart-Allmaras kernels for the context's eqnset (q only matters to the perfect-gas instantiation; the reacting one
// reads the tables pcfd_fr_turb_props filled from the same q, so the two must not be separated by a change of q)
static TurbGas turb_gas(const pcfd_ctx* c) {
  TurbGas g;
  if (c->fr) {
    g.Re = c->prm.Re;
    g.gstride = c->nterms * 3;
    g.goff = (c->neqn - 4) * 3;   // GetVelocityGradLocation() = nspecies (compressibleFR.tcc:681)
  } else {
    g.Re = c->vp.Re / c->vp.mach;
    g.gstride = NTERMS * 3;
    g.goff = 3;
  }
  g.pe = c->tprop_e; g.pb = c->tprop_b; g.pn = c->tprop_n;
  return g;
}
static void turb_launch_edges(pcfd_ctx* c, const double* tvar, const double* tgrad, double* tA) {
  if (c->fr)
    k_turb_edges<true><<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, c->vp, turb_gas(c), c->f[PCFD_F_Q], tvar, tgrad, c->posLR,
                                                                   c->posRL, c->tslots, tA);
  else
    k_turb_edges<false><<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, c->vp, turb_gas(c), c->f[PCFD_F_Q], tvar, tgrad, c->posLR,
                                                                    c->posRL, c->tslots, tA);
}
static void turb_launch_bedges(pcfd_ctx* c, const double* tvar, const double* tgrad, double* tA) {
  if (c->fr)
    k_turb_bedges<true><<<nblk(c->nb, 128), 128, 0, c->stream>>>(c->dm, c->vp, turb_gas(c), c->f[PCFD_F_Q], tvar, tgrad, c->bpos,
                                                                 c->tbslots, tA);
  else
    k_turb_bedges<false><<<nblk(c->nb, 128), 128, 0, c->stream>>>(c->dm, c->vp, turb_gas(c), c->f[PCFD_F_Q], tvar, tgrad, c->bpos,
                                                                  c->tbslots, tA);
}
static void turb_launch_node(pcfd_ctx* c, const double* tvar, double* tb, double* tA) {
  if (c->fr)
    k_turb_node<true><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->vp, turb_gas(c), c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD], tvar,
                                                                  c->f[PCFD_F_WALLDIST], c->f[PCFD_F_TIMESTEP], c->tslots,
                                                                  c->tbslots, c->iau, c->posLR, c->posRL, tb, tA);
  else
    k_turb_node<false><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->vp, turb_gas(c), c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD], tvar,
                                                                   c->f[PCFD_F_WALLDIST], c->f[PCFD_F_TIMESTEP], c->tslots,
                                                                   c->tbslots, c->iau, c->posLR, c->posRL, tb, tA);
}
static void turb_launch_mut(pcfd_ctx* c, const double* tvar) {
  if (c->fr)
    k_turb_mut<true><<<nblk(c->nn, 256), 256, 0, c->stream>>>(c->nn, c->vp, turb_gas(c), c->f[PCFD_F_Q], tvar, c->f[PCFD_F_MUT]);
  else
    k_turb_mut<false><<<nblk(c->nn, 256), 256, 0, c->stream>>>(c->nn, c->vp, turb_gas(c), c->f[PCFD_F_Q], tvar, c->f[PCFD_F_MUT]);
}

int pcfd_turb_compute(pcfd_ctx* c, int nsgs, double* sumsq) {
  if (!c) return 1;
  if (c->prm.turb_model != 1) return fail(c, "pcfd_turb_compute: the context was created without a turbulence model");
  CK(cudaSetDevice(c->device));
  if (comm_on(c)) {
    // across ranks: the phases of pcfd_turb_phase with the reference's exchanges (turb.tcc:185, gradient.tcc:98,
    // crs.tcc:146, turb.tcc:325); block-Jacobi across partitions like CRS::SGS; sumsq = this rank's sum of b^2
    if (nsgs <= 0) return fail(c, "pcfd_turb_compute: nsgs == 0 (explicit turbulence update) is not implemented");
    if (pcfd_turb_phase(c, 0, nullptr) || comm_update(c, PCFD_F_TVAR)) return 1;
    if (pcfd_turb_phase(c, 1, nullptr) || comm_update(c, PCFD_F_TGRAD)) return 1;
    if (pcfd_turb_phase(c, 2, sumsq)) return 1;
    for (int s = 0; s < nsgs; s++)
      if (pcfd_turb_phase(c, 3, nullptr) || comm_update(c, PCFD_F_TURB_X)) return 1;
    if (pcfd_turb_phase(c, 4, nullptr) || comm_update(c, PCFD_F_TVAR)) return 1;
    return pcfd_turb_phase(c, 5, nullptr);
  }
  double *tvar = c->f[PCFD_F_TVAR], *tgrad = c->f[PCFD_F_TGRAD], *tb = c->f[PCFD_F_TURB_B], *tx = c->f[PCFD_F_TURB_X],
         *tA = c->f[PCFD_F_TURB_A];
  // crs.BlankSystem (crs.tcc:417-425); every interior off-diagonal and every diagonal entry is overwritten below
  CK(cudaMemsetAsync(tA, 0, c->fsize[PCFD_F_TURB_A] * sizeof(double), c->stream));
  CK(cudaMemsetAsync(tx, 0, c->fsize[PCFD_F_TURB_X] * sizeof(double), c->stream));
  if (c->ntbnodes) {
    PROF("k_turb_bcs");
    k_turb_bcs<<<nblk(c->ntbnodes, 128), 128, 0, c->stream>>>(c->dm, c->tbnodes, c->ntbnodes, tvar);
    LAUNCH_CHECK();
  }
  PROF("k_turb_gradient");
  k_turb_gradient<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, tvar, c->f[PCFD_F_LSQ_S], tgrad);
  LAUNCH_CHECK();
  if (c->fr && pcfd_fr_turb_props(c)) return 1;
  if (c->nedge) {
    PROF("k_turb_edges");
    turb_launch_edges(c, tvar, tgrad, tA);
    LAUNCH_CHECK();
  }
  if (c->nb) {
    PROF("k_turb_bedges");
    turb_launch_bedges(c, tvar, tgrad, tA);
    LAUNCH_CHECK();
  }
  PROF("k_turb_node");
  turb_launch_node(c, tvar, tb, tA);
  LAUNCH_CHECK();
  if (c->nwall) {
    PROF("k_turb_wall");
    k_turb_wall<<<nblk(c->nwall, 128), 128, 0, c->stream>>>(c->wnodes, c->nwall, c->ia, c->iau, tb, tx, tA);
    LAUNCH_CHECK();
  }
  if (sumsq) {
    PROF("k_sumsq_partial");
    k_sumsq_partial<256, 1><<<RED_BLOCKS, 256, 0, c->stream>>>(tb, c->nnode, c->red);
    LAUNCH_CHECK();
    PROF("k_sumsq_final");
    k_sumsq_final<256, 1><<<1, 256, 0, c->stream>>>(c->red, RED_BLOCKS, c->redout);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(sumsq, c->redout, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  if (nsgs > 0) {
    PROF("k_turb_invdiag");
    k_turb_invdiag<<<nblk(c->nnode, 256), 256, 0, c->stream>>>(c->nnode, c->iau, tA, tb);
    LAUNCH_CHECK();
    if (turb_sweeps(c, nsgs, tA, tb, tx)) return 1;
  } else {
    return fail(c, "pcfd_turb_compute: nsgs == 0 (explicit turbulence update) is not implemented");
  }
  PROF("k_turb_update");
  k_turb_update<<<nblk(c->nnode, 256), 256, 0, c->stream>>>(c->nnode, tx, tvar);
  LAUNCH_CHECK();
  PROF("k_turb_mut");
  turb_launch_mut(c, tvar);
  LAUNCH_CHECK();
  if (sumsq) CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// TurbulenceModel::Compute cut at the reference's exchange points (turb.tcc:183-325) for runs on partitions: the same
// kernels, in the same order, as pcfd_turb_compute; the caller exchanges the named field after each phase.
//   0  blank the system, turbulence BCs                                  -> halo of tvar        (turb.tcc:185)
//   1  LSQ gradient of tvar                                              -> halo of tgrad       (gradient.tcc:98)
//   2  convective / diffusive / source terms, wall rows, sum of b^2 (sumsq, may be NULL), inverse diagonal
//   3  ONE symmetric Gauss-Seidel sweep (forward + backward)             -> halo of turb_x      (crs.tcc:146)
//   4  tvar += x                                                         -> halo of tvar        (turb.tcc:325)
//   5  eddy viscosity of the local and ghost nodes
int pcfd_turb_phase(pcfd_ctx* c, int phase, double* sumsq) {
  if (!c) return 1;
  if (c->prm.turb_model != 1) return fail(c, "pcfd_turb_phase: the context was created without a turbulence model");
  CK(cudaSetDevice(c->device));
  double *tvar = c->f[PCFD_F_TVAR], *tgrad = c->f[PCFD_F_TGRAD], *tb = c->f[PCFD_F_TURB_B], *tx = c->f[PCFD_F_TURB_X],
         *tA = c->f[PCFD_F_TURB_A];
  switch (phase) {
    case 0:
      CK(cudaMemsetAsync(tA, 0, c->fsize[PCFD_F_TURB_A] * sizeof(double), c->stream));
      CK(cudaMemsetAsync(tx, 0, c->fsize[PCFD_F_TURB_X] * sizeof(double), c->stream));
      if (c->ntbnodes) {
        PROF("k_turb_bcs");
        k_turb_bcs<<<nblk(c->ntbnodes, 128), 128, 0, c->stream>>>(c->dm, c->tbnodes, c->ntbnodes, tvar);
        LAUNCH_CHECK();
      }
      return 0;
    case 1:
      PROF("k_turb_gradient");
      k_turb_gradient<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, tvar, c->f[PCFD_F_LSQ_S], tgrad);
      LAUNCH_CHECK();
      return 0;
    case 2:
      if (c->fr && pcfd_fr_turb_props(c)) return 1;
      if (c->nedge) {
        PROF("k_turb_edges");
        turb_launch_edges(c, tvar, tgrad, tA);
        LAUNCH_CHECK();
      }
      if (c->nb) {
        PROF("k_turb_bedges");
        turb_launch_bedges(c, tvar, tgrad, tA);
        LAUNCH_CHECK();
      }
      PROF("k_turb_node");
      turb_launch_node(c, tvar, tb, tA);
      LAUNCH_CHECK();
      if (c->nwall) {
        PROF("k_turb_wall");
        k_turb_wall<<<nblk(c->nwall, 128), 128, 0, c->stream>>>(c->wnodes, c->nwall, c->ia, c->iau, tb, tx, tA);
        LAUNCH_CHECK();
      }
      if (sumsq) {
        PROF("k_sumsq_partial");
        k_sumsq_partial<256, 1><<<RED_BLOCKS, 256, 0, c->stream>>>(tb, c->nnode, c->red);
        LAUNCH_CHECK();
        PROF("k_sumsq_final");
        k_sumsq_final<256, 1><<<1, 256, 0, c->stream>>>(c->red, RED_BLOCKS, c->redout);
        LAUNCH_CHECK();
        CK(cudaMemcpyAsync(sumsq, c->redout, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      }
      PROF("k_turb_invdiag");
      k_turb_invdiag<<<nblk(c->nnode, 256), 256, 0, c->stream>>>(c->nnode, c->iau, tA, tb);
      LAUNCH_CHECK();
      if (sumsq) CK(cudaStreamSynchronize(c->stream));
      return 0;
    case 3:
      return turb_sweeps(c, 1, tA, tb, tx);
    case 4:
      PROF("k_turb_update");
      k_turb_update<<<nblk(c->nnode, 256), 256, 0, c->stream>>>(c->nnode, tx, tvar);
      LAUNCH_CHECK();
      return 0;
    case 5:
      PROF("k_turb_mut");
      turb_launch_mut(c, tvar);
      LAUNCH_CHECK();
      return 0;
    default:
      return fail(c, "pcfd_turb_phase: phase must be 0..5");
  }
}

// One iteration of SolutionSpace::NewtonIterate (solutionSpace.tcc:640-904).  Across ranks (pcfd_comm_connect) the
// reference's halo exchanges run in the reference's places; sumsq / ddq stay THIS rank's sums (the host finishes the
// norms with pcfd_comm_allgather, as ParallelL2Norm does with MPI_Allreduce).
int pcfd_explicit_iterate(pcfd_ctx* c, int refresh_dt, double* sumsq) {
  if (!c) return 1;
  const bool dist = comm_on(c);
  // ComputeTimesteps folded into the residual pass: possible when UpdateBCs does not touch owned nodes (no
  // Dirichlet-type half-edges: the interior-edge states the flux kernel reads are then the ones ComputeTimesteps
  // would have read) and the step is the local CFL one; the half-edge terms are taken here, before UpdateBCs
  const bool ride = refresh_dt && !c->fr && c->eig_fuse && c->nbn == 0 && c->time_local && c->eig;
  if (refresh_dt && !ride && pcfd_timestep(c, nullptr)) return 1;
  if (ride && c->nb) {
    CK(cudaSetDevice(c->device));
    PROF("k_eig_bedges");
    k_eig_bedges<<<nblk(c->nb, 128), 128, 0, c->stream>>>(c->dm, c->prm.gamma, c->f[PCFD_F_Q], c->beig);
    LAUNCH_CHECK();
  }
  if (pcfd_update_bcs(c)) return 1;
  if (dist && comm_q_after_bcs(c)) return 1;                       // solutionSpace.tcc:665
  c->eig_now = ride;
  const int rc = gradient_limiter_residual(c, sumsq);
  c->eig_now = false;
  if (rc) return 1;
  if (pcfd_explicit_solve(c)) return 1;
  if (dist && comm_update(c, PCFD_F_Q)) return 1;                  // :857
  return 0;
}

int pcfd_implicit_iterate(pcfd_ctx* c, int refresh_jac, int nsgs, double* sumsq, double* ddq) {
  if (!c) return 1;
  const bool dist = comm_on(c);
  if (refresh_jac) {
    if (pcfd_timestep(c, nullptr)) return 1;   // PreIterate / PreTimeAdvance (solutionSpace.tcc:510, 629-634)
    if (pcfd_jacobian(c)) return 1;
  }
  if (pcfd_update_bcs(c)) return 1;
  if (dist && comm_q_after_bcs(c)) return 1;                       // solutionSpace.tcc:665
  if (gradient_limiter_residual(c, sumsq)) return 1;
  if (pcfd_prepare_sgs(c)) return 1;
  if (pcfd_blank_x(c)) return 1;
  if (!dist) {
    if (pcfd_sgs(c, nsgs, ddq)) return 1;
  } else {
    // CRS::SGS across ranks (crs.tcc:62-173): block-Jacobi at partition boundaries, a halo of x before the first
    // sweep (:88) and after every sweep (:146); ddq from the last two sweeps of this rank's rows
    // crs.tcc:88: BlankX has just zeroed x, ghost rows included, on every rank: the halo would move zeros onto zeros
    if (!c->comm->elide && comm_update(c, PCFD_F_X)) return 1;
    for (int sweep = 0; sweep < nsgs; sweep++) {
      const bool last = sweep == nsgs - 1;
      if (ddq && sweep >= nsgs - 2 && nsgs >= 2) {
        // every sweep is its own pcfd_sgs call here (the halo of x sits between them), so |xOld - xNorm| is put
        // together from the norms of the closing two sweeps
        double d1 = 0.0;
        if (pcfd_sgs(c, 1, &d1)) return 1;          // d1 = |0 - xNorm| of this sweep
        if (!last) c->sgs_prev_norm = d1; else *ddq = fabs(c->sgs_prev_norm - d1);
      } else {
        double d1 = 0.0;
        if (pcfd_sgs(c, 1, (ddq && last) ? &d1 : nullptr)) return 1;
        if (ddq && last) *ddq = d1;
      }
      if (comm_update(c, PCFD_F_X)) return 1;
    }
  }
  if (pcfd_apply_dq(c)) return 1;
  if (dist && comm_update(c, PCFD_F_Q)) return 1;                  // solutionSpace.tcc:857
  // NewtonIterate updates the turbulence model after the flow update (solutionSpace.tcc:862-866)
  if (c->prm.turb_model == 1) return pcfd_turb_compute(c, nsgs, nullptr);
  return 0;
}

}  // extern "C"

#include "pcfd_crsmatrix.cuh"

int pcfd_internal_create(const pcfd_mesh_desc* mesh, const pcfd_params* params, int device, int neqn, int nvars,
                         int nterms, pcfd_ctx** out) {
  return create_impl(mesh, params, device, neqn, nvars, nterms, out);
}
