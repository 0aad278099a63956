// pcfd_fr.cu -- the reacting eqnset (CompressibleFREqnSet, equationSet = compressibleEulerFR) on the hot path:
// CUDA kernels (sm_100a, FP64) and the host side of every phase entry point for contexts made by pcfd_create_fr.
//
// Same design as pcfd_kernels.cu: expensive per-edge work (HLLC flux, the 19-flux finite-difference Jacobian) is done
// once per edge by edge-parallel kernels that write private slots; every reduction onto a node is an ordered gather
// by the threads that own the node, in the reference's edge order, so there are no atomics and the floating-point
// summation order is the reference's.  The kernels are templates on the species count NS (NEQ = NS+4 equations per
// node, NEQ x NEQ blocks); `fr_dispatch` instantiates the sizes the library ships with.
//
// What differs from the perfect-gas path because the blocks are 9x9 and the state is 21 doubles wide:
//  * rows of a node are shared by NEQ consecutive threads wherever the work is per (node, equation): limiter,
//    residual gather, Jacobian diagonal, SGS -- consecutive lanes then read consecutive doubles of a block row;
//  * the FD Jacobian of an edge is spread over 2*NEQ+1 = 19 lanes, one HLLC flux each (reference state, NEQ left and
//    NEQ right perturbations), which write the two off-diagonal blocks column by column.

#include "pcfd_internal.cuh"
#include "eqnset_fr.cuh"
#include "sgs_tile.cuh"

struct pcfd_fr_state {
  pcfd_fr_params host{};
  pcfd_chem_model* chem_dev = nullptr;
  double *eig = nullptr, *beig = nullptr;   // spectral radius x area per edge / half-edge (time step)
  double* src = nullptr;                    // source term rows (species only) per node
  double *red = nullptr, *redout = nullptr;
  int bad_newton = 0;
  int* dbad = nullptr;
  // compressibleNSFR: per-edge / per-half-edge viscous flux slots (momentum + energy rows)
  bool viscous = false;
  double *vflux = nullptr, *bvflux = nullptr;
  unsigned char* negflag = nullptr;         // nodes whose raw limiter has a negative component (fused clip test)
};

// Register caps: measured on B200 at 10 M cells (tools/time_frjac.py, profiles/r2_ncu_frjac.md).  A cap that doubles the resident
// warps pays where a kernel waits on FP64 latency with 8 warps per SM (kfr_jac_bedges 18.6 -> 14.6 ms at 128 registers,
// kfr_jac_node, kfr_vflux_edges, kfr_vjac_edges, kfr_update_bcs_edges a few per cent); kfr_flux_edges / kfr_clip_edges /
// kfr_gradient are best left to ptxas' own choice (96 registers: 2.32 ms, 128: 2.22 ms, 212: 3.17 ms for the flux kernel).

namespace {

constexpr int FR_RED_BLOCKS = 296;

template <int NS>
struct W {
  static constexpr int NEQ = NS + 4, NV = 3 * NS + 6, NT = 2 * NS + 4, N2 = NEQ * NEQ;
};

// GetGradientsLocation (compressibleFR.tcc:693-711)
template <int NS>
__device__ __forceinline__ int gradloc(int i) { return (i < NS + 4) ? i : (NS + NS + 6 + (i - (NS + 4))); }

template <int NS, int CNT>
__device__ __forceinline__ void load_row(const double* __restrict__ q, int n, double* v) {
  const double* p = q + (size_t)n * W<NS>::NV;
#pragma unroll
  for (int i = 0; i < CNT; i++) v[i] = p[i];
}

// ============================================================ boundary conditions
// UpdateBCs (bc.tcc:1399-1457): one thread per BC half-edge.  Two half-edges of a node interact only through
// ComputeAuxiliaryVariables(QL) at the end of each call: the first half-edge sees the stored aux values, every later
// one sees them recomputed -- which a thread reproduces locally (same argument as k_update_bcs_edges).
template <int NS>
__global__ void __launch_bounds__(64, 8) kfr_update_bcs_edges(DevMesh m, fr::Params<NS> p, const int* __restrict__ list, int n,
                                                            const unsigned char* __restrict__ bfirst,
                                                            const double* __restrict__ beta, double* q) {
  constexpr int NV = W<NS>::NV;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int be = list[t];
  const int2 lr = m.ben[be];
  const bool first = bfirst[be] != 0;
  double QL[NV], QR[NV], av[4];
  load_row<NS, NV>(q, lr.x, QL);
  load_row<NS, NV>(q, lr.y, QR);
  load_avec(m.bea, be, av);
  if (!first) fr::aux(p, QL);
  fr::boundary_variables(p, QL, QR, av, m.bctype[be], beta[lr.x], (m.bctype[be] == PCFD_BC_FARFIELD_VISCOUS) ? m.bubar[be] : 1.0);
  double* qr = q + (size_t)lr.y * NV;
  for (int i = 0; i < NV; i++) qr[i] = QR[i];
  if (first) {
    double* ql = q + (size_t)lr.x * NV;
    for (int i = NS + 4; i < NV; i++) ql[i] = QL[i];
  }
}

// The same update for nodes owning a Dirichlet-type half-edge (Dirichlet, sonic inflow, no-slip: they rewrite QL itself):
// one thread per such node, its BC half-edges in half-edge order -- the reference's sequence, since two half-edges
// interact only through their shared left node (k_update_bcs of the perfect-gas path)
template <int NS>
__global__ void __launch_bounds__(64) kfr_update_bcs_nodes(DevMesh m, fr::Params<NS> p, const int* __restrict__ bnodes, int nb,
                                                            const double* __restrict__ beta, double* q) {
  constexpr int NV = W<NS>::NV;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb) return;
  const int n = bnodes[t];
  double QL[NV];
  load_row<NS, NV>(q, n, QL);
  bool touched = false;
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    const int type = m.bctype[be];
    if (type == PCFD_BC_PARALLEL) continue;
    const int r = a.x & 0x7fffffff;
    double QR[NV], av[4];
    load_row<NS, NV>(q, r, QR);
    load_avec(m.bea, be, av);
    double nT = 0.0, tw = 0.0;
    if (type == PCFD_BC_NOSLIP) { nT = q[(size_t)m.bnormal[be] * NV + NS + 3]; tw = m.btwall[be]; }
    fr::boundary_variables_seq(p, QL, QR, av, type, beta[n], nT, tw, (type == PCFD_BC_FARFIELD_VISCOUS) ? m.bubar[be] : 1.0);
    double* qr = q + (size_t)r * NV;
    for (int i = 0; i < NV; i++) qr[i] = QR[i];
    touched = true;
  }
  if (touched) {
    double* ql = q + (size_t)n * NV;
    for (int i = 0; i < NV; i++) ql[i] = QL[i];
  }
}

// ====================================================================== gradient
// Gradient::Compute (gradient.tcc:57-112, weighted LSQ kernels :251-378, symmetry fix :545-565): ordered gather
template <int NS>
__global__ void __launch_bounds__(128) kfr_gradient(DevMesh m, const double* __restrict__ q, const double* __restrict__ sw,
                                                     double* __restrict__ qgrad) {
  constexpr int NV = W<NS>::NV, NT = W<NS>::NT;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double g[NT * 3], qn[NT], swn[6];
#pragma unroll
  for (int k = 0; k < NT * 3; k++) g[k] = 0.0;
#pragma unroll
  for (int i = 0; i < NT; i++) qn[i] = q[(size_t)n * NV + gradloc<NS>(i)];
#pragma unroll
  for (int k = 0; k < 6; k++) swn[k] = sw[6 * (size_t)n + k];
  const double xn[3] = {m.xyz[3 * n], m.xyz[3 * n + 1], m.xyz[3 * n + 2]};
  const int kbeg = m.adjp[n], kend = m.adjp[n + 1];
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    const int o = a.x & 0x7fffffff;
    const bool right = a.x < 0;
    if (a.y >= m.nedge && !is_ghost(m, o)) continue;
    const double xo[3] = {__ldg(m.xyz + 3 * o), __ldg(m.xyz + 3 * o + 1), __ldg(m.xyz + 3 * o + 2)};
    double dx[3], we[3];
#pragma unroll
    for (int d = 0; d < 3; d++) dx[d] = right ? (xo[d] - xn[d]) : (xn[d] - xo[d]);
    const double dx2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    const double weight = 1.0 / sqrt(dx2);
    dx[0] *= weight; dx[1] *= weight; dx[2] *= weight;
    if (right) { dx[0] = -dx[0]; dx[1] = -dx[1]; dx[2] = -dx[2]; }
    lsq_weights(swn, dx, we);
#pragma unroll
    for (int i = 0; i < NT; i++) {
      const double qo = __ldg(q + (size_t)o * NV + gradloc<NS>(i));
      const double dq = right ? weight * (qn[i] - qo) : weight * (qo - qn[i]);
#pragma unroll
      for (int j = 0; j < 3; j++) {
        if (right) g[3 * i + j] += +we[j] * dq;
        else g[3 * i + j] += -we[j] * dq;
      }
    }
  }
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    if (m.bctype[be] != PCFD_BC_SYMMETRY) continue;
    double av[4];
    load_avec(m.bea, be, av);
#pragma unroll
    for (int i = 0; i < NT; i++) {
      const double dot = g[i * 3] * av[0] + g[i * 3 + 1] * av[1] + g[i * 3 + 2] * av[2];
#pragma unroll
      for (int j = 0; j < 3; j++) g[i * 3 + j] -= dot * av[j];
    }
  }
  double* out = qgrad + (size_t)n * NT * 3;
#pragma unroll
  for (int kk = 0; kk < NT * 3; kk++) out[kk] = g[kk];
}

// Gradient::Compute with Param::gradType == 1: Green-Gauss (gradient.tcc:77-90, kernels :170-248) -- interior edges, then
// all half-edges, the division by the dual volume, the symmetry fix; ordered gather (see k_gradient_gg)
template <int NS>
__global__ void __launch_bounds__(128) kfr_gradient_gg(DevMesh m, const double* __restrict__ q, double* __restrict__ qgrad) {
  constexpr int NV = W<NS>::NV, NT = W<NS>::NT;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double g[NT * 3], qn[NT];
#pragma unroll
  for (int k = 0; k < NT * 3; k++) g[k] = 0.0;
#pragma unroll
  for (int i = 0; i < NT; i++) qn[i] = q[(size_t)n * NV + gradloc<NS>(i)];
  const int kbeg = m.adjp[n], kend = m.adjp[n + 1];
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    const int o = a.x & 0x7fffffff;
    const bool right = a.x < 0;
    double av[4];
    if (a.y < m.nedge) load_avec(m.ea, a.y, av);
    else load_avec(m.bea, a.y - m.nedge, av);
    const double area = av[3];
#pragma unroll
    for (int i = 0; i < NT; i++) {
      const double qo = __ldg(q + (size_t)o * NV + gradloc<NS>(i));
      const double faceavg = 0.5 * (qn[i] + qo);
#pragma unroll
      for (int j = 0; j < 3; j++) {
        if (right) g[3 * i + j] += -faceavg * av[j] * area;
        else g[3 * i + j] += faceavg * av[j] * area;
      }
    }
  }
  const double vol = m.vol[n];
#pragma unroll
  for (int kk = 0; kk < NT * 3; kk++) g[kk] /= vol;
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    if (m.bctype[be] != PCFD_BC_SYMMETRY) continue;
    double av[4];
    load_avec(m.bea, be, av);
#pragma unroll
    for (int i = 0; i < NT; i++) {
      const double dot = g[i * 3] * av[0] + g[i * 3 + 1] * av[1] + g[i * 3 + 2] * av[2];
#pragma unroll
      for (int j = 0; j < 3; j++) g[i * 3 + j] -= dot * av[j];
    }
  }
  double* out = qgrad + (size_t)n * NT * 3;
#pragma unroll
  for (int kk = 0; kk < NT * 3; kk++) out[kk] = g[kk];
}

// ======================================================================= limiter
// Limiter::Compute passes 1+2 (limiters.tcc:53-110), NEQ threads per node (one equation each); unclamped output
template <int NS>
__global__ void __launch_bounds__(W<NS>::NEQ * 16) kfr_limiter(DevMesh m, int type, double chi, const double* __restrict__ q,
                                                                const double* __restrict__ qgrad, double* __restrict__ lim) {
  constexpr int NEQ = W<NS>::NEQ, NV = W<NS>::NV, NT = W<NS>::NT;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = tid / NEQ;
  if (n >= m.nnode + m.gnode) return;
  const int j = tid - n * NEQ;
  double l = 1.0;
  if (n < m.nnode && (type == 1 || type == 2 || type == 3)) {
    double qmin = 0.0, qmax = 0.0;
    const int k0 = m.adjp[n], k1 = m.adjp[n + 1];
    for (int k = k0; k < k1; k++) {
      const int2 a = m.adj[k];
      const int o = a.x & 0x7fffffff;
      if (a.y >= m.nedge && !is_ghost(m, o)) continue;
      const double qo = __ldg(q + (size_t)o * NV + j);
      qmax = fr::maxd(qmax, qo);
      qmin = fr::mind(qmin, qo);
    }
    const double qn = __ldg(q + (size_t)n * NV + j);
    const double* gp = qgrad + (size_t)n * NT * 3 + j * 3;
    const double g0 = gp[0], g1 = gp[1], g2 = gp[2];
    const double xn[3] = {m.xyz[3 * n], m.xyz[3 * n + 1], m.xyz[3 * n + 2]};
    for (int k = k0; k < k1; k++) {
      const int2 a = m.adj[k];
      const int o = a.x & 0x7fffffff;
      if (a.y >= m.nedge && !is_ghost(m, o)) continue;
      const double qo = __ldg(q + (size_t)o * NV + j);
      const double dQ = qo - qn;
      double dx[3];
#pragma unroll
      for (int d = 0; d < 3; d++) dx[d] = 0.5 * (__ldg(m.xyz + 3 * o + d) - xn[d]);
      const double corr = (0.5 * chi * dQ + (1.0 - chi) * (g0 * dx[0] + g1 * dx[1] + g2 * dx[2]));
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
      l = fr::mind(l, t);
    }
  }
  lim[(size_t)n * NEQ + j] = l;
}

// Kernel_PressureClip (limiters.tcc:737-815) as the fixed point on "first edge that zeroes node n" (see
// k_clip_edges in pcfd_kernels.cu); no Roe-state test for this eqnset (EqnSet::RoeVariables returns false)
template <int NS>
__global__ void __launch_bounds__(128) kfr_clip_edges(DevMesh m, fr::Params<NS> p, const double* __restrict__ q,
                                                       const double* __restrict__ qgrad, const double* __restrict__ lim,
                                                       const int* __restrict__ tclip, unsigned char* __restrict__ flag,
                                                       int* __restrict__ any) {
  constexpr int NEQ = W<NS>::NEQ, NT = W<NS>::NT;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const int2 lr = m.en[e];
  const int l = lr.x, r = lr.y;
  double qL[NEQ], qR[NEQ], lm[NEQ], dQ[NEQ], dx[3], Qx[NS + 6], gr[NEQ * 3];
  load_row<NS, NEQ>(q, l, qL);
  load_row<NS, NEQ>(q, r, qR);
  const bool zl = tclip[l] < e, zr = tclip[r] < e;
#pragma unroll
  for (int d = 0; d < 3; d++) dx[d] = 0.5 * (__ldg(m.xyz + 3 * r + d) - __ldg(m.xyz + 3 * l + d));
#pragma unroll
  for (int j = 0; j < NEQ; j++) { dQ[j] = qR[j] - qL[j]; lm[j] = zl ? 0.0 : lim[(size_t)l * NEQ + j]; }
#pragma unroll
  for (int j = 0; j < NEQ * 3; j++) gr[j] = __ldg(qgrad + (size_t)l * NT * 3 + j);
  fr::extrapolate<NS>(p.chi, Qx, qL, dQ, gr, dx, lm);
  fr::aux_pr(p, Qx);
  const bool cl = fr::bad_extrapolation(p, Qx);
#pragma unroll
  for (int d = 0; d < 3; d++) dx[d] = -dx[d];
#pragma unroll
  for (int j = 0; j < NEQ; j++) { dQ[j] = -dQ[j]; lm[j] = zr ? 0.0 : lim[(size_t)r * NEQ + j]; }
#pragma unroll
  for (int j = 0; j < NEQ * 3; j++) gr[j] = __ldg(qgrad + (size_t)r * NT * 3 + j);
  fr::extrapolate<NS>(p.chi, Qx, qR, dQ, gr, dx, lm);
  fr::aux_pr(p, Qx);
  const bool cr = fr::bad_extrapolation(p, Qx);
  const unsigned char f = (cl ? 1 : 0) | (cr ? 2 : 0);
  flag[e] = f;
  if (f) *any = 1;
}

__global__ void kfr_clip_nodes(DevMesh m, const unsigned char* __restrict__ flag, const int* __restrict__ told,
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

__global__ void kfr_limiter_final(int ntot, int nnode, int neqn, const int* __restrict__ tclip, double* __restrict__ lim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot * neqn) return;
  const int n = i / neqn;
  double v = lim[i];
  if (tclip != nullptr && n < nnode && tclip[n] != INT_MAX) v = 0.0;
  if (v < 0.0) v = 0.0;
  lim[i] = v;
}

__global__ void kfr_fill_int(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ====================================================================== residual
// MUSCL reconstruction of one side pair (Kernel_Inviscid_Flux, residual.tcc:192-296): QL / QR hold [0, NS+6)
// CLAMP: `lim` is the raw limiter; negative components are clamped on the fly (limiters.tcc:118-125) and reported
template <int NS, bool CLAMP = false>
__device__ __forceinline__ bool reconstruct(const DevMesh& m, const fr::Params<NS>& p, int l, int r,
                                            const double* __restrict__ qgrad, const double* __restrict__ lim, double* QL,
                                            double* QR) {
  constexpr int NEQ = W<NS>::NEQ, NT = W<NS>::NT;
  bool neg = false;
  double dQ[NEQ], dx[3], gr[NEQ * 3], lm[NEQ], qL[NEQ], qR[NEQ];
#pragma unroll
  for (int j = 0; j < NEQ; j++) { qL[j] = QL[j]; qR[j] = QR[j]; dQ[j] = qR[j] - qL[j]; }
#pragma unroll
  for (int d = 0; d < 3; d++) dx[d] = 0.5 * (__ldg(m.xyz + 3 * r + d) - __ldg(m.xyz + 3 * l + d));
#pragma unroll
  for (int j = 0; j < NEQ * 3; j++) gr[j] = __ldg(qgrad + (size_t)l * NT * 3 + j);
#pragma unroll
  for (int j = 0; j < NEQ; j++) {
    lm[j] = __ldg(lim + (size_t)l * NEQ + j);
    if (CLAMP && lm[j] < 0.0) { lm[j] = 0.0; neg = true; }
  }
  fr::extrapolate<NS>(p.chi, QL, qL, dQ, gr, dx, lm);
#pragma unroll
  for (int j = 0; j < NEQ; j++) dQ[j] = -dQ[j];
#pragma unroll
  for (int d = 0; d < 3; d++) dx[d] = -dx[d];
#pragma unroll
  for (int j = 0; j < NEQ * 3; j++) gr[j] = __ldg(qgrad + (size_t)r * NT * 3 + j);
#pragma unroll
  for (int j = 0; j < NEQ; j++) {
    lm[j] = __ldg(lim + (size_t)r * NEQ + j);
    if (CLAMP && lm[j] < 0.0) { lm[j] = 0.0; neg = true; }
  }
  fr::extrapolate<NS>(p.chi, QR, qR, dQ, gr, dx, lm);
  return neg;
}

// DET = true is the fused fast path of the composite iterations (as k_flux_edges<true> of the perfect-gas path): `lim`
// holds the RAW limiter of kfr_limiter (before Kernel_PressureClip and the clamp of negatives, limiters.tcc:105-125).
// The flux uses the clamped value; the pressure-clip test of kfr_clip_edges (BadExtrapolation of the two extrapolated
// states) is evaluated on the way -- the same states unless a raw component was negative; those edges are left to
// kfr_clip_neg_edges.  If no edge anywhere raises *any, clip(lim) == clamp(lim) and this flux is final; otherwise the
// caller discards it and takes the ordered clip path.
template <int NS, bool DET>
__global__ void __launch_bounds__(128) kfr_flux_edges(DevMesh m, fr::Params<NS> p, const double* __restrict__ q,
                                                       const double* __restrict__ qgrad, const double* __restrict__ lim,
                                                       const double* __restrict__ beta, double* __restrict__ flux,
                                                       int* __restrict__ any) {
  constexpr int NEQ = W<NS>::NEQ;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const int2 lr = m.en[e];
  const int l = lr.x, r = lr.y;
  double av[4], QL[NS + 6], QR[NS + 6], f[NEQ], parts[4];
  load_avec(m.ea, e, av);
  load_row<NS, NS + 6>(q, l, QL);
  load_row<NS, NS + 6>(q, r, QR);
  const double avbeta = 0.5 * (beta[l] + beta[r]);
  bool test = false;
  if (p.sorder > 1) {
    if (DET) {
      // where a raw component is negative (local extrema: common) the clip test needs the raw-limiter states and the
      // flux the clamped ones: those few edges are tested by kfr_clip_neg_edges instead, with a per-node flag
      test = !reconstruct<NS, true>(m, p, l, r, qgrad, lim, QL, QR);
    } else {
      reconstruct<NS, false>(m, p, l, r, qgrad, lim, QL, QR);
    }
    fr::aux_pr(p, QL);
    fr::aux_pr(p, QR);
  }
  fr::numerical_flux(p, QL, QR, av, 0.0, avbeta, f, DET ? parts : nullptr);
  // BadExtrapolation of both states, with the enthalpy terms the flux has just formed
  if (DET && test && (fr::bad_extrapolation(p, QL, parts[0], parts[1]) || fr::bad_extrapolation(p, QR, parts[2], parts[3]))) *any = 1;
#pragma unroll
  for (int j = 0; j < NEQ; j++) flux[(size_t)e * NEQ + j] = f[j];
}

// nodes whose raw limiter has a negative component (owned and ghost rows)
__global__ void kfr_negflag_nodes(int nn, int neqn, const double* __restrict__ lim, unsigned char* __restrict__ negflag) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  bool neg = false;
  for (int j = 0; j < neqn; j++) neg = neg || (lim[(size_t)n * neqn + j] < 0.0);
  negflag[n] = neg ? 1 : 0;
}

// the pressure-clip test of kfr_clip_edges (first pass: nothing zeroed yet) for the edges that touch a flagged node,
// with the RAW limiter as Kernel_PressureClip sees it (limiters.tcc:737-815)
template <int NS>
__global__ void __launch_bounds__(128) kfr_clip_neg_edges(DevMesh m, fr::Params<NS> p, const double* __restrict__ q,
                                                           const double* __restrict__ qgrad, const double* __restrict__ lim,
                                                           const unsigned char* __restrict__ negflag, int* __restrict__ any) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const int2 lr = m.en[e];
  if (!(negflag[lr.x] | negflag[lr.y])) return;
  double QL[NS + 6], QR[NS + 6];
  load_row<NS, NS + 6>(q, lr.x, QL);
  load_row<NS, NS + 6>(q, lr.y, QR);
  reconstruct<NS, false>(m, p, lr.x, lr.y, qgrad, lim, QL, QR);
  fr::aux_pr(p, QL);
  fr::aux_pr(p, QR);
  if (fr::bad_extrapolation(p, QL) || fr::bad_extrapolation(p, QR)) *any = 1;
}

// Bkernel_Inviscid_Flux (residual.tcc:299-387)
template <int NS>
__global__ void __launch_bounds__(128) kfr_flux_bedges(DevMesh m, fr::Params<NS> p, const double* __restrict__ q,
                                                        const double* __restrict__ qgrad, const double* __restrict__ lim,
                                                        const double* __restrict__ beta, double* __restrict__ bflux) {
  constexpr int NEQ = W<NS>::NEQ;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nbedge + m.ngedge) return;
  const int2 lr = m.ben[e];
  const int l = lr.x, r = lr.y;
  double av[4], QL[NS + 6], QR[NS + 6], f[NEQ];
  load_avec(m.bea, e, av);
  load_row<NS, NS + 6>(q, l, QL);
  load_row<NS, NS + 6>(q, r, QR);
  if (p.sorder > 1) {
    if (is_ghost(m, r)) reconstruct<NS>(m, p, l, r, qgrad, lim, QL, QR);
    fr::aux_pr(p, QL);
    fr::aux_pr(p, QR);
  }
  fr::numerical_flux(p, QL, QR, av, 0.0, beta[l], f);
#pragma unroll
  for (int j = 0; j < NEQ; j++) bflux[(size_t)e * NEQ + j] = f[j];
}

// Kernel_Viscous_Flux / Bkernel_Viscous_Flux (residual.tcc:390-562) with CompressibleFREqnSet::ViscousFlux: one thread
// per interior edge (e < nedge) or half-edge; the face gradient (average + directional correction, :430-452) is formed
// for the four terms the flux reads (u, v, w, T); writes the momentum and energy rows of the edge's private slot
template <int NS>
__device__ __forceinline__ void face_gradient_uvwT(const DevMesh& m, int l, int r, bool corrected, const double* QL,
                                                   const double* QR, const double* __restrict__ qgrad, double* g) {
  constexpr int NT = W<NS>::NT;
  const double* gl = qgrad + (size_t)l * NT * 3 + NS * 3;
  const double* gr = qgrad + (size_t)r * NT * 3 + NS * 3;
#pragma unroll
  for (int i = 0; i < 12; i++) g[i] = 0.5 * (__ldg(gl + i) + __ldg(gr + i));
  if (corrected) {
    double dx[3], s2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      dx[i] = (__ldg(m.xyz + 3 * r + i) - __ldg(m.xyz + 3 * l + i));
      s2 += dx[i] * dx[i];
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const double qdots = dx[0] * g[j * 3] + dx[1] * g[j * 3 + 1] + dx[2] * g[j * 3 + 2];
      const double dq = (QR[NS + j] - QL[NS + j] - qdots) / s2;
#pragma unroll
      for (int i = 0; i < 3; i++) g[j * 3 + i] += dq * dx[i];
    }
  }
}

template <int NS>
__global__ void __launch_bounds__(128, 5) kfr_vflux_edges(DevMesh m, fr::Params<NS> p, fr::Transport<NS> t,
                                                        const double* __restrict__ q, const double* __restrict__ qgrad,
                                                        const double* __restrict__ mut, double* __restrict__ vflux,
                                                        double* __restrict__ bvflux) {
  constexpr int NT = W<NS>::NT;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nb = m.nbedge + m.ngedge;
  if (idx >= m.nedge + nb) return;
  const bool interior = idx < m.nedge;
  const int e = interior ? idx : idx - m.nedge;
  const int2 lr = interior ? m.en[e] : m.ben[e];
  const int l = lr.x, r = lr.y;
  double av[4], QL[NS + 4], QR[NS + 4], Q[NS + 4], g[12], f[4], tmut;
  load_avec(interior ? m.ea : m.bea, e, av);
  load_row<NS, NS + 4>(q, l, QL);
  load_row<NS, NS + 4>(q, r, QR);
#pragma unroll
  for (int i = 0; i < NS + 4; i++) Q[i] = 0.5 * (QL[i] + QR[i]);
  if (interior || is_ghost(m, r)) {
    tmut = 0.5 * (mut[l] + mut[r]);
    face_gradient_uvwT<NS>(m, l, r, p.sorder > 1, QL, QR, qgrad, g);
  } else {
    tmut = mut[l];
    const double* gl = qgrad + (size_t)l * NT * 3 + NS * 3;
#pragma unroll
    for (int i = 0; i < 12; i++) g[i] = __ldg(gl + i);
  }
  fr::viscous_flux(p, t, Q, g, av, tmut, f);
  double* dst = interior ? vflux + (size_t)e * 4 : bvflux + (size_t)e * 4;
#pragma unroll
  for (int j = 0; j < 4; j++) dst[j] = f[j];
}

// SourceTerm (compressibleFR.tcc:1276-1316) per node -> src[n*NS]
template <int NS>
__global__ void __launch_bounds__(128) kfr_source(int nnode, fr::Params<NS> p, const double* __restrict__ q,
                                                   const double* __restrict__ vol, double* __restrict__ src) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnode) return;
  double Q[NS + 4], s[NS];
  load_row<NS, NS + 4>(q, n, Q);
  fr::source_term(p, Q, vol[n], s);
#pragma unroll
  for (int i = 0; i < NS; i++) src[(size_t)n * NS + i] = s[i];
}

// DriverScatter (driver.tcc:274-306) as an ordered gather, then the source term (residual.tcc:109-115); NEQ threads
// per node, one equation each
template <int NS>
__global__ void __launch_bounds__(W<NS>::NEQ * 16) kfr_residual_gather(DevMesh m, const double* __restrict__ flux,
                                                                        const double* __restrict__ bflux,
                                                                        const double* __restrict__ vflux,
                                                                        const double* __restrict__ bvflux,
                                                                        const double* __restrict__ src,
                                                                        const unsigned char* __restrict__ wallflag,
                                                                        double* __restrict__ b) {
  constexpr int NEQ = W<NS>::NEQ;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = tid / NEQ;
  if (n >= m.nnode) return;
  const int j = tid - n * NEQ;
  double acc = 0.0;
  const int k0 = m.adjp[n], k1 = m.adjp[n + 1];
  for (int k = k0; k < k1; k++) {
    const int2 a = m.adj[k];
    const double f = (a.y < m.nedge) ? __ldg(flux + (size_t)a.y * NEQ + j) : __ldg(bflux + (size_t)(a.y - m.nedge) * NEQ + j);
    acc += (a.x < 0) ? f : -f;
  }
  // the viscous passes come after both inviscid ones (residual.tcc:88-104); their species rows are zero
  if (vflux && j >= NS) {
    for (int k = k0; k < k1; k++) {
      const int2 a = m.adj[k];
      const double f = (a.y < m.nedge) ? __ldg(vflux + (size_t)a.y * 4 + (j - NS)) : __ldg(bvflux + (size_t)(a.y - m.nedge) * 4 + (j - NS));
      acc += (a.x < 0) ? f : -f;
    }
  }
  acc += (j < NS) ? src[(size_t)n * NS + j] : 0.0;
  // Bkernel_BC_Res_Modify -> ModifyViscousWallResidual (compressibleFR.tcc:2101-2114): momentum and energy rows of no-slip nodes
  if (wallflag && j >= NS && (wallflag[n] & 1)) acc = 0.0;
  b[(size_t)n * NEQ + j] = acc;
}

// TemporalResidual (residual.tcc:125-179, no GCL): the stored state is native, the equations are written for the
// conservative variables (NativeToConservative, compressibleFR.tcc:2117-2130); qold / qoldm1 hold conservative rows
template <int NS>
__global__ void __launch_bounds__(128) kfr_temporal_residual(int nnode, fr::Params<NS> p, double cnp1, double cnm1,
                                                              const double* __restrict__ vol, const double* __restrict__ q,
                                                              const double* __restrict__ qold, const double* __restrict__ qoldm1,
                                                              const unsigned char* __restrict__ wallflag, double* __restrict__ b) {
  constexpr int NEQ = W<NS>::NEQ, NV = W<NS>::NV;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnode) return;
  double Q[NS + 6];
  load_row<NS, NS + 6>(q, n, Q);
  fr::native_to_conservative(p, Q);
  const double dt = cnp1 * vol[n] / p.dt;
  const double dtm1 = cnm1 * vol[n] / p.dt;
#pragma unroll
  for (int j = 0; j < NEQ; j++) {
    const double qo = qold[(size_t)n * NV + j];
    const double dq = Q[j] - qo;
    const double dqm1 = qo - qoldm1[(size_t)n * NV + j];
    double v = b[(size_t)n * NEQ + j];
    v -= dt * dq;
    v -= dtm1 * dqm1;
    if (wallflag && j >= NS && (wallflag[n] & 1)) v = 0.0;   // the wall hook runs after the temporal terms (residual.tcc:40-43)
    b[(size_t)n * NEQ + j] = v;
  }
}

// ====================================================================== time step
// Kernel_Timestep / Bkernel_Timestep (timestep.tcc:80-143): spectral radius x area of the averaged state, per edge
template <int NS>
__global__ void __launch_bounds__(128) kfr_eig_edges(DevMesh m, fr::Params<NS> p, const double* __restrict__ q,
                                                      const double* __restrict__ beta, double* __restrict__ eig,
                                                      double* __restrict__ beig) {
  constexpr int NEQ = W<NS>::NEQ;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nb = m.nbedge + m.ngedge;
  if (t >= m.nedge + nb) return;
  const bool interior = t < m.nedge;
  const int2 lr = interior ? m.en[t] : m.ben[t - m.nedge];
  double av[4], QL[NEQ], QR[NEQ], Q[NS + 6];
  load_avec(interior ? m.ea : m.bea, interior ? t : t - m.nedge, av);
  load_row<NS, NEQ>(q, lr.x, QL);
  load_row<NS, NEQ>(q, lr.y, QR);
#pragma unroll
  for (int i = 0; i < NEQ; i++) Q[i] = 0.5 * (QL[i] + QR[i]);
  fr::aux_pr(p, Q);
  const double bta = interior ? 0.5 * (beta[lr.x] + beta[lr.y]) : beta[lr.x];
  const double me = fr::max_eigenvalue(p, Q, av, 0.0, bta);
  if (interior) eig[t] = me * av[3];
  else beig[t - m.nedge] = me * av[3];
}

// What TurbulenceModel::Compute (turb.tcc:163-339) asks of the eqnset, evaluated once per call for the Spalart-Allmaras
// kernels of pcfd_kernels.cu (which are eqnset-agnostic beyond these numbers): per interior edge / half-edge theta =
// GetTheta of the averaged native state (compressibleFR.tcc:1640-1650) and nu = ComputeViscosity / GetDensity of the averaged
// state after ComputeAuxiliaryVariables (turb.tcc:600-612, :684-692; Wilke-mixed species viscosity, :1572-1587); per local
// and ghost node rho and nu of the stored state (turb.tcc:213-217, 329-334).  props_e / props_b: {theta, nu}; props_n: {rho, nu}.
template <int NS>
__global__ void __launch_bounds__(128) kfr_turb_props(DevMesh m, fr::Params<NS> p, fr::Transport<NS> t,
                                                       const double* __restrict__ q, double* __restrict__ props_e,
                                                       double* __restrict__ props_b, double* __restrict__ props_n) {
  constexpr int NEQ = W<NS>::NEQ, NV = W<NS>::NV;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nb = m.nbedge + m.ngedge, nn = m.nnode + m.gnode;
  if (i < m.nedge + nb) {
    const bool interior = i < m.nedge;
    const int2 lr = interior ? m.en[i] : m.ben[i - m.nedge];
    double av[4], QL[NEQ], QR[NEQ], Q[NS + 6], mu, kc;
    load_avec(interior ? m.ea : m.bea, interior ? i : i - m.nedge, av);
    load_row<NS, NEQ>(q, lr.x, QL);
    load_row<NS, NEQ>(q, lr.y, QR);
#pragma unroll
    for (int k = 0; k < NEQ; k++) Q[k] = 0.5 * (QL[k] + QR[k]);
    const double theta = fr::theta_of<NS>(Q, av, 0.0);
    fr::aux_pr(p, Q);
    fr::mixture_transport(p, t, Q, Q[NS + 3], mu, kc);
    double* out = interior ? props_e + 2 * (size_t)i : props_b + 2 * (size_t)(i - m.nedge);
    out[0] = theta;
    out[1] = mu / Q[NS + 5];
    return;
  }
  const int n = i - (m.nedge + nb);
  if (n >= nn) return;
  const double* Q = q + (size_t)n * NV;
  double rhoi[NS], mu, kc;
#pragma unroll
  for (int k = 0; k < NS; k++) rhoi[k] = Q[k];
  fr::mixture_transport(p, t, rhoi, Q[NS + 3], mu, kc);
  const double rho = Q[NS + 5];
  props_n[2 * (size_t)n] = rho;
  props_n[2 * (size_t)n + 1] = mu / rho;
}

// What Forces asks of the eqnset for the surface node of a BC half-edge (pcfd_forces.cuh): GetPressure, GetCp
// (compressibleFR.tcc:672-678: (P - Pinf) / (V^2 / 2)), ComputeViscosity (Wilke-mixed) and GetDensity
template <int NS>
__global__ void __launch_bounds__(128) kfr_surface_props(DevMesh m, fr::Params<NS> p, fr::Transport<NS> t, double V, bool viscous,
                                                          const double* __restrict__ q, double* __restrict__ props) {
  constexpr int NV = W<NS>::NV;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nbedge) return;
  const double* Q = q + (size_t)m.ben[e].x * NV;
  const double P = Q[NS + 4], Pinf = p.qinf[NS + 4];
  double mu = 0.0, kc;
  if (viscous) {
    double rhoi[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) rhoi[k] = Q[k];
    fr::mixture_transport(p, t, rhoi, Q[NS + 3], mu, kc);
  }
  props[4 * (size_t)e] = P;
  props[4 * (size_t)e + 1] = ((P - Pinf) / (0.5 * V * V));
  props[4 * (size_t)e + 2] = mu;
  props[4 * (size_t)e + 3] = Q[NS + 5];
}

// ComputeTimesteps (timestep.tcc:7-49): dt = CFL * vol / sum, VNN limit from node 1 on
__global__ void __launch_bounds__(128) kfr_timestep(DevMesh m, double cfl, const double* __restrict__ eig,
                                                     const double* __restrict__ beig, const double* __restrict__ vnn23,
                                                     double* __restrict__ dt) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  double s = 0.0;
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    s += (a.y < m.nedge) ? __ldg(eig + a.y) : __ldg(beig + (a.y - m.nedge));
  }
  double d = cfl * (m.vol[n] / s);
  if (vnn23 != nullptr && n > 0) d = fr::mind(d, vnn23[n]);
  dt[n] = d;
}

__global__ void kfr_min_partial(const double* __restrict__ v, int n, double* __restrict__ part) {
  __shared__ double sh[256];
  double a = INFINITY;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) a = fmin(a, v[i]);
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void kfr_min_final(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double sh[256];
  double a = INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) a = fmin(a, part[i]);
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// sum of squares per column and in total: out[0] = total, out[1 + c] = column c
__global__ void kfr_sumsq_partial(const double* __restrict__ v, int nrows, int stride, double* __restrict__ part) {
  __shared__ double sh[256];
  for (int c = 0; c < stride; c++) {
    double a = 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nrows; i += gridDim.x * 256) { const double t = v[(size_t)i * stride + c]; a += t * t; }
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * stride + c] = sh[0];
    __syncthreads();
  }
}
__global__ void kfr_sumsq_final(const double* __restrict__ part, int nparts, int stride, double* __restrict__ out) {
  __shared__ double sh[256];
  double total = 0.0;
  for (int c = 0; c < stride; c++) {
    double a = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) a += part[(size_t)i * stride + c];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) { out[1 + c] = sh[0]; total += sh[0]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = total;
}

// ======================================================================== update
// ExplicitSolve, native-variable branch (solve.tcc:112-130): q is left untouched, x = change of the native variables
template <int NS>
__global__ void __launch_bounds__(128) kfr_explicit(int nnode, fr::Params<NS> p, const double* __restrict__ b,
                                                     const double* __restrict__ dt, const double* __restrict__ vol,
                                                     const double* __restrict__ q, double* __restrict__ x,
                                                     int* __restrict__ bad) {
  constexpr int NEQ = W<NS>::NEQ;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnode) return;
  double Q[NS + 6], qc[NEQ];
  load_row<NS, NS + 6>(q, n, Q);
#pragma unroll
  for (int j = 0; j < NEQ; j++) qc[j] = Q[j];
  fr::native_to_conservative(p, Q);
  const double d = dt[n], v = vol[n];
#pragma unroll
  for (int j = 0; j < NEQ; j++) Q[j] += b[(size_t)n * NEQ + j] * d / v;
  if (!fr::conservative_to_native(p, Q)) atomicAdd(bad, 1);
#pragma unroll
  for (int j = 0; j < NEQ; j++) x[(size_t)n * NEQ + j] = Q[j] - qc[j];
}

// the ApplyDQ loop of NewtonIterate (solutionSpace.tcc:802-804)
template <int NS>
__global__ void __launch_bounds__(128) kfr_apply_dq(int nnode, fr::Params<NS> p, double* __restrict__ x, double* q,
                                                     int* __restrict__ zeroed) {
  constexpr int NEQ = W<NS>::NEQ, NV = W<NS>::NV;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnode) return;
  double Q[NV], dQ[NEQ];
  load_row<NS, NEQ>(q, n, Q);
  bool bad = false;
#pragma unroll
  for (int j = 0; j < NEQ; j++) { dQ[j] = x[(size_t)n * NEQ + j]; bad = bad || !isfinite(dQ[j]); }
  if (bad) {   // solutionSpace.tcc:771-796: the whole update of the node is zeroed, in crs->x as well
#pragma unroll
    for (int j = 0; j < NEQ; j++) { dQ[j] = 0.0; x[(size_t)n * NEQ + j] = 0.0; }
    atomicAdd(zeroed, 1);
  }
  fr::apply_dq(p, dQ, Q);
  double* out = q + (size_t)n * NV;
#pragma unroll
  for (int j = 0; j < NV; j++) out[j] = Q[j];
}

// ====================================================================== Jacobian
// Kernel_NumJac (jacobian.tcc:254-304): one-sided finite differences (h = 1e-8) of the FIRST-ORDER flux: column i of
// A(r,l) = (F_S - F_L,i)/h, column i of A(l,r) = (F_R,i - F_S)/h, 2*NEQ+1 HLLC fluxes per edge.  Most of a flux is
// thermodynamics of ONE state (speed of sound, enthalpy: ~25 IEEE divisions) or of the Roe average of the densities
// and temperatures (~10), and most of the 19 evaluations repeat them: a perturbation of the left state leaves the
// right state's alone, a velocity perturbation leaves all three alone.  So per edge
//   phase A   2*(NS+2) one-state evaluations    (unperturbed, rho_i + h, T + h; per side)
//   phase B   2*(NS+1)+1 Roe-average evaluations (unperturbed, rho_i + h or T + h on one side)
//   phase C0  the unperturbed flux, phase C the 2*NEQ perturbed ones (hllc_assemble from the tables of A and B)
// i.e. 14 + 13 + 19 small pieces instead of 19 x (2 + 1 + 1) for five species; the values are the ones the plain
// evaluation computes (same operations on the same inputs), so A is bit-identical (tests/test_gpu_fr.py).
// Work distribution: a warp owns G consecutive edges and walks each phase's task list 32 tasks at a time (G = 16: 7 / 7
// / 1 / 9 passes with full warps except the last), tables in shared memory, __syncwarp between phases -- no block
// barrier: a block-wide form of the same phases (one thread per task, __syncthreads) spent its time in barrier stalls
// (profiles/r2_ncu_frjac.md).
template <int NS>
__device__ __forceinline__ void jac_perturb(double* Q, int idx, double h) {
#pragma unroll
  for (int i = 0; i < NS + 4; i++) if (i == idx) Q[i] += h;   // selects, not a dynamically indexed (local-memory) array
}
template <int NS, int G, int MINB = 4>
__global__ void __launch_bounds__(128, MINB)
    kfr_jac_edges(DevMesh m, fr::Params<NS> p, const double* __restrict__ q, const double* __restrict__ beta,
                  const int* __restrict__ posLR, const int* __restrict__ posRL, double* __restrict__ A) {
  constexpr int NEQ = W<NS>::NEQ, N2 = W<NS>::N2;
  constexpr int NTH = NS + 2;              // one-state evaluations per side: 0 unperturbed, 1..NS rho_i + h, NS+1 T + h
  constexpr int NROE = 2 * (NS + 1) + 1;   // Roe-average evaluations: 0 unperturbed, then left j = 1..NS+1, then right
  constexpr int WPB = 4;
  __shared__ double sC2[WPB][G][2 * NTH], sHr[WPB][G][2 * NTH], sRoe[WPB][G][NROE], fS[WPB][G][NEQ];
  const double h = 1.0e-8;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e0 = (blockIdx.x * WPB + w) * G;
  if (e0 >= m.nedge) return;
  const int ne = (m.nedge - e0 < G) ? m.nedge - e0 : G;
  for (int t = lane; t < ne * 2 * NTH; t += 32) {
    const int slot = t / (2 * NTH), k = t - slot * (2 * NTH);
    const int side = (k >= NTH) ? 1 : 0, j = k - side * NTH;
    const int2 lr = m.en[e0 + slot];
    double Q[NS + 6], c2, hr;
    load_row<NS, NS + 6>(q, side ? lr.y : lr.x, Q);
    jac_perturb<NS>(Q, (j == 0) ? -1 : ((j <= NS) ? j - 1 : NS + 3), h);
    if (j) fr::aux_pr(p, Q);
    fr::hllc_side_thermo(p, Q, c2, hr);
    sC2[w][slot][k] = c2;
    sHr[w][slot][k] = hr;
  }
  for (int t = lane; t < ne * NROE; t += 32) {
    const int slot = t / NROE, k = t - slot * NROE;
    const int2 lr = m.en[e0 + slot];
    double QL[NS + 6], QR[NS + 6];
    load_row<NS, NS + 6>(q, lr.x, QL);
    load_row<NS, NS + 6>(q, lr.y, QR);
    if (k >= 1 && k <= NS + 1) { jac_perturb<NS>(QL, (k <= NS) ? k - 1 : NS + 3, h); fr::aux_pr(p, QL); }
    else if (k > NS + 1) { const int j = k - (NS + 1); jac_perturb<NS>(QR, (j <= NS) ? j - 1 : NS + 3, h); fr::aux_pr(p, QR); }
    sRoe[w][slot][k] = fr::hllc_roe_c2(p, QL, QR);
  }
  __syncwarp();
  if (lane < ne) {   // the unperturbed flux of each edge
    const int slot = lane, e = e0 + slot;
    const int2 lr = m.en[e];
    double av[4], QL[NS + 6], QR[NS + 6], f[NEQ];
    load_avec(m.ea, e, av);
    load_row<NS, NS + 6>(q, lr.x, QL);
    load_row<NS, NS + 6>(q, lr.y, QR);
    const double avbeta = 0.5 * (beta[lr.x] + beta[lr.y]);
    fr::hllc_assemble(p, QL, QR, av, 0.0, avbeta, sC2[w][slot][0], sHr[w][slot][0], sC2[w][slot][NTH], sHr[w][slot][NTH],
                      sRoe[w][slot][0], f);
#pragma unroll
    for (int j = 0; j < NEQ; j++) fS[w][slot][j] = f[j];
  }
  __syncwarp();
  for (int t = lane; t < ne * 2 * NEQ; t += 32) {
    const int slot = t / (2 * NEQ), role = t - slot * (2 * NEQ);   // role < NEQ: left column role; else right column role - NEQ
    const int e = e0 + slot;
    const bool right = role >= NEQ;
    const int i = right ? role - NEQ : role;
    const int2 lr = m.en[e];
    double av[4], QL[NS + 6], QR[NS + 6], QP[NS + 6], f[NEQ];
    load_avec(m.ea, e, av);
    load_row<NS, NS + 6>(q, lr.x, QL);
    load_row<NS, NS + 6>(q, lr.y, QR);
    const double avbeta = 0.5 * (beta[lr.x] + beta[lr.y]);
#pragma unroll
    for (int k = 0; k < NS + 6; k++) QP[k] = right ? QR[k] : QL[k];
    jac_perturb<NS>(QP, i, h);
    fr::aux_pr(p, QP);
#pragma unroll
    for (int k = 0; k < NS + 6; k++) { if (right) QR[k] = QP[k]; else QL[k] = QP[k]; }
    // which table entries the perturbed side reads; velocity perturbations read the unperturbed ones
    const int jt = (i < NS) ? i + 1 : ((i == NS + 3) ? NS + 1 : 0);
    const int jl = right ? 0 : jt, jr = right ? jt : 0;
    const int kroe = (jt == 0) ? 0 : (right ? NS + 1 + jt : jt);
    fr::hllc_assemble(p, QL, QR, av, 0.0, avbeta, sC2[w][slot][jl], sHr[w][slot][jl], sC2[w][slot][NTH + jr],
                      sHr[w][slot][NTH + jr], sRoe[w][slot][kroe], f);
    double* dst = A + (size_t)(right ? posLR[e] : posRL[e]) * N2 + i;
#pragma unroll
    for (int j = 0; j < NEQ; j++) {
      const double fs = fS[w][slot][j];
      const double num = right ? (f[j] - fs) : (fs - f[j]);
      dst[j * NEQ] = 0.0 + num / h;
    }
  }
}

// Kernel_NumJac_Centered (jacobian.tcc:306-366), Param::fieldJacType == 1: 2*NEQ lanes per edge, two HLLC fluxes each
// (column i perturbed by +h and by -h): lanes 0..NEQ-1 column i of A(r,l) = (F(qL-h) - F(qL+h))/2h, lanes NEQ..2NEQ-1
// column i of A(l,r) = (F(qR+h) - F(qR-h))/2h.
template <int NS, int EPB>
__global__ void __launch_bounds__(2 * W<NS>::NEQ * EPB) kfr_jac_edges_central(DevMesh m, fr::Params<NS> p,
                                                                               const double* __restrict__ q,
                                                                               const double* __restrict__ beta,
                                                                               const int* __restrict__ posLR,
                                                                               const int* __restrict__ posRL,
                                                                               double* __restrict__ A) {
  constexpr int NEQ = W<NS>::NEQ, N2 = W<NS>::N2, LPE = 2 * NEQ;
  const int slot = threadIdx.x / LPE;
  const int role = threadIdx.x - slot * LPE;
  const int e = blockIdx.x * EPB + slot;
  if (e >= m.nedge) return;
  const double h = 1.0e-8;
  const int2 lr = m.en[e];
  const int l = lr.x, r = lr.y;
  const bool left = role < NEQ;
  const int i = left ? role : role - NEQ;
  double av[4], QL[NS + 6], QR[NS + 6], QP[NS + 6], fp[NEQ], fm[NEQ];
  load_avec(m.ea, e, av);
  load_row<NS, NS + 6>(q, l, QL);
  load_row<NS, NS + 6>(q, r, QR);
  const double avbeta = 0.5 * (beta[l] + beta[r]);
#pragma unroll
  for (int k = 0; k < NS + 6; k++) QP[k] = left ? QL[k] : QR[k];
  QP[i] += h;
  fr::aux_pr(p, QP);
  if (left) fr::numerical_flux(p, QP, QR, av, 0.0, avbeta, fp);
  else fr::numerical_flux(p, QL, QP, av, 0.0, avbeta, fp);
#pragma unroll
  for (int k = 0; k < NS + 6; k++) QP[k] = left ? QL[k] : QR[k];
  QP[i] -= h;
  fr::aux_pr(p, QP);
  if (left) fr::numerical_flux(p, QP, QR, av, 0.0, avbeta, fm);
  else fr::numerical_flux(p, QL, QP, av, 0.0, avbeta, fm);
  if (left) {
    double* dst = A + (size_t)posRL[e] * N2 + i;
#pragma unroll
    for (int j = 0; j < NEQ; j++) dst[j * NEQ] = 0.0 + (fm[j] - fp[j]) / (2.0 * h);
  } else {
    double* dst = A + (size_t)posLR[e] * N2 + i;
#pragma unroll
    for (int j = 0; j < NEQ; j++) dst[j * NEQ] = 0.0 + (fp[j] - fm[j]) / (2.0 * h);
  }
}

// Kernel_Viscous_Jac (jacobian.tcc:728-767) with CompressibleFREqnSet::ViscousJacobian: one thread per edge, the part both
// block sides share (Wilke-mixed viscosity / conductivity and cp of the averaged state: most of the work) evaluated once,
// then side 0 (aR added to A(l,r)) and side 1 (-aL added to A(r,l)), after the inviscid finite-difference pass.  The
// species rows of both blocks are zero and are left alone.  (Bkernel_Viscous_Jac ends with size = 0: no boundary part.)
// (Two threads per edge, each with its own copy of the shared part: 10.5 ms at 10 M cells; this form 7.4 ms at its natural 166
// registers, 8.6 ms capped at 128, 18.1 ms capped at 96 -- spills.)
template <int NS>
__global__ void __launch_bounds__(128, 3) kfr_vjac_edges(DevMesh m, fr::Params<NS> p, fr::Transport<NS> t,
                                                          const double* __restrict__ q, const double* __restrict__ mut,
                                                          const int* __restrict__ posLR, const int* __restrict__ posRL,
                                                          double* __restrict__ A) {
  constexpr int NEQ = W<NS>::NEQ, N2 = W<NS>::N2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nedge) return;
  const int2 lr = m.en[e];
  const int l = lr.x, r = lr.y;
  double av[4], QL[NS + 6], QR[NS + 6], dx[3], D[3], s2 = 0.0, a[4 * NEQ];
  load_avec(m.ea, e, av);
  load_row<NS, NS + 6>(q, l, QL);
  load_row<NS, NS + 6>(q, r, QR);
  const double tmut = (mut[l] + mut[r]) / 2.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    dx[i] = (__ldg(m.xyz + 3 * r + i) - __ldg(m.xyz + 3 * l + i));
    s2 += dx[i] * dx[i];
  }
  fr::ViscJacCommon<NS> C;
  fr::viscous_jac_common(p, t, QL, QR, av, tmut, C);
  {
#pragma unroll
    for (int i = 0; i < 3; i++) D[i] = dx[i] / s2;
    fr::viscous_jac_side(p, C, D, QR, av, a);
    double* dst = A + (size_t)posLR[e] * N2 + NS * NEQ;
#pragma unroll
    for (int k = 0; k < 4 * NEQ; k++) dst[k] += a[k];
  }
  {
#pragma unroll
    for (int i = 0; i < 3; i++) D[i] = -dx[i] / s2;
    fr::viscous_jac_side(p, C, D, QL, av, a);
    double* dst = A + (size_t)posRL[e] * N2 + NS * NEQ;
#pragma unroll
    for (int k = 0; k < 4 * NEQ; k++) dst[k] += -a[k];
  }
}

// Bkernel_NumJac (jacobian.tcc:459-544), boundaryJacEval == 0: one thread per half-edge; the boundary state is
// recomputed for every perturbation of the interior state.  Writes q exactly as the reference does (phantom state,
// aux of the left node), the diagonal contribution into the half-edge's bdiag slot and -- ghost half-edges -- A(l,ghost).
template <int NS, int MINB = 8>
__global__ void __launch_bounds__(64, MINB) kfr_jac_bedges(DevMesh m, fr::Params<NS> p, const int* __restrict__ list, int n,
                                                      const unsigned char* __restrict__ bfirst, const double* __restrict__ beta,
                                                      double* q, const int* __restrict__ bpos, double* __restrict__ bdiag,
                                                      double* __restrict__ A) {
  constexpr int NEQ = W<NS>::NEQ, NV = W<NS>::NV, N2 = W<NS>::N2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int be = list[t];
  const int2 lr = m.ben[be];
  const int l = lr.x, r = lr.y;
  const int type = m.bctype[be];
  const bool first = bfirst[be] != 0;
  const bool ghost = is_ghost(m, r);
  const double h = 1.0e-8;
  const double betaL = beta[l];
  double QL[NV], QR[NV], av[4], fS[NEQ], fL[NEQ], fR[NEQ];
  load_row<NS, NV>(q, l, QL);
  load_row<NS, NV>(q, r, QR);
  load_avec(m.bea, be, av);
  if (!first && type != PCFD_BC_PARALLEL) fr::aux(p, QL);
  // viscous far field: ONE copy of the free stream per half-edge, scaled in place by every evaluation (jacobian.tcc:485-506)
  const bool ffv = type == PCFD_BC_FARFIELD_VISCOUS;
  const double ubar = ffv ? m.bubar[be] : 1.0;
  double Qref[NV];
  if (ffv) for (int i = 0; i < NV; i++) Qref[i] = p.qinf[i];
  fr::boundary_variables(p, QL, QR, av, type, betaL, ubar, ffv ? Qref : nullptr);
  if (type != PCFD_BC_PARALLEL) {
    double* qr = q + (size_t)r * NV;
    for (int i = 0; i < NV; i++) qr[i] = QR[i];
    if (first) {
      double* ql = q + (size_t)l * NV;
      for (int i = NS + 4; i < NV; i++) ql[i] = QL[i];
    }
  }
  fr::numerical_flux(p, QL, QR, av, 0.0, betaL, fS);
  double* bd = bdiag + (size_t)be * N2;
  double* ag = ghost ? A + (size_t)bpos[be] * N2 : nullptr;
  for (int i = 0; i < NEQ; i++) {
    double QPL[NV], QPR[NV];
    for (int k = 0; k < NV; k++) { QPL[k] = QL[k]; QPR[k] = QR[k]; }
    QPL[i] += h; QPR[i] += h;
    fr::aux(p, QPL);
    if (!ghost) {
      // the reference also evaluates F(qL, qR + h) here and throws it away (only ghost half-edges own an A(l, r) block)
      for (int k = 0; k < NV; k++) QPR[k] = QR[k];
      fr::boundary_variables(p, QPL, QPR, av, type, betaL, ubar, ffv ? Qref : nullptr);
      fr::numerical_flux(p, QPL, QPR, av, 0.0, betaL, fL);
    } else {
      fr::aux_pr(p, QPR);
      fr::numerical_flux(p, QL, QPR, av, 0.0, betaL, fR);
      fr::numerical_flux(p, QPL, QR, av, 0.0, betaL, fL);
    }
    for (int j = 0; j < NEQ; j++) bd[j * NEQ + i] = (fL[j] - fS[j]) / h;
    if (ghost) for (int j = 0; j < NEQ; j++) ag[j * NEQ + i] = 0.0 + (fR[j] - fS[j]) / h;
  }
}

// (Round 2: a form with NEQ + 1 lanes per half-edge -- lane 0 the reference state, shared through shared memory, lane
// i + 1 perturbation i -- is bit-exact but SLOWER on B200: 34.5 ms against 21.1 ms at 10 M cells.  The pass is bound by
// the local-memory traffic of the boundary-state evaluation itself (ten sub-iterations of the 9x9 eigensystem, ~2 kB of
// stack per thread), and ten times the threads make that worse; profiles/r2_ncu_explicit.md.)
// Bkernel_NumJac for the nodes owning a Dirichlet-type half-edge: one thread per node, ALL its half-edges (ghost ones
// included) in half-edge order with the interior state carried from one to the next (k_jac_bnodes of the perfect-gas path)
template <int NS>
__global__ void __launch_bounds__(64) kfr_jac_bnodes(DevMesh m, fr::Params<NS> p, const int* __restrict__ bnodes, int nb,
                                                      const double* __restrict__ beta, double* q, const int* __restrict__ bpos,
                                                      double* __restrict__ bdiag, double* __restrict__ A) {
  constexpr int NEQ = W<NS>::NEQ, NV = W<NS>::NV, N2 = W<NS>::N2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb) return;
  const int n = bnodes[t];
  const double h = 1.0e-8;
  const double betaL = beta[n];
  double QL[NV];
  load_row<NS, NV>(q, n, QL);
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    const int r = a.x & 0x7fffffff;
    const int type = m.bctype[be];
    const bool ghost = is_ghost(m, r);
    double QR[NV], av[4], fS[NEQ], fL[NEQ], fR[NEQ];
    load_row<NS, NV>(q, r, QR);
    load_avec(m.bea, be, av);
    double nT = 0.0, tw = 0.0;
    if (type == PCFD_BC_NOSLIP) { nT = q[(size_t)m.bnormal[be] * NV + NS + 3]; tw = m.btwall[be]; }
    const bool ffv = type == PCFD_BC_FARFIELD_VISCOUS;
    const double ubar = ffv ? m.bubar[be] : 1.0;
    double Qref[NV];
    if (ffv) for (int i = 0; i < NV; i++) Qref[i] = p.qinf[i];
    fr::boundary_variables_seq(p, QL, QR, av, type, betaL, nT, tw, ubar, ffv ? Qref : nullptr);
    if (type != PCFD_BC_PARALLEL) {
      double* qr = q + (size_t)r * NV;
      for (int i = 0; i < NV; i++) qr[i] = QR[i];
    }
    fr::numerical_flux(p, QL, QR, av, 0.0, betaL, fS);
    double* bd = bdiag + (size_t)be * N2;
    double* ag = ghost ? A + (size_t)bpos[be] * N2 : nullptr;
    for (int i = 0; i < NEQ; i++) {
      double QPL[NV], QPR[NV];
      for (int kk = 0; kk < NV; kk++) { QPL[kk] = QL[kk]; QPR[kk] = QR[kk]; }
      QPL[i] += h; QPR[i] += h;
      fr::aux(p, QPL);
      fr::aux_pr(p, QPR);
      fr::numerical_flux(p, QL, QPR, av, 0.0, betaL, fR);
      if (!ghost) {
        for (int kk = 0; kk < NV; kk++) QPR[kk] = QR[kk];
        fr::boundary_variables_seq(p, QPL, QPR, av, type, betaL, nT, tw, ubar, ffv ? Qref : nullptr);
        fr::numerical_flux(p, QPL, QPR, av, 0.0, betaL, fL);
      } else {
        fr::numerical_flux(p, QPL, QR, av, 0.0, betaL, fL);
      }
      for (int j = 0; j < NEQ; j++) bd[j * NEQ + i] = (fL[j] - fS[j]) / h;
      if (ghost) for (int j = 0; j < NEQ; j++) ag[j * NEQ + i] = 0.0 + (fR[j] - fS[j]) / h;
    }
  }
  double* ql = q + (size_t)n * NV;
  for (int i = 0; i < NV; i++) ql[i] = QL[i];
}

// Bkernel_NumJac_Centered (jacobian.tcc:546-640), Param::boundaryJacType == 1, boundaryJacEval == 0: kfr_jac_bedges with
// central differences.  The BC is re-evaluated for the +h state only (the -h branch tests boundaryJacEval without the
// negation, :604): F(qL-h, QR) is taken against the unperturbed phantom state.
template <int NS>
__global__ void __launch_bounds__(64) kfr_jac_bedges_central(DevMesh m, fr::Params<NS> p, const int* __restrict__ list, int n,
                                                              const unsigned char* __restrict__ bfirst,
                                                              const double* __restrict__ beta, double* q,
                                                              const int* __restrict__ bpos, double* __restrict__ bdiag,
                                                              double* __restrict__ A) {
  constexpr int NEQ = W<NS>::NEQ, NV = W<NS>::NV, N2 = W<NS>::N2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int be = list[t];
  const int2 lr = m.ben[be];
  const int l = lr.x, r = lr.y;
  const int type = m.bctype[be];
  const bool first = bfirst[be] != 0;
  const bool ghost = is_ghost(m, r);
  const double h = 1.0e-8;
  const double betaL = beta[l];
  double QL[NV], QR[NV], av[4], fL[NEQ], fR[NEQ], fLd[NEQ], fRd[NEQ];
  load_row<NS, NV>(q, l, QL);
  load_row<NS, NV>(q, r, QR);
  load_avec(m.bea, be, av);
  if (!first && type != PCFD_BC_PARALLEL) fr::aux(p, QL);
  // viscous far field: ONE copy of the free stream per half-edge, scaled in place by every evaluation (jacobian.tcc:485-506)
  const bool ffv = type == PCFD_BC_FARFIELD_VISCOUS;
  const double ubar = ffv ? m.bubar[be] : 1.0;
  double Qref[NV];
  if (ffv) for (int i = 0; i < NV; i++) Qref[i] = p.qinf[i];
  fr::boundary_variables(p, QL, QR, av, type, betaL, ubar, ffv ? Qref : nullptr);
  if (type != PCFD_BC_PARALLEL) {
    double* qr = q + (size_t)r * NV;
    for (int i = 0; i < NV; i++) qr[i] = QR[i];
    if (first) {
      double* ql = q + (size_t)l * NV;
      for (int i = NS + 4; i < NV; i++) ql[i] = QL[i];
    }
  }
  double* bd = bdiag + (size_t)be * N2;
  double* ag = ghost ? A + (size_t)bpos[be] * N2 : nullptr;
  for (int i = 0; i < NEQ; i++) {
    double QPL[NV], QPR[NV];
    for (int k = 0; k < NV; k++) { QPL[k] = QL[k]; QPR[k] = QR[k]; }
    QPL[i] += h; QPR[i] += h;
    fr::aux(p, QPL);
    fr::aux_pr(p, QPR);
    fr::numerical_flux(p, QL, QPR, av, 0.0, betaL, fR);
    if (!ghost) {
      for (int k = 0; k < NV; k++) QPR[k] = QR[k];
      fr::boundary_variables(p, QPL, QPR, av, type, betaL, ubar, ffv ? Qref : nullptr);
      fr::numerical_flux(p, QPL, QPR, av, 0.0, betaL, fL);
    } else {
      fr::numerical_flux(p, QPL, QR, av, 0.0, betaL, fL);
    }
    for (int k = 0; k < NV; k++) { QPL[k] = QL[k]; QPR[k] = QR[k]; }
    QPL[i] -= h; QPR[i] -= h;
    fr::aux(p, QPL);
    fr::aux_pr(p, QPR);
    fr::numerical_flux(p, QL, QPR, av, 0.0, betaL, fRd);
    fr::numerical_flux(p, QPL, QR, av, 0.0, betaL, fLd);
    for (int j = 0; j < NEQ; j++) bd[j * NEQ + i] = (fL[j] - fLd[j]) / (2.0 * h);
    if (ghost) for (int j = 0; j < NEQ; j++) ag[j * NEQ + i] = 0.0 + (fR[j] - fRd[j]) / (2.0 * h);
  }
}

// kfr_jac_bnodes with central differences (nodes owning a Dirichlet-type half-edge, sequential half-edge walk)
template <int NS>
__global__ void __launch_bounds__(64) kfr_jac_bnodes_central(DevMesh m, fr::Params<NS> p, const int* __restrict__ bnodes, int nb,
                                                              const double* __restrict__ beta, double* q,
                                                              const int* __restrict__ bpos, double* __restrict__ bdiag,
                                                              double* __restrict__ A) {
  constexpr int NEQ = W<NS>::NEQ, NV = W<NS>::NV, N2 = W<NS>::N2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb) return;
  const int n = bnodes[t];
  const double h = 1.0e-8;
  const double betaL = beta[n];
  double QL[NV];
  load_row<NS, NV>(q, n, QL);
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    const int r = a.x & 0x7fffffff;
    const int type = m.bctype[be];
    const bool ghost = is_ghost(m, r);
    double QR[NV], av[4], fL[NEQ], fR[NEQ], fLd[NEQ], fRd[NEQ];
    load_row<NS, NV>(q, r, QR);
    load_avec(m.bea, be, av);
    double nT = 0.0, tw = 0.0;
    if (type == PCFD_BC_NOSLIP) { nT = q[(size_t)m.bnormal[be] * NV + NS + 3]; tw = m.btwall[be]; }
    const bool ffv = type == PCFD_BC_FARFIELD_VISCOUS;
    const double ubar = ffv ? m.bubar[be] : 1.0;
    double Qref[NV];
    if (ffv) for (int i = 0; i < NV; i++) Qref[i] = p.qinf[i];
    fr::boundary_variables_seq(p, QL, QR, av, type, betaL, nT, tw, ubar, ffv ? Qref : nullptr);
    if (type != PCFD_BC_PARALLEL) {
      double* qr = q + (size_t)r * NV;
      for (int i = 0; i < NV; i++) qr[i] = QR[i];
    }
    double* bd = bdiag + (size_t)be * N2;
    double* ag = ghost ? A + (size_t)bpos[be] * N2 : nullptr;
    for (int i = 0; i < NEQ; i++) {
      double QPL[NV], QPR[NV];
      for (int kk = 0; kk < NV; kk++) { QPL[kk] = QL[kk]; QPR[kk] = QR[kk]; }
      QPL[i] += h; QPR[i] += h;
      fr::aux(p, QPL);
      fr::aux_pr(p, QPR);
      fr::numerical_flux(p, QL, QPR, av, 0.0, betaL, fR);
      if (!ghost) {
        for (int kk = 0; kk < NV; kk++) QPR[kk] = QR[kk];
        fr::boundary_variables_seq(p, QPL, QPR, av, type, betaL, nT, tw, ubar, ffv ? Qref : nullptr);
        fr::numerical_flux(p, QPL, QPR, av, 0.0, betaL, fL);
      } else {
        fr::numerical_flux(p, QPL, QR, av, 0.0, betaL, fL);
      }
      for (int kk = 0; kk < NV; kk++) { QPL[kk] = QL[kk]; QPR[kk] = QR[kk]; }
      QPL[i] -= h; QPR[i] -= h;
      fr::aux(p, QPL);
      fr::aux_pr(p, QPR);
      fr::numerical_flux(p, QL, QPR, av, 0.0, betaL, fRd);
      fr::numerical_flux(p, QPL, QR, av, 0.0, betaL, fLd);
      for (int j = 0; j < NEQ; j++) bd[j * NEQ + i] = (fL[j] - fLd[j]) / (2.0 * h);
      if (ghost) for (int j = 0; j < NEQ; j++) ag[j * NEQ + i] = 0.0 + (fR[j] - fRd[j]) / (2.0 * h);
    }
  }
  double* ql = q + (size_t)n * NV;
  for (int i = 0; i < NV; i++) ql[i] = QL[i];
}

// Bkernel_BC_Jac_Modify (jacobian.tcc:247-249, bc.tcc:747-903) -> ModifyViscousWallJacobian (compressibleFR.tcc:2072-2099)
// with CRSMatrix::BlankSubRow (crsmatrix.tcc:524-541): one thread per wall node, its NoSlip half-edges in order
template <int NS>
__global__ void kfr_jac_wall(DevMesh m, const int* __restrict__ wnodes, int nw, const int* __restrict__ ia,
                             const int* __restrict__ ja, const int* __restrict__ iau, double* A) {
  constexpr int NEQ = W<NS>::NEQ, N2 = W<NS>::N2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nw) return;
  const int n = wnodes[t];
  const int r0 = ia[n], r1 = ia[n + 1];
  double* diag = A + (size_t)iau[n] * N2;
  for (int k = m.adjp[n]; k < m.adjp[n + 1]; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const int be = a.y - m.nedge;
    if (m.bctype[be] != PCFD_BC_NOSLIP) continue;
    for (int sub = NS; sub < NS + 4; sub++) {
      for (int kk = r0; kk < r1; kk++)
        for (int j = 0; j < NEQ; j++) A[(size_t)kk * N2 + sub * NEQ + j] = 0.0;
      diag[sub * NEQ + sub] = 1.0;
    }
    if (m.btwall[be] < 0.0) {   // adiabatic: the wall temperature follows the most-normal neighbour
      const int nn_ = m.bnormal[be];
      for (int kk = r0; kk < r1; kk++)
        if (ja[kk] == nn_) { A[(size_t)kk * N2 + (NS + 3) * NEQ + NS + 3] = -1.0; break; }
    }
  }
}

// Kernel_Diag_NumJac (jacobian.tcc:434-456) after the boundary terms: NEQ threads per node, one block row each
template <int NS>
__global__ void __launch_bounds__(W<NS>::NEQ * 16) kfr_jac_diag(DevMesh m, const int* __restrict__ iau,
                                                                 const int* __restrict__ posLR, const int* __restrict__ posRL,
                                                                 const double* __restrict__ bdiag, double* A) {
  constexpr int NEQ = W<NS>::NEQ, N2 = W<NS>::N2;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = tid / NEQ;
  if (n >= m.nnode) return;
  const int j = tid - n * NEQ;
  double d[NEQ];
#pragma unroll
  for (int k = 0; k < NEQ; k++) d[k] = 0.0;
  const int kbeg = m.adjp[n], kend = m.adjp[n + 1];
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y < m.nedge) continue;
    const double* src = bdiag + (size_t)(a.y - m.nedge) * N2 + j * NEQ;
#pragma unroll
    for (int kk = 0; kk < NEQ; kk++) d[kk] += __ldg(src + kk);
  }
  for (int k = kbeg; k < kend; k++) {
    const int2 a = m.adj[k];
    if (a.y >= m.nedge) break;
    const int pos = (a.x < 0) ? posLR[a.y] : posRL[a.y];
    const double* src = A + (size_t)pos * N2 + j * NEQ;
#pragma unroll
    for (int kk = 0; kk < NEQ; kk++) d[kk] += -src[kk];
  }
  double* dg = A + (size_t)iau[n] * N2 + j * NEQ;
#pragma unroll
  for (int k = 0; k < NEQ; k++) dg[k] = d[k];
}

// per node: the source-term Jacobian (EqnSet::SourceTermJacobian, eqnset.tcc:163-187: one-sided FD on the native
// variables, subtracted from the diagonal block, jacobian.tcc:199-208) and ContributeTemporalTerms (:214-250)
template <int NS>
__global__ void __launch_bounds__(64, 8) kfr_jac_node(DevMesh m, fr::Params<NS> p, const int* __restrict__ iau,
                                                    const double* __restrict__ q, const double* __restrict__ dt,
                                                    const double* __restrict__ beta, double cnp1, double* A) {
  constexpr int NEQ = W<NS>::NEQ, N2 = W<NS>::N2;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.nnode) return;
  const double h = 1.0e-8;
  double Q[NS + 6], s0[NS], sP[NS];
  load_row<NS, NS + 6>(q, n, Q);
  double* dg = A + (size_t)iau[n] * N2;
  const double vol = m.vol[n];
  if (p.rxn_on) {
    // NEQ + 1 source terms in the reference.  The source is a function of rho_i and T alone: a density perturbation
    // keeps the temperature, hence every rate constant (all the exp / pow / log of the model) of the unperturbed
    // evaluation; a velocity perturbation reproduces the unperturbed source bit for bit, so its column is
    // (s - s)/h = +0 and the subtraction leaves the block as it is; only T + h needs new rate constants.
    double Kf[PCFD_CHEM_MAX_REACTIONS], Kb[PCFD_CHEM_MAX_REACTIONS];
    chemdev::rate_constants(p.chem, Q[NS + 3] * p.ref_temperature, Kf, Kb);
    fr::source_term_rates(p, Q, vol, Kf, Kb, s0);
    for (int i = 0; i < NS; i++) {
      double QP[NS + 4];
#pragma unroll
      for (int k = 0; k < NS + 4; k++) QP[k] = Q[k];
#pragma unroll
      for (int k = 0; k < NS; k++) if (k == i) QP[k] += h;
      fr::source_term_rates(p, QP, vol, Kf, Kb, sP);
      for (int j = 0; j < NS; j++) dg[j * NEQ + i] -= (sP[j] - s0[j]) / h;
    }
    {
      double QP[NS + 4];
#pragma unroll
      for (int k = 0; k < NS + 4; k++) QP[k] = Q[k];
      QP[NS + 3] += h;
      chemdev::rate_constants(p.chem, QP[NS + 3] * p.ref_temperature, Kf, Kb);
      fr::source_term_rates(p, QP, vol, Kf, Kb, sP);
      for (int j = 0; j < NS; j++) dg[j * NEQ + (NS + 3)] -= (sP[j] - s0[j]) / h;
    }
  }
  fr::temporal_terms(p, Q, vol, cnp1, dt[n], dg, beta[n]);
}

// CRSMatrix::PrepareSGS (crsmatrix.tcc:840-876) -> LU (matrix.h:110-190): k_lu_diag_lanes<NEQ> (pcfd_internal.cuh)

// ========================================================================== SGS
// One level of CRS::SGS (crs.tcc:90-145): NEQ lanes per row (3 rows per warp for 9x9 blocks), lane i owns block-row i,
// accumulates rhs[i] -= (M_k x_k)[i] block after block in ja order, the lanes exchange rhs by shuffle and each runs
// the permuted LuSolve redundantly out of the (L1-broadcast) diagonal block; lane i stores x[i].
template <int NS, int U>
__global__ void __launch_bounds__(128) kfr_sgs_level(const int* __restrict__ rows, int nrows, const int* __restrict__ ia,
                                                      const int* __restrict__ ja, const int* __restrict__ iau,
                                                      const double* __restrict__ A, const int* __restrict__ pv,
                                                      const double* __restrict__ b, double* x) {
  constexpr int NEQ = W<NS>::NEQ, N2 = W<NS>::N2, RPW = 32 / NEQ;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int grp = lane / NEQ;
  const int i = lane - grp * NEQ;
  const int slot = warp * RPW + grp;
  const bool active = (grp < RPW) && (slot < nrows);
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const int row = rows[slot];
  const int k0 = __ldg(ia + row), k1 = __ldg(ia + row + 1);
  double rhs = __ldg(b + (size_t)row * NEQ + i);
  int k = k0 + 1;
  for (; k + U <= k1; k += U) {
    int col[U];
    double a[U][NEQ], xv[U][NEQ];
#pragma unroll
    for (int u = 0; u < U; u++) col[u] = __ldg(ja + k + u);
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < NEQ; j++) a[u][j] = __ldcs(A + (size_t)(k + u) * N2 + i * NEQ + j);
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < NEQ; j++) xv[u][j] = x[(size_t)col[u] * NEQ + j];
#pragma unroll
    for (int u = 0; u < U; u++) {
      double v = a[u][0] * xv[u][0];
#pragma unroll
      for (int j = 1; j < NEQ; j++) v += a[u][j] * xv[u][j];
      rhs -= v;
    }
  }
  for (; k < k1; k++) {
    const double* a = A + (size_t)k * N2 + i * NEQ;
    const double* xv = x + (size_t)__ldg(ja + k) * NEQ;
    double v = __ldcs(a) * xv[0];
#pragma unroll
    for (int j = 1; j < NEQ; j++) v += __ldcs(a + j) * xv[j];
    rhs -= v;
  }
  double bb[NEQ], xx[NEQ];
  int pp[NEQ];
#pragma unroll
  for (int j = 0; j < NEQ; j++) bb[j] = __shfl_sync(mask, rhs, grp * NEQ + j);
#pragma unroll
  for (int j = 0; j < NEQ; j++) pp[j] = __ldg(pv + (size_t)row * NEQ + j);
  const double* d = A + (size_t)__ldg(iau + row) * N2;
#pragma unroll
  for (int r = 0; r < NEQ; r++) {
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < r; j++) sum += d[pp[r] * NEQ + j] * xx[j];
    double bp = bb[0];
#pragma unroll
    for (int j = 1; j < NEQ; j++) bp = (pp[r] == j) ? bb[j] : bp;
    xx[r] = bp - sum;
  }
#pragma unroll
  for (int r = NEQ - 1; r >= 0; r--) {
    double sum = 0.0;
#pragma unroll
    for (int j = NEQ - 1; j > r; j--) sum += d[pp[r] * NEQ + j] * bb[j];
    bb[r] = (xx[r] - sum) / d[pp[r] * NEQ + r];
  }
  double out = bb[0];
#pragma unroll
  for (int j = 1; j < NEQ; j++) out = (i == j) ? bb[j] : out;
  x[(size_t)row * NEQ + i] = out;
}

// ------------------------------------------------------------------ host side
template <int NS>
fr::Params<NS> make_params(const pcfd_ctx* c) {
  const pcfd_fr_state* s = c->fr;
  fr::Params<NS> p{};
  p.chem = s->chem_dev;
  for (int i = 0; i < NS; i++) {
    p.mw[i] = s->host.chem.mw[i];
    p.Rs[i] = chemdev::UNIV_R / s->host.chem.mw[i];
    for (int r = 0; r < 2; r++)
      for (int k = 0; k < 7; k++) p.nasa[i][r][k] = s->host.chem.nasa7[i][r][k];
  }
  p.ref_density = s->host.ref_density; p.ref_velocity = s->host.ref_velocity; p.ref_temperature = s->host.ref_temperature;
  p.ref_pressure = s->host.ref_pressure; p.ref_time = s->host.ref_time; p.ref_specific_enthalpy = s->host.ref_specific_enthalpy;
  p.Pref = s->host.pref; p.dt = s->host.dt;
  p.chi = c->prm.chi; p.cfl = c->prm.cfl;
  p.use_local_dt = s->host.use_local_dt; p.rxn_on = s->host.rxn_on; p.no_cvbc = c->prm.no_cvbc;
  p.sorder = c->prm.sorder; p.limiter = c->prm.limiter;
  for (int i = 0; i < 3 * NS + 6; i++) p.qinf[i] = s->host.qinf[i];
  return p;
}

template <int NS>
fr::Transport<NS> make_transport(const pcfd_ctx* c) {
  const pcfd_fr_state* s = c->fr;
  const pcfd_transport_model& tm = s->host.transport;
  fr::Transport<NS> t{};
  for (int i = 0; i < NS; i++) {
    t.nmu[i] = tm.nmu[i]; t.nk[i] = tm.nk[i];
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 6; k++) { t.mu_fit[i][r][k] = tm.mu_fit[i][r][k]; t.k_fit[i][r][k] = tm.k_fit[i][r][k]; }
    for (int k = 0; k < 4; k++) { t.mu_white[i][k] = tm.mu_white[i][k]; t.k_white[i][k] = tm.k_white[i][k]; }
    for (int j = 0; j < NS; j++) {
      const double MWi = s->host.chem.mw[i], MWj = s->host.chem.mw[j];
      t.pw25[i][j] = pow(MWj / MWi, 0.25);          // chem.tcc:908
      t.pwm05[i][j] = pow((1.0 + MWi / MWj), -0.5);  // chem.tcc:909
    }
  }
  t.sqrt8 = sqrt(8.0);
  t.white_uniform = 1;
  for (int i = 0; i < NS; i++) {
    const double temp = (1.0 + sqrt(1.0) * t.pw25[i][i]);
    t.phi_ii[i] = t.pwm05[i][i] * temp * temp / t.sqrt8;
    for (int k = 0; k < 4; k++)
      if (tm.mu_white[i][k] != tm.mu_white[0][k] || tm.k_white[i][k] != tm.k_white[0][k]) t.white_uniform = 0;
  }
  // one pow serves both properties only if they share T0 and the transition temperature
  if (tm.mu_white[0][1] != tm.k_white[0][1] || tm.mu_white[0][3] != tm.k_white[0][3]) t.white_uniform = 0;
  t.ref_viscosity = s->host.ref_viscosity; t.ref_k = s->host.ref_k;
  t.Re = c->prm.Re; t.PrT = c->prm.PrT;
  return t;
}

int fr_sumsq(pcfd_ctx* c, const double* v, int nrows, int slot, double* host_out) {
  pcfd_fr_state* s = c->fr;
  PROF("kfr_sumsq_partial");
  kfr_sumsq_partial<<<FR_RED_BLOCKS, 256, 0, c->stream>>>(v, nrows, c->neqn, s->red);
  LAUNCH_CHECK();
  PROF("kfr_sumsq_final");
  kfr_sumsq_final<<<1, 256, 0, c->stream>>>(s->red, FR_RED_BLOCKS, c->neqn, s->redout + slot * 32);
  LAUNCH_CHECK();
  if (host_out) {
    CK(cudaMemcpyAsync(host_out, s->redout + slot * 32, (1 + c->neqn) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

template <int NS>
struct Impl {
  using Wd = W<NS>;

  static int update_bcs(pcfd_ctx* c) {
    if (c->nbn) {
      PROF("kfr_update_bcs_nodes");
      kfr_update_bcs_nodes<NS><<<nblk(c->nbn, 64), 64, 0, c->stream>>>(c->dm, make_params<NS>(c), c->bnodes, c->nbn,
                                                                      c->f[PCFD_F_BETA], c->f[PCFD_F_Q]);
      LAUNCH_CHECK();
    }
    if (!c->nblist_bc) return 0;
    PROF("kfr_update_bcs_edges");
    kfr_update_bcs_edges<NS><<<nblk(c->nblist_bc, 64), 64, 0, c->stream>>>(c->dm, make_params<NS>(c), c->blist, c->nblist_bc,
                                                                          c->bfirst, c->f[PCFD_F_BETA], c->f[PCFD_F_Q]);
    LAUNCH_CHECK();
    return 0;
  }
  static int gradient(pcfd_ctx* c) {
    if (c->grad_type == 1) {
      PROF("kfr_gradient_gg");
      kfr_gradient_gg<NS><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD]);
      LAUNCH_CHECK();
      return 0;
    }
    PROF("kfr_gradient");
    kfr_gradient<NS><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->f[PCFD_F_Q], c->f[PCFD_F_LSQ_SW], c->f[PCFD_F_QGRAD]);
    LAUNCH_CHECK();
    return 0;
  }
  static int limiter(pcfd_ctx* c) {
    constexpr int BS = Wd::NEQ * 16;
    const int type = c->prm.limiter;
    double* lim = c->f[PCFD_F_LIMITER];
    const fr::Params<NS> p = make_params<NS>(c);
    PROF("kfr_limiter");
    kfr_limiter<NS><<<nblk((long long)c->nn * Wd::NEQ, BS), BS, 0, c->stream>>>(c->dm, type, c->prm.chi, c->f[PCFD_F_Q],
                                                                               c->f[PCFD_F_QGRAD], lim);
    LAUNCH_CHECK();
    if (type == 0) return 0;
    int cur = 0;
    bool clipped = false;
    PROF("kfr_fill_int");
    kfr_fill_int<<<nblk(c->nnode, 256), 256, 0, c->stream>>>(c->tclip[0], c->nnode, INT_MAX);
    LAUNCH_CHECK();
    for (int it = 0; it < c->nedge + 2; it++) {
      int hflags[2] = {0, 0};
      CK(cudaMemsetAsync(c->dflags, 0, 2 * sizeof(int), c->stream));
      if (c->nedge) {
        PROF("kfr_clip_edges");
        kfr_clip_edges<NS><<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, p, c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD], lim,
                                                                      c->tclip[cur], c->clipflag, c->dflags);
        LAUNCH_CHECK();
      }
      if (it == 0) {
        CK(cudaMemcpyAsync(hflags, c->dflags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (!hflags[0]) break;
        clipped = true;
      }
      PROF("kfr_clip_nodes");
      kfr_clip_nodes<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->clipflag, c->tclip[cur], c->tclip[cur ^ 1],
                                                                 c->dflags + 1);
      LAUNCH_CHECK();
      cur ^= 1;
      CK(cudaMemcpyAsync(hflags, c->dflags, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      if (!hflags[1]) break;
    }
    PROF("kfr_limiter_final");
    kfr_limiter_final<<<nblk((long long)c->nn * Wd::NEQ, 256), 256, 0, c->stream>>>(c->nn, c->nnode, Wd::NEQ,
                                                                                   clipped ? c->tclip[cur] : nullptr, lim);
    LAUNCH_CHECK();
    return 0;
  }
  // Limiter::Compute without Kernel_PressureClip and the clamp: the raw limiter (first half of the fused pair)
  static int limiter_raw(pcfd_ctx* c) {
    constexpr int BS = Wd::NEQ * 16;
    PROF("kfr_limiter");
    kfr_limiter<NS><<<nblk((long long)c->nn * Wd::NEQ, BS), BS, 0, c->stream>>>(c->dm, c->prm.limiter, c->prm.chi, c->f[PCFD_F_Q],
                                                                               c->f[PCFD_F_QGRAD], c->f[PCFD_F_LIMITER]);
    LAUNCH_CHECK();
    return 0;
  }
  static int residual(pcfd_ctx* c, double* sumsq) { return residual_impl(c, sumsq, nullptr); }
  // clip_hit != nullptr: the fused form -- lim holds the raw limiter, the edge kernel clamps on the fly and raises
  // dflags[2] if the pressure clip would act anywhere, kfr_limiter_final clamps in place behind it
  static int residual_impl(pcfd_ctx* c, double* sumsq, bool* clip_hit) {
    constexpr int BS = Wd::NEQ * 16;
    const fr::Params<NS> p = make_params<NS>(c);
    const double* beta = c->f[PCFD_F_BETA];
    const bool fused = clip_hit != nullptr;
    if (fused) {
      CK(cudaMemsetAsync(c->dflags + 2, 0, sizeof(int), c->stream));
      PROF("kfr_negflag_nodes");
      kfr_negflag_nodes<<<nblk(c->nn, 256), 256, 0, c->stream>>>(c->nn, Wd::NEQ, c->f[PCFD_F_LIMITER], c->fr->negflag);
      LAUNCH_CHECK();
      if (c->nedge) {
        PROF("kfr_clip_neg_edges");
        kfr_clip_neg_edges<NS><<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, p, c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD],
                                                                          c->f[PCFD_F_LIMITER], c->fr->negflag, c->dflags + 2);
        LAUNCH_CHECK();
      }
    }
    if (c->nedge) {
      PROF("kfr_flux_edges");
      if (fused)
        kfr_flux_edges<NS, true><<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, p, c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD],
                                                                            c->f[PCFD_F_LIMITER], beta, c->flux, c->dflags + 2);
      else
        kfr_flux_edges<NS, false><<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, p, c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD],
                                                                             c->f[PCFD_F_LIMITER], beta, c->flux, nullptr);
      LAUNCH_CHECK();
    }
    if (fused) {
      CK(cudaMemcpyAsync(c->hflag, c->dflags + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaEventRecord(c->ev_flag, c->stream));
      PROF("kfr_limiter_final");
      kfr_limiter_final<<<nblk((long long)c->nn * Wd::NEQ, 256), 256, 0, c->stream>>>(c->nn, c->nnode, Wd::NEQ, nullptr,
                                                                                     c->f[PCFD_F_LIMITER]);
      LAUNCH_CHECK();
    }
    if (c->nb) {
      PROF("kfr_flux_bedges");
      kfr_flux_bedges<NS><<<nblk(c->nb, 128), 128, 0, c->stream>>>(c->dm, p, c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD],
                                                                  c->f[PCFD_F_LIMITER], beta, c->bflux);
      LAUNCH_CHECK();
    }
    if (c->fr->viscous && c->nedge + c->nb) {
      PROF("kfr_vflux_edges");
      kfr_vflux_edges<NS><<<nblk((long long)c->nedge + c->nb, 128), 128, 0, c->stream>>>(
          c->dm, p, make_transport<NS>(c), c->f[PCFD_F_Q], c->f[PCFD_F_QGRAD], c->f[PCFD_F_MUT], c->fr->vflux, c->fr->bvflux);
      LAUNCH_CHECK();
    }
    PROF("kfr_source");
    kfr_source<NS><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->nnode, p, c->f[PCFD_F_Q], c->vol, c->fr->src);
    LAUNCH_CHECK();
    PROF("kfr_residual_gather");
    kfr_residual_gather<NS><<<nblk((long long)c->nnode * Wd::NEQ, BS), BS, 0, c->stream>>>(
        c->dm, c->flux, c->bflux, c->fr->viscous ? c->fr->vflux : nullptr, c->fr->bvflux, c->fr->src,
        c->nwall ? c->wallflag : nullptr, c->f[PCFD_F_B]);
    LAUNCH_CHECK();
    if (c->torder && c->have_qold) {
      const bool bdf2 = c->iter > 1 && c->torder == 2;
      PROF("kfr_temporal_residual");
      kfr_temporal_residual<NS><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->nnode, p, bdf2 ? 1.5 : 1.0, bdf2 ? -0.5 : 0.0, c->vol,
                                                                           c->f[PCFD_F_Q], c->f[PCFD_F_QOLD],
                                                                           c->f[PCFD_F_QOLDM1], c->nwall ? c->wallflag : nullptr,
                                                                           c->f[PCFD_F_B]);
      LAUNCH_CHECK();
    }
    if (fused) {
      CK(cudaEventSynchronize(c->ev_flag));
      *clip_hit = *c->hflag != 0;
      if (*clip_hit) return 0;      // the caller redoes limiter + residual on the ordered clip path
    }
    if (sumsq) return fr_sumsq(c, c->f[PCFD_F_B], c->nnode, 0, sumsq);
    return 0;
  }
  static int timestep(pcfd_ctx* c, double* dtmin) {
    pcfd_fr_state* s = c->fr;
    PROF("kfr_eig_edges");
    kfr_eig_edges<NS><<<nblk((long long)c->nedge + c->nb, 128), 128, 0, c->stream>>>(c->dm, make_params<NS>(c), c->f[PCFD_F_Q],
                                                                                    c->f[PCFD_F_BETA], s->eig, s->beig);
    LAUNCH_CHECK();
    PROF("kfr_timestep");
    kfr_timestep<<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->dm, c->prm.cfl, s->eig, s->beig, c->vnn23,
                                                             c->f[PCFD_F_TIMESTEP]);
    LAUNCH_CHECK();
    if (dtmin) {
      PROF("kfr_min_partial");
      kfr_min_partial<<<FR_RED_BLOCKS, 256, 0, c->stream>>>(c->f[PCFD_F_TIMESTEP], c->nnode, s->red);
      LAUNCH_CHECK();
      PROF("kfr_min_final");
      kfr_min_final<<<1, 256, 0, c->stream>>>(s->red, FR_RED_BLOCKS, s->redout + 96);
      LAUNCH_CHECK();
      CK(cudaMemcpyAsync(dtmin, s->redout + 96, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
  }
  static int surface_props(pcfd_ctx* c, double V, double* props, double* rho_inf, bool* viscous) {
    *viscous = c->fr->viscous;
    *rho_inf = c->fr->host.qinf[NS + 5];
    if (!c->nbedge) return 0;
    PROF("kfr_surface_props");
    kfr_surface_props<NS><<<nblk(c->nbedge, 128), 128, 0, c->stream>>>(c->dm, make_params<NS>(c), make_transport<NS>(c), V,
                                                                      c->fr->viscous, c->f[PCFD_F_Q], props);
    LAUNCH_CHECK();
    return 0;
  }
  static int turb_props(pcfd_ctx* c) {
    if (!c->fr->viscous) return fail(c, "pcfd_turb_compute: Spalart-Allmaras under the reacting eqnset needs compressibleNSFR");
    const long long nthreads = (long long)c->nedge + c->nb + c->nn;
    PROF("kfr_turb_props");
    kfr_turb_props<NS><<<nblk(nthreads, 128), 128, 0, c->stream>>>(c->dm, make_params<NS>(c), make_transport<NS>(c), c->f[PCFD_F_Q],
                                                                   c->tprop_e, c->tprop_b, c->tprop_n);
    LAUNCH_CHECK();
    return 0;
  }
  static int explicit_solve(pcfd_ctx* c) {
    pcfd_fr_state* s = c->fr;
    CK(cudaMemsetAsync(s->dbad, 0, sizeof(int), c->stream));
    PROF("kfr_explicit");
    kfr_explicit<NS><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->nnode, make_params<NS>(c), c->f[PCFD_F_B],
                                                                c->f[PCFD_F_TIMESTEP], c->vol, c->f[PCFD_F_Q], c->f[PCFD_F_X],
                                                                s->dbad);
    LAUNCH_CHECK();
    // ExplicitSolve leaves q alone for this eqnset; NewtonIterate then applies the update (solutionSpace.tcc:802-804)
    return apply_dq(c);
  }
  static int apply_dq(pcfd_ctx* c) {
    PROF("kfr_apply_dq");
    kfr_apply_dq<NS><<<nblk(c->nnode, 128), 128, 0, c->stream>>>(c->nnode, make_params<NS>(c), c->f[PCFD_F_X], c->f[PCFD_F_Q],
                                                                c->dzeroed);
    LAUNCH_CHECK();
    return 0;
  }
  static int jacobian(pcfd_ctx* c) {
    constexpr int BS = Wd::NEQ * 16;
    const fr::Params<NS> p = make_params<NS>(c);
    double* A = c->f[PCFD_F_A];
    const double* beta = c->f[PCFD_F_BETA];
    // no memset of A: every block is written in full before anything is added to it (see pcfd_jacobian, pcfd_kernels.cu)
    c->ludiag = false;
    if (c->nedge) {
      if (c->field_jac_type == 1) {
        constexpr int EPBC = 4;
        PROF("kfr_jac_edges_central");
        kfr_jac_edges_central<NS, EPBC><<<nblk(c->nedge, EPBC), 2 * Wd::NEQ * EPBC, 0, c->stream>>>(c->dm, p, c->f[PCFD_F_Q], beta,
                                                                                                  c->posLR, c->posRL, A);
      } else {
        PROF("kfr_jac_edges");
        constexpr int G = 16, WPB = 4;   // edges per warp, warps per block (see the kernel)
        // 80 registers (6 blocks of 4 warps per SM): 20.4 ms at 10 M cells; 96: 20.7; 120 (no spills): 21.4
        kfr_jac_edges<NS, G, 6><<<nblk(c->nedge, G * WPB), 32 * WPB, 0, c->stream>>>(c->dm, p, c->f[PCFD_F_Q], beta, c->posLR,
                                                                                    c->posRL, A);
      }
      LAUNCH_CHECK();
    }
    if (c->nbn) {
      if (c->boundary_jac_type == 1) {
        PROF("kfr_jac_bnodes_central");
        kfr_jac_bnodes_central<NS><<<nblk(c->nbn, 64), 64, 0, c->stream>>>(c->dm, p, c->bnodes, c->nbn, beta, c->f[PCFD_F_Q],
                                                                          c->bpos, c->bdiag, A);
      } else {
        PROF("kfr_jac_bnodes");
        kfr_jac_bnodes<NS><<<nblk(c->nbn, 64), 64, 0, c->stream>>>(c->dm, p, c->bnodes, c->nbn, beta, c->f[PCFD_F_Q], c->bpos,
                                                                  c->bdiag, A);
      }
      LAUNCH_CHECK();
    }
    if (c->nblist) {
      if (c->boundary_jac_type == 1) {
        PROF("kfr_jac_bedges_central");
        kfr_jac_bedges_central<NS><<<nblk(c->nblist, 64), 64, 0, c->stream>>>(c->dm, p, c->blist, c->nblist, c->bfirst, beta,
                                                                             c->f[PCFD_F_Q], c->bpos, c->bdiag, A);
      } else {
        PROF("kfr_jac_bedges");
        static const int minb = getenv("PCFD_FRJACB_MINB") ? atoi(getenv("PCFD_FRJACB_MINB")) : 8;
        if (minb == 8)   // 128 registers, 16 warps per SM: 14.6 ms at 10 M cells against 18.6 ms with 255 registers / 8 warps
          kfr_jac_bedges<NS, 8><<<nblk(c->nblist, 64), 64, 0, c->stream>>>(c->dm, p, c->blist, c->nblist, c->bfirst, beta,
                                                                          c->f[PCFD_F_Q], c->bpos, c->bdiag, A);
        else
          kfr_jac_bedges<NS, 4><<<nblk(c->nblist, 64), 64, 0, c->stream>>>(c->dm, p, c->blist, c->nblist, c->bfirst, beta,
                                                                          c->f[PCFD_F_Q], c->bpos, c->bdiag, A);
      }
      LAUNCH_CHECK();
    }
    if (c->fr->viscous && c->nedge) {
      PROF("kfr_vjac_edges");
      kfr_vjac_edges<NS><<<nblk(c->nedge, 128), 128, 0, c->stream>>>(c->dm, p, make_transport<NS>(c), c->f[PCFD_F_Q],
                                                                      c->f[PCFD_F_MUT], c->posLR, c->posRL, A);
      LAUNCH_CHECK();
    }
    PROF("kfr_jac_diag");
    kfr_jac_diag<NS><<<nblk((long long)c->nnode * Wd::NEQ, BS), BS, 0, c->stream>>>(c->dm, c->iau, c->posLR, c->posRL, c->bdiag, A);
    LAUNCH_CHECK();
    PROF("kfr_jac_node");
    kfr_jac_node<NS><<<nblk(c->nnode, 64), 64, 0, c->stream>>>(c->dm, p, c->iau, c->f[PCFD_F_Q], c->f[PCFD_F_TIMESTEP], beta,
                                                               (c->iter > 1 && c->torder == 2) ? 1.5 : 1.0, A);
    LAUNCH_CHECK();
    if (c->nwall) {
      PROF("kfr_jac_wall");
      kfr_jac_wall<NS><<<nblk(c->nwall, 64), 64, 0, c->stream>>>(c->dm, c->wnodes, c->nwall, c->ia, c->ja, c->iau, A);
      LAUNCH_CHECK();
    }
    return 0;
  }
  static int prepare_sgs(pcfd_ctx* c) {
    if (c->ludiag) return 0;
    PROF("kfr_lu_diag");
    constexpr int RPW = 32 / Wd::NEQ;
    k_lu_diag_lanes<Wd::NEQ><<<nblk((long long)((c->nnode + RPW - 1) / RPW) * 32, 128), 128, 0, c->stream>>>(c->nnode, c->iau,
                                                                                                        c->f[PCFD_F_A], c->pv);
    LAUNCH_CHECK();
    c->ludiag = true;
    return 0;
  }
  static int sgs(pcfd_ctx* c, int nsgs, double* ddq) {
    constexpr int RPW = 32 / Wd::NEQ;
    const double* A = c->f[PCFD_F_A];
    double* x = c->f[PCFD_F_X];
    bool prev_tile = false;
    for (int s = 0; s < nsgs; s++) {
      for (int dir = 0; dir < 2; dir++) {
        const std::vector<int>& off = dir ? c->lev_b : c->lev_f;
        const int* rows = dir ? c->rows_b : c->rows_f;
        for (size_t l = 0; l + 1 < off.size(); l++) {
          const int nr = off[l + 1] - off[l];
          const int warps = (nr + RPW - 1) / RPW;
          const int cap = (dir ? c->tile_cap_b : c->tile_cap_f)[l];
          if (cap > 0) {
            // rows of the level are consecutive in memory (colour-sorted numbering): bulk-copy streamed tiles, 16 lanes
            // per row (sgs_tile.cuh)
            const int Wt = c->sgs_tile_warps;
            const int RT = Wt * 2;
            const int capA = (cap * Wd::N2 * 8 + 8 + 15) & ~15;
            const size_t shm = 16 + (size_t)capA;
            const int tiles = (nr + RT - 1) / RT;
            const int pf = c->sgs_pf_dist >= 0 ? c->sgs_pf_dist : (int)((size_t)(24 << 20) / shm);
            const int row0 = dir ? c->lev_first_b[l] : c->lev_first_f[l];
            const int step = dir ? c->lev_step_b[l] : c->lev_step_f[l];
            const bool chain = c->sgs_pdl && !c->prof && prev_tile;   // programmatic dependent launch between levels
            PROF("k_sgs_tile");
#define PCFD_FR_TILE(WW)                                                                                               \
  do {                                                                                                                 \
    static size_t set_##WW = 0;                                                                                        \
    if (shm > set_##WW) {                                                                                              \
      CK(cudaFuncSetAttribute(k_sgs_tile_t<Wd::NEQ, WW, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));  \
      set_##WW = shm;                                                                                                  \
    }                                                                                                                  \
    CK(launch_maybe_pdl(k_sgs_tile_t<Wd::NEQ, WW, 16>, tiles, WW * 32, shm, c->stream, chain, row0, step, nr, c->ia, c->ja, A, \
                        c->pv, c->f[PCFD_F_B], x, pf));                                                                \
  } while (0)
            if (Wt == 1) PCFD_FR_TILE(1); else if (Wt == 4) PCFD_FR_TILE(4); else PCFD_FR_TILE(2);
#undef PCFD_FR_TILE
            LAUNCH_CHECK();
            prev_tile = true;
            continue;
          }
          prev_tile = false;
          PROF("kfr_sgs_level");
          kfr_sgs_level<NS, 2><<<nblk((long long)warps * 32, 128), 128, 0, c->stream>>>(rows + off[l], nr, c->ia, c->ja, c->iau, A,
                                                                                       c->pv, c->f[PCFD_F_B], x);
          LAUNCH_CHECK();
        }
      }
      // the reference halo-updates x here (crs.tcc:147-150); its |xOld - xNorm| monitor needs the last two norms
      if (ddq && s >= nsgs - 2) {
        prev_tile = false;
        if (fr_sumsq(c, x, c->nnode, (s == nsgs - 1) ? 0 : 1, nullptr)) return 1;
      }
    }
    if (ddq) {
      double h[64];
      CK(cudaMemcpyAsync(h, c->fr->redout, 64 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      const double N = (double)c->nnode * Wd::NEQ;
      const double xNorm = (nsgs >= 1) ? sqrt(h[0]) / N : 0.0;
      const double xOld = (nsgs >= 2) ? sqrt(h[32]) / N : 0.0;
      *ddq = fabs(xOld - xNorm);
    }
    return 0;
  }
};

// the species counts the library is built for
#define FR_DISPATCH(c, call)                                                                         \
  switch ((c)->fr->host.chem.nspecies) {                                                             \
    case 5: return Impl<5>::call;                                                                    \
    default: return fail(c, "reacting eqnset: this build carries kernels for 5 species only");       \
  }

}  // namespace

void pcfd_fr_set_time(pcfd_ctx* c, double dt, int use_local) {
  if (!c || !c->fr) return;
  c->fr->host.dt = dt;
  c->fr->host.use_local_dt = use_local;
}

void pcfd_fr_destroy(pcfd_ctx* c) {
  if (!c || !c->fr) return;
  if (c->fr->chem_dev) cudaFree(c->fr->chem_dev);
  delete c->fr;
  c->fr = nullptr;
}
int pcfd_fr_update_bcs(pcfd_ctx* c) { FR_DISPATCH(c, update_bcs(c)); }
int pcfd_fr_gradient(pcfd_ctx* c) { FR_DISPATCH(c, gradient(c)); }
int pcfd_fr_limiter(pcfd_ctx* c) { FR_DISPATCH(c, limiter(c)); }
int pcfd_fr_residual(pcfd_ctx* c, double* sumsq) { FR_DISPATCH(c, residual(c, sumsq)); }
int pcfd_fr_limiter_raw(pcfd_ctx* c) { FR_DISPATCH(c, limiter_raw(c)); }
int pcfd_fr_residual_fused(pcfd_ctx* c, double* sumsq, bool* clip_hit) { FR_DISPATCH(c, residual_impl(c, sumsq, clip_hit)); }
int pcfd_fr_timestep(pcfd_ctx* c, double* dtmin) { FR_DISPATCH(c, timestep(c, dtmin)); }
int pcfd_fr_turb_props(pcfd_ctx* c) { FR_DISPATCH(c, turb_props(c)); }
int pcfd_fr_surface_props(pcfd_ctx* c, double V, double* props, double* rho_inf, bool* viscous) {
  FR_DISPATCH(c, surface_props(c, V, props, rho_inf, viscous));
}
int pcfd_fr_explicit_solve(pcfd_ctx* c) { FR_DISPATCH(c, explicit_solve(c)); }
int pcfd_fr_apply_dq(pcfd_ctx* c) { FR_DISPATCH(c, apply_dq(c)); }
int pcfd_fr_jacobian(pcfd_ctx* c) { FR_DISPATCH(c, jacobian(c)); }
int pcfd_fr_prepare_sgs(pcfd_ctx* c) { FR_DISPATCH(c, prepare_sgs(c)); }
int pcfd_fr_sgs(pcfd_ctx* c, int nsgs, double* ddq) { FR_DISPATCH(c, sgs(c, nsgs, ddq)); }

extern "C" {

int pcfd_create_fr(const pcfd_mesh_desc* mesh, const pcfd_params* params, const pcfd_fr_params* frp, int device,
                   pcfd_ctx** out) {
  pcfd_ctx* c = nullptr;
  if (!mesh || !params || !frp || !out) return fail(c, "pcfd_create_fr: null argument");
  *out = nullptr;
  if (params->eqnset != PCFD_EQNSET_COMPRESSIBLE_EULER_FR && params->eqnset != PCFD_EQNSET_COMPRESSIBLE_NS_FR)
    return fail(c, "pcfd_create_fr: eqnset must be compressibleEulerFR or compressibleNSFR");
  const bool viscous = params->eqnset == PCFD_EQNSET_COMPRESSIBLE_NS_FR;   // param.tcc:401-404
  const int ns = frp->chem.nspecies;
  if (ns != 5) return fail(c, "pcfd_create_fr: this build carries kernels for 5 species only");
  if (frp->chem.nreactions < 0 || frp->chem.nreactions > PCFD_CHEM_MAX_REACTIONS) return fail(c, "pcfd_create_fr: bad reaction count");
  for (int j = 0; j < frp->chem.nreactions; j++) {
    if (frp->chem.nsp[j] < 1 || frp->chem.nsp[j] > ns) return fail(c, "pcfd_create_fr: bad species count in a reaction");
    for (int k = 0; k < frp->chem.nsp[j]; k++)
      if (frp->chem.species[j][k] < 0 || frp->chem.species[j][k] >= ns) return fail(c, "pcfd_create_fr: reaction references an unknown species");
  }
  if (!(frp->ref_density > 0.0 && frp->ref_velocity > 0.0 && frp->ref_temperature > 0.0 && frp->ref_pressure > 0.0 &&
        frp->ref_time > 0.0 && frp->ref_specific_enthalpy > 0.0))
    return fail(c, "pcfd_create_fr: reference values must be positive");
  if (viscous) {
    if (!(params->Re > 0.0 && params->PrT > 0.0 && frp->ref_viscosity > 0.0 && frp->ref_k > 0.0))
      return fail(c, "pcfd_create_fr: compressibleNSFR needs positive Re, PrT, ref_viscosity and ref_k");
    for (int i = 0; i < ns; i++) {
      const pcfd_transport_model& tm = frp->transport;
      if (tm.nmu[i] < 1 || tm.nmu[i] > 3 || tm.nk[i] < 1 || tm.nk[i] > 3)
        return fail(c, "pcfd_create_fr: every species needs 1..3 viscosity and conductivity fit ranges");
    }
  }
  const int nb = mesh->nbedge + mesh->ngedge;
  for (int e = 0; e < nb; e++) {
    const int t = mesh->bedges_bctype[e];
    if (t == PCFD_BC_FARFIELD_VISCOUS && !viscous)
      return fail(c, "pcfd_create_fr: the viscous far-field BC needs compressibleNSFR (Re and the wall distance)");
    if (t == PCFD_BC_NOSLIP && !mesh->bedges_twall)
      return fail(c, "pcfd_create_fr: no-slip walls need bedges_twall (wall temperature / ref_temperature; < 0: adiabatic)");
  }
  if (params->turb_model == 1 && !viscous)
    return fail(c, "pcfd_create_fr: Spalart-Allmaras needs compressibleNSFR");
  pcfd_params prm = *params;
  if (pcfd_internal_create(mesh, &prm, device, ns + 4, 3 * ns + 6, 2 * ns + 4, &c)) return 1;
  c->prm.eqnset = params->eqnset;
  pcfd_fr_state* s = new pcfd_fr_state();
  c->fr = s;
  s->host = *frp;
  s->viscous = viscous;
  struct Guard { pcfd_ctx* c; bool ok = false; ~Guard() { if (!ok) { pcfd_create_err() = c->err; pcfd_destroy(c); } } } guard{c};
  CK(cudaMalloc(reinterpret_cast<void**>(&s->chem_dev), sizeof(pcfd_chem_model)));
  CK(cudaMemcpy(s->chem_dev, &frp->chem, sizeof(pcfd_chem_model), cudaMemcpyHostToDevice));
  if (dev_alloc(c, &s->eig, (size_t)c->nedge)) return 1;
  if (dev_alloc(c, &s->beig, (size_t)c->nb)) return 1;
  if (dev_alloc(c, &s->src, (size_t)c->nnode * ns)) return 1;
  if (dev_alloc(c, &s->red, (size_t)FR_RED_BLOCKS * 32)) return 1;
  if (dev_alloc(c, &s->redout, 128)) return 1;
  if (dev_alloc(c, &s->dbad, 4)) return 1;
  if (dev_alloc(c, &s->negflag, (size_t)c->nn)) return 1;
  if (viscous) {
    if (dev_alloc(c, &s->vflux, (size_t)c->nedge * 4)) return 1;
    if (dev_alloc(c, &s->bvflux, (size_t)c->nb * 4)) return 1;
  }
  // the implicit path always needs the matrix; allocate it lazily like the perfect-gas path does (pcfd_jacobian)
  {   // beta defaults to 1 (no preconditioning) until the host sets the field
    std::vector<double> ones(c->fsize[PCFD_F_BETA], 1.0);
    CK(cudaMemcpy(c->f[PCFD_F_BETA], ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  CK(cudaDeviceSynchronize());
  guard.ok = true;
  *out = c;
  return 0;
}

int pcfd_widths(const pcfd_ctx* c, int* neqn, int* nvars, int* nterms) {
  if (!c) return 1;
  if (neqn) *neqn = c->neqn;
  if (nvars) *nvars = c->nvars;
  if (nterms) *nterms = c->nterms;
  return 0;
}

}  // extern "C"
