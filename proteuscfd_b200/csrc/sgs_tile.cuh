// sgs_tile.cuh -- one level of CRS::SGS (crs.tcc:90-145) with the matrix STREAMED through shared memory by bulk async
// copies (TMA engine: cp.async.bulk + mbarrier), for NQ x NQ blocks (5x5 perfect gas, 9x9 five-species reacting).
// Shared by pcfd_kernels.cu and pcfd_fr.cu.
#pragma once

#include <cuda_runtime.h>

namespace {

// The level with the matrix STREAMED through shared memory by bulk async copies (TMA engine,
// cp.async.bulk + mbarrier) instead of per-lane 8-byte loads.  When the rows of a level are consecutive in
// memory (always the case for a colour-sorted numbering) the blocks of a tile of rows form ONE contiguous byte
// range of A (and of ja): one elected thread issues two bulk copies for the whole tile, the lanes fetch their
// row's b / pv meanwhile and then run the identical ja-ordered arithmetic out of shared memory.  Several CTAs
// per SM keep ~200 KB of matrix in flight per SM, which is what the HBM stream needs; x stays a gather (L2).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                              unsigned long long pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// rows of the tile: row0 + step*slot (step = +1 forward, -1 backward).  Work split inside a row's 5-lane group:
// lane t takes the off-diagonal blocks u = t, t+5, t+10, ... WHOLE: while the bulk copy of the tile's matrix bytes
// is in flight (L2 evict-first: the matrix is read once per sweep and must not push x out of L2) it fetches the
// column index and the 5-vector x[col] of its blocks (one dependent hop for the whole row), then forms
// v_u = M_u x_u (MatVecMult order, matrix.h:63-74) out of shared memory and parks it on top of block u.  After a
// group-level sync lane i accumulates rhs[i] -= v_u[i] for u = 0, 1, 2, ... -- the reference's ja order -- and the
// group runs the permuted LuSolve as in k_sgs_level.
// LPR = lanes per row (5, 10 or 16): more lanes per row = fewer rows (less shared memory) per warp, hence more
// warps per SM to hide the latency of the serial part (ordered accumulation + LuSolve, done by lanes 0..4).
template <int NQ, int WARPS, int LPR>
__global__ void __launch_bounds__(WARPS * 32) k_sgs_tile_t(int row0, int step, int nrows, const int* __restrict__ ia,
                                                          const int* __restrict__ ja, const double* __restrict__ A,
                                                          const int* __restrict__ pv, const double* __restrict__ b,
                                                          double* x, int pf_dist) {
  constexpr int NQ2 = NQ * NQ;
  static_assert(LPR >= NQ, "a row needs at least one lane per equation");
  constexpr int RPW = 32 / LPR;
  constexpr int RT = WARPS * RPW;   // rows per tile
  constexpr int R = (16 + LPR - 1) / LPR;   // rounds held in registers: up to R*LPR neighbour blocks before looping
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem);
  unsigned char* sA = smem + 16;
  const int s0 = blockIdx.x * RT;
  const int s1 = min(s0 + RT, nrows);
  const int rfirst = row0 + step * s0, rlast = row0 + step * (s1 - 1);
  const int rlo = min(rfirst, rlast), rhi = max(rfirst, rlast);
  const int kb0 = __ldg(ia + rlo);
  const unsigned offA = ((unsigned)kb0 * (NQ2 * 8u)) & 15u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    const int kb1 = __ldg(ia + rhi + 1);
    const unsigned bytesA = ((unsigned)(kb1 - kb0) * (NQ2 * 8u) + offA + 15u) & ~15u;
    mbar_expect_tx(bar, bytesA);
    bulk_g2s_hint(sA, reinterpret_cast<const unsigned char*>(A) + (size_t)kb0 * (NQ2 * 8) - offA, bytesA, bar,
                  policy_evict_first());
  }
  if (threadIdx.x == 1 && pf_dist > 0) {
    // pull the matrix / ja / b / pv bytes of the tile that a CTA pf_dist tiles later will need from HBM into L2
    // now, so that its bulk copy and its ja -> x dependent loads are L2 hits (the HBM stream then runs pf_dist
    // tiles ahead of the compute instead of being exposed in every CTA's lifetime)
    const int f0 = (blockIdx.x + pf_dist) * RT;
    if (f0 < nrows) {
      const int f1 = min(f0 + RT, nrows);
      const int qfirst = row0 + step * f0, qlast = row0 + step * (f1 - 1);
      const int qlo = min(qfirst, qlast), qhi = max(qfirst, qlast);
      const int pb0 = __ldg(ia + qlo), pb1 = __ldg(ia + qhi + 1);
      const unsigned oA = ((unsigned)pb0 * (NQ2 * 8u)) & 15u, oJ = (pb0 & 3) * 4u;
      bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(A) + (size_t)pb0 * (NQ2 * 8) - oA,
                       ((unsigned)(pb1 - pb0) * (NQ2 * 8u) + oA + 15u) & ~15u);
      bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(ja) + (size_t)pb0 * 4 - oJ,
                       ((unsigned)(pb1 - pb0) * 4u + oJ + 15u) & ~15u);
      const unsigned oB = ((unsigned)qlo * (NQ * 8u)) & 15u, oP = ((unsigned)qlo * (NQ * 4u)) & 15u;
      bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(b) + (size_t)qlo * (NQ * 8) - oB,
                       ((unsigned)(qhi - qlo + 1) * (NQ * 8u) + oB + 15u) & ~15u);
      bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(pv) + (size_t)qlo * (NQ * 4) - oP,
                       ((unsigned)(qhi - qlo + 1) * (NQ * 4u) + oP + 15u) & ~15u);
    }
  }
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int grp = lane / LPR;
  const int t = lane - grp * LPR;
  const int slot = s0 + warp * RPW + grp;
  const bool active = (grp < RPW) && (slot < s1);
  const unsigned gmask = (LPR == 16 ? 0xffffu : ((1u << LPR) - 1u)) << (grp * LPR);   // the row's own lanes
  int row = 0, k0 = 0, nb = 0;
  int p[NQ], col[R];
  double rhs = 0.0;
  double xr[R][NQ];
  if (active) {
    row = row0 + step * slot;
    k0 = __ldg(ia + row);
    nb = __ldg(ia + row + 1) - k0 - 1;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int u = r * LPR + t;
      col[r] = (u < nb) ? __ldg(ja + k0 + 1 + u) : 0;
    }
#pragma unroll
    for (int j = 0; j < NQ; j++) p[j] = __ldg(pv + (size_t)row * NQ + j);
    if (t < NQ) rhs = __ldg(b + (size_t)row * NQ + t);
  }
  // Programmatic dependent launch (levels of one pcfd_sgs call are launched back to back with
  // cudaLaunchAttributeProgrammaticStreamSerialization): everything above -- the bulk copy of the tile's matrix bytes,
  // the L2 prefetch, ia / ja / pv / b -- is independent of the previous level and overlaps its tail; only x carries the
  // dependency (read below, written at the end), so the wait sits here.  Both instructions are no-ops in a plain launch.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (active) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int u = r * LPR + t;
      if (u < nb) {
        const double* xc = x + (size_t)col[r] * NQ;
#pragma unroll
        for (int j = 0; j < NQ; j++) xr[r][j] = xc[j];
      }
    }
  }
  __syncthreads();        // mbarrier initialised before anyone waits on it
  mbar_wait(bar, 0);
  if (!active) return;
  double* tA = reinterpret_cast<double*>(sA + offA) - (size_t)kb0 * NQ2;   // tA[k*25 + ..], k in the tile
  for (int base = 0;;) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int u = base + r * LPR + t;
      if (u < nb) {
        double* m = tA + (size_t)(k0 + 1 + u) * NQ2;
        double v[NQ];
#pragma unroll
        for (int ii = 0; ii < NQ; ii++) {
          double acc = m[ii * NQ] * xr[r][0];
#pragma unroll
          for (int j = 1; j < NQ; j++) acc += m[ii * NQ + j] * xr[r][j];
          v[ii] = acc;
        }
#pragma unroll
        for (int ii = 0; ii < NQ; ii++) m[ii] = v[ii];   // only this lane ever reads block u's matrix entries
      }
    }
    base += R * LPR;
    if (base >= nb) break;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int u = base + r * LPR + t;
      if (u < nb) {
        const double* xc = x + (size_t)__ldg(ja + k0 + 1 + u) * NQ;
#pragma unroll
        for (int j = 0; j < NQ; j++) xr[r][j] = xc[j];
      }
    }
  }
  __syncwarp(gmask);
  if (t < NQ) {
    const double* vv = tA + (size_t)(k0 + 1) * NQ2 + t;
    for (int u = 0; u < nb; u++) rhs -= vv[(size_t)u * NQ2];
  }
  double bb[NQ], xx[NQ];
#pragma unroll
  for (int j = 0; j < NQ; j++) bb[j] = __shfl_sync(gmask, rhs, grp * LPR + j);
  if (t >= NQ) return;
  const double* d = tA + (size_t)k0 * NQ2;   // the diagonal block is the first of the row (iau == ia)
#pragma unroll
  for (int r = 0; r < NQ; r++) {
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < r; j++) sum += d[p[r] * NQ + j] * xx[j];
    double bp = bb[0];
#pragma unroll
    for (int j = 1; j < NQ; j++) bp = (p[r] == j) ? bb[j] : bp;
    xx[r] = bp - sum;
  }
#pragma unroll
  for (int r = NQ - 1; r >= 0; r--) {
    double sum = 0.0;
#pragma unroll
    for (int j = NQ - 1; j > r; j--) sum += d[p[r] * NQ + j] * bb[j];
    bb[r] = (xx[r] - sum) / d[p[r] * NQ + r];
  }
  double out = bb[0];
#pragma unroll
  for (int j = 1; j < NQ; j++) out = (t == j) ? bb[j] : out;
  x[(size_t)row * NQ + t] = out;
}


// launch with (pdl = true) or without the programmatic-stream-serialization attribute
template <class... KArgs, class... Args>
inline cudaError_t launch_maybe_pdl(void (*kernel)(KArgs...), int grid, int block, size_t shm, cudaStream_t stream, bool pdl,
                                    Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = shm;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace
