// pcfd_crsmatrix.cuh -- CRSMatrix::CRSTranspose (ucs/crsmatrix.tcc:568-599) on the device (included at the end of
// pcfd_kernels.cu).  The adjoint / design path of the reference transposes the assembled Jacobian
// (Compute_dRdQ_Transpose, jacobian.tcc:121-127; derivatives.cpp:682).
//
// The reference transposes every block in place, then swaps the mirror blocks (i, j) <-> (j, i) of local node pairs
// through GetPointer searches, then replaces the blocks of ghost columns by the owner's (PObj::TransposeCommCRS,
// parallel.tcc:54-338).  Here the block positions of both directions of every edge are already known from pcfd_create
// (posLR / posRL, bpos for the parallel half-edges), so the local part is three entry-parallel kernels without a search:
// one thread per entry pair, A(l,r)[k][l'] <-> A(r,l)[l'][k] per interior edge, upper <-> lower triangle for the diagonal
// and the ghost-column blocks.  Pure data movement: bit-identical by construction, HBM-bound (every block read and
// written once: 16 bytes per entry).
// The ghost-column blocks travel through the host (pcfd_crs_ghost_blocks), in the order of the parallel half-edges: the
// operation runs once per adjoint solve, and the host already owns a transport (MPI in ucs.x, torch.distributed in the
// tests); proteuscfd_b200/parallel.py: crs_transpose is the routing of TransposeCommCRS on top of it.
#pragma once

namespace {

// interior edges: block (l, r) at posLR[e], block (r, l) at posRL[e]
template <int N>
__global__ void __launch_bounds__(256) k_crs_transpose_pairs(int nedge, const int* __restrict__ posLR,
                                                             const int* __restrict__ posRL, double* A) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e = t / (N * N);
  if (e >= nedge) return;
  const int kl = (int)(t - e * (N * N));
  const int k = kl / N, l = kl - k * N;
  double* p = A + (size_t)posLR[e] * (N * N) + k * N + l;
  double* q = A + (size_t)posRL[e] * (N * N) + l * N + k;
  const double a = *p, b = *q;
  *p = b;
  *q = a;
}

// blocks that stay where they are (diagonal blocks: pos = iau; ghost-column blocks: pos = bpos of the parallel
// half-edges): upper and lower triangle exchanged
template <int N>
__global__ void __launch_bounds__(256) k_crs_transpose_inplace(int count, const int* __restrict__ pos, double* A) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long s = t / (N * N);
  if (s >= count) return;
  const int kl = (int)(t - s * (N * N));
  const int k = kl / N, l = kl - k * N;
  if (k >= l) return;
  double* blk = A + (size_t)pos[s] * (N * N);
  const double a = blk[k * N + l], b = blk[l * N + k];
  blk[k * N + l] = b;
  blk[l * N + k] = a;
}

// ghost-column blocks <-> a packed buffer, parallel half-edge order
template <int N>
__global__ void __launch_bounds__(256) k_crs_ghost_blocks(int count, const int* __restrict__ pos, int set, double* A, double* buf) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long s = t / (N * N);
  if (s >= count) return;
  const int kl = (int)(t - s * (N * N));
  double* a = A + (size_t)pos[s] * (N * N) + kl;
  if (set) *a = buf[t]; else buf[t] = *a;
}

template <int N>
int crs_transpose_impl(pcfd_ctx* c) {
  double* A = c->f[PCFD_F_A];
  if (c->nedge > 0) {
    PROF("k_crs_transpose_pairs");
    k_crs_transpose_pairs<N><<<nblk((long long)c->nedge * N * N, 256), 256, 0, c->stream>>>(c->nedge, c->posLR, c->posRL, A);
    LAUNCH_CHECK();
  }
  PROF("k_crs_transpose_inplace");
  k_crs_transpose_inplace<N><<<nblk((long long)c->nnode * N * N, 256), 256, 0, c->stream>>>(c->nnode, c->iau, A);
  LAUNCH_CHECK();
  if (c->ngedge > 0) {
    PROF("k_crs_transpose_inplace");
    k_crs_transpose_inplace<N><<<nblk((long long)c->ngedge * N * N, 256), 256, 0, c->stream>>>(c->ngedge, c->bpos + c->nbedge, A);
    LAUNCH_CHECK();
  }
  return 0;
}

template <int N>
int crs_ghost_blocks_impl(pcfd_ctx* c, int set, double* host) {
  const size_t n = (size_t)c->ngedge * N * N;
  if (n == 0) return 0;
  double* buf = nullptr;
  CK(cudaMalloc(reinterpret_cast<void**>(&buf), n * sizeof(double)));
  cudaError_t e = cudaSuccess;
  if (set) e = cudaMemcpyAsync(buf, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    k_crs_ghost_blocks<N><<<nblk((long long)n, 256), 256, 0, c->stream>>>(c->ngedge, c->bpos + c->nbedge, set, c->f[PCFD_F_A], buf);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && !set) e = cudaMemcpyAsync(host, buf, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(buf);
  if (e != cudaSuccess) return fail(c, std::string("pcfd_crs_ghost_blocks: ") + cudaGetErrorString(e));
  return 0;
}

}  // namespace

extern "C" int pcfd_crs_transpose(pcfd_ctx* c) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (!c->f[PCFD_F_A]) return fail(c, "pcfd_crs_transpose: no matrix (pcfd_jacobian or pcfd_set_field(PCFD_F_A) first)");
  if (c->ludiag) return fail(c, "pcfd_crs_transpose: the diagonal blocks have been factored in place (pcfd_prepare_sgs)");
  if (c->neqn == 5) return crs_transpose_impl<5>(c);
  if (c->neqn == 9) return crs_transpose_impl<9>(c);
  return fail(c, "pcfd_crs_transpose: block size not instantiated");
}

extern "C" int pcfd_crs_ghost_blocks(pcfd_ctx* c, int set, double* host) {
  if (!c) return 1;
  CK(cudaSetDevice(c->device));
  if (!c->f[PCFD_F_A]) return fail(c, "pcfd_crs_ghost_blocks: no matrix");
  if (c->ngedge > 0 && !host) return fail(c, "pcfd_crs_ghost_blocks: null buffer");
  if (c->neqn == 5) return crs_ghost_blocks_impl<5>(c, set, host);
  if (c->neqn == 9) return crs_ghost_blocks_impl<9>(c, set, host);
  return fail(c, "pcfd_crs_ghost_blocks: block size not instantiated");
}
