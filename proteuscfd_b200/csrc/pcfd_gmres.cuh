// pcfd_gmres.cuh -- CRS::GMRES (ucs/crs.tcc:176-415) on the device (included by pcfd_kernels.cu).
//
// Restarted GMRES with right preconditioning on the block-CRS system of the context: A (PCFD_F_A, as assembled, NOT yet
// factored by pcfd_prepare_sgs), b (PCFD_F_B), x (PCFD_F_X: initial guess in, solution out).  Preconditioner /
// PrecondBackSolve (crs.tcc:555-641) types 0 (none), 1 (diagonal of the diagonal blocks), 2 (block diagonal, LU with the
// permutation vector of matrix.h:110-190), 3 (local ILU0, crsmatrix.tcc:276-507), 4 (SGS on a copy of the matrix).  The
// reference's flow solver keeps this path behind a comment (solutionSpace.tcc:734-750); its mesh-movement and design
// solvers call it (move.tcc:714).
//
// What runs where: the block-CRS matrix-vector product, the preconditioner solve and every vector update are kernels
// with the reference's arithmetic order per entry (bit-identical); dot products and norms are fixed-tree reductions
// (deterministic, 1e-15-level away from the reference's sequential sums), across ranks summed in rank order through the
// flag pages (pcfd_comm_allgather); the (nSearchDir+1)^2 Hessenberg / Givens algebra is the reference's, on the host.
// Across ranks the vector that needs a halo before every product (vtemp) lives in the field PCFD_F_X, so the library
// exchange applies to it as it is; the solution accumulates in scratch and is copied back at the end.
#pragma once

namespace {

// CRS::MatVecMultiply (crs.tcc:484-508) with MatVecMult (matrix.h:63-74): one thread per (row, component)
template <int N>
__global__ void __launch_bounds__(128) k_gm_spmv(int nnode, const int* __restrict__ ia, const int* __restrict__ ja,
                                                  const double* __restrict__ A, const double* __restrict__ vin,
                                                  double* __restrict__ vout) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = t / N;
  if (row >= nnode) return;
  const int j = t - row * N;
  double acc = 0.0;
  for (int indx = ia[row]; indx < ia[row + 1]; indx++) {
    const double* a1 = A + (size_t)indx * N * N + j * N;
    const double* v1 = vin + (size_t)__ldg(ja + indx) * N;
    double tmp = __ldcs(a1) * v1[0];
#pragma unroll
    for (int k = 1; k < N; k++) tmp += __ldcs(a1 + k) * v1[k];
    acc += tmp;
  }
  vout[(size_t)row * N + j] = acc;
}

// BuildBlockDiagPrecond (crsmatrix.tcc:190-240): copy of the diagonal blocks; type 2: + PrepareSGS (LU, matrix.h:110-190)
template <int N>
__global__ void __launch_bounds__(128) k_gm_precond_build(int nnode, int type, const int* __restrict__ iau,
                                                           const double* __restrict__ A, double* __restrict__ Nd,
                                                           int* __restrict__ pv) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= nnode) return;
  const double* g = A + (size_t)iau[nd] * N * N;
  double a[N * N];
  int p[N];
#pragma unroll
  for (int k = 0; k < N * N; k++) a[k] = g[k];
  for (int i = 0; i < N; i++) p[i] = i;
  if (type == 2) {
    for (int i = 0; i < N; i++) {
      double large = 0.0, lmag = 0.0;   // |large| carried explicitly: see k_lu_diag_lanes (nvcc 12.9 drops the abs otherwise)
      int row = 0;
      for (int j = i; j < N; j++) {
        const double v = a[p[j] * N + i], vmag = fabs(v);
        if (vmag > lmag) { large = v; lmag = vmag; row = j; }
      }
      const int tmp = p[i]; p[i] = p[row]; p[row] = tmp;
      large = 1.0 / large;
      for (int j = i + 1; j < N; j++) a[p[j] * N + i] *= large;
      for (int j = i + 1; j < N; j++)
        for (int k = i + 1; k < N; k++) a[p[j] * N + k] -= a[p[j] * N + i] * a[p[i] * N + k];
    }
  }
  for (int k = 0; k < N * N; k++) Nd[(size_t)nd * N * N + k] = a[k];
  for (int i = 0; i < N; i++) pv[(size_t)nd * N + i] = p[i];
}

// PrecondBackSolve (crs.tcc:592-641): N x = b per node; LuSolve (matrix.h:237-264) for the block-diagonal type
template <int N>
__global__ void __launch_bounds__(128) k_gm_precond_solve(int nnode, int type, const double* __restrict__ Nd,
                                                           const int* __restrict__ pv, const double* __restrict__ b,
                                                           double* __restrict__ x) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= nnode) return;
  const double* a = Nd + (size_t)nd * N * N;
  double bb[N], xx[N];
  for (int i = 0; i < N; i++) bb[i] = b[(size_t)nd * N + i];
  if (type == 1) {
    for (int j = 0; j < N; j++) x[(size_t)nd * N + j] = bb[j] / a[j * N + j];
    return;
  }
  int p[N];
  for (int i = 0; i < N; i++) p[i] = pv[(size_t)nd * N + i];
  for (int i = 0; i < N; i++) {
    double sum = 0.0;
    for (int j = 0; j < i; j++) sum += a[p[i] * N + j] * xx[j];
    xx[i] = bb[p[i]] - sum;
  }
  for (int i = N - 1; i >= 0; i--) {
    double sum = 0.0;
    for (int j = N - 1; j > i; j--) sum += a[p[i] * N + j] * bb[j];
    bb[i] = (xx[i] - sum) / a[p[i] * N + i];
  }
  for (int i = 0; i < N; i++) x[(size_t)nd * N + i] = bb[i];
}

// ---- preconditioner type 3: the local ILU0 (crs.tcc:571-575, 627-630)
//
// CRSMatrix::BuildILU0Local (crsmatrix.tcc:276-428) walks the scalar rows in order; everything the N scalar rows of block
// row r touch is the block row r itself and its mirror blocks (col, r), so two block rows conflict only when they are
// neighbours, and the forward levels of the SGS schedule (no two neighbours in a level, lower-numbered neighbours in
// earlier levels) reproduce the sequential result: one launch per level, one thread per block row of the level, the
// reference's statement order inside it.  The reference finds the blocks below the pivot by scanning all of ja
// (:336-346, quadratic); on the symmetric pattern they are the mirror blocks of the row's upper neighbours.
__device__ __forceinline__ int ilu0_find(const int* __restrict__ ia, const int* __restrict__ ja, int row, int col) {
  for (int indx = ia[row]; indx < ia[row + 1]; indx++)
    if (ja[indx] == col) return indx;
  return -1;
}

template <int N>
__global__ void __launch_bounds__(64) k_ilu0_build(const int* __restrict__ rows, int nr, int nnode, const int* __restrict__ ia,
                                                   const int* __restrict__ ja, const int* __restrict__ iau, double* M) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nr) return;
  const int r = rows[s];
  double* diag = M + (size_t)iau[r] * N * N;
  double temp[N * N];
  for (int c = 0; c < N; c++) {
    const double pivot = 1.0 / diag[c * N + c];   // no pivoting (:322-328)
    for (int j = c + 1; j < N; j++) diag[j * N + c] *= pivot;
    for (int j = ia[r]; j < ia[r + 1]; j++) {     // blocks (col, r) below the pivot: col > r, local rows only
      const int col = ja[j];
      if (col <= r || col >= nnode) continue;
      const int m = ilu0_find(ia, ja, col, r);
      if (m < 0) continue;
      double* block = M + (size_t)m * N * N;
      for (int k = 0; k < N; k++) block[k * N + c] *= pivot;
    }
    for (int j = ia[r]; j < ia[r + 1]; j++) {     // the pivot's block row as stored: diagonal block first
      const int col = ja[j];
      if (col >= nnode) continue;                 // ghost columns stay out of the local factorisation (:359-363)
      double* block2 = M + (size_t)j * N * N;
      const int m = ilu0_find(ia, ja, col, r);
      if (m < 0) continue;                        // (the reference dereferences NULL here: symmetric pattern assumed)
      double* block3 = M + (size_t)m * N * N;
      for (int k = 0; k < N; k++)
        for (int l = 0; l < N; l++) temp[k * N + l] = block2[c * N + l] * block3[k * N + c];
      for (int k = 0; k < N; k++)                 // subtracted from the diagonal block of THIS row (:376-381)
        for (int l = 0; l < N; l++) diag[k * N + l] -= temp[k * N + l];
      for (int k = c + 1; k < N; k++)
        for (int l = 0; l < N; l++) temp[k * N + l] = block2[c * N + l] * diag[k * N + c];
      for (int k = c + 1; k < N; k++)
        for (int l = 0; l < N; l++) block2[k * N + l] -= temp[k * N + l];
      for (int k = 0; k < N; k++)
        for (int l = c + 1; l < N; l++) temp[k * N + l] = block2[c * N + l] * block3[k * N + c];
      for (int k = 0; k < N; k++)
        for (int l = c + 1; l < N; l++) block3[k * N + l] -= temp[k * N + l];
    }
  }
}

// blocks of ghost columns blanked once the factorisation is through (:418-424)
template <int N>
__global__ void k_ilu0_blank_ghost(int nblocks, int nnode, const int* __restrict__ ja, double* M) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long blk = t / (N * N);
  if (blk >= nblocks) return;
  if (ja[blk] >= nnode) M[t] = 0.0;
}

// CRSMatrix::ILU0BackSub (crsmatrix.tcc:430-507), sweep down: one thread per (row of the level, component).  x has been
// blanked; the strictly lower part of the diagonal block multiplies the row's own, still blank, entries (:452-458) --
// the products are kept (a zero times whatever the factor holds) so that a NaN / Inf factor propagates as it does there.
template <int N>
__global__ void __launch_bounds__(128) k_ilu0_fwd(const int* __restrict__ rows, int nr, const int* __restrict__ ia,
                                                   const int* __restrict__ ja, const double* __restrict__ M,
                                                   const double* __restrict__ b, double* x) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = t / N;
  if (s >= nr) return;
  const int k = t - s * N;
  const int i = rows[s];
  double temp = 0.0;
  for (int j = ia[i]; j < ia[i + 1]; j++) {
    const int col = ja[j];
    const double* a1 = M + (size_t)j * N * N + k * N;
    if (col < i) {
      const double* v1 = x + (size_t)col * N;
      double temp2 = a1[0] * v1[0];
      for (int l = 1; l < N; l++) temp2 += a1[l] * v1[l];
      temp += temp2;
    } else if (col == i) {
      for (int l = 0; l < k; l++) temp += a1[l] * 0.0;
    }
  }
  x[(size_t)i * N + k] = b[(size_t)i * N + k] - temp;
}

// sweep up: the strictly upper part of the diagonal block is applied to b (:484-490), then the division by the block's
// own diagonal entry (:493-496).  Ghost columns count as above the row: their blocks and their x rows are zero.
template <int N>
__global__ void __launch_bounds__(128) k_ilu0_bwd(const int* __restrict__ rows, int nr, const int* __restrict__ ia,
                                                   const int* __restrict__ ja, const int* __restrict__ iau,
                                                   const double* __restrict__ M, const double* __restrict__ b, double* x) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = t / N;
  if (s >= nr) return;
  const int k = t - s * N;
  const int i = rows[s];
  double temp = 0.0;
  for (int j = ia[i]; j < ia[i + 1]; j++) {
    const int col = ja[j];
    const double* a1 = M + (size_t)j * N * N + k * N;
    if (col > i) {
      const double* v1 = x + (size_t)col * N;
      double temp2 = a1[0] * v1[0];
      for (int l = 1; l < N; l++) temp2 += a1[l] * v1[l];
      temp += temp2;
    } else if (col == i) {
      for (int l = N - 1; l > k; l--) temp += a1[l] * b[(size_t)i * N + l];
    }
  }
  x[(size_t)i * N + k] = (x[(size_t)i * N + k] - temp) / M[(size_t)iau[i] * N * N + k * N + k];
}

// vector updates of crs.tcc:259-395, one entry per thread, the reference's expression per entry
enum { GM_RES0 = 0, GM_DIVS = 1, GM_ORTHO = 2, GM_SCALE_TO = 3, GM_ACCUM = 4, GM_ADD = 5 };
__global__ void k_gm_vec(int op, int n, double s, const double* __restrict__ a, const double* __restrict__ b, double* y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (op) {
    case GM_RES0: y[i] = b[i] - y[i]; break;            // v0 = b - A x0
    case GM_DIVS: y[i] /= s; break;                     // v0 /= ||r0||
    case GM_ORTHO: y[i] -= (s * a[i]); break;           // uk -= h * vj
    case GM_SCALE_TO: y[i] = a[i] / s; break;           // v_{k+1} = uk / ||uk||
    case GM_ACCUM: y[i] += (a[i] * s); break;           // zk += vj * g[j]
    default: y[i] += a[i]; break;                       // x += N^-1 zk
  }
}

template <int BLOCK>
__global__ void k_gm_dot_partial(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ part) {
  __shared__ double sh[BLOCK];
  double acc = 0.0;
  for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += gridDim.x * BLOCK) acc += a[i] * b[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
template <int BLOCK>
__global__ void k_gm_dot_final(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double sh[BLOCK];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += BLOCK) acc += part[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// sum over this rank's entries of a.b, then over the ranks (rank order) when connected
int gm_dot(pcfd_ctx* c, const double* a, const double* b, int n, double* out) {
  PROF("k_gm_dot_partial");
  k_gm_dot_partial<256><<<RED_BLOCKS, 256, 0, c->stream>>>(a, b, n, c->red);
  LAUNCH_CHECK();
  PROF("k_gm_dot_final");
  k_gm_dot_final<256><<<1, 256, 0, c->stream>>>(c->red, RED_BLOCKS, c->redout + 12);
  LAUNCH_CHECK();
  double* hd = reinterpret_cast<double*>(c->hflag) + 1;   // pinned (c->hflag is a 64-byte pinned block)
  CK(cudaMemcpyAsync(hd, c->redout + 12, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  double local = *hd;
  if (comm_on(c)) {
    double all[COMM_MAXR];
    if (pcfd_comm_allgather(c, &local, 1, all)) return 1;
    local = 0.0;
    for (int r = 0; r < c->nranks; r++) local += all[r];
  }
  *out = local;
  return 0;
}

int gm_vec(pcfd_ctx* c, int op, int n, double s, const double* a, const double* b, double* y) {
  PROF("k_gm_vec");
  k_gm_vec<<<nblk(n, 256), 256, 0, c->stream>>>(op, n, s, a, b, y);
  LAUNCH_CHECK();
  return 0;
}

template <int N>
int gmres_impl(pcfd_ctx* c, int restarts, int nSearchDir, int precondType, double* dq_norm) {
  const double smallnum = 1.0e-15;
  const int nnode = c->nnode, nloc = nnode * N;
  const size_t vstride = (size_t)c->nn * N;
  const bool dist = comm_on(c);
  // scratch: v[0..nSearchDir], uk, the running solution, the block-diagonal preconditioner
  const size_t need = ((size_t)nSearchDir + 3) * vstride + (size_t)nnode * N * N;
  if (c->gm_cap < need) {
    if (c->gm_buf) CK(cudaFree(c->gm_buf));
    if (c->gm_pv) CK(cudaFree(c->gm_pv));
    c->gm_buf = nullptr; c->gm_pv = nullptr; c->gm_cap = 0;
    CK(cudaMalloc(reinterpret_cast<void**>(&c->gm_buf), need * sizeof(double)));
    CK(cudaMalloc(reinterpret_cast<void**>(&c->gm_pv), (size_t)nnode * N * sizeof(int)));
    c->gm_cap = need;
  }
  double* vdat = c->gm_buf;
  double* uk = vdat + ((size_t)nSearchDir + 1) * vstride;
  double* xs = uk + vstride;
  double* Nd = xs + vstride;
  double* vtemp = c->f[PCFD_F_X];      // exchangeable: the library's halo applies to it as it is
  double* const A = c->f[PCFD_F_A];
  double* const b = c->f[PCFD_F_B];
  int* const pvA = c->pv;
  if (precondType == 3 || precondType == 4) {
    // Preconditioner type 4 (crs.tcc:577-581): CopyMatrixStructure + PrepareSGS -- a copy of the whole matrix with its
    // diagonal blocks factored; every application is CRS::SGS(6, N, vtemp, rhs) (:629-632), the previous preconditioned
    // vector being the initial guess.  The sweeps are the context's own (pcfd_sgs) with its matrix / permutation /
    // right-hand-side pointers switched to the copy for the duration of the call.
    const size_t nA = c->fsize[PCFD_F_A];
    if (c->gm_ncap < nA) {
      if (c->gm_n) CK(cudaFree(c->gm_n));
      if (c->gm_npv) CK(cudaFree(c->gm_npv));
      c->gm_n = nullptr; c->gm_npv = nullptr; c->gm_ncap = 0;
      CK(cudaMalloc(reinterpret_cast<void**>(&c->gm_n), (nA + 2) * sizeof(double)));
      CK(cudaMalloc(reinterpret_cast<void**>(&c->gm_npv), (size_t)nnode * N * sizeof(int)));
      c->gm_ncap = nA;
    }
    CK(cudaMemcpyAsync(c->gm_n, A, nA * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    if (precondType == 4) {
      constexpr int RPW = 32 / N;
      PROF("k_lu_diag");
      k_lu_diag_lanes<N><<<nblk((long long)((nnode + RPW - 1) / RPW) * 32, 128), 128, 0, c->stream>>>(nnode, c->iau, c->gm_n, c->gm_npv);
      LAUNCH_CHECK();
    } else {
      // Preconditioner type 3 (crs.tcc:571-575): BuildILU0Local on the copy, level by level
      for (size_t l = 0; l + 1 < c->lev_f.size(); l++) {
        const int nr = c->lev_f[l + 1] - c->lev_f[l];
        if (nr <= 0) continue;
        PROF("k_ilu0_build");
        k_ilu0_build<N><<<nblk(nr, 64), 64, 0, c->stream>>>(c->rows_f + c->lev_f[l], nr, nnode, c->ia, c->ja, c->iau, c->gm_n);
        LAUNCH_CHECK();
      }
      PROF("k_ilu0_blank_ghost");
      k_ilu0_blank_ghost<N><<<nblk((long long)c->nblocks * N * N, 256), 256, 0, c->stream>>>(c->nblocks, nnode, c->ja, c->gm_n);
      LAUNCH_CHECK();
    }
  }
  if (precondType == 1 || precondType == 2) {
    PROF("k_gm_precond_build");
    k_gm_precond_build<N><<<nblk(nnode, 128), 128, 0, c->stream>>>(nnode, precondType, c->iau, A, Nd, c->gm_pv);
    LAUNCH_CHECK();
  }
  auto precond = [&](double* rhs, double* out) -> int {   // N out = rhs
    if (precondType == 4) {   // out is vtemp = field x at both call sites: its halo is the library's
      c->f[PCFD_F_A] = c->gm_n; c->pv = c->gm_npv; c->f[PCFD_F_B] = rhs;
      int rc = 0;
      if (dist) rc = comm_update(c, PCFD_F_X);                     // crs.tcc:88
      for (int s = 0; s < 6 && !rc; s++) {
        rc = pcfd_sgs(c, 1, nullptr);
        if (!rc && dist) rc = comm_update(c, PCFD_F_X);            // crs.tcc:146
      }
      c->f[PCFD_F_A] = A; c->pv = pvA; c->f[PCFD_F_B] = b;
      return rc;
    }
    if (precondType == 0) {
      CK(cudaMemcpyAsync(out, rhs, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      return 0;
    }
    if (precondType == 3) {   // ILU0BackSub: x blanked with its ghost rows, one sweep down, one up
      CK(cudaMemsetAsync(out, 0, vstride * sizeof(double), c->stream));
      for (size_t l = 0; l + 1 < c->lev_f.size(); l++) {
        const int nr = c->lev_f[l + 1] - c->lev_f[l];
        if (nr <= 0) continue;
        PROF("k_ilu0_fwd");
        k_ilu0_fwd<N><<<nblk((long long)nr * N, 128), 128, 0, c->stream>>>(c->rows_f + c->lev_f[l], nr, c->ia, c->ja, c->gm_n, rhs, out);
        LAUNCH_CHECK();
      }
      for (size_t l = 0; l + 1 < c->lev_b.size(); l++) {
        const int nr = c->lev_b[l + 1] - c->lev_b[l];
        if (nr <= 0) continue;
        PROF("k_ilu0_bwd");
        k_ilu0_bwd<N><<<nblk((long long)nr * N, 128), 128, 0, c->stream>>>(c->rows_b + c->lev_b[l], nr, c->ia, c->ja, c->iau, c->gm_n, rhs, out);
        LAUNCH_CHECK();
      }
      return 0;
    }
    PROF("k_gm_precond_solve");
    k_gm_precond_solve<N><<<nblk(nnode, 128), 128, 0, c->stream>>>(nnode, precondType, Nd, c->gm_pv, rhs, out);
    LAUNCH_CHECK();
    return 0;
  };
  auto matvec = [&](const double* vin, double* vout) -> int {
    CK(cudaMemsetAsync(vout, 0, vstride * sizeof(double), c->stream));
    PROF("k_gm_spmv");
    k_gm_spmv<N><<<nblk((long long)nnode * N, 128), 128, 0, c->stream>>>(nnode, c->ia, c->ja, A, vin, vout);
    LAUNCH_CHECK();
    return 0;
  };
  // initial guess: x with its ghost rows current (crs.tcc:249), kept in scratch from here on
  if (dist && comm_update(c, PCFD_F_X)) return 1;
  CK(cudaMemcpyAsync(xs, c->f[PCFD_F_X], vstride * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (precondType == 4)   // vtemp starts blank (AllocateBlankVector, crs.tcc:216): the first SGS application starts from zero
    CK(cudaMemsetAsync(vtemp, 0, vstride * sizeof(double), c->stream));
  std::vector<double> g((size_t)nSearchDir + 2, 0.0), Q((size_t)2 * (nSearchDir + 1), 0.0),
      H((size_t)(nSearchDir + 2) * (nSearchDir + 2), 0.0);
  std::vector<int> Hoffset((size_t)nSearchDir + 1, 0);
  int idir = 0;
  double dot = 0.0;
  for (int irestart = 0; irestart < restarts; irestart++) {
    double* v0 = vdat;
    if (matvec(xs, v0)) return 1;
    if (gm_vec(c, GM_RES0, nloc, 0.0, nullptr, b, v0)) return 1;
    if (gm_dot(c, v0, v0, nloc, &dot)) return 1;
    dot = sqrt(dot);
    if (dot < smallnum) { idir = 0; break; }
    if (gm_vec(c, GM_DIVS, nloc, dot, nullptr, nullptr, v0)) return 1;
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = dot;
    int hpos = 0;
    for (idir = 0; idir < nSearchDir; idir++) {
      double* vk = vdat + (size_t)idir * vstride;
      Hoffset[idir] = hpos;
      if (precond(vk, vtemp)) return 1;
      if (dist && comm_update(c, PCFD_F_X)) return 1;          // crs.tcc:300
      if (matvec(vtemp, uk)) return 1;
      for (int j = 0; j <= idir; j++) {
        const double* vj = vdat + (size_t)j * vstride;
        if (gm_dot(c, uk, vj, nloc, &dot)) return 1;
        H[hpos++] = dot;
        if (gm_vec(c, GM_ORTHO, nloc, dot, vj, nullptr, uk)) return 1;
      }
      if (gm_dot(c, uk, uk, nloc, &dot)) return 1;
      dot = sqrt(dot);
      H[hpos++] = dot;
      if (dot < smallnum) break;
      if (gm_vec(c, GM_SCALE_TO, nloc, dot, uk, nullptr, vdat + (size_t)(idir + 1) * vstride)) return 1;
      for (int jj = 0; jj < idir; jj++) {
        const double cs = Q[jj * 2], sn = Q[jj * 2 + 1];
        const double t1 = H[Hoffset[idir] + jj], t2 = H[Hoffset[idir] + jj + 1];
        H[Hoffset[idir] + jj] = cs * t1 + sn * t2;
        H[Hoffset[idir] + jj + 1] = -sn * t1 + cs * t2;
      }
      const double a2 = H[Hoffset[idir] + idir + 1], a1 = H[Hoffset[idir] + idir];
      const double alpha = sqrt(a1 * a1 + a2 * a2);
      const double cs = a1 / alpha, sn = a2 / alpha;
      Q[idir * 2] = cs; Q[idir * 2 + 1] = sn;
      H[Hoffset[idir] + idir] = alpha;
      H[Hoffset[idir] + idir + 1] = 0.0;
      const double t1 = g[idir], t2 = g[idir + 1];
      g[idir] = cs * t1 + sn * t2;
      g[idir + 1] = -sn * t1 + cs * t2;
    }
    for (int jj = idir - 1; jj >= 0; jj--) {
      double t1 = 0.0;
      for (int ii = jj + 1; ii <= idir - 1; ii++) t1 += H[Hoffset[ii] + jj] * g[ii];
      g[jj] -= t1;
      g[jj] /= H[Hoffset[jj] + jj];
    }
    CK(cudaMemsetAsync(uk, 0, (size_t)nloc * sizeof(double), c->stream));
    for (int jj = 0; jj <= idir - 1; jj++)
      if (gm_vec(c, GM_ACCUM, nloc, g[jj], vdat + (size_t)jj * vstride, nullptr, uk)) return 1;
    if (precond(uk, vtemp)) return 1;
    if (gm_vec(c, GM_ADD, nloc, 0.0, vtemp, nullptr, xs)) return 1;
    if (dist) {   // p->UpdateGeneralVectors(x): through the exchangeable field (vtemp, the next SGS guess, kept aside)
      if (precondType == 4) CK(cudaMemcpyAsync(uk, vtemp, vstride * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      CK(cudaMemcpyAsync(c->f[PCFD_F_X], xs, vstride * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      if (comm_update(c, PCFD_F_X)) return 1;
      CK(cudaMemcpyAsync(xs, c->f[PCFD_F_X], vstride * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      if (precondType == 4) CK(cudaMemcpyAsync(vtemp, uk, vstride * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  CK(cudaMemcpyAsync(c->f[PCFD_F_X], xs, vstride * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (dq_norm) *dq_norm = fabs(g[idir]);
  return 0;
}

}  // namespace

extern "C" int pcfd_gmres(pcfd_ctx* c, int restarts, int nsearch, int precond_type, double* dq_norm) {
  if (!c) return 1;
  if (restarts < 1 || nsearch < 1 || nsearch > 200) return fail(c, "pcfd_gmres: restarts >= 1, 1 <= search directions <= 200");
  if (precond_type < 0 || precond_type > 4)
    return fail(c, "pcfd_gmres: preconditioner 0 (none), 1 (diagonal), 2 (block diagonal), 3 (local ILU0) or 4 (SGS)");
  CK(cudaSetDevice(c->device));
  if (!c->f[PCFD_F_A]) return fail(c, "pcfd_gmres: no matrix (pcfd_jacobian or pcfd_set_field(PCFD_F_A) first)");
  if (c->ludiag) return fail(c, "pcfd_gmres: the diagonal blocks have been factored in place (pcfd_prepare_sgs); GMRES needs the assembled matrix");
  if (c->neqn == 5) return gmres_impl<5>(c, restarts, nsearch, precond_type, dq_norm);
  if (c->neqn == 9) return gmres_impl<9>(c, restarts, nsearch, precond_type, dq_norm);
  return fail(c, "pcfd_gmres: block size not instantiated");
}
