// pcfd_forces.cuh -- surface forces on the device (included by pcfd_kernels.cu).
//
// What the reference does once per iteration on the host (ucs/forces.tcc, called at solutionSpace.tcc:884):
//   ComputeSurfaceAreas (:199-312, once in PreIterate)  per-factag |n_j| area sums; per body the planform area projected
//                                                       on the lift direction
//   FORCE_Kernel (:123-196)   cp per half-edge; per composite body the sums of p n A, r x (p n A) and -- no-slip half-
//                             edges of a viscous run -- of the wall shear stress vector and its moment
//   YpCf_Kernel (:400-478)    y+ and cf per no-slip half-edge
//   ComputeCl (:326-369)      lift / drag / moment coefficients per body
// Here: one thread per BC half-edge for everything local (k_forces_bedges: the reference's operations in the reference's
// order, so cp / y+ / cf and every per-half-edge force term are bit-identical wherever no libm call is involved), then a
// fixed-shape tree sum per body (k_forces_partial / k_forces_final).  The body sums therefore differ from the reference's
// sequential `+=` over half-edges by summation order only (tests: 1e-12 of the sum of magnitudes); they are the same on
// every run.  Across ranks the per-rank sums are added in rank order (the reference: MPI_Allreduce).
#pragma once

struct pcfd_forces {
  int nbodies = 0, num_bcs = 0;
  double liftdir[3], dragdir[3], V = 0.0;
  std::vector<double> moment_pt, moment_axis, surf_area, body_area;
  unsigned* mask = nullptr;      // [nbedge] bit b: the half-edge's surface belongs to body b
  double* cgr = nullptr;         // [nbedge*3] Mesh::cg of the half-edge's phantom node (centroid of its boundary face piece)
  double* mpt = nullptr;         // [nbodies*3]
  double* props = nullptr;       // [nbedge*4] {p, cp, mu, rho} of the surface node's state, per eqnset
  double* terms = nullptr;       // [nbedge*6] pressure force, viscous force
  double *cp = nullptr, *yp = nullptr, *cf = nullptr;   // [nbedge]
  double *partial = nullptr, *sums = nullptr;           // [nbodies*FORCE_BLOCKS*12], [nbodies*12]
  double* hsums = nullptr;       // pinned [nbodies*12]
};

namespace {

constexpr int FORCE_BLOCKS = 64;

// EqnSet::GetPressure / GetCp / ComputeViscosity / GetDensity of the surface node (compressible.tcc:1156, 1174-1180;
// eqnset.h:251-268), perfect gas
__global__ void k_surface_props(DevMesh m, eq::ViscParams vp, double gamma, double V, bool viscous,
                                const double* __restrict__ q, double* __restrict__ props) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nbedge) return;
  const double* Q = q + (size_t)m.ben[e].x * NVARS;
  const double P = Q[6];
  props[4 * (size_t)e] = P;
  props[4 * (size_t)e + 1] = ((P - 1.0 / gamma) / (0.5 * V * V));
  props[4 * (size_t)e + 2] = viscous ? eq::viscosity(vp, Q[5]) : 0.0;
  props[4 * (size_t)e + 3] = Q[0];
}

// ComputeStressVector (compressible.tcc:1113-1153, compressibleFR.tcc:2204-2242); vg = the nine velocity-gradient entries
__device__ __forceinline__ void stress_vector(const double* vg, const double* av, double mu, double reScale, double* stress) {
  const double ux = vg[0], uy = vg[1], uz = vg[2], vx = vg[3], vy = vg[4], vz = vg[5], wx = vg[6], wy = vg[7], wz = vg[8];
  const double div = -2.0 / 3.0 * (ux + vy + wz);
  const double tauxx = 2.0 * ux + div, tauyy = 2.0 * vy + div, tauzz = 2.0 * wz + div;
  const double tauxy = uy + vx, tauxz = uz + wx, tauyz = vz + wy;
  const double tauxn = tauxx * av[0] + tauxy * av[1] + tauxz * av[2];
  const double tauyn = tauxy * av[0] + tauyy * av[1] + tauyz * av[2];
  const double tauzn = tauxz * av[0] + tauyz * av[1] + tauzz * av[2];
  stress[0] = -(mu / reScale) * tauxn;
  stress[1] = -(mu / reScale) * tauyn;
  stress[2] = -(mu / reScale) * tauzn;
}

// FORCE_Kernel + YpCf_Kernel for one BC half-edge (ghost half-edges contribute nothing in the reference)
__global__ void __launch_bounds__(128) k_forces_bedges(DevMesh m, const double* __restrict__ props,
                                                        const double* __restrict__ qgrad, int gstride, int goff,
                                                        double reScale, double V, bool viscous, const int* __restrict__ ia,
                                                        const int* __restrict__ ja, double* __restrict__ terms,
                                                        double* __restrict__ cp, double* __restrict__ yp,
                                                        double* __restrict__ cf) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.nbedge) return;
  const int l = m.ben[e].x;
  double av[4];
  load_avec(m.bea, e, av);
  const double pr = props[4 * (size_t)e];
  cp[e] = props[4 * (size_t)e + 1];
  double* t = terms + 6 * (size_t)e;
#pragma unroll
  for (int j = 0; j < 3; j++) t[j] = pr * av[j] * av[3];
  double ypv = 0.0, cfv = 0.0, vf[3] = {0.0, 0.0, 0.0};
  if (m.bctype[e] == PCFD_BC_NOSLIP && viscous) {
    const double mu = props[4 * (size_t)e + 2], rho = props[4 * (size_t)e + 3];
    double vg[9], stress[3];
#pragma unroll
    for (int i = 0; i < 9; i++) vg[i] = qgrad[(size_t)l * gstride + goff + i];
    stress_vector(vg, av, mu, reScale, stress);
#pragma unroll
    for (int j = 0; j < 3; j++) vf[j] = stress[j] * av[3];
    // the most wall-normal neighbour (psp order = the row's ja order after the diagonal; ties: the later one)
    const double wx = m.xyz[3 * l], wy = m.xyz[3 * l + 1], wz = m.xyz[3 * l + 2];
    double d = 0.0, dotmax = 0.0;
    for (int k = ia[l] + 1; k < ia[l + 1]; k++) {
      const int pt = ja[k];
      const double ex = m.xyz[3 * pt] - wx, ey = m.xyz[3 * pt + 1] - wy, ez = m.xyz[3 * pt + 2] - wz;
      const double mag = sqrt(ex * ex + ey * ey + ez * ez);
      const double nx = ex / mag, ny = ey / mag, nz = ez / mag;
      const double dot = -(nx * av[0] + ny * av[1] + nz * av[2]);
      if (dot >= dotmax) { d = mag; dotmax = dot; }   // Distance(ptx, wallx) is the same sqrt of the same sum
    }
    const double nu = mu / rho;
    const double tauw = sqrt(stress[0] * stress[0] + stress[1] * stress[1] + stress[2] * stress[2]);
    ypv = d * sqrt(tauw / rho) / nu * reScale;
    cfv = (tauw / (0.5 * rho * V * V));
  }
#pragma unroll
  for (int j = 0; j < 3; j++) t[3 + j] = vf[j];
  yp[e] = ypv;
  cf[e] = cfv;
}

// per body: sums of force, viscous force, moment, viscous moment over the half-edges of its surfaces.  Fixed grid,
// strided walk, shared-memory tree: the same sum on every run.
__global__ void __launch_bounds__(256) k_forces_partial(int nbedge, const unsigned* __restrict__ mask,
                                                         const double* __restrict__ terms, const double* __restrict__ cgr,
                                                         const double* __restrict__ mpt, double* __restrict__ partial) {
  __shared__ double sh[256];
  const int body = blockIdx.y;
  double acc[12];
#pragma unroll
  for (int k = 0; k < 12; k++) acc[k] = 0.0;
  const double m0 = mpt[3 * body], m1 = mpt[3 * body + 1], m2 = mpt[3 * body + 2];
  for (int e = blockIdx.x * 256 + threadIdx.x; e < nbedge; e += gridDim.x * 256) {
    if (!((mask[e] >> body) & 1u)) continue;
    const double r0 = cgr[3 * (size_t)e] - m0, r1 = cgr[3 * (size_t)e + 1] - m1, r2 = cgr[3 * (size_t)e + 2] - m2;
    const double* t = terms + 6 * (size_t)e;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const double f0 = t[3 * s], f1 = t[3 * s + 1], f2 = t[3 * s + 2];
      acc[3 * s] += f0; acc[3 * s + 1] += f1; acc[3 * s + 2] += f2;
      acc[6 + 3 * s] += r1 * f2 - f1 * r2;      // CrossProduct(rpos, tforces) (geometry.h:68-73)
      acc[6 + 3 * s + 1] += r2 * f0 - f2 * r0;
      acc[6 + 3 * s + 2] += r0 * f1 - f0 * r1;
    }
  }
  for (int k = 0; k < 12; k++) {
    sh[threadIdx.x] = acc[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[((size_t)body * gridDim.x + blockIdx.x) * 12 + k] = sh[0];
    __syncthreads();
  }
}
__global__ void k_forces_final(int nblocks, const double* __restrict__ partial, double* __restrict__ sums) {
  const int body = blockIdx.x, k = threadIdx.x;
  if (k >= 12) return;
  double a = 0.0;
  for (int b = 0; b < nblocks; b++) a += partial[((size_t)body * nblocks + b) * 12 + k];
  sums[body * 12 + k] = a;
}

// ComputeWallDistOct (ucs/walldist.tcc:116-199): for every local node (owned and ghost) the distance to the nearest viscous
// wall NODE of any rank.  The reference finds it through an octree; here every node looks at every wall point (1.7 M x 14 k
// pairs at 10 M cells: ~2e11 FP64 operations, tens of milliseconds, once per static mesh; the points of a warp's pass come
// out of L1 as a broadcast).  Per pair the arithmetic is `Distance` (geometry.h:30-37); the minimum is taken on the squares
// and the root once: sqrt is monotone and correctly rounded, so sqrt(min s) == min sqrt(s) bit for bit.
__global__ void __launch_bounds__(128) k_wall_distance(int nn, const double* __restrict__ xyz, int npts,
                                                        const double* __restrict__ pts, double* __restrict__ dist) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  const double x = xyz[3 * (size_t)n], y = xyz[3 * (size_t)n + 1], z = xyz[3 * (size_t)n + 2];
  double best = __longlong_as_double(0x7ff0000000000000LL);   // +inf: no viscous wall anywhere
  for (int k = 0; k < npts; k++) {
    const double dx = x - __ldg(pts + 3 * (size_t)k), dy = y - __ldg(pts + 3 * (size_t)k + 1), dz = z - __ldg(pts + 3 * (size_t)k + 2);
    const double s = dx * dx + dy * dy + dz * dz;
    best = (s < best) ? s : best;
  }
  dist[n] = sqrt(best);
}

}  // namespace

// ---------------------------------------------------------------- host side
static int forces_free(pcfd_ctx* c) {
  pcfd_forces* f = c->forces;
  if (!f) return 0;
  void* dev[] = {f->mask, f->cgr, f->mpt, f->props, f->terms, f->cp, f->yp, f->cf, f->partial, f->sums};
  for (void* p : dev) if (p) cudaFree(p);
  if (f->hsums) cudaFreeHost(f->hsums);
  delete f;
  c->forces = nullptr;
  return 0;
}

extern "C" {

int pcfd_forces_configure(pcfd_ctx* c, const pcfd_forces_desc* d) {
  if (!c) return 1;
  if (!d || !d->body_offsets || !d->body_factags || !d->moment_pt || !d->moment_axis || !d->bedges_factag || !d->cg)
    return fail(c, "pcfd_forces_configure: null argument");
  if (d->nbodies < 1 || d->nbodies > 32) return fail(c, "pcfd_forces_configure: 1..32 composite bodies");
  if (d->num_bcs < 0) return fail(c, "pcfd_forces_configure: bad num_bcs");
  if (!(d->velocity > 0.0)) return fail(c, "pcfd_forces_configure: velocity (Param::velocity) must be positive");
  CK(cudaSetDevice(c->device));
  forces_free(c);
  pcfd_forces* f = new pcfd_forces();
  c->forces = f;
  f->nbodies = d->nbodies; f->num_bcs = d->num_bcs; f->V = d->velocity;
  for (int j = 0; j < 3; j++) { f->liftdir[j] = d->liftdir[j]; f->dragdir[j] = d->dragdir[j]; }
  f->moment_pt.assign(d->moment_pt, d->moment_pt + 3 * d->nbodies);
  f->moment_axis.assign(d->moment_axis, d->moment_axis + 3 * d->nbodies);
  const int nbe = c->nbedge;
  std::vector<unsigned> mask(std::max(nbe, 1), 0u);
  std::vector<double> cgr((size_t)std::max(nbe, 1) * 3, 0.0), bea((size_t)std::max(nbe, 1) * 4);
  std::vector<int2> ben(std::max(nbe, 1));
  std::vector<int> bct(std::max(nbe, 1));
  if (nbe) {
    CK(cudaMemcpy(bea.data(), c->bea, (size_t)nbe * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ben.data(), c->ben, (size_t)nbe * sizeof(int2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(bct.data(), c->bctype, (size_t)nbe * sizeof(int), cudaMemcpyDeviceToHost));
  }
  // ComputeSurfaceAreas (forces.tcc:199-312), this rank's half-edges in half-edge order
  f->surf_area.assign((size_t)3 * (d->num_bcs + 1), 0.0);
  f->body_area.assign((size_t)3 * d->nbodies, 0.0);
  for (int e = 0; e < nbe; e++) {
    const int factag = d->bedges_factag[e];
    if (factag < 0 || factag > d->num_bcs) return fail(c, "pcfd_forces_configure: factag outside 0..num_bcs");
    const double* av = &bea[(size_t)e * 4];
    for (int j = 0; j < 3; j++) cgr[(size_t)e * 3 + j] = d->cg[(size_t)ben[e].y * 3 + j];
    if (bct[e] != PCFD_BC_PARALLEL)
      for (int j = 0; j < 3; j++) f->surf_area[(size_t)factag * 3 + j] += fabs(av[j] * av[3]);
    for (int b = 0; b < d->nbodies; b++) {
      bool part = false;
      for (int k = d->body_offsets[b]; k < d->body_offsets[b + 1]; k++) part = part || d->body_factags[k] == factag;
      if (!part) continue;
      mask[e] |= 1u << b;
      if (bct[e] == PCFD_BC_PARALLEL) continue;
      const double dot = d->liftdir[0] * av[0] + d->liftdir[1] * av[1] + d->liftdir[2] * av[2];
      if (dot >= 0.0)
        for (int j = 0; j < 3; j++) f->body_area[(size_t)b * 3 + j] += fabs(dot * av[j] * av[3]);
    }
  }
  auto up = [&](auto** dst, const auto* src, size_t n) -> int {
    CK(cudaMalloc(reinterpret_cast<void**>(dst), std::max<size_t>(n, 1) * sizeof(**dst)));
    if (src && n) CK(cudaMemcpy(*dst, src, n * sizeof(**dst), cudaMemcpyHostToDevice));
    return 0;
  };
  if (up(&f->mask, mask.data(), (size_t)nbe) || up(&f->cgr, cgr.data(), (size_t)nbe * 3) ||
      up(&f->mpt, f->moment_pt.data(), (size_t)3 * d->nbodies) || up(&f->props, (const double*)nullptr, (size_t)nbe * 4) ||
      up(&f->terms, (const double*)nullptr, (size_t)nbe * 6) || up(&f->cp, (const double*)nullptr, (size_t)nbe) ||
      up(&f->yp, (const double*)nullptr, (size_t)nbe) || up(&f->cf, (const double*)nullptr, (size_t)nbe) ||
      up(&f->partial, (const double*)nullptr, (size_t)d->nbodies * FORCE_BLOCKS * 12) ||
      up(&f->sums, (const double*)nullptr, (size_t)d->nbodies * 12))
    return 1;
  CK(cudaMallocHost(reinterpret_cast<void**>(&f->hsums), (size_t)d->nbodies * 12 * sizeof(double)));
  return 0;
}

int pcfd_forces_areas(pcfd_ctx* c, double* surf_area, double* body_area) {
  if (!c) return 1;
  if (!c->forces) return fail(c, "pcfd_forces_areas: pcfd_forces_configure has not been called");
  if (surf_area) memcpy(surf_area, c->forces->surf_area.data(), c->forces->surf_area.size() * sizeof(double));
  if (body_area) memcpy(body_area, c->forces->body_area.data(), c->forces->body_area.size() * sizeof(double));
  return 0;
}

int pcfd_forces_compute(pcfd_ctx* c, double* body, double* coef) {
  if (!c) return 1;
  pcfd_forces* f = c->forces;
  if (!f) return fail(c, "pcfd_forces_compute: pcfd_forces_configure has not been called");
  CK(cudaSetDevice(c->device));
  const int nbe = c->nbedge, nbd = f->nbodies;
  double rho_inf, reScale;
  bool viscous;
  int goff;
  if (c->fr) {
    if (pcfd_fr_surface_props(c, f->V, f->props, &rho_inf, &viscous)) return 1;
    reScale = c->prm.Re;
    goff = (c->neqn - 4) * 3;
  } else {
    viscous = c->viscous;
    rho_inf = c->prm.qinf[0];
    reScale = viscous ? c->vp.Re / c->vp.mach : 1.0;
    goff = 3;
    if (nbe) {
      PROF("k_surface_props");
      k_surface_props<<<nblk(nbe, 128), 128, 0, c->stream>>>(c->dm, c->vp, c->prm.gamma, f->V, viscous, c->f[PCFD_F_Q], f->props);
      LAUNCH_CHECK();
    }
  }
  if (nbe) {
    PROF("k_forces_bedges");
    k_forces_bedges<<<nblk(nbe, 128), 128, 0, c->stream>>>(c->dm, f->props, c->f[PCFD_F_QGRAD], c->nterms * 3, goff, reScale, f->V,
                                                          viscous, c->ia, c->ja, f->terms, f->cp, f->yp, f->cf);
    LAUNCH_CHECK();
  }
  PROF("k_forces_partial");
  k_forces_partial<<<dim3(FORCE_BLOCKS, nbd), 256, 0, c->stream>>>(nbe, f->mask, f->terms, f->cgr, f->mpt, f->partial);
  LAUNCH_CHECK();
  PROF("k_forces_final");
  k_forces_final<<<nbd, 32, 0, c->stream>>>(FORCE_BLOCKS, f->partial, f->sums);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(f->hsums, f->sums, (size_t)nbd * 12 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  std::vector<double> sums(f->hsums, f->hsums + (size_t)nbd * 12), area(f->body_area);
  if (comm_on(c)) {   // MPI_Allreduce of forces.tcc:383-388 and :236-243: per-rank sums added in rank order
    auto reduce = [&](std::vector<double>& v) -> int {
      for (size_t o = 0; o < v.size(); o += COMM_GW) {
        const int n = (int)std::min<size_t>(COMM_GW, v.size() - o);
        double all[COMM_MAXR * COMM_GW];
        if (pcfd_comm_allgather(c, v.data() + o, n, all)) return 1;
        for (int k = 0; k < n; k++) {
          double a = 0.0;
          for (int r = 0; r < c->nranks; r++) a += all[r * n + k];
          v[o + k] = a;
        }
      }
      return 0;
    };
    if (reduce(sums) || reduce(area)) return 1;
  }
  if (body) memcpy(body, sums.data(), sums.size() * sizeof(double));
  if (coef) {   // Forces::ComputeCl (forces.tcc:326-369)
    const double v2 = f->V * f->V;
    for (int b = 0; b < nbd; b++) {
      const double* B = &sums[(size_t)b * 12];
      const double* ax = &f->moment_axis[(size_t)b * 3];
      const double* ar = &area[(size_t)b * 3];
      double lift = f->liftdir[0] * B[0] + f->liftdir[1] * B[1] + f->liftdir[2] * B[2];
      double drag = f->dragdir[0] * B[0] + f->dragdir[1] * B[1] + f->dragdir[2] * B[2];
      double moment = ax[0] * B[6] + ax[1] * B[7] + ax[2] * B[8];
      lift += f->liftdir[0] * B[3] + f->liftdir[1] * B[4] + f->liftdir[2] * B[5];
      drag += f->dragdir[0] * B[3] + f->dragdir[1] * B[4] + f->dragdir[2] * B[5];
      moment += ax[0] * B[9] + ax[1] * B[10] + ax[2] * B[11];
      const double amag = sqrt(ar[0] * ar[0] + ar[1] * ar[1] + ar[2] * ar[2]);
      coef[3 * b] = lift / (0.5 * rho_inf * v2 * amag);
      coef[3 * b + 1] = drag / (0.5 * rho_inf * v2 * amag);
      coef[3 * b + 2] = -moment / (0.5 * rho_inf * v2 * amag * 1.0);
    }
  }
  return 0;
}

int pcfd_wall_distance(pcfd_ctx* c, const double* points, int npoints) {
  if (!c) return 1;
  if (npoints < 0 || (npoints > 0 && !points)) return fail(c, "pcfd_wall_distance: bad argument");
  if (c->fsize[PCFD_F_WALLDIST] == 0)
    return fail(c, "pcfd_wall_distance: the context has no wall-distance field (no turbulence model, no viscous far-field BC)");
  CK(cudaSetDevice(c->device));
  double* dpts = nullptr;
  CK(cudaMalloc(reinterpret_cast<void**>(&dpts), (size_t)std::max(npoints, 1) * 3 * sizeof(double)));
  if (npoints) CK(cudaMemcpy(dpts, points, (size_t)npoints * 3 * sizeof(double), cudaMemcpyHostToDevice));
  PROF("k_wall_distance");
  k_wall_distance<<<nblk(c->nn, 128), 128, 0, c->stream>>>(c->nn, c->xyz, npoints, dpts, c->f[PCFD_F_WALLDIST]);
  LAUNCH_CHECK();
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaFree(dpts));
  if (!c->ffv_edges.empty()) {   // the power-law profile of the viscous far-field BC is tabulated from this field on the host
    std::vector<double> h(c->fsize[PCFD_F_WALLDIST]);
    CK(cudaMemcpy(h.data(), c->f[PCFD_F_WALLDIST], h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    return pcfd_set_field(c, PCFD_F_WALLDIST, h.data(), h.size());
  }
  return 0;
}

int pcfd_forces_get(pcfd_ctx* c, int which, double* out) {
  if (!c) return 1;
  pcfd_forces* f = c->forces;
  if (!f) return fail(c, "pcfd_forces_get: pcfd_forces_configure has not been called");
  if (!out) return fail(c, "pcfd_forces_get: null argument");
  const double* src = which == PCFD_SURF_CP ? f->cp : which == PCFD_SURF_YPLUS ? f->yp : which == PCFD_SURF_CF ? f->cf : nullptr;
  if (!src) return fail(c, "pcfd_forces_get: which must be PCFD_SURF_CP, PCFD_SURF_YPLUS or PCFD_SURF_CF");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (c->nbedge) CK(cudaMemcpy(out, src, (size_t)c->nbedge * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
