// pcfd_host.hpp -- header-only C++ host shim between ProteusCFD's in-process seams and the
// C ABI of libpcfd_b200.so (include/pcfd.h).
//
// pcfd::DropIn<Space> is instantiated with the reference's own SolutionSpace<Real>
// (ucs/solutionSpace.h:94-113).  Each public method is named after, and replaces the body
// of, one reference phase call in SolutionSpace::NewtonIterate / PreTimeAdvance
// (ucs/solutionSpace.tcc:573-904); results land in the reference's own arrays
// (space->qgrad, space->limiter->l, space->crs->b/x, space->crs->A->M/pv, field "timestep",
// space->q), so the rest of ucs.x (norms, output, forces, restart) is untouched.
//
// The template only needs the public members it names; for the reacting eqnsets it also names the
// reference's CompressibleFREqnSet / ChemModel / Species / Reaction (compressibleFR.h, chem.h), so it
// is included where solutionSpace.tcc's own includes are in scope.  It is compiled into the real
// reference by oracle/Makefile (ref_harness_gpu).  Test hook: with PCFD_HOST_DUMP_FR_PARAMS=<file> in
// the environment the pcfd_fr_params handed to pcfd_create_fr is also written to <file>
// (tests/test_dropin_fr_params.py).  C++11, like the reference (make.opts:44).  Error behaviour follows the reference: a failed call is fatal
// (`Abort << ...`, ucs/exceptions.h:30-49); here the handler is pluggable and defaults to
// printing pcfd_last_error and calling std::abort().
#ifndef PCFD_HOST_HPP
#define PCFD_HOST_HPP

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pcfd.h"

namespace pcfd {

typedef void (*ErrorHandler)(const char* where, const char* message);

inline void DefaultErrorHandler(const char* where, const char* message) {
  std::fprintf(stderr, "pcfd: %s failed: %s\n", where, message);
  std::abort();
}

#ifndef PCFD_HOST_NO_MPI
// The exchange of PObj::TransposeCommCRS (parallel.tcc:54-338) on packed ghost-column blocks: block e belongs to the
// parallel half-edge ghost_edges[2e] (local node) -> ghost_edges[2e+1] (ghost id >= nnode) and is replaced by the block
// the ghost's owner holds for the mirrored edge.  A request is the pair (owner's row = gNodeLocalId of my ghost, my
// row) in LOCAL ids; the owner resolves it through its own ghost table.  Counts by MPI_Alltoall, pairs and blocks point
// to point, like the reference.  Plain arrays only, so that it can be exercised without a device
// (oracle/harness/transpose_route_test.cpp, tests/test_crs_transpose.py).
inline bool RouteTransposedGhostBlocks(int nnode, int ngedge, const int* ghost_edges, const int* gNodeOwner,
                                       const int* gNodeLocalId, int n2, double* blocks) {
  int np = 1;
  MPI_Comm_size(MPI_COMM_WORLD, &np);
  std::vector<int> recvc(np, 0), sendc(np, 0), roff(np + 1, 0), soff(np + 1, 0), owner((size_t)(ngedge > 0 ? ngedge : 1));
  for (int e = 0; e < ngedge; e++) {
    owner[e] = gNodeOwner[ghost_edges[2 * e + 1] - nnode];
    recvc[owner[e]]++;
  }
  MPI_Alltoall(recvc.data(), 1, MPI_INT, sendc.data(), 1, MPI_INT, MPI_COMM_WORLD);
  for (int p = 0; p < np; p++) { roff[p + 1] = roff[p] + recvc[p]; soff[p + 1] = soff[p] + sendc[p]; }
  const size_t nr = (size_t)(roff[np] > 0 ? roff[np] : 1), ns = (size_t)(soff[np] > 0 ? soff[np] : 1);
  std::vector<int> ask(2 * nr), slot(nr), fill(roff.begin(), roff.end() - 1), asked(2 * ns);
  for (int e = 0; e < ngedge; e++) {      // what I ask of each owner, in half-edge order, and the slot the answer fills
    const int k = fill[owner[e]]++;
    ask[2 * k] = gNodeLocalId[ghost_edges[2 * e + 1] - nnode];
    ask[2 * k + 1] = ghost_edges[2 * e];
    slot[k] = e;
  }
  std::vector<MPI_Request> rq((size_t)np), sq((size_t)np);
  for (int p = 0; p < np; p++) if (sendc[p]) MPI_Irecv(&asked[2 * soff[p]], 2 * sendc[p], MPI_INT, p, 0, MPI_COMM_WORLD, &rq[p]);
  for (int p = 0; p < np; p++) if (recvc[p]) MPI_Isend(&ask[2 * roff[p]], 2 * recvc[p], MPI_INT, p, 0, MPI_COMM_WORLD, &sq[p]);
  for (int p = 0; p < np; p++) {
    if (sendc[p]) MPI_Wait(&rq[p], MPI_STATUS_IGNORE);
    if (recvc[p]) MPI_Wait(&sq[p], MPI_STATUS_IGNORE);
  }
  // serve: (my row j, asking rank p, p's local node i) -> my half-edge whose ghost p owns under the local id i
  bool ok = true;
  std::vector<double> out(ns * n2, 0.0), in(nr * n2);
  for (int p = 0; p < np; p++)
    for (int k = soff[p]; k < soff[p + 1]; k++) {
      int found = -1;
      for (int e = 0; e < ngedge && found < 0; e++) {
        const int g = ghost_edges[2 * e + 1] - nnode;
        if (ghost_edges[2 * e] == asked[2 * k] && gNodeOwner[g] == p && gNodeLocalId[g] == asked[2 * k + 1]) found = e;
      }
      if (found < 0) { ok = false; continue; }      // keep the exchange going: every rank must reach its waits
      std::memcpy(&out[(size_t)k * n2], &blocks[(size_t)found * n2], sizeof(double) * n2);
    }
  for (int p = 0; p < np; p++) if (recvc[p]) MPI_Irecv(&in[(size_t)roff[p] * n2], recvc[p] * n2, MPI_DOUBLE, p, 1, MPI_COMM_WORLD, &rq[p]);
  for (int p = 0; p < np; p++) if (sendc[p]) MPI_Isend(&out[(size_t)soff[p] * n2], sendc[p] * n2, MPI_DOUBLE, p, 1, MPI_COMM_WORLD, &sq[p]);
  for (int p = 0; p < np; p++) {
    if (recvc[p]) MPI_Wait(&rq[p], MPI_STATUS_IGNORE);
    if (sendc[p]) MPI_Wait(&sq[p], MPI_STATUS_IGNORE);
  }
  for (int k = 0; k < roff[np]; k++) std::memcpy(&blocks[(size_t)slot[k] * n2], &in[(size_t)k * n2], sizeof(double) * n2);
  return ok;
}
#endif

#ifndef PCFD_HOST_ROUTING_ONLY   // (the device-free routing test includes the header without the reference's types)
template <class Space>
class DropIn {
 public:
  // SolutionSpace ctor + Init (solutionSpace.tcc:26-114, 148-376) have already built the mesh,
  // metrics, eqnset and CRS; flatten what the hot path reads and create the device context.
  explicit DropIn(Space* space, int device = 0, ErrorHandler onError = DefaultErrorHandler)
      : s_(space), ctx_(NULL), onError_(onError) {
    nnode_ = s_->m->GetNumNodes();
    gnode_ = s_->m->GetNumParallelNodes();
    nbnode_ = s_->m->GetNumBoundaryNodes();
    nedge_ = s_->m->GetNumEdges();
    nbedge_ = s_->m->GetNumBoundaryEdges();
    ngedge_ = s_->m->GetNumParallelEdges();
    neqn_ = s_->eqnset->neqn;
    nvars_ = neqn_ + s_->eqnset->nauxvars;
    nterms_ = s_->grad->GetNterms();
    const int nb = nbedge_ + ngedge_;

    std::vector<int> en(2 * (size_t)nedge_), bn(2 * (size_t)nb), bt((size_t)nb);
    std::vector<double> ea(4 * (size_t)nedge_), ba(4 * (size_t)nb), tw((size_t)(nbedge_ > 0 ? nbedge_ : 1));
    for (int e = 0; e < nedge_; e++) {          // Edges<Type>, uns_base.h:12-22
      en[2 * e] = s_->m->edges[e].n[0];
      en[2 * e + 1] = s_->m->edges[e].n[1];
      for (int k = 0; k < 4; k++) ea[4 * (size_t)e + k] = s_->m->edges[e].a[k];
    }
    for (int e = 0; e < nb; e++) {              // HalfEdges<Type>, uns_base.h:28-37
      bn[2 * e] = s_->m->bedges[e].n[0];
      bn[2 * e + 1] = s_->m->bedges[e].n[1];
      for (int k = 0; k < 4; k++) ba[4 * (size_t)e + k] = s_->m->bedges[e].a[k];
      bt[e] = s_->bc->GetBCType(s_->m->bedges[e].factag);   // bc.tcc:70-74
      // wall temperature as bc.tcc:1283-1286 non-dimensionalises it (read for NoSlip surfaces only)
      if (e < nbedge_) tw[e] = s_->bc->GetBCObj(s_->m->bedges[e].factag)->twall / s_->param->ref_temperature;
    }
    pcfd_mesh_desc md;
    md.nnode = nnode_; md.gnode = gnode_; md.nbnode = nbnode_;
    md.nedge = nedge_; md.nbedge = nbedge_; md.ngedge = ngedge_;
    md.edges_n = en.data(); md.edges_a = ea.data();
    md.bedges_n = bn.data(); md.bedges_a = ba.data(); md.bedges_bctype = bt.data();
    md.xyz = s_->m->xyz; md.vol = s_->m->vol; md.ipsp = s_->m->ipsp; md.psp = s_->m->psp;
    md.bedges_twall = tw.data();

    pcfd_params pr;
    pr.eqnset = s_->param->eqnset_id;
    pr.sorder = s_->param->sorder;
    pr.limiter = s_->param->limiter;
    pr.no_cvbc = s_->param->no_cvbc;
    pr.gamma = s_->param->gamma;
    pr.chi = s_->param->chi;
    pr.cfl = s_->param->GetCFL();
    for (int k = 0; k < 10; k++) pr.qinf[k] = (k < nvars_) ? s_->eqnset->Qinf[k] : 0.0;
    pr.enable_vnn = s_->param->enableVNN ? 1 : 0;
    pr.vnn = s_->param->VNN;
    pr.Re = s_->param->Re; pr.Pr = s_->param->Pr; pr.PrT = s_->param->PrT;
    pr.tref = s_->param->ref_temperature;
    pr.mach = s_->param->GetVelocity(s_->iter);
    pr.turb_model = s_->param->viscous ? s_->param->turbModel : 0;

    const bool reacting = pr.eqnset == PCFD_EQNSET_COMPRESSIBLE_EULER_FR || pr.eqnset == PCFD_EQNSET_COMPRESSIBLE_NS_FR;
    if (reacting) {
      // CompressibleFREqnSet (compressibleFR.h:12-23): chemistry tables, reference values, all nvars of Qinf
      std::vector<pcfd_fr_params> fpv(1);   // ~100 kB: keep it off the stack
      FillFrParams(fpv[0]);
      if (const char* path = std::getenv("PCFD_HOST_DUMP_FR_PARAMS")) {   // test hook: the struct as handed to pcfd_create_fr
        if (FILE* f = std::fopen(path, "wb")) { std::fwrite(&fpv[0], sizeof(pcfd_fr_params), 1, f); std::fclose(f); }
      }
      if (pcfd_create_fr(&md, &pr, &fpv[0], device, &ctx_) != 0) onError_("pcfd_create_fr", pcfd_last_error(NULL));
      if (!ctx_) return;
      // preconditioning field "beta" (solutionSpace.tcc:235-247); rows beyond the local + ghost nodes stay 1
      std::vector<double> beta(pcfd_field_size(ctx_, PCFD_F_BETA), 1.0);
      const double* hb = s_->GetFieldData("beta", FIELDS::STATE_NONE);
      for (int i = 0; i < nnode_ + gnode_ && (size_t)i < beta.size(); i++) beta[i] = hb[i];
      Check(pcfd_set_field(ctx_, PCFD_F_BETA, beta.data(), beta.size()), "set beta");
    } else
    if (pcfd_create(&md, &pr, device, &ctx_) != 0) onError_("pcfd_create", pcfd_last_error(NULL));
    if (!ctx_) return;   // a non-aborting handler gets here: leave an empty shim rather than touch a null context
    // Param::dt / useLocalTimeStepping / torder / iter as they stand now (steady runs: dt < 0); PushTimeIntegration
    // refreshes them per step
    if (s_->param->dt != 0.0)
      Check(pcfd_set_time_integration(ctx_, s_->param->dt, s_->param->useLocalTimeStepping ? 1 : 0,
                                      (s_->param->torder == 2) ? 2 : 1, s_->iter), "pcfd_set_time_integration");
    // Param::gradType (gradient.tcc:68-90) and Param::fieldJacType / boundaryJacType (jacobian.tcc:140-176)
    Check(pcfd_set_gradient_type(ctx_, s_->param->gradType), "pcfd_set_gradient_type");
    Check(pcfd_set_jacobian_type(ctx_, s_->param->fieldJacType, s_->param->boundaryJacType), "pcfd_set_jacobian_type");
    // Mesh::s / Mesh::sw were filled by ComputeNodeLSQCoefficients during Init: reuse them
    Check(pcfd_set_field(ctx_, PCFD_F_LSQ_S, s_->m->s, 6 * (size_t)(nnode_ + gnode_)), "set s");
    Check(pcfd_set_field(ctx_, PCFD_F_LSQ_SW, s_->m->sw, 6 * (size_t)(nnode_ + gnode_)), "set sw");
  }
  ~DropIn() { if (ctx_) pcfd_destroy(ctx_); }

  pcfd_ctx* Context() { return ctx_; }

  // ---- state in / out, in the reference's own arrays
  void PushQ() { Check(pcfd_set_field(ctx_, PCFD_F_Q, s_->q, NQ()), "PushQ"); }
  void PullQ() { Check(pcfd_get_field(ctx_, PCFD_F_Q, s_->q, NQ()), "PullQ"); }
  void PushB() { Check(pcfd_set_field(ctx_, PCFD_F_B, s_->crs->b, (size_t)nnode_ * neqn_), "PushB"); }
  void PullB() { Check(pcfd_get_field(ctx_, PCFD_F_B, s_->crs->b, (size_t)nnode_ * neqn_), "PullB"); }
  void PullX() { Check(pcfd_get_field(ctx_, PCFD_F_X, s_->crs->x, (size_t)(nnode_ + gnode_) * neqn_), "PullX"); }
  void PullGradient() {
    Check(pcfd_get_field(ctx_, PCFD_F_QGRAD, s_->qgrad, (size_t)(nnode_ + gnode_) * nterms_ * 3), "PullGradient");
  }
  void PullLimiter() {
    Check(pcfd_get_field(ctx_, PCFD_F_LIMITER, s_->limiter->l, (size_t)(nnode_ + gnode_) * neqn_), "PullLimiter");
  }
  void PullTimestep(double* dt) { Check(pcfd_get_field(ctx_, PCFD_F_TIMESTEP, dt, (size_t)nnode_), "PullTimestep"); }
  void PullMatrix() {   // CRSMatrix::M and ::pv
    Check(pcfd_get_field(ctx_, PCFD_F_A, s_->crs->A->M, (size_t)s_->crs->A->nblocks * neqn_ * neqn_), "PullMatrix");
    Check(pcfd_get_crs(ctx_, NULL, NULL, NULL, s_->crs->A->pv), "PullMatrix pv");
  }

  // ---- phase replacements (same order and meaning as the reference calls they stand for)
  void ComputeNodeLSQCoefficients() { Check(pcfd_lsq_coefficients(ctx_), "ComputeNodeLSQCoefficients"); }  // gradient.tcc:115
  void UpdateBCs() { Check(pcfd_update_bcs(ctx_), "UpdateBCs"); }                                          // bc.tcc:1399
  void GradientCompute() { Check(pcfd_gradient(ctx_), "Gradient::Compute"); }                              // gradient.tcc:57
  void LimiterCompute() { Check(pcfd_limiter(ctx_), "Limiter::Compute"); }                                 // limiters.tcc:53
  // residual.tcc:13-63: returns [resGlobal, res_0 .. res_{neqn-1}] with the reference's norm sqrt(sum)/N
  // (parallel.h:160-219); single rank here -- with several ranks the caller all-reduces the sums first
  std::vector<double> ComputeResiduals() {
    double ss[1 + PCFD_CHEM_MAX_SPECIES + 4];
    Check(pcfd_residual(ctx_, ss), "ComputeResiduals");
    std::vector<double> res(1 + neqn_);
    res[0] = std::sqrt(ss[0]) / ((double)nnode_ * neqn_);
    for (int k = 0; k < neqn_; k++) res[1 + k] = std::sqrt(ss[1 + k]) / (double)nnode_;
    return res;
  }
  double ComputeTimesteps() {                                                                              // timestep.tcc:7
    double dtmin = 0.0;
    Check(pcfd_set_cfl(ctx_, s_->param->GetCFL()), "set CFL");
    Check(pcfd_timestep(ctx_, &dtmin), "ComputeTimesteps");
    return dtmin;
  }
  void ComputeJacobians() { Check(pcfd_jacobian(ctx_), "ComputeJacobians"); }                              // jacobian.tcc:13
  void PrepareSGS() { Check(pcfd_prepare_sgs(ctx_), "CRSMatrix::PrepareSGS"); }                            // crsmatrix.tcc:840
  void BlankX() { Check(pcfd_blank_x(ctx_), "CRS::BlankX"); }                                              // crs.tcc:448
  double SGS(int nSgs) {                                                                                   // crs.tcc:62
    double ddq = 0.0;
    Check(pcfd_sgs(ctx_, nSgs, &ddq), "CRS::SGS");
    return ddq;
  }
  // TurbulenceModel::Compute (turb.tcc:163-339), Spalart-Allmaras: state in / out through the reference's own
  // TurbulenceModel::tvar, field "wallDistance" and field "mut"; returns the reference's residual norm sqrt(sum)/N
  void PushTurbulence() {
    Check(pcfd_set_field(ctx_, PCFD_F_TVAR, s_->turb->tvar, (size_t)(nnode_ + gnode_ + nbnode_)), "push tvar");
    Check(pcfd_set_field(ctx_, PCFD_F_WALLDIST, s_->GetFieldData("wallDistance", FIELDS::STATE_NONE),
                         (size_t)(nnode_ + gnode_)), "push wallDistance");
  }
  double TurbulenceCompute(int nSgs) {
    double ss = 0.0;
    Check(pcfd_turb_compute(ctx_, nSgs, &ss), "TurbulenceModel::Compute");
    return std::sqrt(ss) / (double)nnode_;
  }
  void PullTurbulence() {
    Check(pcfd_get_field(ctx_, PCFD_F_TVAR, s_->turb->tvar, (size_t)(nnode_ + gnode_ + nbnode_)), "pull tvar");
    std::vector<double> mut((size_t)(nnode_ + gnode_ + nbnode_));
    Check(pcfd_get_field(ctx_, PCFD_F_MUT, mut.data(), mut.size()), "pull mut");
    double* dst = s_->GetFieldData("mut", FIELDS::STATE_NONE);
    for (int i = 0; i < nnode_ + gnode_; i++) dst[i] = mut[i];
  }
  // ComputeWallDistOct (walldist.tcc:116-199; SolutionSpace::Init for viscous runs): the field "wallDistance" from an exact
  // search on the device.  The viscous wall nodes -- left nodes of the no-slip half-edges, half-edge order -- of all ranks
  // are gathered with the host's MPI as SyncParallelPoint does (walldist.tcc:24-113).
  void ComputeWallDistance() {
    auto* m = s_->m;
    std::vector<double> pts;
    for (int e = 0; e < m->GetNumBoundaryEdges(); e++) {
      if (s_->bc->GetBCType(m->bedges[e].factag) != Proteus_NoSlip) continue;
      const int l = m->bedges[e].n[0];
      for (int k = 0; k < 3; k++) pts.push_back(m->xyz[l * 3 + k]);
    }
#ifndef PCFD_HOST_NO_MPI
    int np = 1;
    MPI_Comm_size(MPI_COMM_WORLD, &np);
    if (np > 1) {
      // (MPI_Allgather on padded blocks: the reference itself uses no MPI_Allgatherv, nor does the test shim provide one)
      int mine = (int)pts.size(), maxc = 0;
      std::vector<int> counts(np);
      MPI_Allgather(&mine, 1, MPI_INT, &counts[0], 1, MPI_INT, MPI_COMM_WORLD);
      for (int r = 0; r < np; r++) maxc = std::max(maxc, counts[r]);
      std::vector<double> padded((size_t)std::max(maxc, 1), 0.0), blocks((size_t)std::max(maxc, 1) * np), all;
      std::copy(pts.begin(), pts.end(), padded.begin());
      MPI_Allgather(&padded[0], std::max(maxc, 1), MPI_DOUBLE, &blocks[0], std::max(maxc, 1), MPI_DOUBLE, MPI_COMM_WORLD);
      for (int r = 0; r < np; r++)
        all.insert(all.end(), blocks.begin() + (size_t)r * std::max(maxc, 1), blocks.begin() + (size_t)r * std::max(maxc, 1) + counts[r]);
      pts.swap(all);
    }
#endif
    Check(pcfd_wall_distance(ctx_, pts.empty() ? NULL : &pts[0], (int)(pts.size() / 3)), "ComputeWallDistOct");
    Check(pcfd_get_field(ctx_, PCFD_F_WALLDIST, s_->GetFieldData("wallDistance", FIELDS::STATE_NONE),
                         (size_t)(nnode_ + gnode_)), "pull wallDistance");
  }
  // Forces::Compute (forces.tcc:315-324, called at solutionSpace.tcc:884, 924) on the device-resident q and qgrad: body
  // sums, coefficients and the per-half-edge cp / y+ / cf land where the reference keeps them (Forces::bodies, cp, yp, cf;
  // scalar fields CL / CM).  The bodies, Param::liftdir / dragdir and Mesh::cg are handed over on the first call
  // (pcfd_forces_configure also evaluates ComputeSurfaceAreas; the areas are written to Forces::surfArea / bodies[].surfArea
  // -- single rank: as they are, several ranks: the host sums them as forces.tcc:236-243 does).
  void ComputeForces() {
    auto* f = s_->forces;
    auto* m = s_->m;
    const int nbodies = f->num_bodies, nbedge = m->GetNumBoundaryEdges();
    if (nbodies < 1) return;
    if (!forces_ready_) {
      std::vector<int> offs(1, 0), tags, factag((size_t)std::max(nbedge, 1));
      std::vector<double> mpt, max_;
      for (int i = 1; i <= nbodies; i++) {
        for (int k = 0; k < f->bodies[i].nsurfs; k++) tags.push_back(f->bodies[i].list[k]);
        offs.push_back((int)tags.size());
        for (int k = 0; k < 3; k++) { mpt.push_back(f->bodies[i].momentPt[k]); max_.push_back(f->bodies[i].momentAxis[k]); }
      }
      if (tags.empty()) tags.push_back(-1);
      for (int e = 0; e < nbedge; e++) factag[e] = m->bedges[e].factag;
      pcfd_forces_desc d;
      d.nbodies = nbodies; d.num_bcs = f->num_bcs;
      d.body_offsets = offs.data(); d.body_factags = tags.data();
      d.moment_pt = mpt.data(); d.moment_axis = max_.data();
      d.bedges_factag = factag.data(); d.cg = m->cg;
      for (int k = 0; k < 3; k++) { d.liftdir[k] = s_->param->liftdir[k]; d.dragdir[k] = s_->param->dragdir[k]; }
      d.velocity = s_->param->GetVelocity(s_->iter);
      Check(pcfd_forces_configure(ctx_, &d), "pcfd_forces_configure");
      std::vector<double> ba((size_t)3 * nbodies);
      Check(pcfd_forces_areas(ctx_, f->surfArea, ba.data()), "pcfd_forces_areas");
      for (int i = 1; i <= nbodies; i++)
        for (int k = 0; k < 3; k++) f->bodies[i].surfArea[k] = ba[(size_t)3 * (i - 1) + k];
      forces_ready_ = true;
    }
    std::vector<double> body((size_t)12 * nbodies), coef((size_t)3 * nbodies);
    Check(pcfd_forces_compute(ctx_, body.data(), coef.data()), "Forces::Compute");
    for (int i = 1; i <= nbodies; i++) {
      const double* B = &body[(size_t)12 * (i - 1)];
      for (int k = 0; k < 3; k++) {
        f->bodies[i].forces[k] = B[k]; f->bodies[i].vforces[k] = B[3 + k];
        f->bodies[i].moments[k] = B[6 + k]; f->bodies[i].vmoments[k] = B[9 + k];
      }
      f->bodies[i].cl = coef[(size_t)3 * (i - 1)]; f->bodies[i].cd = coef[(size_t)3 * (i - 1) + 1];
      f->bodies[i].cm = coef[(size_t)3 * (i - 1) + 2];
    }
    Check(pcfd_forces_get(ctx_, PCFD_SURF_CP, f->cp), "pull cp");
    Check(pcfd_forces_get(ctx_, PCFD_SURF_YPLUS, f->yp), "pull y+");
    Check(pcfd_forces_get(ctx_, PCFD_SURF_CF, f->cf), "pull cf");
    s_->GetScalarField("CL").SetField(f->bodies[1].cl);
    s_->GetScalarField("CM").SetField(f->bodies[1].cm);
  }
  // The body of SolutionSpace::NewtonIterate as ONE library call each (solutionSpace.tcc:640-904): the phases above in the
  // reference's order, state resident on the device, halos inside when the ranks are connected.  Returns what the
  // phase-wise calls return: this rank's sum of b^2 (and, implicit, |xOld - xNorm| of the last two sweeps through ddq).
  double NewtonIterateExplicit(bool refreshTimesteps) {
    double ss = 0.0;
    Check(pcfd_explicit_iterate(ctx_, refreshTimesteps ? 1 : 0, &ss), "NewtonIterate (explicit)");
    return ss;
  }
  double NewtonIterateImplicit(int nSgs, bool refreshJacobian, double* ddq = NULL) {
    double ss = 0.0, d = 0.0;
    Check(pcfd_implicit_iterate(ctx_, refreshJacobian ? 1 : 0, nSgs, &ss, &d), "NewtonIterate (implicit)");
    if (ddq) *ddq = d;
    return ss;
  }
  // CRS::GMRES(restarts, nSearchDir, precondType, ...) (crs.tcc:176-415) on the matrix of ComputeJacobians (not yet
  // factored by PrepareSGS), b and x of the context; precondType 0 none, 1 diagonal, 2 block diagonal, 3 local ILU0, 4 SGS
  double GMRES(int restarts, int nSearchDir, int precondType) {
    double dq = 0.0;
    Check(pcfd_gmres(ctx_, restarts, nSearchDir, precondType, &dq), "CRS::GMRES");
    return dq;
  }
  // nodes whose NaN / Inf update ApplyDQ zeroed so far (solutionSpace.tcc:771-796 prints this count)
  long long ZeroedUpdates() { return pcfd_zeroed_updates(ctx_); }
  void ExplicitSolve() { Check(pcfd_explicit_solve(ctx_), "ExplicitSolve"); }                              // solve.tcc:71
  void ApplyDQ() { Check(pcfd_apply_dq(ctx_), "ApplyDQ"); }                                                // solutionSpace.tcc:802

#ifndef PCFD_HOST_NO_MPI
  // ---- several ranks, one process (and one GPU) each.  PObj keeps its comm maps private, so ConnectRanks rebuilds them
  // the way PObj::BuildCommMaps does (parallel.tcc:461-554) from the mesh's public ghost tables (Mesh::gNodeOwner /
  // gNodeLocalId, mesh.h:205-210) with the host's own MPI: receive counts per owner, MPI_Alltoall for the send counts,
  // the requested local ids point to point.  The device-resident exchange then needs ONE more collective, ever: an
  // MPI_Allgather of the pcfd_comm blobs (CUDA-IPC handles of the fields and of a flag page); after pcfd_comm_connect
  // every PObj::UpdateGeneralVectors(v, n) of the hot path becomes UpdateGeneralVectors(PCFD_F_*): a put kernel and a
  // wait kernel, no MPI, no host copy (parallel.tcc:779-873).
  void ConnectRanks() {
    int rank = 0, np = 1;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &np);
    std::vector<int> recvc(np, 0), sendc(np, 0), roff(np + 1, 0), soff(np + 1, 0);
    for (int g = 0; g < gnode_; g++) recvc[s_->m->gNodeOwner[g]]++;                    // parallel.tcc:482-484
    MPI_Alltoall(recvc.data(), 1, MPI_INT, sendc.data(), 1, MPI_INT, MPI_COMM_WORLD);  // :497
    for (int p = 0; p < np; p++) { roff[p + 1] = roff[p] + recvc[p]; soff[p + 1] = soff[p] + sendc[p]; }
    std::vector<int> list((size_t)(soff[np] > 0 ? soff[np] : 1));
    std::vector<MPI_Request> rq((size_t)np), sq((size_t)np);
    for (int p = 0; p < np; p++)                                                       // :523-541: who wants which of my nodes
      if (sendc[p]) MPI_Irecv(&list[soff[p]], sendc[p], MPI_INT, p, 0, MPI_COMM_WORLD, &rq[p]);
    for (int p = 0; p < np; p++)
      if (recvc[p]) MPI_Isend(&s_->m->gNodeLocalId[roff[p]], recvc[p], MPI_INT, p, 0, MPI_COMM_WORLD, &sq[p]);
    for (int p = 0; p < np; p++) {
      if (sendc[p]) MPI_Wait(&rq[p], MPI_STATUS_IGNORE);
      if (recvc[p]) MPI_Wait(&sq[p], MPI_STATUS_IGNORE);
    }
    Check(pcfd_halo_configure(ctx_, rank, np, sendc.data(), list.data(), recvc.data()), "pcfd_halo_configure");
    if (np == 1) return;
    const size_t bs = pcfd_comm_blob_size();
    std::vector<char> mine(bs), all(bs * (size_t)np);
    Check(pcfd_comm_export(ctx_, mine.data()), "pcfd_comm_export");
    MPI_Allgather(mine.data(), (int)bs, MPI_BYTE, all.data(), (int)bs, MPI_BYTE, MPI_COMM_WORLD);
    Check(pcfd_comm_connect(ctx_, all.data()), "pcfd_comm_connect");
    MPI_Barrier(MPI_COMM_WORLD);       // nobody posts before everybody is connected
  }
  // CRSMatrix::CRSTranspose (crsmatrix.tcc:568-599) of the device-resident Jacobian, for Compute_dRdQ_Transpose
  // (jacobian.tcc:121-127).  The local part is three kernels; across ranks the blocks of the ghost columns are replaced by
  // the owner's block of the mirrored cut edge as PObj::TransposeCommCRS does (parallel.tcc:54-338), with requests in LOCAL
  // ids (RouteTransposedGhostBlocks above).  Once per adjoint solve, so the blocks go through the host.
  void CRSTranspose() {
    Check(pcfd_crs_transpose(ctx_), "CRSMatrix::CRSTranspose");
    int np = 1;
    MPI_Comm_size(MPI_COMM_WORLD, &np);
    if (np == 1 || ngedge_ == 0) return;
    const int n2 = neqn_ * neqn_;
    std::vector<double> blocks((size_t)ngedge_ * n2);
    std::vector<int> ge(2 * (size_t)ngedge_);
    for (int e = 0; e < ngedge_; e++) {
      ge[2 * e] = s_->m->bedges[nbedge_ + e].n[0];
      ge[2 * e + 1] = s_->m->bedges[nbedge_ + e].n[1];
    }
    Check(pcfd_crs_ghost_blocks(ctx_, 0, blocks.data()), "pcfd_crs_ghost_blocks (get)");
    if (!RouteTransposedGhostBlocks(nnode_, ngedge_, ge.data(), s_->m->gNodeOwner, s_->m->gNodeLocalId, n2, blocks.data()))
      onError_("CRSTranspose", "a requested cut edge has no mirror on its owner");
    Check(pcfd_crs_ghost_blocks(ctx_, 1, blocks.data()), "pcfd_crs_ghost_blocks (set)");
  }
  // PObj::UpdateGeneralVectors(v, n) for a device-resident field (PCFD_F_Q, _QGRAD, _LIMITER, _X, _LSQ_S, _LSQ_SW, ...)
  void UpdateGeneralVectors(int field) { Check(pcfd_comm_update(ctx_, field), "UpdateGeneralVectors"); }
  // the two halves, for ghost-independent work in between
  void PostHalo(int field) { Check(pcfd_comm_post(ctx_, field), "PostHalo"); }
  void WaitHalo(int field) { Check(pcfd_comm_wait(ctx_, field), "WaitHalo"); }
  // ParallelL2Norm / StridedParallelL2Norm (parallel.h:160-219) of the residual across ranks: sums of squares from
  // ComputeResidualSums all-reduced with the host's MPI
  std::vector<double> ComputeResidualsParallel() {
    double ss[1 + PCFD_CHEM_MAX_SPECIES + 4];
    Check(pcfd_residual(ctx_, ss), "ComputeResiduals");
    int nn = nnode_;
    MPI_Allreduce(MPI_IN_PLACE, ss, 1 + neqn_, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    MPI_Allreduce(MPI_IN_PLACE, &nn, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    std::vector<double> res(1 + neqn_);
    res[0] = std::sqrt(ss[0]) / ((double)nn * neqn_);
    for (int k = 0; k < neqn_; k++) res[1 + k] = std::sqrt(ss[1 + k]) / (double)nn;
    return res;
  }
#endif
  // Param::dt, useLocalTimeStepping, torder and SolutionSpace::iter change from step to step: hand them over before
  // ComputeResiduals / ComputeJacobians of an unsteady run, together with q^n and q^{n-1} (solutionSpace.tcc:582-592)
  void PushTimeIntegration(bool with_history) {
    Check(pcfd_set_time_integration(ctx_, s_->param->dt, s_->param->useLocalTimeStepping ? 1 : 0, s_->param->torder, s_->iter),
          "pcfd_set_time_integration");
    if (with_history) {
      Check(pcfd_set_field(ctx_, PCFD_F_QOLD, s_->qold, (size_t)nnode_ * nvars_), "push qold");
      Check(pcfd_set_field(ctx_, PCFD_F_QOLDM1, s_->qoldm1, (size_t)nnode_ * nvars_), "push qoldm1");
    }
  }

 private:
  // pcfd_fr_params from the reference's own objects: ChemModel / Species / Reaction (chem.h, species.h:35-49,
  // reaction.h:51-95) as CompressibleFREqnSet holds them, Param's reference values (param.tcc:352-398)
  void FillFrParams(pcfd_fr_params& fp) {
    std::memset(&fp, 0, sizeof(fp));
    CompressibleFREqnSet<double>* fr = dynamic_cast<CompressibleFREqnSet<double>*>(s_->eqnset);
    if (!fr) { onError_("DropIn", "eqnset_id names a reacting eqnset but eqnset is not a CompressibleFREqnSet"); return; }
    ChemModel<double>& chem = *fr->chem;
    const int ns = chem.nspecies, nr = chem.nreactions;
    if (ns > PCFD_CHEM_MAX_SPECIES || nr > PCFD_CHEM_MAX_REACTIONS) {
      onError_("DropIn", "chemistry model exceeds PCFD_CHEM_MAX_SPECIES / PCFD_CHEM_MAX_REACTIONS");
      return;
    }
    pcfd_chem_model& cm = fp.chem;
    cm.nspecies = ns;
    cm.nreactions = nr;
    for (int i = 0; i < ns; i++) {
      Species<double>& sp = chem.species[i];
      cm.mw[i] = sp.MW;
      for (int k = 0; k < 7; k++) { cm.nasa7[i][0][k] = sp.thermo_coeff[0][k]; cm.nasa7[i][1][k] = sp.thermo_coeff[1][k]; }
      fp.transport.nmu[i] = sp.mu_coeff_curves;
      fp.transport.nk[i] = sp.k_coeff_curves;
      for (int r = 0; r < sp.mu_coeff_curves && r < 3; r++) for (int k = 0; k < 6; k++) fp.transport.mu_fit[i][r][k] = sp.mu_coeff[r][k];
      for (int r = 0; r < sp.k_coeff_curves && r < 3; r++) for (int k = 0; k < 6; k++) fp.transport.k_fit[i][r][k] = sp.k_coeff[r][k];
      for (int k = 0; k < 3; k++) { fp.transport.mu_white[i][k] = sp.mu_coeff_White[k]; fp.transport.k_white[i][k] = sp.k_coeff_White[k]; }
      fp.transport.mu_white[i][3] = sp.mu_transition_White;
      fp.transport.k_white[i][3] = sp.k_transition_White;
    }
    for (int j = 0; j < nr; j++) {
      Reaction<double>& r = chem.reactions[j];
      cm.rxn_type[j] = r.rxnType;
      cm.third_body[j] = r.thirdBodiesPresent ? 1 : 0;
      cm.backward_given[j] = r.backwardRateGiven ? 1 : 0;
      cm.rxn_type_b[j] = r.backwardRateGiven ? r.rxnTypeBackward : 0;
      cm.nsp[j] = r.GetNspecies();
      cm.A[j] = r.A; cm.EA[j] = r.EA; cm.n[j] = r.n;
      if (r.backwardRateGiven) { cm.Ab[j] = r.Ab; cm.EAb[j] = r.EAb; cm.nb[j] = r.nb; }
      for (int k = 0; k < r.GetNspecies(); k++) {
        cm.species[j][k] = r.globalIndx[k];
        cm.nup[j][k] = r.Nup[k];
        cm.nupp[j][k] = r.Nupp[k];
        cm.tbeff[j][k] = (r.thirdBodiesPresent && (size_t)k < r.TBEff.size()) ? r.TBEff[k] : 1.0;
      }
    }
    fp.ref_density = s_->param->ref_density;
    fp.ref_velocity = s_->param->ref_velocity;
    fp.ref_temperature = s_->param->ref_temperature;
    fp.ref_pressure = s_->param->ref_pressure;
    fp.ref_time = s_->param->ref_time;
    fp.ref_specific_enthalpy = s_->param->ref_specific_enthalpy;
    fp.pref = fr->Pref;
    fp.dt = s_->param->dt;
    fp.use_local_dt = s_->param->useLocalTimeStepping ? 1 : 0;
    fp.rxn_on = s_->param->rxnOn ? 1 : 0;
    for (int k = 0; k < nvars_; k++) fp.qinf[k] = s_->eqnset->Qinf[k];
    fp.ref_viscosity = s_->param->ref_viscosity;
    fp.ref_k = s_->param->ref_k;
  }
  size_t NQ() const { return (size_t)(nnode_ + gnode_ + nbnode_) * nvars_; }
  void Check(int rc, const char* where) {
    if (rc != 0) onError_(where, pcfd_last_error(ctx_));
  }
  DropIn(const DropIn&);
  DropIn& operator=(const DropIn&);

  Space* s_;
  pcfd_ctx* ctx_;
  ErrorHandler onError_;
  bool forces_ready_ = false;
  int nnode_, gnode_, nbnode_, nedge_, nbedge_, ngedge_, neqn_, nvars_, nterms_;
};
#endif  // PCFD_HOST_ROUTING_ONLY

}  // namespace pcfd

#endif
