/*
 * pcfd.h -- C ABI of libpcfd_b200.so: the B200 (sm_100a) implementation of
 * ProteusCFD's edge-based finite-volume hot path.
 *
 * The reference has no FFI: its "plugin surface" is in-process C++ (SURVEY.md
 * 8b).  Each entry point below replaces the body of one reference phase
 * function; the citation gives the reference call site (paths relative to
 * /root/reference/ucs).  INTEGRATION.md shows the C++ shim a ucs.x maintainer
 * adds at those call sites.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; the message is
 *    available from pcfd_last_error().  Nothing throws across this boundary.
 *  - the caller owns all host buffers; the library owns all device memory.
 *  - host arrays use exactly the reference's layouts (AoS, row-major):
 *      q     [(nnode+gnode+nbnode) * nvars]   (solutionSpace.h:94-113)
 *      qgrad [(nnode+gnode) * nterms*3]
 *      lim   [(nnode+gnode) * neqn]           (limiters.h: Limiter::l)
 *      b     [nnode*neqn], x [(nnode+gnode)*neqn]   (crs.h: CRS::b, CRS::x)
 *      A     [nblocks*neqn*neqn], diagonal block first in each row (crsmatrix.tcc:48-97)
 *  - one context per GPU; calls on one context are serialised by the caller.
 *  - there is NO CPU fallback: pcfd_create fails if no sm_100-class device is
 *    usable.
 */
#ifndef PCFD_H
#define PCFD_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCFD_ABI_VERSION 9

/* eqnset ids follow eqnset_defines.h / create_functions.h:17-45 */
enum { PCFD_EQNSET_COMPRESSIBLE_EULER_FR = 0, PCFD_EQNSET_COMPRESSIBLE_NS_FR = 1, PCFD_EQNSET_COMPRESSIBLE_EULER = 2,
       PCFD_EQNSET_COMPRESSIBLE_NS = 3 };

/* BC types: bc_defines.h:4-30 (the value bc->GetBCType(factag) returns) */
enum { PCFD_BC_PARALLEL = 0, PCFD_BC_DIRICHLET = 1, PCFD_BC_NEUMANN = 2, PCFD_BC_IMPERMEABLE_WALL = 3,
       PCFD_BC_NOSLIP = 4, PCFD_BC_FARFIELD_VISCOUS = 5, PCFD_BC_FARFIELD = 6, PCFD_BC_SONIC_INFLOW = 7,
       PCFD_BC_SONIC_OUTFLOW = 8, PCFD_BC_SYMMETRY = 9 };

/* fields that can be moved across the boundary with pcfd_set_field/pcfd_get_field */
enum {
  PCFD_F_Q = 0,        /* SolutionSpace::q        (nnode+gnode+nbnode)*nvars */
  PCFD_F_QGRAD = 1,    /* SolutionSpace::qgrad    (nnode+gnode)*nterms*3 */
  PCFD_F_LIMITER = 2,  /* Limiter::l              (nnode+gnode)*neqn */
  PCFD_F_B = 3,        /* CRS::b                  nnode*neqn */
  PCFD_F_X = 4,        /* CRS::x                  (nnode+gnode)*neqn */
  PCFD_F_TIMESTEP = 5, /* field "timestep"        nnode */
  PCFD_F_BETA = 6,     /* field "beta"            nnode+gnode+nbnode (unused by the perfect-gas eqnset) */
  PCFD_F_LSQ_S = 7,    /* Mesh::s                 (nnode+gnode)*6 */
  PCFD_F_LSQ_SW = 8,   /* Mesh::sw                (nnode+gnode)*6 */
  PCFD_F_A = 9,        /* CRSMatrix::M            nblocks*neqn*neqn */
  PCFD_F_MUT = 10,     /* field "mut" (eddy viscosity) nnode+gnode+nbnode; zero for laminar flow */
  /* Spalart-Allmaras model (turb.h:24-30): TurbulenceModel::tvar, ::tgrad, field "wallDistance", and the scalar
     system TurbulenceModel::crs (b, x, A->M with one double per block on the flow's CRS pattern) */
  PCFD_F_TVAR = 11,      /* nnode+gnode+nbnode */
  PCFD_F_TGRAD = 12,     /* (nnode+gnode)*3 */
  PCFD_F_WALLDIST = 13,  /* nnode+gnode */
  PCFD_F_TURB_B = 14,    /* nnode */
  PCFD_F_TURB_X = 15,    /* nnode+gnode */
  PCFD_F_TURB_A = 16,    /* nblocks */
  /* time integration (TemporalResidual, residual.tcc:125-179): CONSERVATIVE variables at t^n and t^{n-1}, rows of
     nvars doubles for the owned nodes, as SolutionSpace::qold / qoldm1 hold them (solutionSpace.tcc:582-592).
     Allocated on first use (unsteady runs only). */
  PCFD_F_QOLD = 17,      /* nnode*nvars */
  PCFD_F_QOLDM1 = 18,    /* nnode*nvars */
  PCFD_F_COUNT = 19
};

/* Mesh::edges / bedges / xyz / vol / ipsp / psp as flat arrays (uns_base.h:12-37, mesh.h:199-254) */
typedef struct {
  int nnode, gnode, nbnode;
  int nedge, nbedge, ngedge;
  const int* edges_n;        /* [2*nedge]   Edges::n           */
  const double* edges_a;     /* [4*nedge]   Edges::a (unit normal, area) */
  const int* bedges_n;       /* [2*(nbedge+ngedge)]  HalfEdges::n */
  const double* bedges_a;    /* [4*(nbedge+ngedge)]  HalfEdges::a */
  const int* bedges_bctype;  /* [nbedge+ngedge]  bc->GetBCType(HalfEdges::factag) */
  const double* xyz;         /* [3*(nnode+gnode)] */
  const double* vol;         /* [nnode] */
  const int* ipsp;           /* [nnode+1] */
  const int* psp;            /* [ipsp[nnode]]  order preserved: it fixes the CRS column order */
  const double* bedges_twall;/* [nbedge] bcobj->twall / Param::ref_temperature of the half-edge's surface (bc.tcc:1283-1286),
                                read for NoSlip half-edges only; < 0: adiabatic wall.  NULL: 1/tref (bcobj.tcc:31) */
} pcfd_mesh_desc;

/* the Param<Type> fields the hot path reads (param.tcc:84-229) */
typedef struct {
  int eqnset;                /* PCFD_EQNSET_* */
  int sorder;                /* spatialOrder 1|2 */
  int limiter;               /* 0 none, 1 Barth, 2 Venkatakrishnan, 3 modified Venkatakrishnan (limiters.tcc:534-735) */
  int no_cvbc;
  double gamma, chi, cfl;
  double qinf[10];           /* free-stream state incl. aux vars (bc.tcc: Qinf) */
  /* viscous terms, read when eqnset == PCFD_EQNSET_COMPRESSIBLE_NS (param.tcc:401-404 sets Param::viscous) */
  int enable_vnn;            /* Param::enableVNN: Von Neumann time-step limit (timestep.tcc:37-41) */
  double vnn;                /* Param::VNN */
  double Re, Pr, PrT;        /* Param::Re (compressible.tcc:85-89), prandtlNumber, turbulent Prandtl number */
  double tref;               /* Param::ref_temperature [K] (Sutherland constant 110.4/tref, eqnset.h:259) */
  double mach;               /* Param::GetVelocity(iter): Re is rescaled by it (compressible.tcc:753-755) */
  int turb_model;            /* Param::turbModel: 0 laminar, 1 Spalart-Allmaras (turb.h:91-107) */
} pcfd_params;

typedef struct pcfd_ctx pcfd_ctx;

int pcfd_abi_version(void);
/* message of the last failed call on ctx (ctx may be NULL for pcfd_create failures) */
const char* pcfd_last_error(const pcfd_ctx* ctx);

/* SolutionSpace ctor + Init (solutionSpace.tcc:26-114, 148-376): uploads the mesh,
   builds the node->edge gather lists, the block-CRS pattern (CRSMatrix::Init,
   crsmatrix.tcc:48-97) and the SGS level schedule. */
int pcfd_create(const pcfd_mesh_desc* mesh, const pcfd_params* params, int device, pcfd_ctx** out);
int pcfd_destroy(pcfd_ctx* ctx);
/* run all subsequent work of ctx on an existing cudaStream_t (NULL = the context's own stream) */
int pcfd_set_stream(pcfd_ctx* ctx, void* cuda_stream);
int pcfd_synchronize(pcfd_ctx* ctx);
int pcfd_set_cfl(pcfd_ctx* ctx, double cfl);           /* Param::UpdateCFL (solutionSpace.tcc:629) */
/* Time integration state read by ComputeResiduals / ComputeJacobians: Param::dt (< 0: steady), Param::useLocalTimeStepping,
   Param::torder (1 | 2) and SolutionSpace::iter (BDF2 coefficients from the second step on, residual.tcc:139-149,
   jacobian.tcc:228-230).  Default: steady, first order, iter = 1.  TemporalResidual is evaluated once field
   PCFD_F_QOLD has been set (before that q^n == q^{n+1} and its contribution is an exact zero); the diagonal terms
   cnp1 V/dt + V/dtau (eqnset.tcc:195-208, compressibleFR.tcc:1326-1331) follow these values at once. */
int pcfd_set_time_integration(pcfd_ctx* ctx, double dt, int use_local_time_stepping, int torder, int iter);
/* Param::gradType (param.tcc, read at gradient.tcc:68-90): 0 weighted least squares (default), 1 Green-Gauss
   (Kernel_Green_Gauss_Gradient / Bkernel_Green_Gauss_Gradient gradient.tcc:170-248, divided by the dual volume :83-89).
   pcfd_gradient and the composite iterations follow it at once.  (ABI v6) */
int pcfd_set_gradient_type(pcfd_ctx* ctx, int type);
/* Param::fieldJacType / boundaryJacType (jacobian.tcc:140-176): 0 one-sided finite differences (default,
   Kernel_NumJac :254-304 / Bkernel_NumJac :459-544), 1 central differences (Kernel_NumJac_Centered :306-366 /
   Bkernel_NumJac_Centered :546-640, whose -h branch does not re-evaluate the BC); field type 2: complex step
   (Kernel_NumJac_Complex :370-433: the Roe flux on a complex state perturbed by i*1e-11, perfect-gas eqnsets only, ABI
   v9).  Boundary type 2 (Bkernel_NumJac_Complex) and field type 2 on a reacting context are rejected.  (ABI v6) */
int pcfd_set_jacobian_type(pcfd_ctx* ctx, int field_type, int boundary_type);

size_t pcfd_field_size(const pcfd_ctx* ctx, int field); /* number of doubles */
int pcfd_set_field(pcfd_ctx* ctx, int field, const double* host, size_t n);
int pcfd_get_field(pcfd_ctx* ctx, int field, double* host, size_t n);
void* pcfd_field_device_ptr(pcfd_ctx* ctx, int field);  /* device-resident access (same layout) */
/* CRSMatrix::{ia,ja,iau,pv}; any pointer may be NULL */
int pcfd_crs_sizes(const pcfd_ctx* ctx, int* nrows, int* nblocks);
int pcfd_get_crs(pcfd_ctx* ctx, int* ia, int* ja, int* iau, int* pv);

/* Gradient::ComputeNodeLSQCoefficients (gradient.tcc:115-138) -> Mesh::s, Mesh::sw */
int pcfd_lsq_coefficients(pcfd_ctx* ctx);
/* UpdateBCs (bc.tcc:1399-1457): phantom-node states of q */
int pcfd_update_bcs(pcfd_ctx* ctx);
/* Gradient::Compute (gradient.tcc:57-112), weighted LSQ */
int pcfd_gradient(pcfd_ctx* ctx);
/* Limiter::Compute (limiters.tcc:53-132) */
int pcfd_limiter(pcfd_ctx* ctx);
/* ComputeResiduals -> SpatialResidual (residual.tcc:13-122); resnorm (may be NULL) receives
   [ParallelL2Norm(b), StridedParallelL2Norm(b, eq) for eq < neqn] of this rank's nodes
   as sum-of-squares (the caller finishes sqrt(sum)/N across ranks, parallel.h:160-219) */
int pcfd_residual(pcfd_ctx* ctx, double* sumsq);
/* ComputeTimesteps (timestep.tcc:7-77), local time stepping; dtmin may be NULL */
int pcfd_timestep(pcfd_ctx* ctx, double* dtmin);
/* ExplicitSolve (solve.tcc:71-140): x = b*dt/vol, q <- ApplyDQ(x) */
int pcfd_explicit_solve(pcfd_ctx* ctx);
/* ComputeJacobians (jacobian.tcc:13-18, 130-250), one-sided FD field + boundary Jacobians */
int pcfd_jacobian(pcfd_ctx* ctx);
/* CRSMatrix::PrepareSGS (crsmatrix.tcc:840-876) */
int pcfd_prepare_sgs(pcfd_ctx* ctx);
/* CRS::BlankX (crs.tcc:448-460) */
int pcfd_blank_x(pcfd_ctx* ctx);
/* CRS::SGS (crs.tcc:62-173); ddq (may be NULL) receives |xOld - xNorm| */
int pcfd_sgs(pcfd_ctx* ctx, int nsgs, double* ddq);
/* CRS::GMRES (crs.tcc:176-415): restarted GMRES (restarts x nsearch directions) with right preconditioning on the
   block-CRS system of the context: A = field PCFD_F_A as ASSEMBLED (before pcfd_prepare_sgs factors its diagonal in
   place; an error otherwise), b = PCFD_F_B, x = PCFD_F_X (initial guess in, solution out).  precond_type as
   CRS::Preconditioner (:555-590): 0 none, 1 diagonal, 2 block diagonal (LU), 3 local ILU0 (BuildILU0Local /
   ILU0BackSub, crsmatrix.tcc:276-507: pivot-free, ghost columns left out, factored level by level on a copy), 4 SGS
   (six sweeps on a copy of the matrix per application); the copy of 3 and 4 costs another matrix of device memory.
   dq_norm (may be NULL) receives the reference's return value |g[idir]|.  On connected contexts the vector that
   needs a halo before every product travels through pcfd_comm and the dot products are summed across ranks.  The
   reference's flow solver keeps this solver behind a comment (solutionSpace.tcc:734-750); move.tcc:714 calls it. (ABI v7) */
int pcfd_gmres(pcfd_ctx* ctx, int restarts, int nsearch, int precond_type, double* dq_norm);

/* CRSMatrix::CRSTranspose (crsmatrix.tcc:568-599) on PCFD_F_A as ASSEMBLED (an error after pcfd_prepare_sgs): the
   adjoint path's transposed Jacobian (Compute_dRdQ_Transpose, jacobian.tcc:121-127).  Every block is transposed and the
   mirror blocks (i, j) <-> (j, i) of local node pairs change places; blocks of ghost columns are transposed where they
   are.  On a partition they must then be replaced by the owning rank's block of the mirrored cut edge
   (PObj::TransposeCommCRS, parallel.tcc:54-338): pcfd_crs_ghost_blocks moves them between the device and a host buffer
   of ngedge * neqn^2 doubles (set = 0: device -> host, 1: host -> device), block e being A(bedges_n[2(nbedge+e)],
   bedges_n[2(nbedge+e)+1]) -- the order of the parallel half-edges of pcfd_mesh_desc -- and the host routes them with the
   transport it owns (MPI in ucs.x; proteuscfd_b200/parallel.py: crs_transpose).  (ABI v9) */
int pcfd_crs_transpose(pcfd_ctx* ctx);
int pcfd_crs_ghost_blocks(pcfd_ctx* ctx, int set, double* host);

/* Forces (ucs/forces.tcc; called once per iteration at solutionSpace.tcc:884 and :924).  (ABI v8)
   pcfd_forces_configure takes the composite bodies the .bc file declares ("body #k = [factags]", composite.tcc:78-170;
   0-based here, at most 32), Param::liftdir / dragdir / velocity, the factag of every BC half-edge and Mesh::cg, and
   evaluates ComputeSurfaceAreas (forces.tcc:199-312) for this rank's half-edges (pcfd_forces_areas reads the result:
   surf_area [(num_bcs+1)*3] per factag, body_area [nbodies*3]; a multi-rank host sums them as the reference does).
   pcfd_forces_compute = Forces::Compute on the device-resident q and qgrad: FORCE_Kernel (:123-196), YpCf_Kernel
   (:400-478), ComputeCl (:326-369).  body (may be NULL) [nbodies*12] = forces, vforces, moments, vmoments of each body;
   coef (may be NULL) [nbodies*3] = cl, cd, cm.  On connected contexts (pcfd_comm_connect) the sums and the body areas
   are added over the ranks in rank order.  cp / y+ / cf per BC half-edge [nbedge]: pcfd_forces_get.  The per-half-edge
   terms follow the reference's arithmetic; the body sums are tree sums (same on every run; they differ from the
   reference's sequential += by summation order only). */
typedef struct {
  int nbodies, num_bcs;            /* num_bcs = BoundaryConditions::largest_bc_id */
  const int* body_offsets;         /* [nbodies+1] into body_factags */
  const int* body_factags;
  const double* moment_pt;         /* [3*nbodies] CompositeBody::momentPt */
  const double* moment_axis;       /* [3*nbodies] CompositeBody::momentAxis */
  const int* bedges_factag;        /* [nbedge] */
  const double* cg;                /* Mesh::cg [(nnode+gnode+nbnode)*3] */
  double liftdir[3], dragdir[3];   /* as given in the .param file (not normalised) */
  double velocity;                 /* Param::GetVelocity(iter) */
} pcfd_forces_desc;
enum { PCFD_SURF_CP = 0, PCFD_SURF_YPLUS = 1, PCFD_SURF_CF = 2 };
int pcfd_forces_configure(pcfd_ctx* ctx, const pcfd_forces_desc* desc);
int pcfd_forces_areas(pcfd_ctx* ctx, double* surf_area, double* body_area);
int pcfd_forces_compute(pcfd_ctx* ctx, double* body, double* coef);
int pcfd_forces_get(pcfd_ctx* ctx, int which, double* out);

/* ComputeWallDistOct (ucs/walldist.tcc:116-199): fills field PCFD_F_WALLDIST [nnode+gnode] with the distance of every
   local node to the nearest of `points` [npoints*3] -- the viscous wall nodes of ALL ranks (the left nodes of the no-slip
   half-edges; the reference gathers them with MPI, walldist.tcc:24-113, and so does the host here).  Exact search with the
   reference's `Distance` arithmetic: equals the reference's field wherever its octree returns the true nearest node.
   npoints == 0: +inf everywhere.  Needs a context with that field (turbulence model or viscous far-field BC).  (ABI v8) */
int pcfd_wall_distance(pcfd_ctx* ctx, const double* points, int npoints);

/* Limiter::Compute + ComputeResiduals in two halves for multi-rank hosts (all eqnsets).  pcfd_limiter_raw runs
   passes 1+2 of Limiter::Compute (limiters.tcc:53-110) and leaves the UNCLAMPED limiter in field PCFD_F_LIMITER; the
   host exchanges that field (limiters.tcc:128); pcfd_residual_fused then clamps it (:118-125), evaluates the residual
   and performs the test of Kernel_PressureClip (:737-815) on the way.  *clip_hit == 0: limiter and residual are final
   (the pressure clip would not have changed anything).  *clip_hit != 0 (rare; on ANY rank): discard, then every rank
   calls pcfd_limiter, exchanges the limiter and calls pcfd_residual. */
int pcfd_limiter_raw(pcfd_ctx* ctx);
/* how many times the fused limiter / residual pair of this context had to fall back to the ordered clip path */
long long pcfd_clip_fallbacks(const pcfd_ctx* ctx);
/* nodes whose update pcfd_apply_dq zeroed since creation because a component was NaN / Inf -- what NewtonIterate does
   before ApplyDQ (solutionSpace.tcc:771-796; the reference prints the count); waits for the stream; -1 on error. (ABI v8) */
long long pcfd_zeroed_updates(pcfd_ctx* ctx);
int pcfd_residual_fused(pcfd_ctx* ctx, double* sumsq, int* clip_hit);

/* loop over nnode of EqnSet::ApplyDQ (solutionSpace.tcc:802-804) */
int pcfd_apply_dq(pcfd_ctx* ctx);

/* One explicit iteration as SolutionSpace::NewtonIterate runs it with nSgs == 0
   (solutionSpace.tcc:640-904): UpdateBCs, Gradient, Limiter, ComputeResiduals,
   ExplicitSolve.  ComputeTimesteps is run first when refresh_dt != 0. */
int pcfd_explicit_iterate(pcfd_ctx* ctx, int refresh_dt, double* sumsq);
/* One implicit iteration (nSgs > 0): [ComputeJacobians + ComputeTimesteps when
   refresh_jac != 0], UpdateBCs, Gradient, Limiter, ComputeResiduals, PrepareSGS,
   BlankX, SGS(nsgs), ApplyDQ, and -- with a turbulence model -- TurbulenceModel::Compute
   (solutionSpace.tcc:862-866). */
int pcfd_implicit_iterate(pcfd_ctx* ctx, int refresh_jac, int nsgs, double* sumsq, double* ddq);

/* TurbulenceModel::Compute (turb.tcc:163-339) for the Spalart-Allmaras model (spalart.tcc), turbulenceSpatialOrder = 1:
   BCs, unweighted LSQ gradient of nu~, convective + diffusive + source terms with their Jacobians, nsgs scalar SGS
   sweeps (nsgs == 0: explicit update), nu~ update with the clip at zero, eddy viscosity into field "mut".
   Reads q, qgrad and timestep as the flow iteration left them.  sumsq (may be NULL) receives sum(b^2). */
int pcfd_turb_compute(pcfd_ctx* ctx, int nsgs, double* sumsq);
/* The same computation cut at the reference's exchange points (turb.tcc:183-325), for runs on partitions: the caller
   exchanges the named field after each phase (pcfd_halo_pack / pcfd_halo_recv_ptr know them).
     0 blank + turbulence BCs -> halo PCFD_F_TVAR (:185);  1 gradient of tvar -> halo PCFD_F_TGRAD (gradient.tcc:98);
     2 assembly, wall rows, sumsq = sum(b^2) of this rank's nodes (may be NULL), inverse diagonal;
     3 ONE symmetric sweep -> halo PCFD_F_TURB_X (crs.tcc:146), repeated nSgs times;
     4 tvar += x -> halo PCFD_F_TVAR (:325);  5 eddy viscosity (local + ghost nodes).
   Phases 0..5 in order with nSgs repeats of phase 3 and no exchange are pcfd_turb_compute(ctx, nSgs, sumsq).  (ABI v6) */
int pcfd_turb_phase(pcfd_ctx* ctx, int phase, double* sumsq);

/* ---- halo exchange: PObj (parallel.tcc).  The send lists are persistent (the reference re-sends the index
   lists on every call, parallel.tcc:809-827).  send_list = PObj::nodePackingList concatenated in peer order
   (local ids peer p wants, in p's ghost order); recv_counts = PObj::commCountsRecv; the rows received from
   peer p land at ghost rows [commOffsetsRecv[p], +recv_counts[p]) (parallel.tcc:848-864). */
int pcfd_halo_configure(pcfd_ctx* ctx, int rank, int nranks, const int* send_counts, const int* send_list,
                        const int* recv_counts);
int pcfd_halo_width(const pcfd_ctx* ctx, int field);      /* doubles per node of a field (0: not exchangeable) */
int pcfd_halo_send_total(const pcfd_ctx* ctx);            /* rows this rank sends per exchange, all peers */
/* gather the rows owed to `peer` (peer < 0: all peers in rank order) into device memory dst; dst may be a
   local staging buffer or a peer GPU's ghost segment mapped with CUDA IPC (direct put over NVLink) */
int pcfd_halo_pack(pcfd_ctx* ctx, int field, int peer, void* dst);
/* device address of the ghost rows filled by `peer` (peer < 0: start of the ghost segment) */
void* pcfd_halo_recv_ptr(pcfd_ctx* ctx, int field, int peer);

/* CUDA-IPC mapping of a field for the direct-put exchange between one-process-per-GPU ranks: export a 64-byte
   handle here, open it in the peer process, and pass (opened base + ghost offset) as dst of pcfd_halo_pack. */
int pcfd_ipc_export(pcfd_ctx* ctx, int field, void* handle64);
int pcfd_ipc_open(pcfd_ctx* ctx, const void* handle64, void** devptr);
int pcfd_ipc_close(pcfd_ctx* ctx, void* devptr);

/* ---- collective-free exchange between one-process-per-GPU ranks (ABI v7): the device-resident replacement of
   PObj::UpdateGeneralVectors (parallel.tcc:779-873).  After pcfd_halo_configure every rank exports one fixed-size blob
   (CUDA-IPC handles of its exchangeable fields and of a flag page, its node count and receive offsets); the host
   all-gathers the blobs with its own transport (MPI_Allgather in ucs.x) and hands the table, in rank order, to
   pcfd_comm_connect.  From then on
     pcfd_comm_post(field)   ONE kernel: neighbour handshake on epoch flags, gather of nodePackingList rows, stores
                             straight into the neighbours' ghost segments over NVLink, done flags;
     pcfd_comm_wait(field)   a one-block kernel that waits for the neighbours' done flags of that post;
     pcfd_comm_update(field) = post + wait (what UpdateGeneralVectors does).
   Nothing in them synchronises the host or calls a collective; ghost-independent work queued between post and wait
   overlaps the transfer.  While connected, pcfd_lsq_coefficients, pcfd_explicit_iterate, pcfd_implicit_iterate and
   pcfd_turb_compute run the reference's multi-rank sequence themselves (exchanges at solutionSpace.tcc:665, 857,
   gradient.tcc:98,131-134, limiters.tcc:128, crs.tcc:88,146, turb.tcc:185,325), the pressure-clip decision being taken
   globally through pcfd_comm_allgather's flag words.  Blobs published by contexts of the SAME process are wired with
   plain pointers (several ranks as threads of one process: the tests' single-GPU harness). */
#define PCFD_COMM_MAX_RANKS 64
size_t pcfd_comm_blob_size(void);
int pcfd_comm_export(pcfd_ctx* ctx, void* blob);
int pcfd_comm_connect(pcfd_ctx* ctx, const void* blobs /* nranks * pcfd_comm_blob_size() bytes, rank order */);
int pcfd_comm_disconnect(pcfd_ctx* ctx);
int pcfd_comm_connected(const pcfd_ctx* ctx);
/* diagnostics: out[0..nranks) = ready epochs, out[nranks..2 nranks) = done epochs seen from each rank, out[2 nranks] = 1
   if a flag wait of this rank ran into its time limit (20 s; PCFD_COMM_SPIN_SECONDS overrides at connect time) */
int pcfd_comm_debug_flags(pcfd_ctx* ctx, unsigned long long* out);
int pcfd_comm_post(pcfd_ctx* ctx, int field);
int pcfd_comm_wait(pcfd_ctx* ctx, int field);
int pcfd_comm_update(pcfd_ctx* ctx, int field);
/* the reference's small reductions (ParallelL2Norm parallel.h:160-219, MPI_MIN of dtmin solutionSpace.tcc:712-714):
   every rank contributes n <= 8 doubles, out[r*n + k] = value k of rank r (host buffers; waits for the stream) */
int pcfd_comm_allgather(pcfd_ctx* ctx, const double* vals, int n, double* out);

/* number of CUDA kernels this context has launched since creation */
long long pcfd_launch_count(const pcfd_ctx* ctx);

/* Per-kernel device timing (CUDA events on the launch stream around every kernel), the
   GPU-side counterpart of the reference's TimerList (timer.h:29-89; solutionSpace.tcc:34-43).
   enable(1) starts recording, count() drains pending events and returns the number of
   distinct kernels seen, get(i) returns the i-th kernel's name, summed milliseconds and launches. */
int pcfd_profile_enable(pcfd_ctx* ctx, int on);
int pcfd_profile_reset(pcfd_ctx* ctx);
int pcfd_profile_count(pcfd_ctx* ctx);
int pcfd_profile_get(pcfd_ctx* ctx, int i, const char** name, double* total_ms, long long* launches);

/* ------------------------------------------------------------------------------------------------------------
 * Finite-rate chemistry source term of the reacting eqnset (compressibleFR): ChemModel / Reaction / Species
 * (chem.h:32-120, reaction.h:40-100, species.h:20-60).  Independent of pcfd_ctx (which carries the 5-equation
 * perfect-gas system): a model handle plus node-parallel entry points.
 */
#define PCFD_CHEM_MAX_SPECIES 16
#define PCFD_CHEM_MAX_REACTIONS 32

/* the tables the reference builds from <case>.rxn and chemdb.hdf5, as flat arrays */
typedef struct {
  int nspecies, nreactions;
  double mw[PCFD_CHEM_MAX_SPECIES];               /* Species::MW after GetDBInfo's /1000 (species.tcc:160-166) */
  double nasa7[PCFD_CHEM_MAX_SPECIES][2][7];      /* Species::thermo_coeff[0] (T <= 1000 K) and [1] (T > 1000 K) */
  int rxn_type[PCFD_CHEM_MAX_REACTIONS];          /* 0 Arrhenius, 1 ModArrhenius, 2 GuptaModArrhenius, 3 Power */
  int third_body[PCFD_CHEM_MAX_REACTIONS];        /* Reaction::thirdBodiesPresent */
  int backward_given[PCFD_CHEM_MAX_REACTIONS];    /* Reaction::backwardRateGiven (+ rxn_type_b, Ab, EAb, nb) */
  int rxn_type_b[PCFD_CHEM_MAX_REACTIONS];
  int nsp[PCFD_CHEM_MAX_REACTIONS];               /* species taking part (third-body-only ones included), local order */
  int species[PCFD_CHEM_MAX_REACTIONS][PCFD_CHEM_MAX_SPECIES];   /* Reaction::globalIndx */
  double A[PCFD_CHEM_MAX_REACTIONS], EA[PCFD_CHEM_MAX_REACTIONS], n[PCFD_CHEM_MAX_REACTIONS];
  double Ab[PCFD_CHEM_MAX_REACTIONS], EAb[PCFD_CHEM_MAX_REACTIONS], nb[PCFD_CHEM_MAX_REACTIONS];
  double nup[PCFD_CHEM_MAX_REACTIONS][PCFD_CHEM_MAX_SPECIES];    /* Reaction::Nup, Nupp, TBEff by local index */
  double nupp[PCFD_CHEM_MAX_REACTIONS][PCFD_CHEM_MAX_SPECIES];
  double tbeff[PCFD_CHEM_MAX_REACTIONS][PCFD_CHEM_MAX_SPECIES];
} pcfd_chem_model;

typedef struct pcfd_chem pcfd_chem;

int pcfd_chem_create(const pcfd_chem_model* model, int device, pcfd_chem** out);
int pcfd_chem_destroy(pcfd_chem* chem);
const char* pcfd_chem_last_error(const pcfd_chem* chem);   /* chem may be NULL for pcfd_chem_create failures */
/* ChemModel::GetMassProductionRates (chem.tcc:575-583) -> Reaction::GetMassProductionRate (reaction.tcc:763-856)
   for n states.  HOST buffers: rhoi [n*nspecies] kg/m^3, T [n] K; wdot [n*nspecies] kg/(m^3 s). */
int pcfd_chem_mass_production(pcfd_chem* chem, int n, const double* rhoi, const double* T, double* wdot);
/* CompressibleFREqnSet::SourceTerm (compressibleFR.tcc:1276-1316; rxnOn, gravity off) for n nodes.  Q rows of
   `stride` doubles in the eqnset's non-dimensional variables [rho_0..rho_{ns-1}, u, v, w, T, ...]; source rows of
   nspecies+4 doubles (species rows = vol*wdot_i/(ref_density/ref_time), momentum and energy rows 0).
   DEVICE pointers, launched on `cuda_stream` (NULL: the default stream) -- the residency entry point. */
int pcfd_chem_source_term_device(pcfd_chem* chem, int n, int stride, const void* d_Q, const void* d_vol,
                                 double ref_density, double ref_time, double ref_temperature, void* d_source,
                                 void* cuda_stream);
/* the same with HOST buffers (copies in and out) */
int pcfd_chem_source_term(pcfd_chem* chem, int n, int stride, const double* Q, const double* vol, double ref_density,
                          double ref_time, double ref_temperature, double* source);

/* ------------------------------------------------------------------------------------------------------------
 * The reacting eqnset on the full hot path: CompressibleFREqnSet (compressibleFR.h/.tcc), equationSet =
 * compressibleEulerFR.  Native variables [rho_1..rho_ns, u, v, w, T]; with ns species
 *    neqn = ns+4, nvars = 3ns+6 [rho_i | u v w | T | P | rho | cv_i | mol_i], nterms = 2ns+4
 * (compressibleFR.tcc:14-31, 43-44, 693-711), so q / qgrad / limiter / b / x / A rows have these widths and A holds
 * (ns+4)x(ns+4) blocks.  HLLC flux with preconditioned wave speeds (:301-548), NASA-7 thermodynamics, characteristic
 * far-field and slip-wall BCs (:940-1134), finite-rate source term (:1276-1316) and its finite-difference Jacobian
 * (eqnset.tcc:163-187), dense temporal terms (:1319-1463), native <-> conservative explicit update (solve.tcc:112-130).
 * Every pcfd_* phase entry point above works on such a context; pcfd_turb_compute does not.
 */
/* Species transport data as Species holds it (species.h:35-49; species.tcc:13-22, 241-310): Sutherland law (White) up
   to the transition temperature, NASA RP-1311 fits above (rows [Tlo, Thi, A, B, C, D], as the reference's chemDB.py
   stores chemdata/trans.inp under /species/<sym>/mu and /species/<sym>/k). */
typedef struct {
  int nmu[PCFD_CHEM_MAX_SPECIES], nk[PCFD_CHEM_MAX_SPECIES];          /* Species::mu_coeff_curves, k_coeff_curves (<= 3) */
  double mu_fit[PCFD_CHEM_MAX_SPECIES][3][6], k_fit[PCFD_CHEM_MAX_SPECIES][3][6];
  double mu_white[PCFD_CHEM_MAX_SPECIES][4], k_white[PCFD_CHEM_MAX_SPECIES][4];   /* value, T0, S, transition T */
} pcfd_transport_model;

typedef struct {
  pcfd_chem_model chem;      /* ChemModel tables (species order = the model's) */
  /* Param reference values (param.tcc:352-398) */
  double ref_density, ref_velocity, ref_temperature, ref_pressure, ref_time, ref_specific_enthalpy;
  double pref;               /* CompressibleFREqnSet::Pref = GetPressure(Qinf) (compressibleFR.tcc:1515) */
  double dt;                 /* Param::dt (negative: steady); enters ContributeTemporalTerms (:1326-1331) */
  int use_local_dt;          /* Param::useLocalTimeStepping */
  int rxn_on;                /* Param::rxnOn */
  double qinf[3 * PCFD_CHEM_MAX_SPECIES + 6];   /* EqnSet::Qinf, all nvars entries */
  /* compressibleNSFR (params->eqnset == PCFD_EQNSET_COMPRESSIBLE_NS_FR, param.tcc:401-404): viscous flux
     (compressibleFR.tcc:551-637) and analytic viscous Jacobian (:1713-2040) with Wilke-mixed species transport
     (chem.tcc:876-938, species.tcc:393-479).  Re and PrT come from pcfd_params; the eddy viscosity from field PCFD_F_MUT. */
  pcfd_transport_model transport;
  double ref_viscosity, ref_k;                  /* Param::ref_viscosity, ref_k (param.tcc:210-211) */
} pcfd_fr_params;

/* params->eqnset must be PCFD_EQNSET_COMPRESSIBLE_EULER_FR or PCFD_EQNSET_COMPRESSIBLE_NS_FR; sorder, limiter, chi, cfl, no_cvbc, enable_vnn / vnn are
   read from params, gamma / qinf / the viscous fields are not.  The preconditioning field "beta"
   (solutionSpace.tcc:235-247) is set with pcfd_set_field(PCFD_F_BETA). */
int pcfd_create_fr(const pcfd_mesh_desc* mesh, const pcfd_params* params, const pcfd_fr_params* fr, int device,
                   pcfd_ctx** out);
/* system widths of a context: 5 / 10 / 9 for the perfect-gas eqnsets */
int pcfd_widths(const pcfd_ctx* ctx, int* neqn, int* nvars, int* nterms);

#ifdef __cplusplus
}
#endif
#endif
