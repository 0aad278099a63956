/*
 * pcfd_oracle.h -- plain-C CPU restatement of ProteusCFD's edge-based
 * finite-volume hot path.  TEST INFRASTRUCTURE ONLY: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it; the
 * product (proteuscfd_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against tests/golden/ npz files, which tools/make_golden.py produced by running
 * the unmodified reference (oracle/_ref/ref_harness) -- including the
 * reference's own unit-test fixture cube_LowFi.0.h5.
 *
 * Array layouts are exactly the reference's (AoS, ucs/solutionSpace.h:94-113):
 *   q     [(nnode+gnode+nbnode) * nvars]   nvars = 10  [rho,ru,rv,rw,rE | T,P,u,v,w]
 *   qgrad [(nnode+gnode) * nterms*3]       nterms = 9  (vars 0,1,2,3,4,5,7,8,9)
 *   lim   [(nnode+gnode) * neqn]           neqn = 5
 *   b     [nnode * neqn],  x [(nnode+gnode) * neqn]
 *   A     block-CRS, diagonal block first in each row (ucs/crsmatrix.tcc:48-97)
 */
#ifndef PCFD_ORACLE_H
#define PCFD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NEQN 5
#define ORC_NVARS 10
#define ORC_NTERMS 9

/* ucs/bc_defines.h:4-30 */
enum { ORC_BC_PARALLEL = 0, ORC_BC_DIRICHLET = 1, ORC_BC_NEUMANN = 2, ORC_BC_IMPERMEABLE_WALL = 3,
       ORC_BC_NOSLIP = 4, ORC_BC_FARFIELD_VISCOUS = 5, ORC_BC_FARFIELD = 6, ORC_BC_SONIC_INFLOW = 7,
       ORC_BC_SONIC_OUTFLOW = 8, ORC_BC_SYMMETRY = 9 };

typedef struct {
  int nnode, gnode, nbnode;
  int nedge, nbedge, ngedge;
  const int* edges_n;        /* [2*nedge]            uns_base.h:12-22 */
  const double* edges_a;     /* [4*nedge]  unit normal + area */
  const int* bedges_n;       /* [2*(nbedge+ngedge)]  uns_base.h:28-37 */
  const double* bedges_a;    /* [4*(nbedge+ngedge)] */
  const int* bedges_bctype;  /* [nbedge+ngedge]      bc->GetBCType(factag) */
  const double* xyz;         /* [3*(nnode+gnode)] */
  const double* vol;         /* [nnode] */
  const int* ipsp;           /* [nnode+1] */
  const int* psp;            /* [ipsp[nnode]] */
  double gamma, chi, cfl;
  int limiter;               /* 0 none, 1 Barth, 2 Venkatakrishnan, 3 modified Venkatakrishnan */
  int sorder;                /* 1 or 2 */
  int no_cvbc;
  double qinf[ORC_NVARS];
  /* viscous terms -- compressibleNS (param.h: viscous, Re, Pr, PrT, ref_temperature, velocity;
     timestep.tcc:37-41 enableVNN / VNN) */
  int viscous, enable_vnn;
  double Re, Pr, PrT, tref, mach, vnn;
  const double* bedges_twall;  /* [nbedge] bcobj->twall / ref_temperature of the half-edge's surface
                                  (read for NoSlip only; < 0: adiabatic).  NULL: 1/tref (bcobj.tcc:31) */
  const double* mut;           /* field "mut" [nnode+gnode+nbnode]; NULL: zero (laminar) */
  /* time integration (Param::dt, useLocalTimeStepping, torder; SolutionSpace::iter, qold, qoldm1).  qold == NULL:
     q^n == q^{n+1} (first Newton iteration of a steady step): TemporalResidual contributes exact zeros and is skipped */
  double dt_param;             /* Param::dt (< 0: steady) -- perfect-gas eqnsets; the reacting one reads orc_fr_params */
  int use_local_dt, torder, iter;
  const double* qold;          /* [nnode*nvars] conservative variables at t^n */
  const double* qoldm1;        /* [nnode*nvars] ... at t^{n-1} */
  const double* walldist;      /* field "wallDistance" [nnode+gnode]; read by the FarFieldViscous BC (bc.tcc:1092-1108) */
  int field_jac_type, boundary_jac_type;   /* Param::fieldJacType / boundaryJacType: 0 one-sided FD, 1 central FD
                                              (jacobian.tcc:306-366, 546-640) */
  int grad_type;               /* Param::gradType: 0 weighted least squares, 1 Green-Gauss (gradient.tcc:77-90, 170-248) */
} orc_case;

/* gradient.tcc:115-138, 381-542 : s and sw, each [(nnode+gnode)*6] */
void orc_lsq_coefficients(const orc_case* c, double* s, double* sw);
/* gradient.tcc:57-112, 251-378, 545-565 */
void orc_gradient(const orc_case* c, const double* q, const double* sw, double* qgrad);
/* limiters.tcc:53-132 (+ kernels) */
void orc_limiter(const orc_case* c, const double* q, const double* qgrad, double* lim);
/* bc.tcc:1399-1457, 1058-1120; compressible.tcc:1246-1475 */
void orc_update_bcs(const orc_case* c, double* q, const double* beta);
/* residual.tcc:66-122, 192-387 ; returns nothing, b is overwritten */
void orc_residual(const orc_case* c, const double* q, const double* qgrad, const double* lim,
		  const double* beta, double* b);
/* timestep.tcc:7-143 (local time stepping branch); returns dtmin */
double orc_timestep(const orc_case* c, const double* q, const double* beta, double* dt);
/* solve.tcc:71-98 + compressible.tcc:929-993 */
void orc_explicit_solve(const orc_case* c, double* q, const double* b, const double* dt, double* x);
void orc_apply_dq(const orc_case* c, double* q, const double* x);

/* crsmatrix.tcc:48-97 : ia[nnode+1], ja[ia[nnode]], iau[nnode] */
void orc_crs_init(const orc_case* c, int* ia, int* ja, int* iau);
/* jacobian.tcc:13-18, 130-250, 254-304, 437-544 (one-sided FD, h = 1e-8).
   q is written (phantom nodes) exactly as the reference does. */
void orc_jacobian(const orc_case* c, double* q, const double* beta, const double* dt,
		  const int* ia, const int* ja, const int* iau, double* A);
/* crsmatrix.tcc:840-876 + matrix.h:110-190 */
void orc_prepare_sgs(const orc_case* c, const int* iau, double* A, int* pv);
/* crs.tcc:62-173 (single rank); returns |xOld - xNorm| */
double orc_sgs(const orc_case* c, int nsgs, const int* ia, const int* ja, const int* iau,
	       const double* A, const int* pv, const double* b, double* x);

/* compressible.tcc:713-795 ViscousFlux and :1633-1893 ViscousJacobian -- exposed for unit tests.
   Q, QL, QR are full nvars states (aux vars valid). */
void orc_viscous_flux(const orc_case* c, const double* Q, const double* grad, const double* avec, double mut,
		      double* flux);
void orc_viscous_jacobian(const orc_case* c, const double* QL, const double* QR, const double* dx, double s2,
			  const double* avec, double mut, double* aL, double* aR);
/* PowerLawU (ucs/powerLaw.h:11-29) */
double orc_power_law_u(double uinf, double wallDist, double Re);
/* most-normal neighbour of a wall node (bc.tcc:1182-1206) for half-edge e */
int orc_normal_node(const orc_case* c, int e);

/* Spalart-Allmaras one-equation model, one TurbulenceModel::Compute (turb.tcc:163-339 with spalart.tcc:141-361),
   turbulenceSpatialOrder = 1 (the reference's order-2 path calls the 5-equation ExtrapolateVariables on
   1-variable arrays, limiters.tcc:412 -- out-of-bounds, not reproducible).  In/out tvar [(nnode+gnode+nbnode)];
   q, qgrad, dt as left by the flow iteration; s = unweighted LSQ sums (Mesh::s); dist = field "wallDistance".
   Out: tgrad [(nnode+gnode)*3], b [nnode], A [nblocks] (diagonal entries inverted by PrepareSGS,
   crsmatrix.tcc:852-858), x [nnode+gnode], mut [nnode+gnode].  Returns ParallelL2Norm(b). */
double orc_turb_sa(const orc_case* c, int nsgs, const double* q, const double* qgrad, const double* s,
		   const double* dist, const double* dt, const int* ia, const int* ja, const int* iau,
		   double* tvar, double* tgrad, double* b, double* A, double* x, double* mut);
/* the eqnset as TurbulenceModel sees it (EqnSet::GetTheta / ComputeAuxiliaryVariables / GetDensity / ComputeViscosity /
   GetRe / GetVelocityGradLocation): perfect gas in pcfd_oracle.c, compressibleNSFR in pcfd_oracle_fr.c */
typedef struct orc_gas {
  int nvars, nterms, vloc;      /* row widths of q and qgrad/3, offset of the velocity gradient in a qgrad row */
  double Re;
  const void* ctx;
  double (*theta_avg)(const struct orc_gas* g, const double* qL, const double* qR, const double* avec);
  void (*rho_nu_avg)(const struct orc_gas* g, const double* qL, const double* qR, double* rho, double* nu);
  void (*rho_nu_node)(const struct orc_gas* g, const double* Q, double* rho, double* nu);
  /* what Forces asks in addition (forces.tcc:123-196, 400-478): Param::GetVelocity, GetDensity(Qinf), EqnSet::GetCp,
     GetPressure, ComputeViscosity + GetDensity of a stored state; set by orc_forces / orc_fr_forces only */
  double V, rho_inf;
  double (*cp_of)(const struct orc_gas* g, const double* Q);
  double (*pressure_of)(const struct orc_gas* g, const double* Q);
  void (*mu_rho_node)(const struct orc_gas* g, const double* Q, double* mu, double* rho);
} orc_gas;

/* ---- surface forces: ComputeSurfaceAreas (forces.tcc:199-312), Forces::Compute (:315-478).
   Composite bodies as the .bc file declares them ("body #k = [factags]", composite.tcc), 0-based here. */
typedef struct {
  int nbodies, num_bcs;            /* num_bcs = bc->largest_bc_id */
  const int* body_offsets;         /* [nbodies+1] into body_factags */
  const int* body_factags;
  const double* moment_pt;         /* [3*nbodies] CompositeBody::momentPt */
  const double* moment_axis;       /* [3*nbodies] CompositeBody::momentAxis */
  const int* bedges_factag;        /* [nbedge+ngedge] */
  const double* cg;                /* Mesh::cg [(nnode+gnode+nbnode)*3]: phantom nodes carry the face-piece centroid */
  double liftdir[3], dragdir[3];   /* Param::liftdir / dragdir, as given (not normalised) */
} orc_forces_desc;
/* surf_area [(num_bcs+1)*3], body_area [nbodies*3] */
void orc_surface_areas(const orc_case* c, const orc_forces_desc* d, double* surf_area, double* body_area);
/* cp, yp, cf [nbedge]; body [nbodies*12] = forces, vforces, moments, vmoments; coef [nbodies*3] = cl, cd, cm */
void orc_forces_gas(const orc_case* c, const orc_gas* gas, const orc_forces_desc* d, const double* q, const double* qgrad,
		    const double* body_area, double* cp, double* yp, double* cf, double* body, double* coef);
void orc_forces(const orc_case* c, const orc_forces_desc* d, const double* q, const double* qgrad,
		const double* body_area, double* cp, double* yp, double* cf, double* body, double* coef);
double orc_turb_sa_phase_gas(const orc_case* c, const orc_gas* gas, int phase, int nsgs, const double* q, const double* qgrad,
			     const double* s, const double* dist, const double* dt, const int* ia, const int* ja, const int* iau,
			     double* tvar, double* tgrad, double* b, double* A, double* x, double* mut);
/* the same update one phase at a time (0..5, see pcfd_oracle.c), for per-rank replays with halos in between; phase 2
   returns the sum of b^2 */
double orc_turb_sa_phase(const orc_case* c, int phase, int nsgs, const double* q, const double* qgrad, const double* s,
		   const double* dist, const double* dt, const int* ia, const int* ja, const int* iau,
		   double* tvar, double* tgrad, double* b, double* A, double* x, double* mut);

/* compressible.tcc:93-230 -- exposed for unit tests */
void orc_roe_flux(const double* QL, const double* QR, const double* avec, double vdotn, double gamma,
		  double* flux);

/* ---- finite-rate chemistry (compressibleFR source term).  Tables as the reference's ChemModel holds them
   (same layout as pcfd_chem_model in include/pcfd.h; kept separate so the oracle includes nothing of the product). */
#define ORC_CHEM_MAX_SPECIES 16
#define ORC_CHEM_MAX_REACTIONS 32
typedef struct {
  int nspecies, nreactions;
  double mw[ORC_CHEM_MAX_SPECIES];
  double nasa7[ORC_CHEM_MAX_SPECIES][2][7];
  int rxn_type[ORC_CHEM_MAX_REACTIONS], third_body[ORC_CHEM_MAX_REACTIONS], backward_given[ORC_CHEM_MAX_REACTIONS];
  int rxn_type_b[ORC_CHEM_MAX_REACTIONS], nsp[ORC_CHEM_MAX_REACTIONS];
  int species[ORC_CHEM_MAX_REACTIONS][ORC_CHEM_MAX_SPECIES];
  double A[ORC_CHEM_MAX_REACTIONS], EA[ORC_CHEM_MAX_REACTIONS], n[ORC_CHEM_MAX_REACTIONS];
  double Ab[ORC_CHEM_MAX_REACTIONS], EAb[ORC_CHEM_MAX_REACTIONS], nb[ORC_CHEM_MAX_REACTIONS];
  double nup[ORC_CHEM_MAX_REACTIONS][ORC_CHEM_MAX_SPECIES];
  double nupp[ORC_CHEM_MAX_REACTIONS][ORC_CHEM_MAX_SPECIES];
  double tbeff[ORC_CHEM_MAX_REACTIONS][ORC_CHEM_MAX_SPECIES];
} orc_chem_model;

/* ChemModel::GetMassProductionRates (chem.tcc:575-583; reaction.tcc:607-856; species.tcc:96-138) for n states:
   rhoi [n*ns] kg/m^3, T [n] K -> wdot [n*ns] kg/(m^3 s).  wscale (may be NULL) [n*ns] receives
   MW_i sum_r |nu_ir| Gamma_r (|k_f prod_f| + |k_b prod_b|): the magnitude the net rates cancel from, i.e. the
   scale against which a rounding-level comparison of wdot is meaningful.  kf, kb (may be NULL) [n*nr]. */
void orc_chem_mass_production(const orc_chem_model* m, int n, const double* rhoi, const double* T, double* wdot,
			      double* wscale, double* kf, double* kb);
/* CompressibleFREqnSet::SourceTerm (compressibleFR.tcc:1276-1316), rxnOn, gravity off */
void orc_chem_source_term(const orc_chem_model* m, int n, int stride, const double* Q, const double* vol,
			  double ref_density, double ref_time, double ref_temperature, double* source);

/* ---- reacting eqnset (CompressibleFREqnSet, compressibleFR.tcc), oracle/pcfd_oracle_fr.c.  The mesh, chi, cfl,
   limiter, sorder, no_cvbc and the VNN fields of orc_case are used; gamma and qinf[10] are not.  ns = nspecies:
   neqn = ns+4, nvars = 3ns+6 [rho_i | u v w | T | P | rho | cv_i | mol_i], nterms = 2ns+4. */
/* species transport data as Species holds it (species.h:35-49, species.tcc:13-22, 241-310): Sutherland law below the
   transition temperature (White), NASA RP-1311 fits [Tlo, Thi, A, B, C, D] above */
typedef struct {
  int nmu[ORC_CHEM_MAX_SPECIES], nk[ORC_CHEM_MAX_SPECIES];
  double mu_fit[ORC_CHEM_MAX_SPECIES][3][6], k_fit[ORC_CHEM_MAX_SPECIES][3][6];
  double mu_white[ORC_CHEM_MAX_SPECIES][4], k_white[ORC_CHEM_MAX_SPECIES][4];   /* ref value, T0, S, transition T */
} orc_transport;

typedef struct {
  const orc_chem_model* chem;
  double ref_density, ref_velocity, ref_temperature, ref_pressure, ref_time, ref_specific_enthalpy;  /* param.tcc:352-398 */
  double Pref;             /* CompressibleFREqnSet::Pref = GetPressure(Qinf) (compressibleFR.tcc:1515) */
  double dt_param;         /* Param::dt (negative: steady) */
  int use_local_dt;        /* Param::useLocalTimeStepping */
  int rxn_on;              /* Param::rxnOn */
  double qinf[3*ORC_CHEM_MAX_SPECIES + 6];
  /* compressibleNSFR (orc_case.viscous, Re, PrT, mut are used as for compressibleNS) */
  const orc_transport* transport;   /* NULL unless viscous */
  double ref_viscosity, ref_k;      /* param.tcc:210-211 */
} orc_fr_params;

/* ChemModel::GetViscosity / GetThermalConductivity (chem.tcc:876-938): Wilke-mixed, dimensional (rhoi kg/m^3, T K) */
double orc_fr_mixture_viscosity(const orc_fr_params* p, const double* rhoi, double T);
double orc_fr_mixture_conductivity(const orc_fr_params* p, const double* rhoi, double T);
/* CompressibleFREqnSet::ViscousFlux (compressibleFR.tcc:551-637) and ViscousJacobian (:1713-2040); Q, QL, QR full states */
void orc_fr_viscous_flux(const orc_case* c, const orc_fr_params* p, const double* Q, const double* grad, const double* avec,
			 double mut, double* flux);
void orc_fr_viscous_jacobian(const orc_case* c, const orc_fr_params* p, const double* QL, const double* QR, const double* dx,
			     double s2, const double* avec, double mut, double* aL, double* aR);

void orc_fr_update_bcs(const orc_case* c, const orc_fr_params* p, double* q, const double* beta);
void orc_fr_gradient(const orc_case* c, const orc_fr_params* p, const double* q, const double* sw, double* qgrad);
void orc_fr_limiter(const orc_case* c, const orc_fr_params* p, const double* q, const double* qgrad, double* lim);
void orc_fr_residual(const orc_case* c, const orc_fr_params* p, const double* q, const double* qgrad,
		     const double* lim, const double* beta, double* b);
/* Forces::Compute under the reacting eqnset; V = Param::velocity */
void orc_fr_forces(const orc_case* c, const orc_fr_params* p, const orc_forces_desc* d, double V, const double* q,
		   const double* qgrad, const double* body_area, double* cp, double* yp, double* cf, double* body, double* coef);
double orc_fr_timestep(const orc_case* c, const orc_fr_params* p, const double* q, const double* beta, double* dt);
/* returns the number of nodes whose ConservativeToNative Newton iteration did not converge (the reference aborts) */
int orc_fr_explicit_solve(const orc_case* c, const orc_fr_params* p, double* q, const double* b, const double* dt,
			  double* x);
void orc_fr_apply_dq(const orc_case* c, const orc_fr_params* p, double* q, const double* x);
void orc_fr_jacobian(const orc_case* c, const orc_fr_params* p, double* q, const double* beta, const double* dt,
		     const int* ia, const int* ja, const int* iau, double* A);
void orc_fr_prepare_sgs(const orc_case* c, const orc_fr_params* p, const int* iau, double* A, int* pv);
double orc_fr_sgs(const orc_case* c, const orc_fr_params* p, int nsgs, const int* ia, const int* ja, const int* iau,
		  const double* A, const int* pv, const double* b, double* x);
void orc_fr_compute_aux(const orc_fr_params* p, double* Q);
void orc_fr_hllc_flux(const orc_fr_params* p, const double* QL, const double* QR, const double* avec, double vdotn,
		      double beta, double* flux);

#ifdef __cplusplus
}
#endif
/* Spalart-Allmaras under compressibleNSFR: the eqnset-agnostic TurbulenceModel::Compute of pcfd_oracle.c with the reacting
   eqnset's accessors (Wilke-mixed viscosity, density from the aux variables, native velocities); arrays as orc_turb_sa with
   q rows of 3 ns + 6 and qgrad rows of (2 ns + 4) x 3 doubles */
double orc_fr_turb_sa_phase(const orc_case* c, const orc_fr_params* p, int phase, int nsgs, const double* q, const double* qgrad,
			    const double* s, const double* dist, const double* dt, const int* ia, const int* ja, const int* iau,
			    double* tvar, double* tgrad, double* b, double* A, double* x, double* mut);
double orc_fr_turb_sa(const orc_case* c, const orc_fr_params* p, int nsgs, const double* q, const double* qgrad, const double* s,
		      const double* dist, const double* dt, const int* ia, const int* ja, const int* iau,
		      double* tvar, double* tgrad, double* b, double* A, double* x, double* mut);

/* Kernel_NumJac_Complex (jacobian.tcc:370-433) over the interior edges, added into A (pcfd_oracle_cs.c) */
void orc_jac_edges_complex(int nedge, const int* edges_n, const double* edges_a, const double* q, int nvars, double gamma_r,
			   const int* ia, const int* ja, double* A);

/* CRSMatrix::CRSTranspose (crsmatrix.tcc:568-599) up to its parallel sync (PObj::TransposeCommCRS replaces the ghost-column
   blocks afterwards; tests/test_crs_transpose.py routes them) */
void orc_crs_transpose_local(int nnode, int neqn, const int* ia, const int* ja, double* A);

/* the local ILU0 preconditioner of CRS::GMRES: CRSMatrix::BuildILU0Local (crsmatrix.tcc:276-428) in place on a copy N of
   the assembled matrix, and CRSMatrix::ILU0BackSub (crsmatrix.tcc:430-507), N x = b with x blanked first */
void orc_ilu0_build(int nnode, int neqn, const int* ia, const int* ja, const int* iau, double* N);
void orc_ilu0_backsub(int nnode, int gnode, int neqn, const int* ia, const int* ja, const int* iau, const double* N,
		      double* x, const double* b);

/* CRS::GMRES (crs.tcc:176-415) on one rank: restarts x nSearchDir, right preconditioner type 0 / 1 / 2 / 3 / 4 (none,
   diagonal, block diagonal LU, local ILU0, SGS; crs.tcc:555-641); any block size neqn <= 32.  A: assembled matrix before PrepareSGS; x: initial guess
   in, solution out, (nnode+gnode)*neqn; returns the reference's dqNorm. */
double orc_gmres(int nnode, int gnode, int neqn, int restarts, int nSearchDir, int precondType, const int* ia,
		 const int* ja, const int* iau, const double* A, const double* b, double* x);

#endif
