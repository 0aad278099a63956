/*
 * ref_harness.cpp -- parity/timing harness around the UNMODIFIED reference
 * (ngcurrier/ProteusCFD, /root/reference/ucs).  TEST INFRASTRUCTURE ONLY.
 *
 * This TU contains no solver arithmetic of its own: it builds a
 * SolutionSpace<Real> exactly like ucs/main.cpp:167-223 does, injects a smooth
 * analytic state, then calls the reference's own phase entry
 * points in the order of SolutionSpace::NewtonIterate
 * (ucs/solutionSpace.tcc:640-865) and dumps every intermediate array as a raw
 * little-endian binary so the CUDA path can be compared with it.
 *
 *   ref_harness <case> <outdir> dump          one pass, dump everything
 *   ref_harness <case> <outdir> time <reps>   time each phase, print one JSON line
 *
 * Compiled by oracle/Makefile from the sources where they lie in
 * /root/reference; output goes to oracle/_ref/ only.
 */
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <sys/time.h>
#include <mpi.h>

// the reference keeps its halo maps private (ucs/parallel.h:109-121); the
// harness only reads them.
#define private public
#define protected public
#include "general.h"
#include "exceptions.h"
#include "bc.h"
#include "mesh.h"
#include "param.h"
#include "eqnset.h"
#include "create_functions.h"
#include "parallel.h"
#include "timer.h"
#include "customics.h"
#include "derivatives.h"
#include "portFileio.h"
#include "move.h"
#include "composite.h"
#include "solutionSpaceBase.h"
#include "solutionSpace.h"
#include "dataInfo.h"
#include "solve.h"
#include "solutionOrdering.h"
#include "fluid_structure.h"
#include "temporalControl.h"
#include "pythonInterface.h"
#include "timestep.h"
#include "residual.h"
#include "jacobian.h"
#include "gradient.h"
#include "limiters.h"
#include "compressibleFR.h"
#include "chem.h"
#include "reaction.h"
#include "species.h"
#undef private
#undef protected

#ifdef PCFD_DROPIN
// the same harness with every phase call swapped for the B200 drop-in shim: this is what
// INTEGRATION.md tells a ucs.x maintainer to do, compiled against the real reference headers
#include "pcfd_host.hpp"
#endif

static std::string g_out;
static int g_rank = 0;

static double Now()
{
  timeval tv; gettimeofday(&tv, NULL);
  return tv.tv_sec + 1e-6*tv.tv_usec;
}

template <class T>
static void Dump(const std::string& name, const T* data, size_t n)
{
  std::ostringstream fn;
  fn << g_out << "/" << name << "." << g_rank << ".bin";
  FILE* f = fopen(fn.str().c_str(), "wb");
  if(!f){ perror(fn.str().c_str()); exit(2); }
  if(n) fwrite(data, sizeof(T), n, f);
  fclose(f);
}

// smooth non-uniform state of SURVEY.md 8d, in the reference's own
// non-dimensionalisation (rho_inf = 1, c_inf = 1, p_inf = 1/gamma)
static void InjectState(SolutionSpace<Real>* space)
{
  Mesh<Real>* m = space->m;
  EqnSet<Real>* eqnset = space->eqnset;
  Param<Real>* param = space->param;
  Int neqn = eqnset->neqn;
  Int nvars = neqn + eqnset->nauxvars;
  Int nnode = m->GetNumNodes();
  const Real twopi = 2.0*3.14159265358979323846;
  Real gamma = param->gamma;
  Real mach = param->GetVelocity(space->iter);
  if(param->eqnset_id == CompressibleEulerFR || param->eqnset_id == CompressibleNSFR){
    // reacting eqnset: native variables [rho_i..., u, v, w, T] (compressibleFR.tcc:14-31); species densities,
    // velocity and temperature perturbed around the free stream so that every reaction is active
    CompressibleFREqnSet<Real>* fr = dynamic_cast<CompressibleFREqnSet<Real>*>(eqnset);
    Int ns = fr->nspecies;
    for(Int i = 0; i < nnode; i++){
      Real x = m->xyz[3*i + 0], y = m->xyz[3*i + 1], z = m->xyz[3*i + 2];
      Real* q = &space->q[i*nvars];
      for(Int k = 0; k < ns; k++){
	q[k] = eqnset->Qinf[k]*(1.0 + 0.1*sin(twopi*x)*cos(twopi*y))*(1.0 + 0.05*sin(twopi*(z + 0.17*k)));
      }
      q[ns+0] = mach*param->flowdir[0] + 0.05*sin(twopi*y);
      q[ns+1] = mach*param->flowdir[1] + 0.05*sin(twopi*z);
      q[ns+2] = mach*param->flowdir[2] + 0.05*sin(twopi*x);
      q[ns+3] = eqnset->Qinf[ns+3]*(1.0 + 0.1*cos(twopi*z));
      eqnset->ComputeAuxiliaryVariables(q);
    }
    return;
  }
  for(Int i = 0; i < nnode; i++){
    Real x = m->xyz[3*i + 0], y = m->xyz[3*i + 1], z = m->xyz[3*i + 2];
    Real rho = 1.0 + 0.1*sin(twopi*x)*cos(twopi*y);
    Real u = mach*param->flowdir[0] + 0.05*sin(twopi*y);
    Real v = mach*param->flowdir[1] + 0.05*sin(twopi*z);
    Real w = mach*param->flowdir[2] + 0.05*sin(twopi*x);
    Real p = 1.0/gamma*(1.0 + 0.1*cos(twopi*z));
    Real* q = &space->q[i*nvars];
    q[0] = rho;
    q[1] = rho*u;
    q[2] = rho*v;
    q[3] = rho*w;
    q[4] = p/(gamma - 1.0) + 0.5*rho*(u*u + v*v + w*w);
    eqnset->ComputeAuxiliaryVariables(q);
  }
}

static void DumpMesh(SolutionSpace<Real>* space)
{
  Mesh<Real>* m = space->m;
  PObj<Real>* p = space->p;
  Int nnode = m->GetNumNodes(), gnode = m->GetNumParallelNodes(), nbnode = m->GetNumBoundaryNodes();
  Int nedge = m->GetNumEdges(), nbedge = m->GetNumBoundaryEdges(), ngedge = m->GetNumParallelEdges();
  Int nb = nbedge + ngedge;

  std::vector<Int> en(2*(size_t)nedge);
  std::vector<Real> ea(4*(size_t)nedge);
  for(Int e = 0; e < nedge; e++){
    en[2*e] = m->edges[e].n[0]; en[2*e+1] = m->edges[e].n[1];
    for(Int k = 0; k < 4; k++) ea[4*e+k] = m->edges[e].a[k];
  }
  Dump("edges_n", en.data(), en.size());
  Dump("edges_a", ea.data(), ea.size());
  std::vector<Int> bn(2*(size_t)nb), bf(nb), bt(nb);
  std::vector<Real> ba(4*(size_t)nb);
  for(Int e = 0; e < nb; e++){
    bn[2*e] = m->bedges[e].n[0]; bn[2*e+1] = m->bedges[e].n[1];
    for(Int k = 0; k < 4; k++) ba[4*e+k] = m->bedges[e].a[k];
    bf[e] = m->bedges[e].factag;
    bt[e] = space->bc->GetBCType(m->bedges[e].factag);
  }
  Dump("bedges_n", bn.data(), bn.size());
  Dump("bedges_a", ba.data(), ba.size());
  Dump("bedges_factag", bf.data(), bf.size());
  {
    // non-dimensional wall temperature of each BC half-edge's surface, as bc.tcc:1283-1286 forms it
    std::vector<Real> tw(nbedge);
    for(Int e = 0; e < nbedge; e++){
      tw[e] = space->bc->GetBCObj(m->bedges[e].factag)->twall / space->param->ref_temperature;
    }
    Dump("bedges_twall", tw.data(), tw.size());
  }
  Dump("bedges_bctype", bt.data(), bt.size());
  Dump("xyz", m->xyz, 3*(size_t)(nnode+gnode));
  if(getenv("PCFD_RCM")){
    // Mesh::ReorderMeshCuthillMcKee (mesh.tcc:2412-2494) on the built neighbour lists: only the permutation is taken
    // (ordering[new] = old), the mesh itself stays as it is; PCFD_RCM=2: reversed
    std::vector<Int> keep(m->ordering, m->ordering + nnode);
    m->ReorderMeshCuthillMcKee(atoi(getenv("PCFD_RCM")) == 2 ? 1 : 0);
    Dump("rcm_ordering", m->ordering, (size_t)nnode);
    for(Int k = 0; k < nnode; k++) m->ordering[k] = keep[k];
  }
  if(getenv("PCFD_DUMP_ELEMENTS")){
    // element list in the reference's internal winding (etypes.h: TRI 0 .. HEX 5): type, factag, 8 node slots (-1 padded)
    std::vector<Int> et, ef, enodes;
    for(size_t k = 0; k < m->elementList.size(); k++){
      Element<Real>& el = *m->elementList[k];
      Int* nodes = NULL;
      Int nn = el.GetNodes(&nodes);
      et.push_back(el.GetType());
      ef.push_back(el.GetFactag());
      for(Int j = 0; j < 8; j++) enodes.push_back(j < nn ? nodes[j] : -1);
    }
    Dump("elem_type", et.data(), et.size());
    Dump("elem_factag", ef.data(), ef.size());
    Dump("elem_nodes", enodes.data(), enodes.size());
  }
  Dump("vol", m->vol, (size_t)nnode);
  Dump("ipsp", m->ipsp, (size_t)nnode+1);
  Dump("psp", m->psp, (size_t)m->ipsp[nnode]);
  Dump("lsq_s", m->s, 6*(size_t)(nnode+gnode));
  Dump("lsq_sw", m->sw, 6*(size_t)(nnode+gnode));
  Dump("gNodeOwner", m->gNodeOwner, (size_t)gnode);
  Dump("gNodeLocalId", m->gNodeLocalId, (size_t)gnode);
  Int np = p->GetNp();
  Dump("commCountsSend", p->commCountsSend, (size_t)np);
  Dump("commCountsRecv", p->commCountsRecv, (size_t)np);
  Dump("commOffsetsRecv", p->commOffsetsRecv, (size_t)np);
  {
    std::vector<Int> pack;
    for(Int r = 0; r < np; r++){
      for(Int j = 0; j < p->commCountsSend[r]; j++) pack.push_back(p->nodePackingList[r][j]);
    }
    Dump("nodePackingList", pack.data(), pack.size());
  }
  EqnSet<Real>* eqnset = space->eqnset;
  Int nvars = eqnset->neqn + eqnset->nauxvars;
  Dump("qinf", eqnset->Qinf, (size_t)nvars);
  Dump("beta", space->GetFieldData("beta", FIELDS::STATE_NONE), (size_t)(nnode+gnode));
  if(space->param->viscous){
    // eddy viscosity seen by the flow's viscous flux / Jacobian (zero unless a turbulence model initialised it)
    Dump("mut", space->GetFieldData("mut", FIELDS::STATE_NONE), (size_t)(nnode+gnode));
    // wall distance (read by the FarFieldViscous BC and the turbulence model)
    Dump("wallDistance", space->GetFieldData("wallDistance", FIELDS::STATE_NONE), (size_t)(nnode+gnode));
  }

  Param<Real>* param = space->param;
  std::ostringstream fn;
  fn << g_out << "/meta." << g_rank << ".txt";
  std::ofstream f(fn.str().c_str());
  f << std::setprecision(17);
  f << "rank " << p->GetRank() << "\nnp " << np << "\n";
  f << "nnode " << nnode << "\ngnode " << gnode << "\nnbnode " << nbnode << "\n";
  f << "nedge " << nedge << "\nnbedge " << nbedge << "\nngedge " << ngedge << "\n";
  f << "neqn " << eqnset->neqn << "\nnvars " << nvars << "\nnterms " << space->grad->GetNterms() << "\n";
  f << "gamma " << param->gamma << "\nchi " << param->chi << "\nsorder " << param->sorder << "\n";
  f << "limiter " << param->limiter << "\nflux_id " << param->flux_id << "\nnSgs " << param->nSgs << "\n";
  f << "cfl " << param->GetCFL() << "\nvelocity " << param->GetVelocity(space->iter) << "\n";
  f << "viscous " << param->viscous << "\nRe " << param->Re << "\nPr " << param->Pr << "\nPrT " << param->PrT << "\n";
  f << "useLocalTimeStepping " << param->useLocalTimeStepping << "\ndt " << param->dt << "\n";
  f << "torder " << param->torder << "\nfieldJacType " << param->fieldJacType << "\n";
  f << "boundaryJacType " << param->boundaryJacType << "\nboundaryJacEval " << param->boundaryJacEval << "\n";
  f << "no_cvbc " << param->no_cvbc << "\nsymmetry2D " << param->symmetry2D << "\n";
  f << "eqnset_id " << param->eqnset_id << "\ngradType " << param->gradType << "\n";
  f << "ref_temperature " << param->ref_temperature << "\nenableVNN " << param->enableVNN << "\nVNN " << param->VNN << "\n";
  f << "turbModel " << param->turbModel << "\nturbModelSorder " << param->turbModelSorder << "\n";
  f << "iter " << space->iter << "\nnFirstOrderSteps " << param->nFirstOrderSteps << "\n";
  f << "ref_density " << param->ref_density << "\nref_velocity " << param->ref_velocity << "\n";
  f << "ref_pressure " << param->ref_pressure << "\nref_time " << param->ref_time << "\n";
  f << "ref_length " << param->ref_length << "\nref_specific_enthalpy " << param->ref_specific_enthalpy << "\n";
  f << "rxnOn " << param->rxnOn << "\ngravity_on " << param->gravity_on << "\n";
  f << "ref_viscosity " << param->ref_viscosity << "\nref_k " << param->ref_k << "\n";
  if(param->eqnset_id == CompressibleEulerFR || param->eqnset_id == CompressibleNSFR){
    CompressibleFREqnSet<Real>* fr = dynamic_cast<CompressibleFREqnSet<Real>*>(eqnset);
    f << "nspecies " << fr->nspecies << "\nPref " << fr->Pref << "\n";
  }
  f.close();
  if(param->eqnset_id == CompressibleEulerFR || param->eqnset_id == CompressibleNSFR){
    // the chemistry tables exactly as the reference's ChemModel holds them (same layout as ref_chem.cpp)
    CompressibleFREqnSet<Real>* fr = dynamic_cast<CompressibleFREqnSet<Real>*>(eqnset);
    ChemModel<Real>& chem = *fr->chem;
    Int ns = chem.nspecies, nr = chem.nreactions;
    std::vector<Real> mw(ns), coeff(ns*14), Rs(ns);
    for(Int i = 0; i < ns; i++){
      mw[i] = chem.species[i].MW;
      Rs[i] = chem.species[i].R;
      for(Int k = 0; k < 7; k++){
	coeff[i*14 + k] = chem.species[i].thermo_coeff[0][k];
	coeff[i*14 + 7 + k] = chem.species[i].thermo_coeff[1][k];
      }
    }
    Dump("species_mw", mw.data(), mw.size());
    Dump("species_R", Rs.data(), Rs.size());
    Dump("species_nasa7", coeff.data(), coeff.size());
    {
      // transport data as Species holds it (species.h:35-49): Sutherland coefficients below the transition
      // temperature, NASA RP-1311 fits [Tlo, Thi, A, B, C, D] above (up to 3 ranges, unused rows zero)
      std::vector<Real> mufit(ns*18, 0.0), kfit(ns*18, 0.0), white(ns*8);
      std::vector<Int> counts(ns*2);
      for(Int i = 0; i < ns; i++){
	Species<Real>& sp = chem.species[i];
	counts[2*i] = sp.mu_coeff_curves;
	counts[2*i+1] = sp.k_coeff_curves;
	for(Int r = 0; r < sp.mu_coeff_curves; r++) for(Int k = 0; k < 6; k++) mufit[i*18 + r*6 + k] = sp.mu_coeff[r][k];
	for(Int r = 0; r < sp.k_coeff_curves; r++) for(Int k = 0; k < 6; k++) kfit[i*18 + r*6 + k] = sp.k_coeff[r][k];
	for(Int k = 0; k < 3; k++){ white[i*8 + k] = sp.mu_coeff_White[k]; white[i*8 + 4 + k] = sp.k_coeff_White[k]; }
	white[i*8 + 3] = sp.mu_transition_White;
	white[i*8 + 7] = sp.k_transition_White;
      }
      Dump("species_mu_fit", mufit.data(), mufit.size());
      Dump("species_k_fit", kfit.data(), kfit.size());
      Dump("species_white", white.data(), white.size());
      Dump("species_fit_counts", counts.data(), counts.size());
    }
    std::vector<Real> rk(nr*3), nup(nr*ns, 0.0), nupp(nr*ns, 0.0), tbeff(nr*ns, 1.0);
    std::vector<Int> flags(nr*4), order(nr*ns, -1);
    for(Int j = 0; j < nr; j++){
      Reaction<Real>& r = chem.reactions[j];
      rk[j*3] = r.A; rk[j*3+1] = r.EA; rk[j*3+2] = r.n;
      flags[j*4] = r.rxnType; flags[j*4+1] = r.thirdBodiesPresent; flags[j*4+2] = r.backwardRateGiven;
      flags[j*4+3] = r.GetNspecies();
      for(Int k = 0; k < r.GetNspecies(); k++){
	order[j*ns + k] = r.globalIndx[k];
	nup[j*ns + k] = r.Nup[k];
	nupp[j*ns + k] = r.Nupp[k];
	if(r.thirdBodiesPresent && (size_t)k < r.TBEff.size()) tbeff[j*ns + k] = r.TBEff[k];
      }
    }
    Dump("rxn_A_EA_n", rk.data(), rk.size());
    Dump("rxn_flags", flags.data(), flags.size());
    Dump("rxn_species", order.data(), order.size());
    Dump("rxn_nup", nup.data(), nup.size());
    Dump("rxn_nupp", nupp.data(), nupp.size());
    Dump("rxn_tbeff", tbeff.data(), tbeff.size());
    Int dims[2] = {ns, nr};
    Dump("chem_dims", dims, 2);
  }
}


// Forces / surface post-processing (PCFD_FORCES set; bodies come from the "body #k = [...]" lines of the .bc file): the
// reference's own ComputeSurfaceAreas (forces.tcc:199-262) and Forces::Compute (:317-418: FORCE_Kernel, YpCf_Kernel,
// ComputeCl) on the state and gradient the iteration left behind
static void InjectForcesGeometry(SolutionSpace<Real>* space)
{
  Forces<Real>* f = space->forces;
  if(f->num_bodies >= 1){
    f->bodies[1].momentPt[0] = 0.25; f->bodies[1].momentPt[1] = 0.1; f->bodies[1].momentPt[2] = -0.05;
  }
  if(f->num_bodies >= 2){
    f->bodies[2].momentAxis[0] = 0.0; f->bodies[2].momentAxis[1] = 1.0; f->bodies[2].momentAxis[2] = 0.0;
  }
}

static void DumpForcesResult(SolutionSpace<Real>* space)
{
  Mesh<Real>* m = space->m;
  Param<Real>* param = space->param;
  Forces<Real>* f = space->forces;
  Int nbedge = m->GetNumBoundaryEdges();
  Dump("forces_cp", f->cp, (size_t)nbedge);
  Dump("forces_yp", f->yp, (size_t)nbedge);
  Dump("forces_cf", f->cf, (size_t)nbedge);
  Dump("forces_surfArea", f->surfArea, (size_t)(f->num_bcs+1)*3);
  std::vector<int> lists;
  std::vector<Real> body, geom;
  for(Int i = 1; i <= f->num_bodies; i++){
    CompositeBody<Real>& b = f->bodies[i];
    lists.push_back(b.nsurfs);
    for(Int k = 0; k < b.nsurfs; k++) lists.push_back(b.list[k]);
    for(Int k = 0; k < 3; k++) body.push_back(b.forces[k]);
    for(Int k = 0; k < 3; k++) body.push_back(b.vforces[k]);
    for(Int k = 0; k < 3; k++) body.push_back(b.moments[k]);
    for(Int k = 0; k < 3; k++) body.push_back(b.vmoments[k]);
    for(Int k = 0; k < 3; k++) body.push_back(b.surfArea[k]);
    body.push_back(b.cl); body.push_back(b.cd); body.push_back(b.cm);
    for(Int k = 0; k < 3; k++) geom.push_back(b.momentPt[k]);
    for(Int k = 0; k < 3; k++) geom.push_back(b.momentAxis[k]);
  }
  Dump("forces_body_lists", lists.data(), lists.size());
  Dump("forces_body", body.data(), body.size());
  Dump("forces_body_geom", geom.data(), geom.size());
  Real dirs[6] = {param->liftdir[0], param->liftdir[1], param->liftdir[2], param->dragdir[0], param->dragdir[1], param->dragdir[2]};
  Dump("forces_dirs", dirs, 6);
}

static void DumpForces(SolutionSpace<Real>* space)
{
  if(!getenv("PCFD_FORCES")) return;
  Mesh<Real>* m = space->m;
  Int nnode = m->GetNumNodes(), gnode = m->GetNumParallelNodes(), nbnode = m->GetNumBoundaryNodes();
  Int nvars = space->eqnset->neqn + space->eqnset->nauxvars;
  InjectForcesGeometry(space);
  Dump("forces_q", space->q, (size_t)(nnode+gnode+nbnode)*nvars);
  Dump("forces_qgrad", space->qgrad, (size_t)(nnode+gnode)*space->grad->GetNterms()*3);
  Dump("forces_cg", m->cg, (size_t)(nnode+gnode+nbnode)*3);
  ComputeSurfaceAreas(space, 0);
  space->forces->Compute();
  DumpForcesResult(space);
}

// Spalart-Allmaras (turbulenceModel = 1): inject a smooth positive nu~ field and run the reference's own
// TurbulenceModel::Compute (turb.tcc:163-339) on the state the flow iteration left behind
static void DumpTurbulence(SolutionSpace<Real>* space)
{
  Mesh<Real>* m = space->m;
  Param<Real>* param = space->param;
  if(!param->viscous || param->turbModel != 1) return;
  TurbulenceModel<Real>* turb = space->turb;
  Int nnode = m->GetNumNodes(), gnode = m->GetNumParallelNodes(), nbnode = m->GetNumBoundaryNodes();
  Int nb = m->GetNumBoundaryEdges() + m->GetNumParallelEdges();
  const Real twopi = 2.0*3.14159265358979323846;
  for(Int i = 0; i < nnode; i++){
    Real x = m->xyz[3*i + 0], y = m->xyz[3*i + 1], z = m->xyz[3*i + 2];
    turb->tvar[i] = turb->tvarinf[0]*(1.0 + 0.3*sin(twopi*x)*cos(twopi*z) + 0.2*sin(twopi*y));
  }
  space->p->UpdateGeneralVectors(turb->tvar, 1);
  for(Int e = 0; e < nb; e++){
    if(!m->IsGhostNode(m->bedges[e].n[1])) turb->tvar[m->bedges[e].n[1]] = turb->tvar[m->bedges[e].n[0]];
  }
  memcpy(turb->tvarold, turb->tvar, sizeof(Real)*(size_t)(nnode+gnode+nbnode));
  memcpy(turb->tvaroldm1, turb->tvar, sizeof(Real)*(size_t)(nnode+gnode+nbnode));
  Dump("turb_tvar0", turb->tvar, (size_t)(nnode+gnode+nbnode));
  Dump("turb_q", space->q, (size_t)(nnode+gnode+nbnode)*(space->eqnset->neqn + space->eqnset->nauxvars));
  Dump("turb_qgrad", space->qgrad, (size_t)(nnode+gnode)*space->grad->GetNterms()*3);
  Dump("turb_dt", space->GetFieldData("timestep", FIELDS::STATE_NONE), (size_t)nnode);
  Dump("wallDistance", space->GetFieldData("wallDistance", FIELDS::STATE_NONE), (size_t)(nnode+gnode));
  Real res = turb->Compute();
  Dump("turb_res", &res, 1);
  Dump("turb_tgrad", turb->tgrad, (size_t)(nnode+gnode)*3);
  Dump("turb_b", turb->crs.b, (size_t)nnode);
  Dump("turb_x", turb->crs.x, (size_t)(nnode+gnode));
  Dump("turb_A", turb->crs.A->M, (size_t)turb->crs.A->nblocks);
  Dump("turb_tvar1", turb->tvar, (size_t)(nnode+gnode+nbnode));
  Dump("turb_mut", space->GetFieldData("mut", FIELDS::STATE_NONE), (size_t)(nnode+gnode));
}

int main(int argc, char* argv[])
{
  MPI_Init(&argc, &argv);
  if(argc < 4){
    std::cerr << "usage: " << argv[0] << " <case> <outdir> dump|time [reps]" << std::endl;
    return 1;
  }
  std::string casestring = argv[1];
  g_out = argv[2];
  std::string mode = argv[3];
  Int reps = (argc > 4) ? atoi(argv[4]) : 1;
  mkdir(g_out.c_str(), 0777);

  std::vector<Param<Real>*> paramList;
  SolutionOrdering<Real> operations;
  TemporalControl<Real> temporalControl;

  PObj<Real> pobj;
  g_rank = pobj.GetRank();

  size_t pos = casestring.rfind('/');
  std::string pathname;
  if(pos != std::string::npos){
    pathname = casestring.substr(0, pos+1);
    casestring = casestring.substr(pos);
  }
  else{
    pathname = "./";
  }
  Abort.rootDirectory = pathname;

  // quiet the chatter of all ranks but keep stderr of rank 0
  {
    std::ostringstream lo;
    lo << g_out << "/log." << g_rank << ".txt";
    freopen(lo.str().c_str(), "w", stdout);
    if(g_rank != 0) freopen("/dev/null", "w", stderr);
  }

  HDF_TurnOffErrorHandling();
  if(ReadParamFile(paramList, casestring, pathname)){ Abort << "Error in param read"; return -1; }
  if(operations.Read(casestring, pathname)){ Abort << "Error in solution ordering read"; return -1; }
  if(temporalControl.Read(casestring, pathname)){ Abort << "Error in temporal control read"; return -1; }
  for(UInt i = 0; i < paramList.size(); ++i) paramList[i]->mode = 0;

  std::vector<SolutionSpaceBase<Real>*> solSpaces = CreateSolutionSpaces(paramList, pobj, temporalControl);
  SolutionSpace<Real>* space = dynamic_cast<SolutionSpace<Real>*>(solSpaces[0]);
  Mesh<Real>* m = space->m;
  EqnSet<Real>* eqnset = space->eqnset;
  Param<Real>* param = space->param;
  PObj<Real>* p = space->p;
  Int neqn = eqnset->neqn;
  Int nvars = neqn + eqnset->nauxvars;
  Int nnode = m->GetNumNodes(), gnode = m->GetNumParallelNodes(), nbnode = m->GetNumBoundaryNodes();
  Int nterms = space->grad->GetNterms();

  // same bookkeeping as SolutionSpace::PreIterate (solutionSpace.tcc:510-545)
  space->iter = param->nFirstOrderSteps + 1;
  param->UpdateCFL(space->iter, 9999.0);

  InjectState(space);
  // a steady run has q^n == q^{n+1} at the first Newton iteration of a step
  // (fields.EvolveInTime, solutionSpace.tcc:582-592): make TemporalResidual
  // (residual.tcc:125-179) see a zero time difference
  memcpy(space->qold, space->q, sizeof(Real)*(size_t)nnode*nvars);
  memcpy(space->qoldm1, space->q, sizeof(Real)*(size_t)nnode*nvars);
  for(Int i = 0; i < nnode; i++){
    eqnset->NativeToConservative(&space->qold[i*nvars]);
    eqnset->NativeToConservative(&space->qoldm1[i*nvars]);
  }
  if(getenv("PCFD_UNSTEADY")){
    // unsteady fixture: a third (or later) time step of a BDF run -- q^n and q^{n-1} differ from q^{n+1} and from each
    // other, so that TemporalResidual (residual.tcc:125-179) and the cnp1 V / dt diagonal terms (jacobian.tcc:214-250)
    // contribute; iter > 2 selects the BDF2 coefficients when timeOrder = 2
    const Real twopi = 2.0*3.14159265358979323846;
    space->iter = (space->iter > 3) ? space->iter : 3;
    for(Int i = 0; i < nnode; i++){
      Real x = m->xyz[3*i + 0], y = m->xyz[3*i + 1], z = m->xyz[3*i + 2];
      for(Int j = 0; j < neqn; j++){
	space->qold[i*nvars + j] *= (1.0 - 0.010*sin(twopi*(x + 0.1*j))*cos(twopi*z));
	space->qoldm1[i*nvars + j] *= (1.0 - 0.017*cos(twopi*(y + 0.07*j))*sin(twopi*x));
      }
    }
  }
  if(mode == "dump"){
    Dump("qold", space->qold, (size_t)nnode*nvars);
    Dump("qoldm1", space->qoldm1, (size_t)nnode*nvars);
  }
  // NewtonIterate head (solutionSpace.tcc:662-665)
  UpdateBCs(space);
  p->UpdateGeneralVectors(space->q, nvars);
  // the BC update is an iterated map on the phantom states (10 sub-iterations
  // per call, compressible.tcc:1258,1383): dump the state before and after one
  // more call so the parity tests can replay exactly one UpdateBCs
  if(mode == "dump") Dump("q_pre", space->q, (size_t)(nnode+gnode+nbnode)*nvars);
  UpdateBCs(space);
  p->UpdateGeneralVectors(space->q, nvars);

#ifdef PCFD_DROPIN
  if(mode == "dump"){
    DumpMesh(space);
    Dump("q0", space->q, (size_t)(nnode+gnode+nbnode)*nvars);
    pcfd::DropIn<SolutionSpace<Real> > gpu(space, 0);
    // several ranks: halo maps from the mesh's ghost tables, CUDA-IPC connection of the ranks' device contexts; every
    // p->UpdateGeneralVectors of the CPU branch below becomes gpu.UpdateGeneralVectors(field) on device-resident data
    const bool multi = p->GetNp() > 1;
    if(multi) gpu.ConnectRanks();
    gpu.PushQ();

    Real dtmin = gpu.ComputeTimesteps();
    gpu.PullTimestep(space->GetFieldData("timestep", FIELDS::STATE_NONE));
    Dump("timestep", space->GetFieldData("timestep", FIELDS::STATE_NONE), (size_t)nnode);
    Dump("dtmin", &dtmin, 1);

    gpu.GradientCompute();
    if(multi) gpu.UpdateGeneralVectors(PCFD_F_QGRAD);        // gradient.tcc:98
    gpu.PullGradient();
    Dump("qgrad", space->qgrad, (size_t)(nnode+gnode)*nterms*3);

    if(param->limiter){
      gpu.LimiterCompute();
      if(multi) gpu.UpdateGeneralVectors(PCFD_F_LIMITER);    // limiters.tcc:128
      gpu.PullLimiter();
    }
    Dump("limiter", space->limiter->l, (size_t)(nnode+gnode)*neqn);

    std::vector<Real> res = multi ? gpu.ComputeResidualsParallel() : gpu.ComputeResiduals();
    gpu.PullB();
    Dump("b", space->crs->b, (size_t)nnode*neqn);
    Dump("resnorm", res.data(), res.size());

    if(param->nSgs > 0){
      gpu.ComputeJacobians();
      gpu.PullMatrix();
      CRSMatrix<Real>* A = space->crs->A;
      Dump("ia", A->ia, (size_t)nnode+1);
      Dump("ja", A->ja, (size_t)A->nblocks);
      Dump("iau", A->iau, (size_t)nnode);
      Dump("A", A->M, (size_t)A->nblocks*neqn*neqn);
      if(getenv("PCFD_TRANSPOSE")){
	// CRSMatrix::CRSTranspose on the device-resident matrix (ghost-column blocks through the host's MPI), and back
	gpu.CRSTranspose();
	gpu.PullMatrix();
	Dump("A_T", A->M, (size_t)A->nblocks*neqn*neqn);
	gpu.CRSTranspose();
      }
      gpu.PrepareSGS();
      gpu.PullMatrix();
      Dump("A_lu", A->M, (size_t)A->nblocks*neqn*neqn);
      Dump("pv", A->pv, (size_t)nnode*neqn);
      gpu.BlankX();
      Real ddq = 0.0;
      if(multi){
	// CRS::SGS across ranks (crs.tcc:88,146): a halo of x before the first and after every sweep
	gpu.UpdateGeneralVectors(PCFD_F_X);
	for(Int isgs = 0; isgs < param->nSgs; isgs++){
	  ddq = gpu.SGS(1);
	  gpu.UpdateGeneralVectors(PCFD_F_X);
	}
      }
      else ddq = gpu.SGS(param->nSgs);
      gpu.PullX();
      Dump("x", space->crs->x, (size_t)(nnode+gnode)*neqn);
      Dump("sgs_ddq", &ddq, 1);
      gpu.ApplyDQ();
    }
    else{
      gpu.ExplicitSolve();
      gpu.PullX();
      Dump("x", space->crs->x, (size_t)nnode*neqn);
    }
    if(multi) gpu.UpdateGeneralVectors(PCFD_F_Q);            // solutionSpace.tcc:857, on the device
    gpu.PullQ();
    if(!multi) p->UpdateGeneralVectors(space->q, nvars);
    Dump("q1", space->q, (size_t)(nnode+gnode+nbnode)*nvars);
    if(getenv("PCFD_FORCES") && !multi){
      // Forces::Compute (solutionSpace.tcc:884) on the device-resident state
      InjectForcesGeometry(space);
      Dump("forces_q", space->q, (size_t)(nnode+gnode+nbnode)*nvars);
      Dump("forces_qgrad", space->qgrad, (size_t)(nnode+gnode)*space->grad->GetNterms()*3);
      Dump("forces_cg", m->cg, (size_t)(nnode+gnode+nbnode)*3);
      gpu.ComputeForces();
      DumpForcesResult(space);
    }
  }
#else
  if(mode == "dump"){
    DumpMesh(space);
    Dump("q0", space->q, (size_t)(nnode+gnode+nbnode)*nvars);

    Real dtmin = ComputeTimesteps(space);
    Dump("timestep", space->GetFieldData("timestep", FIELDS::STATE_NONE), (size_t)nnode);
    Dump("dtmin", &dtmin, 1);

    space->grad->Compute();
    Dump("qgrad", space->qgrad, (size_t)(nnode+gnode)*nterms*3);

    if(param->limiter){
      space->limiter->Compute(space);
    }
    Dump("limiter", space->limiter->l, (size_t)(nnode+gnode)*neqn);

    std::vector<Real> res = ComputeResiduals(space);
    Dump("b", space->crs->b, (size_t)nnode*neqn);
    Dump("resnorm", res.data(), res.size());

    if(param->nSgs > 0){
      ComputeJacobians(space);
      CRSMatrix<Real>* A = space->crs->A;
      Dump("ia", A->ia, (size_t)nnode+1);
      Dump("ja", A->ja, (size_t)A->nblocks);
      Dump("iau", A->iau, (size_t)nnode);
      Dump("A", A->M, (size_t)A->nblocks*neqn*neqn);
      if(getenv("PCFD_GMRES")){
	// CRS::GMRES (crs.tcc:176-415) on the assembled system, BEFORE the diagonal is factored in place: restarted GMRES
	// with right preconditioning (PCFD_GMRES = the reference's precondType: 0 none, 1 diagonal, 2 block diagonal)
	Int ptype = atoi(getenv("PCFD_GMRES"));
	Int ndir = getenv("PCFD_GMRES_NDIR") ? atoi(getenv("PCFD_GMRES_NDIR")) : 10;
	Int nrest = getenv("PCFD_GMRES_RESTARTS") ? atoi(getenv("PCFD_GMRES_RESTARTS")) : 1;
	size_t nx = (size_t)(nnode+gnode)*neqn;
	Real* xg = new Real[nx];
	for(size_t k = 0; k < nx; k++) xg[k] = 0.0;
	Real dq = space->crs->GMRES(nrest, ndir, ptype, NULL, xg, NULL);
	Real cfg[3] = {(Real)ptype, (Real)ndir, (Real)nrest};
	Dump("gmres_x", xg, nx);
	Dump("gmres_dq", &dq, 1);
	Dump("gmres_cfg", cfg, 3);
	delete [] xg;
      }
      if(getenv("PCFD_TRANSPOSE")){
	// CRSMatrix::CRSTranspose (crsmatrix.tcc:568-599, with PObj::TransposeCommCRS across ranks) on the assembled matrix;
	// the matrix is put back afterwards so that every later dump is what it is without this block
	size_t nA = (size_t)A->nblocks*neqn*neqn;
	Real* keep = new Real[nA];
	memcpy(keep, A->M, sizeof(Real)*nA);
	A->CRSTranspose();
	Dump("A_T", A->M, nA);
	memcpy(A->M, keep, sizeof(Real)*nA);
	delete [] keep;
      }
      A->PrepareSGS();
      Dump("A_lu", A->M, (size_t)A->nblocks*neqn*neqn);
      Dump("pv", A->pv, (size_t)nnode*neqn);
      space->crs->BlankX();
      Real ddq = space->crs->SGS(param->nSgs, NULL, NULL, NULL);
      Dump("x", space->crs->x, (size_t)(nnode+gnode)*neqn);
      Dump("sgs_ddq", &ddq, 1);
      for(Int j = 0; j < nnode; j++){
	eqnset->ApplyDQ(&space->crs->x[j*neqn], &space->q[j*nvars], &m->xyz[j*3]);
      }
    }
    else{
      ExplicitSolve(*space);
      Dump("x", space->crs->x, (size_t)nnode*neqn);
      if(!eqnset->varsConservative){
	// ExplicitSolve leaves q untouched for native-variable eqnsets (solve.tcc:112-130); the update is applied by
	// the node loop of NewtonIterate (solutionSpace.tcc:797-812)
	for(Int j = 0; j < nnode; j++){
	  eqnset->ApplyDQ(&space->crs->x[j*neqn], &space->q[j*nvars], &m->xyz[j*3]);
	}
      }
    }
    p->UpdateGeneralVectors(space->q, nvars);
    Dump("q1", space->q, (size_t)(nnode+gnode+nbnode)*nvars);
    DumpTurbulence(space);
    DumpForces(space);
  }
#endif
  else if(mode == "time"){
    // CPU baseline: wall time of each reference phase (the same regions the
    // reference's own GradientTimer/ResidualTimer/JacobianAssembleTimer/
    // LinearSolveTimer bracket, solutionSpace.tcc:593-759), max over ranks.
    double tTs = 0, tGrad = 0, tLim = 0, tRes = 0, tJac = 0, tLU = 0, tSGS = 0, tUpd = 0;
    Real* qsave = new Real[(size_t)(nnode+gnode+nbnode)*nvars];
    memcpy(qsave, space->q, sizeof(Real)*(size_t)(nnode+gnode+nbnode)*nvars);
    for(Int r = 0; r < reps; r++){
      double t0;
      MPI_Barrier(MPI_COMM_WORLD);
      t0 = Now(); ComputeTimesteps(space); tTs += Now() - t0;
      t0 = Now(); UpdateBCs(space); p->UpdateGeneralVectors(space->q, nvars); tUpd += Now() - t0;
      t0 = Now(); space->grad->Compute(); tGrad += Now() - t0;
      t0 = Now(); if(param->limiter) space->limiter->Compute(space); tLim += Now() - t0;
      t0 = Now(); ComputeResiduals(space); tRes += Now() - t0;
      if(param->nSgs > 0){
	t0 = Now(); ComputeJacobians(space); tJac += Now() - t0;
	t0 = Now(); space->crs->A->PrepareSGS(); tLU += Now() - t0;
	space->crs->BlankX();
	t0 = Now(); space->crs->SGS(param->nSgs, NULL, NULL, NULL); tSGS += Now() - t0;
	t0 = Now();
	for(Int j = 0; j < nnode; j++){
	  eqnset->ApplyDQ(&space->crs->x[j*neqn], &space->q[j*nvars], &m->xyz[j*3]);
	}
	tUpd += Now() - t0;
      }
      else{
	t0 = Now(); ExplicitSolve(*space); tUpd += Now() - t0;
      }
      // keep every repetition on the same state
      memcpy(space->q, qsave, sizeof(Real)*(size_t)(nnode+gnode+nbnode)*nvars);
    }
    delete [] qsave;
    double t[8] = {tTs, tGrad, tLim, tRes, tJac, tLU, tSGS, tUpd};
    MPI_Allreduce(MPI_IN_PLACE, t, 8, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    Int counts[3] = {nnode, m->GetNumEdges(), m->GetNumParallelEdges()};
    MPI_Allreduce(MPI_IN_PLACE, counts, 3, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    if(g_rank == 0){
      std::ostringstream fn;
      fn << g_out << "/timing.json";
      std::ofstream f(fn.str().c_str());
      f << std::setprecision(9);
      // every cut edge is a half-edge on both ranks: count it once
      f << "{\"np\": " << p->GetNp() << ", \"reps\": " << reps
	<< ", \"nnode\": " << counts[0] << ", \"nedge\": " << counts[1] + counts[2]/2
	<< ", \"nsgs\": " << param->nSgs
	<< ", \"t_timestep\": " << t[0]/reps << ", \"t_gradient\": " << t[1]/reps
	<< ", \"t_limiter\": " << t[2]/reps << ", \"t_residual\": " << t[3]/reps
	<< ", \"t_jacobian\": " << t[4]/reps << ", \"t_lu\": " << t[5]/reps
	<< ", \"t_sgs\": " << t[6]/reps << ", \"t_update\": " << t[7]/reps << "}" << std::endl;
    }
  }
  else{
    std::cerr << "unknown mode " << mode << std::endl;
  }

  fflush(stdout);
  MPI_Finalize();
  // the reference's destructors close HDF5 files etc.; nothing there matters for the dumps
  _exit(0);
}
