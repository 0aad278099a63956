/*
 * pcfd_oracle_fr.c -- plain-C CPU restatement of ProteusCFD's hot path for the REACTING eqnset
 * (CompressibleFREqnSet, ucs/compressibleFR.tcc: native variables [rho_1..rho_ns, u, v, w, T], HLLC flux with
 * preconditioned wave speeds, finite-rate source term, dense temporal terms).  TEST INFRASTRUCTURE ONLY -- see
 * pcfd_oracle.h.  The edge / node loops are the reference's sequential loops (same summation order); every function
 * cites the reference lines it restates (paths relative to /root/reference/ucs).
 *
 * Parity status: PINNED against tests/golden/box5_fr_explicit.npz and box4_fr_implicit.npz, produced by running the
 * unmodified reference (oracle/_ref/ref_harness) on its own chemModels/5speciesAir.rxn.
 *
 * Layouts (ns = nspecies): neqn = ns+4, nvars = 3ns+6, nterms = 2ns+4
 *   q row  [rho_i (ns) | u v w | T | P | rho | cv_i (ns) | mol_i (ns)]       compressibleFR.tcc:14-31
 *   qgrad  terms = vars 0..ns+3 then the ns concentrations                   compressibleFR.tcc:693-711
 */
#include "pcfd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXS ORC_CHEM_MAX_SPECIES
#define MAXE (MAXS + 4)
#define MAXV (3*MAXS + 6)
#define MAXT (2*MAXS + 4)
#define UNIV_R 8.31447215   /* chem_constants.h:5 */

static double MAXD(double x, double y){ return (x > y) ? x : y; }
static double MIND(double x, double y){ return (x < y) ? x : y; }
static int is_ghost(const orc_case* c, int n){ return n >= c->nnode && n < c->nnode + c->gnode; }

/* ------------------------------------------------------------- thermodynamics */

/* species.tcc:96-138 (the pinned temperature is local to GetThermoCoeff: the polynomial sees the caller's T) */
static const double* thermo_coeff(const orc_fr_params* p, int sp, double T)
{
  if(T < 200.0) return p->chem->nasa7[sp][0];
  if(T > 6000.0) return p->chem->nasa7[sp][1];
  return (T > 1000.0) ? p->chem->nasa7[sp][1] : p->chem->nasa7[sp][0];
}
static double sp_R(const orc_fr_params* p, int i){ return UNIV_R/p->chem->mw[i]; }   /* species.tcc:315 */
/* species.tcc:43-53; GetdHdT :337-349 is the same polynomial */
static double sp_cp(const orc_fr_params* p, int i, double T)
{
  const double* a = thermo_coeff(p, i, T);
  double cp_R = a[0] + T*(a[1] + T*(a[2] + T*(a[3] + T*a[4])));
  return cp_R*sp_R(p, i);
}
/* species.tcc:55-71 (href == 0) */
static double sp_h(const orc_fr_params* p, int i, double T)
{
  const double* a = thermo_coeff(p, i, T);
  double h_R = a[5] + T*(a[0] + T*(a[1]/2.0 + T*(a[2]/3.0 + T*(a[3]/4.0 + T*a[4]/5.0))));
  double h = h_R*sp_R(p, i);
  h -= 0.0;
  return h;
}
/* chem.tcc:989-999 with IdealGasEOS::GetP (EOS.tcc:36-40) */
static double chem_P(const orc_fr_params* p, const double* rhoiDim, double T)
{
  int i, ns = p->chem->nspecies;
  double P = 0.0;
  for(i = 0; i < ns; i++) P += rhoiDim[i]*sp_R(p, i)*T;
  return P;
}
/* chem.tcc:586-595 */
static double chem_specific_enthalpy(const orc_fr_params* p, const double* X, double T)
{
  int i, ns = p->chem->nspecies;
  double h = 0.0;
  for(i = 0; i < ns; i++) h += sp_h(p, i, T)*X[i];
  return h;
}

/* CompressibleFREqnSet::ComputeAuxiliaryVariables compressibleFR.tcc:755-814 */
static void fr_aux(const orc_fr_params* p, double* Q)
{
  int i, ns = p->chem->nspecies;
  double* rho = &Q[ns+5];
  double* cvi = &Q[ns+6];
  double* mol = &Q[ns+ns+6];
  double T = Q[ns+3], rhoiDim[MAXS], TDim, pDim, s_ref;
  *rho = 0.0;
  for(i = 0; i < ns; i++){
    *rho += Q[i];
    rhoiDim[i] = Q[i]*p->ref_density;
  }
  TDim = T*p->ref_temperature;
  pDim = chem_P(p, rhoiDim, TDim);
  Q[ns+4] = pDim/p->ref_pressure;
  s_ref = (p->ref_velocity*p->ref_velocity/p->ref_temperature);
  for(i = 0; i < ns; i++){
    double cpiDim = sp_cp(p, i, TDim);
    cvi[i] = (cpiDim - sp_R(p, i))/s_ref;   /* IdealGasEOS::GetCv EOS.tcc:79-91 */
  }
  for(i = 0; i < ns; i++) mol[i] = Q[i]/p->chem->mw[i]/1000.0;   /* chem.tcc:980-986 */
}

/* CompressibleFREqnSet::GetFluidProperties compressibleFR.tcc:1523-1570 -> ChemModel::GetFluidProperties
   chem.tcc:545-572 with GetCp/GetCv :1002-1044 */
static void fr_fluid_props(const orc_fr_params* p, const double* rhoi, double T, double* cv, double* cp, double* R,
			   double* gamma, double* c2)
{
  int i, ns = p->chem->nspecies;
  double rhoiDim[MAXS], X[MAXS];
  double s_ref = (p->ref_velocity*p->ref_velocity/p->ref_temperature);
  double Tdim, RDim = 0.0, PDim = 0.0, rhoDim = 0.0, cpDim = 0.0, cvDim = 0.0, g, c2Dim;
  for(i = 0; i < ns; i++) rhoiDim[i] = rhoi[i]*p->ref_density;
  Tdim = T*p->ref_temperature;
  for(i = 0; i < ns; i++) rhoDim += rhoiDim[i];
  for(i = 0; i < ns; i++){
    double Rs = sp_R(p, i);
    RDim += rhoiDim[i]*Rs;
    PDim += rhoiDim[i]*Rs*Tdim;
  }
  RDim /= rhoDim;
  for(i = 0; i < ns; i++) X[i] = rhoiDim[i]/rhoDim;
  for(i = 0; i < ns; i++) cpDim += sp_cp(p, i, Tdim)*X[i];
  for(i = 0; i < ns; i++){
    double cpi = sp_cp(p, i, Tdim);
    cvDim += X[i]*(cpi - sp_R(p, i));
  }
  g = cpDim/cvDim;
  c2Dim = g*RDim*Tdim;
  (void)PDim;
  *gamma = (cpDim)/(cvDim);
  *cv = cvDim/s_ref;
  *cp = cpDim/s_ref;
  *R = RDim/s_ref;
  *c2 = c2Dim/(p->ref_velocity*p->ref_velocity);
}

/* compressibleFR.tcc:1249-1273 */
static double fr_total_enthalpy(const orc_fr_params* p, const double* Q)
{
  int i, ns = p->chem->nspecies;
  double X[MAXS], T = Q[ns+3], rho = Q[ns+5], u = Q[ns], v = Q[ns+1], w = Q[ns+2];
  double v2 = u*u + v*v + w*w, h;
  for(i = 0; i < ns; i++) X[i] = Q[i]/rho;
  h = chem_specific_enthalpy(p, X, T*p->ref_temperature)/p->ref_specific_enthalpy;
  return h*rho + 0.5*rho*v2;
}
/* compressibleFR.tcc:1219-1246 */
static double fr_total_energy(const orc_fr_params* p, const double* Q)
{
  int i, ns = p->chem->nspecies;
  double X[MAXS], T = Q[ns+3], P = Q[ns+4], rho = Q[ns+5], u = Q[ns], v = Q[ns+1], w = Q[ns+2];
  double v2 = u*u + v*v + w*w, h, E;
  for(i = 0; i < ns; i++) X[i] = Q[i]/rho;
  h = chem_specific_enthalpy(p, X, T*p->ref_temperature)/p->ref_specific_enthalpy;
  E = h*rho - P;
  return E + 0.5*rho*v2;
}
/* compressibleFR.tcc:1640-1644 */
static double fr_theta(int ns, const double* Q, const double* avec, double vdotn)
{
  return (Q[ns]*avec[0] + Q[ns+1]*avec[1] + Q[ns+2]*avec[2] + vdotn);
}

/* CompressibleFREqnSet::HLLCFlux compressibleFR.tcc:301-548 (RoeFlux :293-298 forwards here; symmetry2D off).
   The Roe-averaged state only feeds GetFluidProperties (rho_i, T), so its auxiliary variables are not needed. */
static int fr_hllc_flux(const orc_fr_params* p, const double* QL, const double* QR, const double* avec, double vdotn,
			double* flux, double beta)
{
  int i, ns = p->chem->nspecies;
  const double* rhoiL = &QL[0];
  double uL = QL[ns], vL = QL[ns+1], wL = QL[ns+2], TL = QL[ns+3], pL = QL[ns+4];
  double pgL = pL - p->Pref, rhoL = QL[ns+5];
  double cvL, cpL, RL, gammaL, c2L, HTL, ETL, thetaL;
  const double* rhoiR = &QR[0];
  double uR = QR[ns], vR = QR[ns+1], wR = QR[ns+2], TR = QR[ns+3], pR = QR[ns+4];
  double pgR = pR - p->Pref, rhoR = QR[ns+5];
  double cvR, cpR, RR, gammaR, c2R, HTR, ETR, thetaR;
  double rho, sigma, roeQ[MAXV], theta, cv, cp, R, gammaT, c2;
  double oneMBeta, thetaPrime, cPrime, thetaLPrime, thetaRPrime, cLPrime, cRPrime, eig5L, eig4R, eig4, eig5;
  double SL, SR, SM, Qflux[MAXE], pStar = 0.0, omega, const1, const2, area, Et, thetabar;
  fr_fluid_props(p, rhoiL, TL, &cvL, &cpL, &RL, &gammaL, &c2L);
  HTL = fr_total_enthalpy(p, QL);
  ETL = HTL - pL;
  thetaL = fr_theta(ns, QL, avec, vdotn);
  fr_fluid_props(p, rhoiR, TR, &cvR, &cpR, &RR, &gammaR, &c2R);
  HTR = fr_total_enthalpy(p, QR);
  ETR = HTR - pR;
  thetaR = fr_theta(ns, QR, avec, vdotn);

  rho = sqrt(rhoL*rhoR);
  sigma = rho/(rhoL + rho);
  for(i = 0; i < ns; i++) roeQ[i] = rhoiL[i] + sigma*(rhoiR[i] - rhoiL[i]);
  roeQ[ns+0] = uL + sigma*(uR - uL);
  roeQ[ns+1] = vL + sigma*(vR - vL);
  roeQ[ns+2] = wL + sigma*(wR - wL);
  roeQ[ns+3] = TL + sigma*(TR - TL);
  theta = fr_theta(ns, roeQ, avec, vdotn);
  fr_fluid_props(p, roeQ, roeQ[ns+3], &cv, &cp, &R, &gammaT, &c2);

  oneMBeta = 1.0 - beta;
  thetaPrime = theta*(1.0 + beta)*0.5;
  cPrime = 0.5*sqrt(theta*theta*(oneMBeta*oneMBeta) + 4.0*beta*c2);
  thetaLPrime = thetaL*(1.0 + beta)*0.5;
  thetaRPrime = thetaR*(1.0 + beta)*0.5;
  cLPrime = 0.5*sqrt(thetaL*thetaL*(oneMBeta*oneMBeta) + 4.0*beta*c2L);
  cRPrime = 0.5*sqrt(thetaR*thetaR*(oneMBeta*oneMBeta) + 4.0*beta*c2R);
  eig5L = thetaLPrime - cLPrime;
  eig4R = thetaRPrime + cRPrime;
  eig4 = thetaPrime + cPrime;
  eig5 = thetaPrime - cPrime;
  SL = MIND(eig5L, eig5);
  SR = MAXD(eig4R, eig4);
  SM = (pgR - pgL + rhoL*thetaL*(SL - thetaL) - rhoR*thetaR*(SR-thetaR))/
    (rhoL*(SL-thetaL) - rhoR*(SR-thetaR));

  if(SL >= 0.0){
    for(i = 0; i < ns; i++) Qflux[i] = rhoiL[i];
    Qflux[ns] = rhoL*uL; Qflux[ns+1] = rhoL*vL; Qflux[ns+2] = rhoL*wL; Qflux[ns+3] = ETL;
    pStar = pgL;
    SM = thetaL;
  }
  else if(SR <= 0.0){
    for(i = 0; i < ns; i++) Qflux[i] = rhoiR[i];
    Qflux[ns] = rhoR*uR; Qflux[ns+1] = rhoR*vR; Qflux[ns+2] = rhoR*wR; Qflux[ns+3] = ETR;
    pStar = pgR;
    SM = thetaR;
  }
  else if((SL <= 0.0) && (SM >= 0.0)){
    pStar = pgL + rhoL*(thetaL - SL)*(thetaL - SM);
    omega = 1.0/(SL - SM);
    const1 = SL - thetaL;
    const2 = pStar - pgL;
    for(i = 0; i < ns; i++) Qflux[i] = omega*const1*rhoiL[i];
    Qflux[ns] = omega*(const1*rhoL*uL + const2*avec[0]);
    Qflux[ns+1] = omega*(const1*rhoL*vL + const2*avec[1]);
    Qflux[ns+2] = omega*(const1*rhoL*wL + const2*avec[2]);
    Qflux[ns+3] = omega*(const1*ETL - pgL*thetaL + (pStar*SM)) + p->Pref*(SM - thetaL)*omega;
  }
  else if((SM <= 0.0) && (SR >= 0.0)){
    pStar = pgR + rhoR*(thetaR - SR)*(thetaR - SM);
    omega = 1.0/(SR - SM);
    const1 = SR - thetaR;
    const2 = pStar - pgR;
    for(i = 0; i < ns; i++) Qflux[i] = omega*const1*rhoiR[i];
    Qflux[ns] = omega*(const1*rhoR*uR + const2*avec[0]);
    Qflux[ns+1] = omega*(const1*rhoR*vR + const2*avec[1]);
    Qflux[ns+2] = omega*(const1*rhoR*wR + const2*avec[2]);
    Qflux[ns+3] = omega*(const1*ETR - pgR*thetaR + (pStar*SM)) + p->Pref*(SM-thetaR)*omega;
  }
  else{
    /* "HLLC: Should never be here" (NaN wave speeds): the reference soft-aborts and then reads the uninitialised
       Qflux; the flux is reported as NaN and kneecapped by the caller like any other NaN */
    for(i = 0; i < ns+4; i++) Qflux[i] = NAN;
  }
  area = avec[3];
  Et = Qflux[ns+3];
  theta = SM;
  thetabar = theta - vdotn;
  for(i = 0; i < ns; i++) flux[i] = area*Qflux[i]*theta;
  flux[ns] = area*(Qflux[ns]*theta + pStar*avec[0]);
  flux[ns+1] = area*(Qflux[ns+1]*theta + pStar*avec[1]);
  flux[ns+2] = area*(Qflux[ns+2]*theta + pStar*avec[2]);
  flux[ns+3] = area*(Et*theta + pStar*thetabar) + p->Pref*thetabar*area;
  return 0;
}

/* EqnSet::NumericalFlux / BoundaryFlux eqnset.tcc:21-90: NaN components are zeroed */
static void fr_numerical_flux(const orc_fr_params* p, const double* QL, const double* QR, const double* avec,
			      double vdotn, double* flux, double beta)
{
  int i, neqn = p->chem->nspecies + 4;
  fr_hllc_flux(p, QL, QR, avec, vdotn, flux, beta);
  for(i = 0; i < neqn; i++) if(isnan(flux[i])) flux[i] = 0.0;
}

/* compressibleFR.tcc:1603-1637 */
static double fr_max_eigenvalue(const orc_fr_params* p, const double* Q, const double* avec, double vdotn, double beta)
{
  int ns = p->chem->nspecies;
  double cv, cp, R, g, c2, theta, oneMBeta, thetaPrime, cPrime, eig4, eig5;
  fr_fluid_props(p, Q, Q[ns+3], &cv, &cp, &R, &g, &c2);
  theta = fr_theta(ns, Q, avec, vdotn);
  oneMBeta = 1.0 - beta;
  thetaPrime = theta*(1.0 + beta)*0.5;
  cPrime = 0.5*sqrt(theta*theta*(oneMBeta*oneMBeta) + 4.0*beta*c2);
  eig4 = thetaPrime + cPrime;
  eig5 = thetaPrime - cPrime;
  return MAXD(fabs(eig4), fabs(eig5));
}

/* eqnset.h:231-238 + compressibleFR.tcc:734-752 */
static void fr_extrapolate(double chi, int neqn, double* Qho, const double* Q, const double* dQedge,
			   const double* gradQ, const double* dx, const double* limiter)
{
  int i;
  for(i = 0; i < neqn; i++){
    double corr = (0.5*chi*dQedge[i] + (1.0 - chi)*(gradQ[i*3]*dx[0] + gradQ[i*3+1]*dx[1] + gradQ[i*3+2]*dx[2]));
    Qho[i] = Q[i] + corr*limiter[i];
  }
}

/* compressibleFR.tcc:714-731 */
static int fr_bad_extrapolation(const orc_fr_params* p, const double* Q)
{
  int i, ns = p->chem->nspecies;
  for(i = 0; i < ns; i++) if(Q[i] < 0.0) return 1;
  if(fr_total_energy(p, Q) <= 0.0) return 1;
  if(Q[ns+4] < 1.0e-10) return 1;
  if(Q[ns+3] < 1.0e-10) return 1;
  return 0;
}

/* ---------------------------------------------------------- linear algebra */

/* matrix.h:63-74 */
static void matvec(const double* a, const double* v, double* vout, int n)
{
  int i, j;
  for(i = 0; i < n; i++){
    vout[i] = a[i*n + 0]*v[0];
    for(j = 1; j < n; j++) vout[i] += a[i*n + j]*v[j];
  }
}
/* matrix.h:110-190 */
static int lu(double* a, int* p, int n)
{
  int i, j, k, row = 0, temp;
  double large;
  for(i = 0; i < n; i++) p[i] = i;
  for(i = 0; i < n; i++){
    large = 0.0;
    for(j = i; j < n; j++){
      if(fabs(a[p[j]*n + i]) > fabs(large)){ large = a[p[j]*n + i]; row = j; }
    }
    temp = p[i]; p[i] = p[row]; p[row] = temp;
    large = 1.0/large;
    for(j = i+1; j < n; j++) a[p[j]*n + i] *= large;
    for(j = i+1; j < n; j++){
      for(k = i+1; k < n; k++) a[p[j]*n + k] -= a[p[j]*n + i]*a[p[i]*n + k];
    }
  }
  return 0;
}
/* matrix.h:237-264 */
static void lu_solve(const double* a, double* b, const int* p, double* x, int n)
{
  int i, j;
  double sum;
  for(i = 0; i < n; i++){
    sum = 0.0;
    for(j = 0; j < i; j++) sum += a[p[i]*n + j]*x[j];
    x[i] = b[p[i]] - sum;
  }
  for(i = n-1; i >= 0; i--){
    sum = 0.0;
    for(j = n-1; j > i; j--) sum += a[p[i]*n + j]*b[j];
    b[i] = (x[i] - sum)/a[p[i]*n + i];
  }
}

/* ------------------------------------------------------ boundary conditions */

/* geometry.h:101-129 */
static void perp_vectors(const double* n, double* v1, double* v2)
{
  double dot, mag;
  v1[0] = v1[1] = v1[2] = 0.0;
  if(fabs(dot = n[0]) < 0.95) v1[0] = 1.0;
  else if(fabs(dot = n[1]) < 0.95) v1[1] = 1.0;
  else{ dot = n[2]; v1[2] = 1.0; }
  v1[0] -= dot*n[0];
  v1[1] -= dot*n[1];
  v1[2] -= dot*n[2];
  mag = sqrt(v1[0]*v1[0] + v1[1]*v1[1] + v1[2]*v1[2]);
  v1[0] = v1[0]/mag; v1[1] = v1[1]/mag; v1[2] = v1[2]/mag;
  v2[0] = n[1]*v1[2] - v1[1]*n[2];
  v2[1] = n[2]*v1[0] - v1[2]*n[0];
  v2[2] = n[0]*v1[1] - v1[0]*n[1];
  mag = sqrt(v2[0]*v2[0] + v2[1]*v2[1] + v2[2]*v2[2]);
  v2[0] = v2[0]/mag; v2[1] = v2[1]/mag; v2[2] = v2[2]/mag;
}

/* CompressibleFREqnSet::Eigensystem compressibleFR.tcc:150-290 (c2i is overwritten by the bulk c2, :176-180) */
static void fr_eigensystem(const orc_fr_params* p, const double* Q, const double* avec, double vdotn,
			   double* eigenvalues, double* T, double* Tinv, double beta)
{
  int i, ns = p->chem->nspecies, neqn = ns + 4;
  double nx = avec[0], ny = avec[1], nz = avec[2];
  double bm1 = beta - 1.0, bp1 = beta + 1.0, oneMBeta = 1.0 - beta;
  double theta = fr_theta(ns, Q, avec, vdotn);
  const double* rhoi = &Q[0];
  double rho = Q[ns+5];
  double cv, cp, R, g, c2, thetaPrime, cPrime, m[3], l[3], lx, ly, lz, mx, my, mz, betam, Xp, Xm;
  int uloc = ns, vloc = ns+1, wloc = ns+2, tloc = ns+3;
  fr_fluid_props(p, rhoi, Q[ns+3], &cv, &cp, &R, &g, &c2);
  thetaPrime = 0.5*bp1*theta;
  cPrime = 0.5*sqrt(theta*theta*(oneMBeta*oneMBeta) + 4.0*beta*c2);
  perp_vectors(avec, l, m);
  lx = l[0]; ly = l[1]; lz = l[2]; mx = m[0]; my = m[1]; mz = m[2];
  for(i = 0; i < ns+2; i++) eigenvalues[i] = theta;
  eigenvalues[wloc] = thetaPrime + cPrime;
  eigenvalues[tloc] = thetaPrime - cPrime;
  betam = oneMBeta*0.5;
  Xp = theta*betam + cPrime;
  Xm = theta*betam - cPrime;
  for(i = 0; i < neqn*neqn; i++) T[i] = 0.0;
  for(i = 0; i < ns; i++){
    T[neqn*i + i] = 1.0;
    T[neqn*i + wloc] = -(rhoi[i]*(c2 + bm1*c2 - theta*bm1*Xm))/(c2*Xm);
    T[neqn*i + tloc] =  (rhoi[i]*(c2 + bm1*c2 - theta*bm1*Xp))/(c2*Xp);
  }
  T[neqn*(ns+0) + (ns+0)] = lx; T[neqn*(ns+0) + (ns+1)] = mx; T[neqn*(ns+0) + (ns+2)] = nx; T[neqn*(ns+0) + (ns+3)] = -nx;
  T[neqn*(ns+1) + (ns+0)] = ly; T[neqn*(ns+1) + (ns+1)] = my; T[neqn*(ns+1) + (ns+2)] = ny; T[neqn*(ns+1) + (ns+3)] = -ny;
  T[neqn*(ns+2) + (ns+0)] = lz; T[neqn*(ns+2) + (ns+1)] = mz; T[neqn*(ns+2) + (ns+2)] = nz; T[neqn*(ns+2) + (ns+3)] = -nz;
  T[neqn*(ns+3) + wloc] = -rho*Xm;
  T[neqn*(ns+3) + tloc] = rho*Xp;
  for(i = 0; i < neqn*neqn; i++) Tinv[i] = 0.0;
  for(i = 0; i < ns; i++){
    double KK = -rhoi[i]*(c2*(Xm + Xp) - bm1*Xm*Xp*theta + bm1*c2*(Xm + Xp));
    Tinv[i*neqn + i] = 1.0;
    Tinv[i*neqn + uloc] = -((ly*mz-lz*my)*KK)/(c2*Xm*Xp);
    Tinv[i*neqn + vloc] =  ((lx*mz-lz*mx)*KK)/(c2*Xm*Xp);
    Tinv[i*neqn + wloc] = -((lx*my-ly*mx)*KK)/(c2*Xm*Xp);
    Tinv[i*neqn + tloc] =  (rhoi[i]*(c2 + bm1*c2))/(rho*c2*Xm*Xp);
  }
  Tinv[neqn*uloc + uloc] =  my*nz-mz*ny;
  Tinv[neqn*uloc + vloc] =-(mx*nz-mz*nx);
  Tinv[neqn*uloc + wloc] =  mx*ny-my*nx;
  Tinv[neqn*uloc + tloc] =  0.0;
  Tinv[neqn*vloc + uloc] =-(ly*nz-lz*ny);
  Tinv[neqn*vloc + vloc] =  lx*nz-lz*nx;
  Tinv[neqn*vloc + wloc] =-(lx*ny-ly*nx);
  Tinv[neqn*vloc + tloc] =  0.0;
  Tinv[neqn*wloc + uloc] =  ((Xp)*(ly*mz-lz*my))/(2.0*cPrime);
  Tinv[neqn*wloc + vloc] = -((Xp)*(lx*mz-lz*mx))/(2.0*cPrime);
  Tinv[neqn*wloc + wloc] =  ((Xp)*(lx*my-ly*mx))/(2.0*cPrime);
  Tinv[neqn*wloc + tloc] =  1.0/(2.0*rho*cPrime);
  Tinv[neqn*tloc + uloc] =  ((Xm)*(ly*mz-lz*my))/(2.0*cPrime);
  Tinv[neqn*tloc + vloc] = -((Xm)*(lx*mz-lz*mx))/(2.0*cPrime);
  Tinv[neqn*tloc + wloc] =  ((Xm)*(lx*my-ly*mx))/(2.0*cPrime);
  Tinv[neqn*tloc + tloc] =  1.0/(2.0*rho*cPrime);
}

/* compressibleFR.tcc:2343-2378 */
static double fr_newton_T_given_P(const orc_fr_params* p, const double* rhoi, double Pgoal, double Tinit)
{
  int i, j, ns = p->chem->nspecies, maxit = 28;
  double tol = 1.0e-15, TDim = Tinit*p->ref_temperature, PgoalDim = Pgoal*p->ref_pressure, rhoiDim[MAXS], dT = 0.0;
  for(i = 0; i < ns; i++) rhoiDim[i] = rhoi[i]*p->ref_density;
  for(j = 0; j < maxit; j++){
    double TpDim = TDim + 1.0e-8;
    double PDim = chem_P(p, rhoiDim, TDim);
    double PpDim = chem_P(p, rhoiDim, TpDim);
    double zpoint = PgoalDim - PDim;
    double zpointp = PgoalDim - PpDim;
    double dzdT = (zpointp - zpoint)/(TpDim - TDim);
    dT = -zpoint/dzdT;
    if(fabs(dT/p->ref_temperature) < tol) break;
    else TDim += dT;
  }
  return TDim/p->ref_temperature;
}

#define N_SUBIT 10   /* compressibleFR.tcc:32 */

/* compressibleFR.tcc:940-1039 */
static void fr_farfield_bc(const orc_case* c, const orc_fr_params* p, const double* QL, double* QR, const double* Qinf,
			   const double* avec, double vdotn, double beta)
{
  int i, j, subit, ns = p->chem->nspecies, neqn = ns + 4, nvars = 3*ns + 6;
  double qavg[MAXV], eig[MAXE], Tinv[MAXE*MAXE], T[MAXE*MAXE], rhs[MAXE], ql[MAXE], qinf[MAXE];
  for(subit = 0; subit < N_SUBIT; subit++){
    for(i = 0; i < neqn; i++) qavg[i] = 0.5*(QL[i] + QR[i]);
    fr_aux(p, qavg);
    fr_eigensystem(p, qavg, avec, vdotn, eig, T, Tinv, beta);
    if(c->no_cvbc){
      if(eig[0] >= 0.0) memcpy(QR, QL, nvars*sizeof(double));
      else memcpy(QR, Qinf, nvars*sizeof(double));
    }
    else{
      double Tguess, pgoal;
      memcpy(ql, QL, sizeof(double)*neqn);
      memcpy(qinf, Qinf, sizeof(double)*neqn);
      Tguess = ql[neqn-1];
      ql[neqn-1] = QL[ns+4];
      qinf[neqn-1] = Qinf[ns+4];
      for(i = 0; i < neqn; i++){
	rhs[i] = 0.0;
	for(j = 0; j < neqn; j++) rhs[i] += Tinv[i*neqn + j]*(eig[i] >= 0.0 ? ql[j] : qinf[j]);
      }
      matvec(T, rhs, QR, neqn);
      for(i = 0; i < ns; i++) if(QR[i] < 0.0) QR[i] = 0.0;
      pgoal = QR[neqn-1];
      QR[neqn-1] = fr_newton_T_given_P(p, QR, pgoal, Tguess);
    }
  }
}

/* compressibleFR.tcc:1042-1134 */
static void fr_inviscid_wall_bc(const orc_case* c, const orc_fr_params* p, const double* QL, double* QR,
				const double* avec, double vdotn, double beta)
{
  int i, j, subit, ns = p->chem->nspecies, neqn = ns + 4, nvars = 3*ns + 6;
  if(!c->no_cvbc){
    double qavg[MAXV], eig[MAXE], Tinv[MAXE*MAXE], T[MAXE*MAXE], rhs[MAXE], scr[MAXE], ql[MAXE];
    int pv[MAXE];
    for(subit = 0; subit < N_SUBIT; subit++){
      double Tguess, pgoal;
      for(i = 0; i < neqn; i++) qavg[i] = 0.5*(QL[i] + QR[i]);
      fr_aux(p, qavg);
      fr_eigensystem(p, qavg, avec, vdotn, eig, T, Tinv, beta);
      memcpy(ql, QL, sizeof(double)*neqn);
      Tguess = ql[neqn-1];
      ql[neqn-1] = QL[ns+4];
      for(i = 0; i < neqn; i++){
	rhs[i] = 0.0;
	for(j = 0; j < neqn; j++) rhs[i] += Tinv[i*neqn + j]*ql[j];
      }
      for(i = 0; i < ns; i++) Tinv[(neqn-1)*neqn + i] = 0.0;
      Tinv[(neqn-1)*neqn + ns] = avec[0];
      Tinv[(neqn-1)*neqn + ns+1] = avec[1];
      Tinv[(neqn-1)*neqn + ns+2] = avec[2];
      Tinv[(neqn-1)*neqn + ns+3] = 0.0;
      rhs[neqn-1] = vdotn;
      lu(Tinv, pv, neqn);
      lu_solve(Tinv, rhs, pv, scr, neqn);
      memcpy(QR, rhs, neqn*sizeof(double));
      pgoal = QR[neqn-1];
      QR[neqn-1] = fr_newton_T_given_P(p, QR, pgoal, Tguess);
    }
  }
  else{
    double QLmod[MAXV], dot;
    memcpy(QLmod, QL, sizeof(double)*nvars);
    QLmod[ns] += vdotn*avec[0];
    QLmod[ns+1] += vdotn*avec[1];
    QLmod[ns+2] += vdotn*avec[2];
    fr_aux(p, QLmod);
    for(i = 0; i < neqn; i++) QR[i] = QLmod[i];
    dot = 2.0*(QLmod[ns]*avec[0] + QLmod[ns+1]*avec[1] + QLmod[ns+2]*avec[2]);   /* MirrorVector geometry.h:469-476 */
    QR[ns] = QLmod[ns] - dot*avec[0];
    QR[ns+1] = QLmod[ns+1] - dot*avec[1];
    QR[ns+2] = QLmod[ns+2] - dot*avec[2];
  }
}

/* bc.tcc:1058-1397 for the BC types of the reacting configs; the reference passes bcobj->GetQref == Qinf */
/* CompressibleFREqnSet::GetViscousWallBoundaryVariables compressibleFR.tcc:2048-2070 */
static void fr_viscous_wall_bc(int ns, double* QL, double* QR, const double* vel, const double* normalQ, double Twall)
{
  int i;
  if(Twall < 0.0){
    for(i = 0; i < ns; i++) QR[i] = QL[i];
    QR[ns+3] = QL[ns+3] = normalQ[ns+3];
  }
  else{
    for(i = 0; i < ns; i++) QR[i] = QL[i];
    QR[ns+3] = QL[ns+3] = Twall;
  }
  QR[ns] = QL[ns] = vel[0];
  QR[ns+1] = QL[ns+1] = vel[1];
  QR[ns+2] = QL[ns+2] = vel[2];
}

/* bcobj->twall / ref_temperature of the half-edge's surface (bc.tcc:1283-1286); default bcobj.tcc:31 */
static double fr_wall_temperature(const orc_case* c, const orc_fr_params* p, int e)
{
  return c->bedges_twall ? c->bedges_twall[e] : 1.0/p->ref_temperature;
}

/* e, q: the half-edge and the whole state array (the no-slip wall reads its most-normal neighbour, bc.tcc:1182-1206) */
/* qref: see boundary_variables in pcfd_oracle.c (one free-stream copy per half-edge in Bkernel_NumJac, scaled in place by
   the FarFieldViscous branch at every re-evaluation; NULL: a fresh copy, as in BC_Kernel) */
static void fr_boundary_variables(const orc_case* c, const orc_fr_params* p, double* QL, double* QR, const double* avec,
				  int bctype, double betaL, int e, const double* q, double* qref)
{
  int i, ns = p->chem->nspecies, neqn = ns + 4, nvars = 3*ns + 6;
  double vdotn = 0.0;
  switch(bctype){
  case ORC_BC_PARALLEL: return;
  case ORC_BC_SONIC_INFLOW: case ORC_BC_DIRICHLET:
    for(i = 0; i < nvars; i++) QR[i] = QL[i] = p->qinf[i];
    break;
  case ORC_BC_SONIC_OUTFLOW: case ORC_BC_NEUMANN:
    for(i = 0; i < neqn; i++) QR[i] = QL[i];
    break;
  case ORC_BC_FARFIELD:
    fr_farfield_bc(c, p, QL, QR, p->qinf, avec, vdotn, betaL);
    break;
  case ORC_BC_FARFIELD_VISCOUS: {   /* bc.tcc:1092-1108; GetMomentumLocation() == nspecies */
    double fresh[MAXV], ubar = orc_power_law_u(1.0, c->walldist[c->bedges_n[2*e]], c->Re);
    double* Qinf = qref ? qref : fresh;
    if(!qref) for(i = 0; i < nvars; i++) fresh[i] = p->qinf[i];
    if(ubar < 1.0){
      for(i = 0; i < 3; i++) Qinf[ns+i] = ubar*Qinf[ns+i];
    }
    fr_farfield_bc(c, p, QL, QR, Qinf, avec, vdotn, betaL);
    break;
  }
  case ORC_BC_IMPERMEABLE_WALL: case ORC_BC_SYMMETRY:
    fr_inviscid_wall_bc(c, p, QL, QR, avec, vdotn, betaL);
    break;
  case ORC_BC_NOSLIP: {   /* bc.tcc:1182-1291, static wall */
    double vel[3] = {0.0, 0.0, 0.0}, nQ[MAXV];
    int normalNode = orc_normal_node(c, e);
    vel[0] += 0.0; vel[1] += 0.0; vel[2] += 0.0;   /* velw */
    for(i = 0; i < nvars; i++) nQ[i] = q[(size_t)normalNode*nvars + i];
    fr_viscous_wall_bc(ns, QL, QR, vel, nQ, fr_wall_temperature(c, p, e));
    break;
  }
  default: break;
  }
  fr_aux(p, QR);
  fr_aux(p, QL);
}

/* bc.tcc:1399-1457 */
void orc_fr_update_bcs(const orc_case* c, const orc_fr_params* p, double* q, const double* beta)
{
  int e, nb = c->nbedge + c->ngedge, nvars = 3*p->chem->nspecies + 6;
  for(e = 0; e < nb; e++){
    int l = c->bedges_n[2*e], r = c->bedges_n[2*e+1];
    fr_boundary_variables(c, p, &q[(size_t)l*nvars], &q[(size_t)r*nvars], &c->bedges_a[4*e], c->bedges_bctype[e], beta[l], e, q, NULL);
  }
}

/* ------------------------------------------------------- gradient / limiter */

/* gradient.tcc:141-168 */
static void lsq_weights(const double* s, const double* dxbar, double* we)
{
  double r11 = s[0], r12 = s[1], r13 = s[2], s22 = s[3], s23 = s[4], s33 = s[5];
  double r12_r11 = (r11 == 0.0) ? 0.0 : r12/r11;
  double r22 = s22 - r12*r12_r11;
  double r23 = s23 - r12_r11*r13;
  double r13_r11 = (r11 == 0.0) ? 0.0 : r13/r11;
  double r23_r22 = (r22 == 0.0) ? 0.0 : r23/r22;
  double r33 = s33 - r13*r13_r11 - r23*r23_r22;
  double dykdx = (dxbar[1] - (r12_r11)*dxbar[0]);
  we[2] = (r33 == 0.0) ? 0.0 : (dxbar[2] - r13_r11*dxbar[0] - r23_r22*dykdx)/r33;
  we[1] = (r22 == 0.0) ? 0.0 : (dykdx - r23*we[2])/r22;
  we[0] = (r11 == 0.0) ? 0.0 : (dxbar[0] - r12*we[1] - r13*we[2])/r11;
}

/* GetGradientsLocation compressibleFR.tcc:693-711 */
static int fr_gradloc(int ns, int i){ return (i < ns + 4) ? i : (ns + ns + 6 + (i - (ns + 4))); }

/* gradient.tcc:57-112 (weighted LSQ), kernels :251-378, symmetry fix :545-565 */
void orc_fr_gradient(const orc_case* c, const orc_fr_params* p, const double* q, const double* sw, double* qgrad)
{
  int e, i, j, ns = p->chem->nspecies, nvars = 3*ns + 6, nterms = 2*ns + 4;
  int nb = c->nbedge + c->ngedge;
  size_t k, ntot = (size_t)(c->nnode + c->gnode)*nterms*3;
  for(k = 0; k < ntot; k++) qgrad[k] = 0.0;
  if(c->grad_type == 1){
    /* Green-Gauss: gradient.tcc:170-248 and the division by the dual volume (:83-89) */
    for(e = 0; e < c->nedge; e++){
      int l = c->edges_n[2*e], r = c->edges_n[2*e+1];
      const double* avec = &c->edges_a[4*e];
      double area = avec[3];
      for(i = 0; i < nterms; i++){
	int loc = fr_gradloc(ns, i);
	double faceavg = 0.5*(q[(size_t)l*nvars + loc] + q[(size_t)r*nvars + loc]);
	for(j = 0; j < 3; j++){
	  qgrad[(size_t)r*nterms*3 + 3*i + j] += -faceavg*avec[j]*area;
	  qgrad[(size_t)l*nterms*3 + 3*i + j] += faceavg*avec[j]*area;
	}
      }
    }
    for(e = 0; e < nb; e++){
      int l = c->bedges_n[2*e], r = c->bedges_n[2*e+1];
      const double* avec = &c->bedges_a[4*e];
      double area = avec[3];
      for(i = 0; i < nterms; i++){
	int loc = fr_gradloc(ns, i);
	double faceavg = 0.5*(q[(size_t)l*nvars + loc] + q[(size_t)r*nvars + loc]);
	for(j = 0; j < 3; j++) qgrad[(size_t)l*nterms*3 + 3*i + j] += faceavg*avec[j]*area;
      }
    }
    for(i = 0; i < c->nnode; i++) for(j = 0; j < nterms*3; j++) qgrad[(size_t)i*nterms*3 + j] /= c->vol[i];
  }
  else
  for(e = 0; e < c->nedge + nb; e++){
    int interior = e < c->nedge;
    int l = interior ? c->edges_n[2*e] : c->bedges_n[2*(e - c->nedge)];
    int r = interior ? c->edges_n[2*e+1] : c->bedges_n[2*(e - c->nedge)+1];
    double dx[3], weL[3], weR[3], dx2, weight, dq;
    const double *qL, *qR;
    if(!interior && !is_ghost(c, r)) continue;
    dx[0] = c->xyz[3*l] - c->xyz[3*r];
    dx[1] = c->xyz[3*l+1] - c->xyz[3*r+1];
    dx[2] = c->xyz[3*l+2] - c->xyz[3*r+2];
    qL = &q[(size_t)l*nvars]; qR = &q[(size_t)r*nvars];
    dx2 = dx[0]*dx[0] + dx[1]*dx[1] + dx[2]*dx[2];
    weight = 1.0/sqrt(dx2);
    dx[0] *= weight; dx[1] *= weight; dx[2] *= weight;
    lsq_weights(&sw[6*l], dx, weL);
    if(interior){
      dx[0] = -dx[0]; dx[1] = -dx[1]; dx[2] = -dx[2];
      lsq_weights(&sw[6*r], dx, weR);
      for(i = 0; i < nterms; i++){
	dq = weight*(qR[fr_gradloc(ns, i)] - qL[fr_gradloc(ns, i)]);
	for(j = 0; j < 3; j++) qgrad[(size_t)r*nterms*3 + 3*i + j] += +weR[j]*dq;
      }
    }
    for(i = 0; i < nterms; i++){
      dq = weight*(qR[fr_gradloc(ns, i)] - qL[fr_gradloc(ns, i)]);
      for(j = 0; j < 3; j++) qgrad[(size_t)l*nterms*3 + 3*i + j] += -weL[j]*dq;
    }
  }
  for(e = 0; e < nb; e++){
    if(c->bedges_bctype[e] == ORC_BC_SYMMETRY){
      int l = c->bedges_n[2*e];
      const double* avec = &c->bedges_a[4*e];
      double* g = &qgrad[(size_t)l*nterms*3];
      for(i = 0; i < nterms; i++){
	double dot = g[i*3]*avec[0] + g[i*3+1]*avec[1] + g[i*3+2]*avec[2];
	for(j = 0; j < 3; j++) g[i*3 + j] -= dot*avec[j];
      }
    }
  }
}

static double limiter_fn(int type, double temp)
{
  if(type == 1){             /* Barth, limiters.tcc:263-266 */
    temp = MAXD(0.0, temp);
    temp = MIND(1.0, temp);
    return temp;
  }
  return (temp*temp + 2.0*temp)/(temp*temp + temp + 2.0);   /* Venkatakrishnan :440 */
}

static void fr_limit_side(const orc_case* c, int neqn, int nvars, int nterms, const double* q, const double* qgrad,
			  const double* qmin, const double* qmax, double* lim, int me, int other, int variant)
{
  int j;
  double QL[MAXE], dQ[MAXE], dx[3], ones[MAXE];
  const double* qL = &q[(size_t)me*nvars];
  const double* qR = &q[(size_t)other*nvars];
  for(j = 0; j < neqn; j++){ dQ[j] = qR[j] - qL[j]; ones[j] = 1.0; }
  dx[0] = 0.5*(c->xyz[3*other] - c->xyz[3*me]);
  dx[1] = 0.5*(c->xyz[3*other+1] - c->xyz[3*me+1]);
  dx[2] = 0.5*(c->xyz[3*other+2] - c->xyz[3*me+2]);
  fr_extrapolate(c->chi, neqn, QL, qL, dQ, &qgrad[(size_t)me*nterms*3], dx, ones);
  if(c->limiter == 3){
    /* Kernel_VenkatMod / Bkernel_VenkatMod (limiters.tcc:534-735); variant 0 left, 2 right, 1 ghost half-edge; see
       pcfd_oracle.c for the DP == unset case */
    const double Pi = 3.141592653589793;
    double L3 = 6.0*Pi*c->vol[me], K3 = 1.0*1.0*1.0;
    for(j = 0; j < neqn; j++){
      double DM = QL[j] - qL[j], DP = 0.0, ep2, temp;
      if(variant == 1){
	if(QL[j] > qL[j]) DP = qmax[(size_t)me*neqn + j] - QL[j];
	else if(QL[j] <= qL[j]) DP = qmin[(size_t)me*neqn + j] - QL[j];
      }
      else if(variant == 2){
	if(QL[j] > qL[j]) DP = qmax[(size_t)me*neqn + j] - qL[j];
	if(QL[j] <= qL[j]) DP = qmin[(size_t)me*neqn + j] - qL[j];
      }
      else{
	if(QL[j] > qL[j]) DP = qmax[(size_t)me*neqn + j] - qL[j];
	else if(QL[j] < qL[j]) DP = qmin[(size_t)me*neqn + j] - qL[j];
      }
      ep2 = L3*K3;
      temp = (DP*DP + ep2 + 2.0*DM*DP)/(DP*DP + 2.0*DM*DM + DM*DP + ep2);
      lim[(size_t)me*neqn + j] = MIND(lim[(size_t)me*neqn + j], temp);
    }
    return;
  }
  for(j = 0; j < neqn; j++){
    double temp = 1.0;
    if(QL[j] > qL[j]) temp = (qmax[(size_t)me*neqn + j] - qL[j])/(QL[j] - qL[j]);
    else if(QL[j] < qL[j]) temp = (qmin[(size_t)me*neqn + j] - qL[j])/(QL[j] - qL[j]);
    temp = limiter_fn(c->limiter, temp);
    lim[(size_t)me*neqn + j] = MIND(lim[(size_t)me*neqn + j], temp);
  }
}

/* limiters.tcc:53-132; the pressure clip :737-815 has no Roe-state test here (EqnSet::RoeVariables returns false,
   eqnset.h:47-51) */
void orc_fr_limiter(const orc_case* c, const orc_fr_params* p, const double* q, const double* qgrad, double* lim)
{
  int e, i, j, ns = p->chem->nspecies, neqn = ns + 4, nvars = 3*ns + 6, nterms = 2*ns + 4;
  int nnode = c->nnode, nb = c->nbedge + c->ngedge;
  double* qmin = (double*)calloc((size_t)nnode*neqn, sizeof(double));
  double* qmax = (double*)calloc((size_t)nnode*neqn, sizeof(double));
  for(i = 0; i < (nnode + c->gnode)*neqn; i++) lim[i] = 1.0;
  for(e = 0; e < c->nedge; e++){
    int l = c->edges_n[2*e], r = c->edges_n[2*e+1];
    for(j = 0; j < neqn; j++){
      qmax[(size_t)l*neqn+j] = MAXD(qmax[(size_t)l*neqn+j], q[(size_t)r*nvars+j]);
      qmin[(size_t)l*neqn+j] = MIND(qmin[(size_t)l*neqn+j], q[(size_t)r*nvars+j]);
    }
    for(j = 0; j < neqn; j++){
      qmax[(size_t)r*neqn+j] = MAXD(qmax[(size_t)r*neqn+j], q[(size_t)l*nvars+j]);
      qmin[(size_t)r*neqn+j] = MIND(qmin[(size_t)r*neqn+j], q[(size_t)l*nvars+j]);
    }
  }
  for(e = 0; e < nb; e++){
    int l = c->bedges_n[2*e], r = c->bedges_n[2*e+1];
    if(!is_ghost(c, r)) continue;
    for(j = 0; j < neqn; j++){
      qmax[(size_t)l*neqn+j] = MAXD(qmax[(size_t)l*neqn+j], q[(size_t)r*nvars+j]);
      qmin[(size_t)l*neqn+j] = MIND(qmin[(size_t)l*neqn+j], q[(size_t)r*nvars+j]);
    }
  }
  if(c->limiter == 1 || c->limiter == 2 || c->limiter == 3){
    for(e = 0; e < c->nedge; e++){
      int l = c->edges_n[2*e], r = c->edges_n[2*e+1];
      fr_limit_side(c, neqn, nvars, nterms, q, qgrad, qmin, qmax, lim, l, r, 0);
      fr_limit_side(c, neqn, nvars, nterms, q, qgrad, qmin, qmax, lim, r, l, 2);
    }
    for(e = 0; e < nb; e++){
      int l = c->bedges_n[2*e], r = c->bedges_n[2*e+1];
      if(!is_ghost(c, r)) continue;
      fr_limit_side(c, neqn, nvars, nterms, q, qgrad, qmin, qmax, lim, l, r, 1);
    }
  }
  if(c->limiter != 0){
    for(e = 0; e < c->nedge; e++){
      int l = c->edges_n[2*e], r = c->edges_n[2*e+1];
      double QL[MAXV], QR[MAXV], dQ[MAXE], dx[3];
      const double* qL = &q[(size_t)l*nvars];
      const double* qR = &q[(size_t)r*nvars];
      dx[0] = 0.5*(c->xyz[3*r] - c->xyz[3*l]);
      dx[1] = 0.5*(c->xyz[3*r+1] - c->xyz[3*l+1]);
      dx[2] = 0.5*(c->xyz[3*r+2] - c->xyz[3*l+2]);
      for(j = 0; j < neqn; j++) dQ[j] = qR[j] - qL[j];
      fr_extrapolate(c->chi, neqn, QL, qL, dQ, &qgrad[(size_t)l*nterms*3], dx, &lim[(size_t)l*neqn]);
      fr_aux(p, QL);
      if(fr_bad_extrapolation(p, QL)) for(i = 0; i < neqn; i++) lim[(size_t)l*neqn+i] = 0.0;
      dx[0] = -dx[0]; dx[1] = -dx[1]; dx[2] = -dx[2];
      for(j = 0; j < neqn; j++) dQ[j] = -dQ[j];
      fr_extrapolate(c->chi, neqn, QR, qR, dQ, &qgrad[(size_t)r*nterms*3], dx, &lim[(size_t)r*neqn]);
      fr_aux(p, QR);
      if(fr_bad_extrapolation(p, QR)) for(i = 0; i < neqn; i++) lim[(size_t)r*neqn+i] = 0.0;
    }
    for(i = 0; i < (nnode + c->gnode)*neqn; i++) if(lim[i] < 0.0) lim[i] = 0.0;
  }
  free(qmin); free(qmax);
}

/* ------------------------------------------------------------- transport + viscous terms (compressibleNSFR) */

static double chem_dRmixdRhoi(const orc_fr_params* p, const double* rhoi, double rho, int i);
static void fr_native_to_conservative(const orc_fr_params* p, double* Q);

/* Species::GetViscosity / GetThermalConductivity species.tcc:393-479: Sutherland (White) up to the transition
   temperature, NASA RP-1311 fit above; the range search keeps the LAST range containing T (no break) */
static double sp_transport(const double white[4], const double fit[3][6], int nfit, double T, double convFact)
{
  int i, range = -1;
  if(T <= white[3]){
    double v0 = white[0], T0 = white[1], S = white[2];
    return v0*(pow(T/T0, 1.5))*((T0 + S)/(T + S));
  }
  for(i = 0; i < nfit; i++){
    if(T >= fit[i][0] && T <= fit[i][1]) range = i;
  }
  if(range == -1) return NAN;      /* the reference aborts */
  {
    double A = fit[range][2], B = fit[range][3], Cc = fit[range][4], D = fit[range][5];
    double logv = A*log(T) + B/T + Cc/(T*T) + D;
    return exp(logv)*convFact;
  }
}
static double sp_visc(const orc_fr_params* p, int i, double T)
{ return sp_transport(p->transport->mu_white[i], p->transport->mu_fit[i], p->transport->nmu[i], T, 1.0e-7); }
static double sp_cond(const orc_fr_params* p, int i, double T)
{ return sp_transport(p->transport->k_white[i], p->transport->k_fit[i], p->transport->nk[i], T, 0.0001); }

/* ChemModel::WilkesMixtureRule chem.tcc:876-916 with MassFractionToMoleFraction :941-958.  Note the species
   VISCOSITIES weight the mixing of either property. */
static double wilke(const orc_fr_params* p, const double* rhoi, const double* property, double T)
{
  int i, j, ns = p->chem->nspecies;
  double mixtureProp = 0.0, sqrt8 = sqrt(8.0), rho = 0.0, summ = 0.0;
  double massfrac[MAXS], molefrac[MAXS], visc[MAXS];
  for(i = 0; i < ns; i++) rho += rhoi[i];
  for(i = 0; i < ns; i++) massfrac[i] = rhoi[i]/rho;
  for(i = 0; i < ns; i++){
    molefrac[i] = (massfrac[i])/p->chem->mw[i];
    summ += molefrac[i];
  }
  for(i = 0; i < ns; i++) molefrac[i] /= summ;
  summ = 0.0;
  for(i = 0; i < ns - 1; i++) summ += molefrac[i];
  molefrac[ns-1] = 1.0 - summ;
  for(i = 0; i < ns; i++) visc[i] = sp_visc(p, i, T);
  for(i = 0; i < ns; i++){
    double wi = 0.0, fraci = molefrac[i], MWi = p->chem->mw[i];
    for(j = 0; j < ns; j++){
      double fracj = molefrac[j], MWj = p->chem->mw[j];
      double temp = (1.0 + sqrt(visc[i]/visc[j])*pow(MWj/MWi, 0.25));
      double phi = pow((1.0 + MWi/MWj), -0.5)*temp*temp/sqrt8;
      wi += fracj*phi;
    }
    mixtureProp += (fraci/wi)*property[i];
  }
  return mixtureProp;
}

double orc_fr_mixture_viscosity(const orc_fr_params* p, const double* rhoi, double T)
{
  int i, ns = p->chem->nspecies;
  double visc[MAXS];
  for(i = 0; i < ns; i++) visc[i] = sp_visc(p, i, T);
  return wilke(p, rhoi, visc, T);
}

double orc_fr_mixture_conductivity(const orc_fr_params* p, const double* rhoi, double T)
{
  int i, ns = p->chem->nspecies;
  double k[MAXS];
  for(i = 0; i < ns; i++) k[i] = sp_cond(p, i, T);
  return wilke(p, rhoi, k, T);
}

/* CompressibleFREqnSet::GetMolecularViscosity / GetThermalConductivity compressibleFR.tcc:1580-1599 */
static double fr_molecular_viscosity(const orc_fr_params* p, const double* rhoi, double T)
{
  int i, ns = p->chem->nspecies;
  double rhoidim[MAXS];
  for(i = 0; i < ns; i++) rhoidim[i] = rhoi[i]*p->ref_density;
  return orc_fr_mixture_viscosity(p, rhoidim, T*p->ref_temperature)/p->ref_viscosity;
}
static double fr_thermal_conductivity(const orc_fr_params* p, const double* rhoi, double T)
{
  int i, ns = p->chem->nspecies;
  double rhoidim[MAXS];
  for(i = 0; i < ns; i++) rhoidim[i] = rhoi[i]*p->ref_density;
  return orc_fr_mixture_conductivity(p, rhoidim, T*p->ref_temperature)/p->ref_k;
}

/* CompressibleFREqnSet::ViscousFlux compressibleFR.tcc:551-637 (symmetry2D off) */
void orc_fr_viscous_flux(const orc_case* c, const orc_fr_params* p, const double* Q, const double* grad, const double* avec,
			 double mut, double* flux)
{
  int i, ns = p->chem->nspecies, offset = ns*3;
  double ux = grad[offset + 0], uy = grad[offset + 1], uz = grad[offset + 2];
  double vx = grad[offset + 3], vy = grad[offset + 4], vz = grad[offset + 5];
  double wx = grad[offset + 6], wy = grad[offset + 7], wz = grad[offset + 8];
  double Tx = grad[offset + 9], Ty = grad[offset + 10], Tz = grad[offset + 11];
  double u = Q[ns], v = Q[ns+1], w = Q[ns+2];
  const double* rhoi = &Q[0];
  double T = Q[ns+3];
  double cv, cp, R, gamma, c2;
  double mu, tmut, fact = 2.0/3.0, tauxx, tauyy, tauzz, tauxy, tauxz, tauyz, tauxn, tauyn, tauzn;
  double Re = c->Re, RK, RKT, k, Tn, kT;
  fr_fluid_props(p, rhoi, T, &cv, &cp, &R, &gamma, &c2);
  mu = fr_molecular_viscosity(p, rhoi, T);
  tmut = (mu + mut);
  tauxx = 2.0*fact*ux - fact*vy - fact*wz;
  tauyy = 2.0*fact*vy - fact*ux - fact*wz;
  tauzz = 2.0*fact*wz - fact*ux - fact*vy;
  tauxy = uy + vx;
  tauxz = uz + wx;
  tauyz = vz + wy;
  RK = avec[3]/Re;
  RKT = RK*tmut;
  k = -fr_thermal_conductivity(p, rhoi, T);
  Tn = Tx*avec[0] + Ty*avec[1] + Tz*avec[2];
  kT = (cp*mut)/c->PrT;
  k -= kT;
  tauxn = -tauxx*avec[0] - tauxy*avec[1] - tauxz*avec[2];
  tauyn = -tauxy*avec[0] - tauyy*avec[1] - tauyz*avec[2];
  tauzn = -tauxz*avec[0] - tauyz*avec[1] - tauzz*avec[2];
  for(i = 0; i < ns; i++) flux[i] = 0.0;
  flux[ns+0] = RKT*(tauxn);
  flux[ns+1] = RKT*(tauyn);
  flux[ns+2] = RKT*(tauzn);
  flux[ns+3] = RKT*(tauxn*u + tauyn*v + tauzn*w) + RK*k*Tn;
}

/* one side of CompressibleFREqnSet::ViscousJacobian (compressibleFR.tcc:1796-1960 right, :1963-2010 left): D = -/+dx/s2 */
static void fr_viscous_jac_side(const orc_fr_params* p, const double* D, const double* Qs, double u, double v, double w,
				const double* avec, double RK, double RKT, double c1, double* a)
{
  int i, j, ns = p->chem->nspecies, neqn = ns + 4;
  const double* rhois = &Qs[0];
  double rhos = Qs[ns+5], us = Qs[ns], vs = Qs[ns+1], ws = Qs[ns+2], Ps = Qs[ns+4], Ts = Qs[ns+3];
  double cvs, cps, Rs, gammas, c2s;
  double Drho[3];
  double dux, duy, duz, dvx, dvy, dvz, dwx, dwy, dwz, dfact, dtauxx, dtauyy, dtauzz, dtauxy, dtauxz, dtauyz;
  double dtauxn, dtauyn, dtauzn, c43 = 4.0/3.0, mc23 = -2.0/3.0;
  double dR2_drhou, dR2_drhov, dR2_drhow, dR3_drhou, dR3_drhov, dR3_drhow, dR4_drhou, dR4_drhov, dR4_drhow;
  double dT_dP, dTdRhoi[MAXS], rhoidim[MAXS], dP_dru, dP_drv, dP_drw, dP_dret, Tn;
  double* row;
  fr_fluid_props(p, rhois, Ts, &cvs, &cps, &Rs, &gammas, &c2s);
  for(i = 0; i < 3; i++) Drho[i] = D[i]/rhos;
  dux = -u*Drho[0]; duy = -u*Drho[1]; duz = -u*Drho[2];
  dvx = -v*Drho[0]; dvy = -v*Drho[1]; dvz = -v*Drho[2];
  dwx = -w*Drho[0]; dwy = -w*Drho[1]; dwz = -w*Drho[2];
  dfact = -2.0/3.0*(dux + dvy + dwz);
  dtauxx = (2.0*dux + dfact);
  dtauyy = (2.0*dvy + dfact);
  dtauzz = (2.0*dwz + dfact);
  dtauxy = duy + dvx;
  dtauxz = duz + dwx;
  dtauyz = dvz + dwy;
  dtauxn = dtauxx*avec[0] + dtauxy*avec[1] + dtauxz*avec[2];
  dtauyn = dtauxy*avec[0] + dtauyy*avec[1] + dtauyz*avec[2];
  dtauzn = dtauxz*avec[0] + dtauyz*avec[1] + dtauzz*avec[2];
  for(i = 0; i < ns; i++){
    row = &a[i*neqn];
    for(j = 0; j < neqn; j++) row[j] = 0.0;
  }
  row = &a[ns*neqn];
  dR2_drhou = (c43*Drho[0]*avec[0] + Drho[1]*avec[1] + Drho[2]*avec[2]);
  dR2_drhov = (mc23*Drho[1]*avec[0] + Drho[0]*avec[1]);
  dR2_drhow = (mc23*Drho[2]*avec[0] + Drho[0]*avec[2]);
  for(i = 0; i < ns; i++) row[i] = -RKT*(dtauxn);
  row[ns+0] = -RKT*dR2_drhou;
  row[ns+1] = -RKT*dR2_drhov;
  row[ns+2] = -RKT*dR2_drhow;
  row[ns+3] = 0.0;
  row = &a[(ns+1)*neqn];
  dR3_drhou = (mc23*Drho[0]*avec[1] + Drho[1]*avec[0]);
  dR3_drhov = (Drho[0]*avec[0] + c43*Drho[1]*avec[1] + Drho[2]*avec[2]);
  dR3_drhow = (mc23*Drho[2]*avec[1] + Drho[1]*avec[2]);
  for(i = 0; i < ns; i++) row[i] = -RKT*(dtauyn);
  row[ns+0] = -RKT*dR3_drhou;
  row[ns+1] = -RKT*dR3_drhov;
  row[ns+2] = -RKT*dR3_drhow;
  row[ns+3] = 0.0;
  row = &a[(ns+2)*neqn];
  dR4_drhou = (mc23*Drho[0]*avec[2] + Drho[2]*avec[0]);
  dR4_drhov = (mc23*Drho[1]*avec[2] + Drho[2]*avec[1]);
  dR4_drhow = (Drho[0]*avec[0] + Drho[1]*avec[1] + c43*Drho[2]*avec[2]);
  for(i = 0; i < ns; i++) row[i] = -RKT*(dtauzn);
  row[ns+0] = -RKT*dR4_drhou;
  row[ns+1] = -RKT*dR4_drhov;
  row[ns+2] = -RKT*dR4_drhow;
  row[ns+3] = 0.0;
  /* IdealGasEOS::GetdT_dP EOS.tcc:42-45; ChemModel::dTdRhoi chem.tcc:687-704 on dimensional densities / pressure */
  dT_dP = (1.0/(rhos*Rs));
  {
    double rho = 0.0, Rmix = 0.0, Pdim = Ps*p->ref_pressure, Td, dTdRho, dTdR;
    for(i = 0; i < ns; i++) rhoidim[i] = Qs[i]*p->ref_density;
    for(i = 0; i < ns; i++){
      rho += rhoidim[i];
      Rmix += rhoidim[i]*sp_R(p, i);
    }
    Rmix /= rho;
    Td = Pdim/(rho*Rmix);
    (void)Td;
    dTdRho = (-Pdim/(Rmix*rho*rho));
    dTdR = (-Pdim/(rho*Rmix*Rmix));
    for(i = 0; i < ns; i++) dTdRhoi[i] = dTdR*chem_dRmixdRhoi(p, rhoidim, rho, i) + dTdRho;
  }
  dP_dru = Rs/cvs*us;
  dP_drv = Rs/cvs*vs;
  dP_drw = Rs/cvs*ws;
  dP_dret = Rs/cvs;
  Tn = (D[0]*avec[0] + D[1]*avec[1] + D[2]*avec[2])*c1;
  row = &a[(ns+3)*neqn];
  for(i = 0; i < ns; i++){
    double dT_drho = dTdRhoi[i]/(p->ref_temperature/p->ref_density);
    row[i] = -RKT*(dtauxn*u + dtauyn*v + dtauzn*w) + RK*Tn*dT_drho;
  }
  row[ns+0] = -RKT*(dR2_drhou*u + dR3_drhou*v + dR4_drhou*w) + RK*Tn*dT_dP*dP_dru;
  row[ns+1] = -RKT*(dR2_drhov*u + dR3_drhov*v + dR4_drhov*w) + RK*Tn*dT_dP*dP_drv;
  row[ns+2] = -RKT*(dR2_drhow*u + dR3_drhow*v + dR4_drhow*w) + RK*Tn*dT_dP*dP_drw;
  row[ns+3] =                                                  RK*Tn*dT_dP*dP_dret;
}

/* CompressibleFREqnSet::ViscousJacobian compressibleFR.tcc:1713-2040 */
void orc_fr_viscous_jacobian(const orc_case* c, const orc_fr_params* p, const double* QL, const double* QR, const double* dx,
			     double s2, const double* avec, double mut, double* aL, double* aR)
{
  int i, ns = p->chem->nspecies, neqn = ns + 4;
  double Q[MAXV], DxL[3], DxR[3];
  double T, mu, tmut, RK, RKT, cv, cp, R, gamma, c2, u, v, w, k, kT, c1;
  for(i = 0; i < neqn; i++) Q[i] = 0.5*(QL[i] + QR[i]);
  fr_aux(p, Q);
  T = Q[ns+3];
  mu = fr_molecular_viscosity(p, Q, T);
  tmut = (mu + mut);
  RK = avec[3]/c->Re;
  RKT = RK*tmut;
  fr_fluid_props(p, Q, T, &cv, &cp, &R, &gamma, &c2);
  u = 0.5*(QL[ns] + QR[ns]);
  v = 0.5*(QL[ns+1] + QR[ns+1]);
  w = 0.5*(QL[ns+2] + QR[ns+2]);
  for(i = 0; i < 3; i++){
    DxL[i] = -dx[i]/s2;
    DxR[i] = dx[i]/s2;
  }
  k = fr_thermal_conductivity(p, Q, T);
  kT = mut/c->PrT*cp;
  c1 = -(k + kT);
  fr_viscous_jac_side(p, DxR, QR, u, v, w, avec, RK, RKT, c1, aR);
  fr_viscous_jac_side(p, DxL, QL, u, v, w, avec, RK, RKT, c1, aL);
  for(i = 0; i < neqn*neqn; i++) aL[i] = -aL[i];
}

/* the face gradient of Kernel_Viscous_Flux residual.tcc:430-452 for nterms gradient terms */
static void fr_face_gradient(const orc_case* c, int ns, const double* qL, const double* qR, const double* gradL,
			     const double* gradR, const double* xL, const double* xR, double* grad)
{
  int i, j, nterms = 2*ns + 4;
  for(i = 0; i < nterms*3; i++) grad[i] = 0.5*(gradL[i] + gradR[i]);
  if(c->sorder > 1){
    double dx[3], s2 = 0.0;
    for(i = 0; i < 3; i++){
      dx[i] = (xR[i] - xL[i]);
      s2 += dx[i]*dx[i];
    }
    for(j = 0; j < nterms; j++){
      double qdots = dx[0]*grad[j*3] + dx[1]*grad[j*3+1] + dx[2]*grad[j*3+2];
      int loc = fr_gradloc(ns, j);
      double dq = (qR[loc] - qL[loc] - qdots)/s2;
      for(i = 0; i < 3; i++) grad[j*3 + i] += dq*dx[i];
    }
  }
}

/* ------------------------------------------------------------- residual */

/* SourceTerm compressibleFR.tcc:1276-1316 (gravity off) */
static void fr_source_term(const orc_fr_params* p, const double* Q, double vol, double* source)
{
  int i, neqn = p->chem->nspecies + 4, nvars = 3*p->chem->nspecies + 6;
  if(p->rxn_on) orc_chem_source_term(p->chem, 1, nvars, Q, &vol, p->ref_density, p->ref_time, p->ref_temperature, source);
  else for(i = 0; i < neqn; i++) source[i] = 0.0;
}

/* residual.tcc:13-122 with Kernel_Inviscid_Flux :192-296, Bkernel_Inviscid_Flux :299-387.  TemporalResidual
   (:125-179, torder = 1) subtracts cnp1*vol/dt*(q - qold) with q == qold: an exact zero. */
void orc_fr_residual(const orc_case* c, const orc_fr_params* p, const double* q, const double* qgrad,
		     const double* lim, const double* beta, double* b)
{
  int e, i, j, ns = p->chem->nspecies, neqn = ns + 4, nvars = 3*ns + 6, nterms = 2*ns + 4;
  int nb = c->nbedge + c->ngedge;
  for(i = 0; i < c->nnode*neqn; i++) b[i] = 0.0;
  for(e = 0; e < c->nedge; e++){
    int l = c->edges_n[2*e], r = c->edges_n[2*e+1];
    const double* avec = &c->edges_a[4*e];
    const double* qL = &q[(size_t)l*nvars];
    const double* qR = &q[(size_t)r*nvars];
    double QL[MAXV], QR[MAXV], dQ[MAXE], dx[3], flux[MAXE];
    double avbeta = 0.5*(beta[l] + beta[r]);
    memcpy(QL, qL, sizeof(double)*nvars);
    memcpy(QR, qR, sizeof(double)*nvars);
    if(c->sorder > 1){
      dx[0] = 0.5*(c->xyz[3*r] - c->xyz[3*l]);
      dx[1] = 0.5*(c->xyz[3*r+1] - c->xyz[3*l+1]);
      dx[2] = 0.5*(c->xyz[3*r+2] - c->xyz[3*l+2]);
      for(j = 0; j < neqn; j++) dQ[j] = qR[j] - qL[j];
      fr_extrapolate(c->chi, neqn, QL, qL, dQ, &qgrad[(size_t)l*nterms*3], dx, &lim[(size_t)l*neqn]);
      for(j = 0; j < neqn; j++) dQ[j] = -dQ[j];
      dx[0] = -dx[0]; dx[1] = -dx[1]; dx[2] = -dx[2];
      fr_extrapolate(c->chi, neqn, QR, qR, dQ, &qgrad[(size_t)r*nterms*3], dx, &lim[(size_t)r*neqn]);
      fr_aux(p, QL);
      fr_aux(p, QR);
    }
    fr_numerical_flux(p, QL, QR, avec, 0.0, flux, avbeta);
    for(i = 0; i < neqn; i++) b[(size_t)r*neqn + i] += flux[i];
    for(i = 0; i < neqn; i++) b[(size_t)l*neqn + i] += -flux[i];
  }
  for(e = 0; e < nb; e++){
    int l = c->bedges_n[2*e], r = c->bedges_n[2*e+1];
    const double* avec = &c->bedges_a[4*e];
    const double* qL = &q[(size_t)l*nvars];
    const double* qR = &q[(size_t)r*nvars];
    double QL[MAXV], QR[MAXV], dQ[MAXE], dx[3], flux[MAXE];
    memcpy(QL, qL, sizeof(double)*nvars);
    memcpy(QR, qR, sizeof(double)*nvars);
    if(c->sorder > 1){
      if(is_ghost(c, r)){
	for(j = 0; j < neqn; j++) dQ[j] = qR[j] - qL[j];
	dx[0] = 0.5*(c->xyz[3*r] - c->xyz[3*l]);
	dx[1] = 0.5*(c->xyz[3*r+1] - c->xyz[3*l+1]);
	dx[2] = 0.5*(c->xyz[3*r+2] - c->xyz[3*l+2]);
	fr_extrapolate(c->chi, neqn, QL, qL, dQ, &qgrad[(size_t)l*nterms*3], dx, &lim[(size_t)l*neqn]);
	for(j = 0; j < neqn; j++) dQ[j] = -dQ[j];
	dx[0] = -dx[0]; dx[1] = -dx[1]; dx[2] = -dx[2];
	fr_extrapolate(c->chi, neqn, QR, qR, dQ, &qgrad[(size_t)r*nterms*3], dx, &lim[(size_t)r*neqn]);
      }
      fr_aux(p, QL);
      fr_aux(p, QR);
    }
    fr_numerical_flux(p, QL, QR, avec, 0.0, flux, beta[l]);
    for(i = 0; i < neqn; i++) b[(size_t)l*neqn + i] += -flux[i];
  }
  if(c->viscous){
    /* Kernel_Viscous_Flux residual.tcc:390-466 */
    for(e = 0; e < c->nedge; e++){
      int l = c->edges_n[2*e], r = c->edges_n[2*e+1];
      const double* qL = &q[(size_t)l*nvars];
      const double* qR = &q[(size_t)r*nvars];
      double Qavg[MAXV], grad[MAXT*3], flux[MAXE], tmut;
      for(i = 0; i < neqn; i++) Qavg[i] = (qL[i] + qR[i])/2.0;
      fr_aux(p, Qavg);
      tmut = c->mut ? 0.5*(c->mut[l] + c->mut[r]) : 0.0;
      fr_face_gradient(c, ns, qL, qR, &qgrad[(size_t)l*nterms*3], &qgrad[(size_t)r*nterms*3], &c->xyz[3*l], &c->xyz[3*r], grad);
      orc_fr_viscous_flux(c, p, Qavg, grad, &c->edges_a[4*e], tmut, flux);
      for(i = 0; i < neqn; i++) b[(size_t)r*neqn + i] += flux[i];
      for(i = 0; i < neqn; i++) b[(size_t)l*neqn + i] += -flux[i];
    }
    /* Bkernel_Viscous_Flux residual.tcc:468-562 */
    for(e = 0; e < nb; e++){
      int l = c->bedges_n[2*e], r = c->bedges_n[2*e+1];
      const double* qL = &q[(size_t)l*nvars];
      const double* qR = &q[(size_t)r*nvars];
      double Qavg[MAXV], grad[MAXT*3], flux[MAXE], tmut;
      for(i = 0; i < neqn; i++) Qavg[i] = 0.5*(qL[i] + qR[i]);
      fr_aux(p, Qavg);
      if(is_ghost(c, r)){
	tmut = c->mut ? (c->mut[l] + c->mut[r])/2.0 : 0.0;
	fr_face_gradient(c, ns, qL, qR, &qgrad[(size_t)l*nterms*3], &qgrad[(size_t)r*nterms*3], &c->xyz[3*l], &c->xyz[3*r], grad);
      }
      else{
	tmut = c->mut ? c->mut[l] : 0.0;
	memcpy(grad, &qgrad[(size_t)l*nterms*3], sizeof(double)*3*nterms);
      }
      orc_fr_viscous_flux(c, p, Qavg, grad, &c->bedges_a[4*e], tmut, flux);
      for(i = 0; i < neqn; i++) b[(size_t)l*neqn + i] += -flux[i];
    }
  }
  for(i = 0; i < c->nnode; i++){
    double source[MAXE];
    fr_source_term(p, &q[(size_t)i*nvars], c->vol[i], source);
    for(j = 0; j < neqn; j++) b[(size_t)i*neqn + j] += source[j];
  }
  /* TemporalResidual residual.tcc:125-179 (no GCL): the stored state is native, the equations conservative */
  if(c->torder && c->qold){
    double cnp1 = 1.0, cnm1 = 0.0;
    if(c->iter > 1 && c->torder == 2){ cnp1 = 1.5; cnm1 = -0.5; }
    for(i = 0; i < c->nnode; i++){
      double dt = cnp1*c->vol[i]/p->dt_param;
      double dtm1 = cnm1*c->vol[i]/p->dt_param;
      double Q[MAXV], dq[MAXE], dqm1[MAXE];
      memcpy(Q, &q[(size_t)i*nvars], sizeof(double)*nvars);
      fr_native_to_conservative(p, Q);
      for(j = 0; j < neqn; j++){
	dq[j] = Q[j] - c->qold[(size_t)i*nvars + j];
	dqm1[j] = c->qold[(size_t)i*nvars + j] - c->qoldm1[(size_t)i*nvars + j];
      }
      for(j = 0; j < neqn; j++){
	b[(size_t)i*neqn + j] -= dt*dq[j];
	b[(size_t)i*neqn + j] -= dtm1*dqm1[j];
      }
    }
  }
  /* Bkernel_BC_Res_Modify (residual.tcc:40-43, bc.tcc:905-1056) -> ModifyViscousWallResidual compressibleFR.tcc:2101-2114 */
  for(e = 0; e < nb; e++){
    if(c->bedges_bctype[e] == ORC_BC_NOSLIP){
      double* res = &b[(size_t)c->bedges_n[2*e]*neqn];
      res[ns] = 0.0; res[ns+1] = 0.0; res[ns+2] = 0.0; res[ns+3] = 0.0;
    }
  }
}

/* timestep.tcc:7-49 (local time stepping), kernels :80-143 */
double orc_fr_timestep(const orc_case* c, const orc_fr_params* p, const double* q, const double* beta, double* dt)
{
  int e, i, ns = p->chem->nspecies, neqn = ns + 4, nvars = 3*ns + 6;
  int nb = c->nbedge + c->ngedge;
  double dtmin;
  for(i = 0; i < c->nnode; i++) dt[i] = 0.0;
  for(e = 0; e < c->nedge + nb; e++){
    int interior = e < c->nedge;
    int l = interior ? c->edges_n[2*e] : c->bedges_n[2*(e - c->nedge)];
    int r = interior ? c->edges_n[2*e+1] : c->bedges_n[2*(e - c->nedge)+1];
    const double* avec = interior ? &c->edges_a[4*e] : &c->bedges_a[4*(e - c->nedge)];
    double Q[MAXV], maxeig, bta;
    for(i = 0; i < neqn; i++) Q[i] = 0.5*(q[(size_t)l*nvars + i] + q[(size_t)r*nvars + i]);
    fr_aux(p, Q);
    bta = interior ? 0.5*(beta[l] + beta[r]) : beta[l];
    maxeig = fr_max_eigenvalue(p, Q, avec, 0.0, bta);
    if(interior) dt[r] += maxeig*avec[3];
    dt[l] += maxeig*avec[3];
  }
  dt[0] = c->cfl*(c->vol[0]/dt[0]);
  dtmin = dt[0];
  for(i = 1; i < c->nnode; i++){
    dt[i] = c->cfl*(c->vol[i]/dt[i]);
    if(c->enable_vnn) dt[i] = MIND(dt[i], c->vnn*pow(c->vol[i], 2.0/3.0));
    dtmin = MIND(dtmin, dt[i]);
  }
  return dtmin;
}

/* --------------------------------------------------------------- update */

/* compressibleFR.tcc:886-937 */
static void fr_apply_dq(const orc_fr_params* p, const double* dQ, double* Q)
{
  int i, ns = p->chem->nspecies;
  double minT = 1.0e-10, projectedT;
  for(i = 0; i < ns; i++){
    double rho = Q[i] + dQ[i];
    if(rho < 0.0){ /* refuse to update */ }
    else Q[i] += dQ[i];
  }
  projectedT = Q[ns+3] + dQ[ns+3];
  if(projectedT < 0.0) Q[ns+3] = minT;
  else Q[ns+3] = projectedT;
  Q[ns] += dQ[ns];
  Q[ns+1] += dQ[ns+1];
  Q[ns+2] += dQ[ns+2];
  fr_aux(p, Q);
}

void orc_fr_apply_dq(const orc_case* c, const orc_fr_params* p, double* q, const double* x)
{
  int i, neqn = p->chem->nspecies + 4, nvars = 3*p->chem->nspecies + 6;
  for(i = 0; i < c->nnode; i++) fr_apply_dq(p, &x[(size_t)i*neqn], &q[(size_t)i*nvars]);
}

/* compressibleFR.tcc:2117-2130 */
static void fr_native_to_conservative(const orc_fr_params* p, double* Q)
{
  int ns = p->chem->nspecies;
  double rho = Q[ns+5];
  double Et = fr_total_energy(p, Q);
  Q[ns] *= rho;
  Q[ns+1] *= rho;
  Q[ns+2] *= rho;
  Q[ns+3] = Et;
}

/* compressibleFR.tcc:2133-2201; returns 1 if the Newton iteration did not converge (the reference aborts) */
static int fr_conservative_to_native(const orc_fr_params* p, double* Q)
{
  int i, j, ns = p->chem->nspecies, maxit = 20;
  double tol = 1.0e-12, rho = 0.0, Y[MAXS], rhoiDim[MAXS], R = 0.0, u, v, w, v2, res, P, T, dT = 0.0;
  for(i = 0; i < ns; i++){
    rho += Q[i];
    rhoiDim[i] = Q[i]*p->ref_density;
    R += rhoiDim[i]*sp_R(p, i);
  }
  R /= p->ref_density*rho;
  for(i = 0; i < ns; i++) Y[i] = Q[i]/rho;
  u = Q[ns]/rho; v = Q[ns+1]/rho; w = Q[ns+2]/rho;
  v2 = u*u + v*v + w*w;
  res = Q[ns+3] - 0.5*v2*rho;
  P = Q[ns+4];
  T = ((P*p->ref_pressure)/(R*(rho*p->ref_density)))/p->ref_temperature;   /* IdealGasEOS::GetT EOS.tcc:22-27 */
  for(j = 0; j < maxit; j++){
    double Tp = T + 1.0e-8;
    double H = rho*(chem_specific_enthalpy(p, Y, T*p->ref_temperature)/p->ref_specific_enthalpy);
    double Hp = rho*(chem_specific_enthalpy(p, Y, Tp*p->ref_temperature)/p->ref_specific_enthalpy);
    double Pn = chem_P(p, rhoiDim, T*(p->ref_temperature))/p->ref_pressure;
    double Pp = chem_P(p, rhoiDim, Tp*(p->ref_temperature))/p->ref_pressure;
    double E = H - Pn;
    double Ep = Hp - Pp;
    double zpoint = res - E;
    double zpointp = res - Ep;
    double dzdT = (zpointp - zpoint)/(Tp - T);
    dT = -zpoint/dzdT;
    T += dT;
    if(fabs(dT) < tol) break;
  }
  Q[ns] = u; Q[ns+1] = v; Q[ns+2] = w;
  Q[ns+3] = T;
  return j == maxit;
}

/* solve.tcc:71-140, native-variable branch :112-130 (q is restored; x = change of the native variables) */
int orc_fr_explicit_solve(const orc_case* c, const orc_fr_params* p, double* q, const double* b, const double* dt,
			  double* x)
{
  int i, j, bad = 0, neqn = p->chem->nspecies + 4, nvars = 3*p->chem->nspecies + 6;
  double qc[MAXE];
  for(i = 0; i < c->nnode; i++){
    double* Q = &q[(size_t)i*nvars];
    memcpy(qc, Q, sizeof(double)*neqn);
    fr_native_to_conservative(p, Q);
    for(j = 0; j < neqn; j++) x[(size_t)i*neqn + j] = b[(size_t)i*neqn + j]*dt[i]/c->vol[i];
    for(j = 0; j < neqn; j++) Q[j] += x[(size_t)i*neqn + j];
    bad += fr_conservative_to_native(p, Q);
    for(j = 0; j < neqn; j++) x[(size_t)i*neqn + j] = Q[j] - qc[j];
    memcpy(Q, qc, sizeof(double)*neqn);
  }
  return bad;
}

/* ---------------------------------------------------- block-CRS + solve */

static double* get_block(const int* ia, const int* ja, double* A, int row, int col, int n2)
{
  int k;
  for(k = ia[row]; k < ia[row+1]; k++) if(ja[k] == col) return &A[(size_t)k*n2];
  return NULL;
}

/* chem.tcc:861-873 */
static double chem_dRmixdRhoi(const orc_fr_params* p, const double* rhoi, double rho, int i)
{
  int j, ns = p->chem->nspecies;
  double d = sp_R(p, i)*(rho - rhoi[i])/(rho*rho);
  for(j = 0; j < ns; j++){
    if(j == i) continue;
    d -= sp_R(p, j)*rhoi[j]/(rho*rho);
  }
  return d;
}

/* chem.tcc:829-858 with the IdealGasEOS derivatives EOS.tcc:42-76 */
static double chem_dEtdP_dEtdRhoi(const orc_fr_params* p, const double* rhoi, double T, double P, double v2,
				  double* dEtdRhoi)
{
  int i, ns = p->chem->nspecies;
  double hv2 = 0.5*v2, rhomix = 0.0, Rmix = 0.0, dPdrhomix, dPdRmix, dEtdP = 0.0, dTdP;
  (void)P;
  for(i = 0; i < ns; i++){
    rhomix += rhoi[i];
    Rmix += rhoi[i]*sp_R(p, i);
  }
  Rmix /= rhomix;
  dPdrhomix = (Rmix*T);
  dPdRmix = (rhomix*T);
  for(i = 0; i < ns; i++) dEtdRhoi[i] = hv2 + sp_h(p, i, T) - (dPdrhomix + dPdRmix*chem_dRmixdRhoi(p, rhoi, rhomix, i));
  dTdP = (1.0/(rhomix*Rmix));
  for(i = 0; i < ns; i++) dEtdP += rhoi[i]*(sp_cp(p, i, T))*dTdP;
  dEtdP -= 1.0;
  return dEtdP;
}

/* CompressibleFREqnSet::ContributeTemporalTerms compressibleFR.tcc:1319-1463 */
static void fr_temporal_terms(const orc_fr_params* p, const double* Q, double vol, double cnp1, double dt, double dtau,
			      double* A, double beta)
{
  int i, j, ns = p->chem->nspecies, neqn = ns + 4;
  int uloc = ns, vloc = ns+1, wloc = ns+2, tloc = ns+3;
  double dEtdRhoi[MAXS], rhoiDim[MAXS], Yi[MAXS], thetaOBetai[MAXS];
  double vOverDt, rho = Q[ns+5], T = Q[tloc], P = Q[ns+4], u = Q[uloc], v = Q[vloc], w = Q[wloc];
  double rvOverDt, v2, TDim, PDim, precond, cv, cp, R, gamma, c2, s_ref, rhoDim, dEtdP, ref_detdrho, ref_detdP, dPdT;
  double oneOBeta, oneMbeta;
  if(p->use_local_dt) vOverDt = cnp1*vol/dt + vol/dtau;
  else vOverDt = cnp1*vol/dtau;
  rvOverDt = rho*vOverDt;
  v2 = u*u + v*v + w*w;
  fr_fluid_props(p, Q, T, &cv, &cp, &R, &gamma, &c2);
  s_ref = (p->ref_velocity*p->ref_velocity/p->ref_temperature);
  rhoDim = rho*p->ref_density;
  TDim = T*p->ref_temperature;
  PDim = P*p->ref_pressure;
  v2 *= p->ref_velocity*p->ref_velocity;
  for(i = 0; i < ns; i++){
    rhoiDim[i] = Q[i]*p->ref_density;
    Yi[i] = Q[i]/rho;
  }
  dEtdP = chem_dEtdP_dEtdRhoi(p, rhoiDim, TDim, PDim, v2, dEtdRhoi);
  ref_detdrho = p->ref_specific_enthalpy*p->ref_density/p->ref_density;
  for(i = 0; i < ns; i++) dEtdRhoi[i] /= ref_detdrho;
  ref_detdP = p->ref_specific_enthalpy*p->ref_density/p->ref_pressure;
  dEtdP /= ref_detdP;
  dPdT = (rhoDim*(R*s_ref));
  dPdT /= (p->ref_pressure/p->ref_temperature);
  oneOBeta = 1.0/beta;
  oneMbeta = 1.0 - beta;
  for(i = 0; i < ns; i++) thetaOBetai[i] = (Yi[i]*oneMbeta/c2)*oneOBeta;
  for(i = 0; i < ns; i++){
    A[neqn*i + i] += vOverDt;
    A[neqn*i + tloc] += thetaOBetai[i]*vOverDt*dPdT;
  }
  precond = 0.0;
  for(j = 0; j < ns; j++){ A[neqn*uloc + j] += u*vOverDt; precond += thetaOBetai[j]*u; }
  A[neqn*uloc + uloc] += rvOverDt;
  A[neqn*uloc + tloc] += precond*vOverDt*dPdT;
  precond = 0.0;
  for(j = 0; j < ns; j++){ A[neqn*vloc + j] += v*vOverDt; precond += thetaOBetai[j]*v; }
  A[neqn*vloc + vloc] += rvOverDt;
  A[neqn*vloc + tloc] += precond*vOverDt*dPdT;
  precond = 0.0;
  for(j = 0; j < ns; j++){ A[neqn*wloc + j] += w*vOverDt; precond += thetaOBetai[j]*w; }
  A[neqn*wloc + wloc] += rvOverDt;
  A[neqn*wloc + tloc] += precond*vOverDt*dPdT;
  precond = 0.0;
  for(j = 0; j < ns; j++){ A[neqn*tloc + j] += dEtdRhoi[j]*vOverDt; precond += dEtdRhoi[j]*thetaOBetai[j]; }
  A[neqn*tloc + uloc] += u*rvOverDt;
  A[neqn*tloc + vloc] += v*rvOverDt;
  A[neqn*tloc + wloc] += w*rvOverDt;
  precond += dEtdP*(oneOBeta);
  A[neqn*tloc + tloc] += precond*vOverDt*dPdT;
}

/* jacobian.tcc:130-250: blank, Driver(NumJac :254-304), Bdriver(BNumJac :459-544), Driver(Diag :434-456), source-term
   Jacobian (eqnset.tcc:163-187), temporal terms.  q is written (phantom nodes) exactly as the reference does. */
void orc_fr_jacobian(const orc_case* c, const orc_fr_params* p, double* q, const double* beta, const double* dt,
		     const int* ia, const int* ja, const int* iau, double* A)
{
  int e, i, j, ns = p->chem->nspecies, neqn = ns + 4, nvars = 3*ns + 6, n2 = neqn*neqn;
  int nb = c->nbedge + c->ngedge;
  size_t k;
  const double h = 1.0e-8;
  (void)iau;
  for(k = 0; k < (size_t)ia[c->nnode]*n2; k++) A[k] = 0.0;
  for(e = 0; e < c->nedge; e++){
    int l = c->edges_n[2*e], r = c->edges_n[2*e+1], kk;
    const double* avec = &c->edges_a[4*e];
    const double* QL = &q[(size_t)l*nvars];
    const double* QR = &q[(size_t)r*nvars];
    double QPL[MAXV], QPR[MAXV], fluxS[MAXE], fluxL[MAXE], fluxR[MAXE], tempL[MAXE*MAXE], tempR[MAXE*MAXE];
    double *pR, *pL, avbeta = 0.5*(beta[l] + beta[r]);
    if(c->field_jac_type == 1){
      /* Kernel_NumJac_Centered jacobian.tcc:306-366 */
      double fluxLd[MAXE], fluxRd[MAXE];
      for(i = 0; i < neqn; i++){
	memcpy(QPL, QL, sizeof(double)*nvars);
	memcpy(QPR, QR, sizeof(double)*nvars);
	QPL[i] += h; QPR[i] += h;
	fr_aux(p, QPL); fr_aux(p, QPR);
	fr_numerical_flux(p, QPL, QR, avec, 0.0, fluxL, avbeta);
	fr_numerical_flux(p, QL, QPR, avec, 0.0, fluxR, avbeta);
	memcpy(QPL, QL, sizeof(double)*nvars);
	memcpy(QPR, QR, sizeof(double)*nvars);
	QPL[i] -= h; QPR[i] -= h;
	fr_aux(p, QPL); fr_aux(p, QPR);
	fr_numerical_flux(p, QPL, QR, avec, 0.0, fluxLd, avbeta);
	fr_numerical_flux(p, QL, QPR, avec, 0.0, fluxRd, avbeta);
	for(j = 0; j < neqn; j++) tempL[j*neqn + i] = (fluxLd[j] - fluxL[j])/(2.0*h);
	for(j = 0; j < neqn; j++) tempR[j*neqn + i] = (fluxR[j] - fluxRd[j])/(2.0*h);
      }
    }
    else{
    fr_numerical_flux(p, QL, QR, avec, 0.0, fluxS, avbeta);
    for(i = 0; i < neqn; i++){
      memcpy(QPL, QL, sizeof(double)*nvars);
      memcpy(QPR, QR, sizeof(double)*nvars);
      QPL[i] += h; QPR[i] += h;
      fr_aux(p, QPL); fr_aux(p, QPR);
      fr_numerical_flux(p, QPL, QR, avec, 0.0, fluxL, avbeta);
      fr_numerical_flux(p, QL, QPR, avec, 0.0, fluxR, avbeta);
      for(j = 0; j < neqn; j++) tempL[j*neqn + i] = (fluxS[j] - fluxL[j])/h;
      for(j = 0; j < neqn; j++) tempR[j*neqn + i] = (fluxR[j] - fluxS[j])/h;
    }
    }
    pR = get_block(ia, ja, A, l, r, n2);
    pL = get_block(ia, ja, A, r, l, n2);
    for(kk = 0; kk < n2; kk++) pR[kk] += tempR[kk];
    for(kk = 0; kk < n2; kk++) pL[kk] += tempL[kk];
  }
  for(e = 0; e < nb; e++){
    int l = c->bedges_n[2*e], r = c->bedges_n[2*e+1], kk;
    const double* avec = &c->bedges_a[4*e];
    double* QL = &q[(size_t)l*nvars];
    double* QR = &q[(size_t)r*nvars];
    int bctype = c->bedges_bctype[e];
    double QPL[MAXV], QPR[MAXV], fluxS[MAXE], fluxL[MAXE], fluxR[MAXE], tempL[MAXE*MAXE], tempR[MAXE*MAXE], Qref[MAXV];
    double *pL, betaL = beta[l];
    for(i = 0; i < nvars; i++) Qref[i] = p->qinf[i];
    fr_boundary_variables(c, p, QL, QR, avec, bctype, betaL, e, q, Qref);
    if(c->boundary_jac_type == 1){
      /* Bkernel_NumJac_Centered jacobian.tcc:546-640, boundaryJacEval == 0 (BC re-evaluated for the +h state only) */
      double fluxLd[MAXE], fluxRd[MAXE];
      for(i = 0; i < neqn; i++){
	memcpy(QPL, QL, sizeof(double)*nvars);
	memcpy(QPR, QR, sizeof(double)*nvars);
	QPL[i] += h; QPR[i] += h;
	fr_aux(p, QPL); fr_aux(p, QPR);
	fr_numerical_flux(p, QL, QPR, avec, 0.0, fluxR, betaL);
	if(!is_ghost(c, r)){
	  memcpy(QPR, QR, sizeof(double)*nvars);
	  fr_aux(p, QPR);
	  fr_boundary_variables(c, p, QPL, QPR, avec, bctype, betaL, e, q, Qref);
	  fr_numerical_flux(p, QPL, QPR, avec, 0.0, fluxL, betaL);
	}
	else{
	  fr_numerical_flux(p, QPL, QR, avec, 0.0, fluxL, betaL);
	}
	memcpy(QPL, QL, sizeof(double)*nvars);
	memcpy(QPR, QR, sizeof(double)*nvars);
	QPL[i] -= h; QPR[i] -= h;
	fr_aux(p, QPL); fr_aux(p, QPR);
	fr_numerical_flux(p, QL, QPR, avec, 0.0, fluxRd, betaL);
	fr_numerical_flux(p, QPL, QR, avec, 0.0, fluxLd, betaL);
	for(j = 0; j < neqn; j++){
	  tempL[j*neqn + i] = (fluxL[j] - fluxLd[j])/(2.0*h);
	  tempR[j*neqn + i] = (fluxR[j] - fluxRd[j])/(2.0*h);
	}
      }
    }
    else{
    fr_numerical_flux(p, QL, QR, avec, 0.0, fluxS, betaL);
    for(i = 0; i < neqn; i++){
      memcpy(QPL, QL, sizeof(double)*nvars);
      memcpy(QPR, QR, sizeof(double)*nvars);
      QPL[i] += h; QPR[i] += h;
      fr_aux(p, QPL); fr_aux(p, QPR);
      fr_numerical_flux(p, QL, QPR, avec, 0.0, fluxR, betaL);
      if(!is_ghost(c, r)){
	memcpy(QPR, QR, sizeof(double)*nvars);
	fr_aux(p, QPR);
	fr_boundary_variables(c, p, QPL, QPR, avec, bctype, betaL, e, q, Qref);
	fr_numerical_flux(p, QPL, QPR, avec, 0.0, fluxL, betaL);
      }
      else{
	fr_numerical_flux(p, QPL, QR, avec, 0.0, fluxL, betaL);
      }
      for(j = 0; j < neqn; j++){
	tempL[j*neqn + i] = (fluxL[j] - fluxS[j])/h;
	tempR[j*neqn + i] = (fluxR[j] - fluxS[j])/h;
      }
    }
    }
    if(is_ghost(c, r)){
      double* pR = get_block(ia, ja, A, l, r, n2);
      for(kk = 0; kk < n2; kk++) pR[kk] += tempR[kk];
    }
    pL = get_block(ia, ja, A, l, l, n2);
    for(kk = 0; kk < n2; kk++) pL[kk] += tempL[kk];
  }
  /* Kernel_Viscous_Jac jacobian.tcc:728-767 (Bkernel_Viscous_Jac :770-848 ends with size = 0: no-op) */
  if(c->viscous){
    for(e = 0; e < c->nedge; e++){
      int l = c->edges_n[2*e], r = c->edges_n[2*e+1], kk;
      double tmut = c->mut ? (c->mut[l] + c->mut[r])/2.0 : 0.0;
      double dx[3], s2 = 0.0, tempL[MAXE*MAXE], tempR[MAXE*MAXE];
      double *pL, *pR;
      for(i = 0; i < 3; i++){
	dx[i] = (c->xyz[3*r+i] - c->xyz[3*l+i]);
	s2 += dx[i]*dx[i];
      }
      orc_fr_viscous_jacobian(c, p, &q[(size_t)l*nvars], &q[(size_t)r*nvars], dx, s2, &c->edges_a[4*e], tmut, tempL, tempR);
      pL = get_block(ia, ja, A, r, l, n2);
      pR = get_block(ia, ja, A, l, r, n2);
      for(kk = 0; kk < n2; kk++) pR[kk] += tempR[kk];
      for(kk = 0; kk < n2; kk++) pL[kk] += tempL[kk];
    }
  }
  for(e = 0; e < c->nedge; e++){
    int l = c->edges_n[2*e], r = c->edges_n[2*e+1], kk;
    double* dL = get_block(ia, ja, A, l, l, n2);
    double* dR = get_block(ia, ja, A, r, r, n2);
    const double* jacL = get_block(ia, ja, A, r, l, n2);
    const double* jacR = get_block(ia, ja, A, l, r, n2);
    for(kk = 0; kk < n2; kk++) dR[kk] += -jacR[kk];
    for(kk = 0; kk < n2; kk++) dL[kk] += -jacL[kk];
  }
  /* source-term Jacobian: one-sided FD on the native variables, subtracted from the diagonal block */
  for(i = 0; i < c->nnode; i++){
    const double* Q = &q[(size_t)i*nvars];
    double source[MAXE], sourceP[MAXE], QP[MAXV], SJ[MAXE*MAXE];
    double* jac = get_block(ia, ja, A, i, i, n2);
    int kk, m;
    for(kk = 0; kk < n2; kk++) SJ[kk] = 0.0;
    fr_source_term(p, Q, c->vol[i], source);
    for(m = 0; m < neqn; m++){
      memcpy(QP, Q, sizeof(double)*neqn);
      QP[m] += h;
      fr_aux(p, QP);
      fr_source_term(p, QP, c->vol[i], sourceP);
      for(j = 0; j < neqn; j++) SJ[j*neqn + m] = (sourceP[j] - source[j])/h;
    }
    for(kk = 0; kk < n2; kk++) jac[kk] -= SJ[kk];
  }
  /* ContributeTemporalTerms jacobian.tcc:214-250 */
  {
    double cnp1 = (c->iter > 1 && c->torder == 2) ? 1.5 : 1.0;
    for(i = 0; i < c->nnode; i++){
      fr_temporal_terms(p, &q[(size_t)i*nvars], c->vol[i], cnp1, p->dt_param, dt[i], get_block(ia, ja, A, i, i, n2), beta[i]);
    }
  }
  /* Bkernel_BC_Jac_Modify (jacobian.tcc:247-249, bc.tcc:747-903) -> ModifyViscousWallJacobian compressibleFR.tcc:2072-2099
     with CRSMatrix::BlankSubRow (crsmatrix.tcc:524-541) */
  for(e = 0; e < nb; e++){
    if(c->bedges_bctype[e] == ORC_BC_NOSLIP){
      int cv = c->bedges_n[2*e], sub, kk;
      double Twall = fr_wall_temperature(c, p, e);
      double* diag = get_block(ia, ja, A, cv, cv, n2);
      for(sub = ns; sub < ns + 4; sub++){
	for(kk = ia[cv]; kk < ia[cv+1]; kk++) for(j = 0; j < neqn; j++) A[(size_t)kk*n2 + sub*neqn + j] = 0.0;
	diag[sub*neqn + sub] = 1.0;
      }
      if(Twall < 0.0){
	double* off = get_block(ia, ja, A, cv, orc_normal_node(c, e), n2);
	off[(ns+3)*neqn + ns+3] = -1.0;
      }
    }
  }
}

/* crsmatrix.tcc:840-876 */
void orc_fr_prepare_sgs(const orc_case* c, const orc_fr_params* p, const int* iau, double* A, int* pv)
{
  int i, neqn = p->chem->nspecies + 4;
  for(i = 0; i < c->nnode; i++) lu(&A[(size_t)iau[i]*neqn*neqn], &pv[(size_t)i*neqn], neqn);
}

/* crs.tcc:62-173 on one rank */
double orc_fr_sgs(const orc_case* c, const orc_fr_params* p, int nsgs, const int* ia, const int* ja, const int* iau,
		  const double* A, const int* pv, const double* b, double* x)
{
  int isgs, i, k, indx, dir, neqn = p->chem->nspecies + 4, n2 = neqn*neqn;
  double rhs[MAXE], vout[MAXE], temp[MAXE];
  double xOld = 0.0, xNorm = 0.0;
  int n = c->nnode;
  for(isgs = 0; isgs < nsgs; isgs++){
    for(dir = 0; dir < 2; dir++){
      for(k = 0; k < n; k++){
	i = dir ? (n - 1 - k) : k;
	memcpy(rhs, &b[(size_t)i*neqn], sizeof(double)*neqn);
	for(indx = ia[i]+1; indx < ia[i+1]; indx++){
	  int j, node2 = ja[indx];
	  matvec(&A[(size_t)indx*n2], &x[(size_t)node2*neqn], vout, neqn);
	  for(j = 0; j < neqn; j++) rhs[j] -= vout[j];
	}
	lu_solve(&A[(size_t)iau[i]*n2], rhs, &pv[(size_t)i*neqn], temp, neqn);
	memcpy(&x[(size_t)i*neqn], rhs, sizeof(double)*neqn);
      }
    }
    xOld = xNorm;
    {
      double s = 0.0;
      for(i = 0; i < n*neqn; i++) s += x[i]*x[i];
      xNorm = sqrt(s)/(double)(n*neqn);
    }
  }
  return fabs(xOld - xNorm);
}

/* unit-test hooks */
void orc_fr_compute_aux(const orc_fr_params* p, double* Q){ fr_aux(p, Q); }
void orc_fr_hllc_flux(const orc_fr_params* p, const double* QL, const double* QR, const double* avec, double vdotn,
		      double beta, double* flux)
{
  fr_numerical_flux(p, QL, QR, avec, vdotn, flux, beta);
}


/* ---- Spalart-Allmaras under the viscous reacting eqnset (turbulenceModel = 1 with compressibleNSFR): TurbulenceModel::
   Compute (turb.tcc:163-339) is eqnset-agnostic; what it asks of CompressibleFREqnSet: GetTheta (:1640-1650: native
   velocities), ComputeAuxiliaryVariables on the averaged native state (:755-814), GetDensity (:640: the aux entry),
   ComputeViscosity (:1572-1587: Wilke-mixed species viscosity / ref_viscosity), EqnSet::GetRe (eqnset.h:137: Param::Re),
   GetVelocityGradLocation (:681: nspecies). */
static double frg_theta_avg(const orc_gas* g, const double* qL, const double* qR, const double* avec)
{
  const orc_fr_params* p = (const orc_fr_params*)g->ctx;
  int i, ns = p->chem->nspecies, neqn = ns + 4;
  double qa[MAXV];
  for(i = 0; i < neqn; i++) qa[i] = 0.5*(qL[i] + qR[i]);
  return fr_theta(ns, qa, avec, 0.0);
}
static void frg_rho_nu_avg(const orc_gas* g, const double* qL, const double* qR, double* rho, double* nu)
{
  const orc_fr_params* p = (const orc_fr_params*)g->ctx;
  int i, ns = p->chem->nspecies, neqn = ns + 4;
  double qavg[MAXV], mu;
  for(i = 0; i < g->nvars; i++) qavg[i] = 0.0;
  for(i = 0; i < neqn; i++) qavg[i] = 0.5*(qL[i] + qR[i]);
  fr_aux(p, qavg);
  *rho = qavg[ns+5];
  mu = fr_molecular_viscosity(p, qavg, qavg[ns+3]);
  *nu = mu/(*rho);
}
static void frg_rho_nu_node(const orc_gas* g, const double* Q, double* rho, double* nu)
{
  const orc_fr_params* p = (const orc_fr_params*)g->ctx;
  int ns = p->chem->nspecies;
  double mu = fr_molecular_viscosity(p, Q, Q[ns+3]);
  *rho = Q[ns+5];
  *nu = mu/(*rho);
}

double orc_fr_turb_sa_phase(const orc_case* c, const orc_fr_params* p, int phase, int nsgs, const double* q, const double* qgrad,
			    const double* s, const double* dist, const double* dt, const int* ia, const int* ja, const int* iau,
			    double* tvar, double* tgrad, double* b, double* A, double* x, double* mut)
{
  int ns = p->chem->nspecies;
  orc_gas gas;
  gas.nvars = 3*ns + 6; gas.nterms = 2*ns + 4; gas.vloc = 3*ns;
  gas.Re = c->Re;
  gas.ctx = p;
  gas.theta_avg = frg_theta_avg; gas.rho_nu_avg = frg_rho_nu_avg; gas.rho_nu_node = frg_rho_nu_node;
  return orc_turb_sa_phase_gas(c, &gas, phase, nsgs, q, qgrad, s, dist, dt, ia, ja, iau, tvar, tgrad, b, A, x, mut);
}

double orc_fr_turb_sa(const orc_case* c, const orc_fr_params* p, int nsgs, const double* q, const double* qgrad, const double* s,
		      const double* dist, const double* dt, const int* ia, const int* ja, const int* iau,
		      double* tvar, double* tgrad, double* b, double* A, double* x, double* mut)
{
  int ph, isgs;
  double ss = 0.0;
  for(ph = 0; ph <= 5; ph++){
    int reps = (ph == 3 && nsgs > 0) ? nsgs : 1;
    for(isgs = 0; isgs < reps; isgs++){
      double r = orc_fr_turb_sa_phase(c, p, ph, nsgs, q, qgrad, s, dist, dt, ia, ja, iau, tvar, tgrad, b, A, x, mut);
      if(ph == 2) ss = r;
    }
  }
  return sqrt(ss)/(double)c->nnode;
}


/* ---- surface forces under the reacting eqnset: GetCp (compressibleFR.tcc:672-678: (P - Pinf)/(V^2/2)), GetPressure,
   ComputeViscosity (Wilke-mixed), GetDensity, ComputeStressVector (:2204-2242: mu/Re), GetCf (:2245-2250) */
static double frg_cp_of(const orc_gas* g, const double* Q)
{
  const orc_fr_params* p = (const orc_fr_params*)g->ctx;
  int ns = p->chem->nspecies;
  double P = Q[ns+4], Pinf = p->qinf[ns+4];
  return ((P - Pinf)/(0.5*g->V*g->V));
}
static double frg_pressure_of(const orc_gas* g, const double* Q)
{
  const orc_fr_params* p = (const orc_fr_params*)g->ctx;
  return Q[p->chem->nspecies+4];
}
static void frg_mu_rho_node(const orc_gas* g, const double* Q, double* mu, double* rho)
{
  const orc_fr_params* p = (const orc_fr_params*)g->ctx;
  int ns = p->chem->nspecies;
  *mu = fr_molecular_viscosity(p, Q, Q[ns+3]);
  *rho = Q[ns+5];
}

void orc_fr_forces(const orc_case* c, const orc_fr_params* p, const orc_forces_desc* d, double V, const double* q,
		   const double* qgrad, const double* body_area, double* cp, double* yp, double* cf, double* body, double* coef)
{
  int ns = p->chem->nspecies;
  orc_gas gas;
  gas.nvars = 3*ns + 6; gas.nterms = 2*ns + 4; gas.vloc = 3*ns;
  gas.Re = c->Re;
  gas.ctx = p;
  gas.theta_avg = frg_theta_avg; gas.rho_nu_avg = frg_rho_nu_avg; gas.rho_nu_node = frg_rho_nu_node;
  gas.V = V; gas.rho_inf = p->qinf[ns+5];
  gas.cp_of = frg_cp_of; gas.pressure_of = frg_pressure_of; gas.mu_rho_node = frg_mu_rho_node;
  orc_forces_gas(c, &gas, d, q, qgrad, body_area, cp, yp, cf, body, coef);
}
