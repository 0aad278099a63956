// eqnset_fr.cuh -- device-side physics of the reacting eqnset (CompressibleFREqnSet, ucs/compressibleFR.tcc) for
// sm_100a, FP64.  NS = number of species is a template parameter: NEQ = NS+4 equations, NV = 3NS+6 variables per node
// [rho_i | u v w | T | P | rho | cv_i | mol_i] (compressibleFR.tcc:14-31), NT = 2NS+4 gradient terms (:693-711).
//
// Parity contract (same as eqnset_compressible.cuh): every expression keeps the reference's operand order and the
// library is built with --fmad=false, so +, -, *, /, sqrt round exactly as the reference's x86-64 build does.  The
// chemistry source term (chem_device.cuh) calls exp / log / pow, which are CUDA libm here and glibc there: those rows
// agree to rounding (1e-12 of the rate scale), everything else is bit-identical.
#pragma once

#include "chem_device.cuh"

namespace fr {

template <int NS>
struct Params {
  const pcfd_chem_model* chem;   // device copy of the reaction tables (source term)
  double Rs[NS];                 // Species::R = UNIV_R / MW (species.tcc:315)
  double mw[NS];
  double nasa[NS][2][7];         // Species::thermo_coeff
  double ref_density, ref_velocity, ref_temperature, ref_pressure, ref_time, ref_specific_enthalpy;
  double Pref, dt;
  double chi, cfl;
  int use_local_dt, rxn_on, no_cvbc, sorder, limiter;
  double qinf[3 * NS + 6];
};

__device__ __forceinline__ double maxd(double x, double y) { return (x > y) ? x : y; }
__device__ __forceinline__ double mind(double x, double y) { return (x < y) ? x : y; }

// Species::GetThermoCoeff (species.tcc:96-138): the pinned temperature is local to that function
__device__ __forceinline__ int thermo_range(double T) {
  if (T < 200.0) return 0;
  if (T > 6000.0) return 1;
  return (T > 1000.0) ? 1 : 0;
}
// Species::GetCp (species.tcc:43-53); GetdHdT (:337-349) is the same polynomial
template <int NS>
__device__ __forceinline__ double sp_cp(const Params<NS>& p, int i, int rng, double T) {
  const double* a = p.nasa[i][rng];
  const double cp_R = a[0] + T * (a[1] + T * (a[2] + T * (a[3] + T * a[4])));
  return cp_R * p.Rs[i];
}
// Species::GetH (species.tcc:55-71), href == 0
template <int NS>
__device__ __forceinline__ double sp_h(const Params<NS>& p, int i, int rng, double T) {
  const double* a = p.nasa[i][rng];
  const double h_R = a[5] + T * (a[0] + T * (a[1] / 2.0 + T * (a[2] / 3.0 + T * (a[3] / 4.0 + T * a[4] / 5.0))));
  return h_R * p.Rs[i];
}
// ChemModel::GetP (chem.tcc:989-999), IdealGasEOS::GetP (EOS.tcc:36-40)
template <int NS>
__device__ __forceinline__ double chem_P(const Params<NS>& p, const double* rhoiDim, double T) {
  double P = 0.0;
#pragma unroll
  for (int i = 0; i < NS; i++) P += rhoiDim[i] * p.Rs[i] * T;
  return P;
}
// ChemModel::GetSpecificEnthalpy (chem.tcc:586-595)
template <int NS>
__device__ __forceinline__ double chem_h(const Params<NS>& p, const double* X, double T) {
  const int rng = thermo_range(T);
  double h = 0.0;
#pragma unroll
  for (int i = 0; i < NS; i++) h += sp_h(p, i, rng, T) * X[i];
  return h;
}

// ComputeAuxiliaryVariables (compressibleFR.tcc:755-814): P and rho only -- what the fluxes read
template <int NS>
__device__ __forceinline__ void aux_pr(const Params<NS>& p, double* Q) {
  double rho = 0.0, rhoiDim[NS];
#pragma unroll
  for (int i = 0; i < NS; i++) { rho += Q[i]; rhoiDim[i] = Q[i] * p.ref_density; }
  const double TDim = Q[NS + 3] * p.ref_temperature;
  Q[NS + 4] = chem_P(p, rhoiDim, TDim) / p.ref_pressure;
  Q[NS + 5] = rho;
}
// ... and the stored ones as well (cv_i, molar concentrations)
template <int NS>
__device__ __forceinline__ void aux(const Params<NS>& p, double* Q) {
  aux_pr(p, Q);
  const double TDim = Q[NS + 3] * p.ref_temperature;
  const double s_ref = (p.ref_velocity * p.ref_velocity / p.ref_temperature);
  const int rng = thermo_range(TDim);
#pragma unroll
  for (int i = 0; i < NS; i++) {
    const double cpiDim = sp_cp(p, i, rng, TDim);
    Q[NS + 6 + i] = (cpiDim - p.Rs[i]) / s_ref;           // IdealGasEOS::GetCv (EOS.tcc:79-91)
  }
#pragma unroll
  for (int i = 0; i < NS; i++) Q[NS + NS + 6 + i] = Q[i] / p.mw[i] / 1000.0;   // MassToMole (chem.tcc:980-986)
}

// GetFluidProperties (compressibleFR.tcc:1523-1570) -> ChemModel::GetFluidProperties (chem.tcc:545-572): mixture gas
// constant and speed of sound squared (non-dimensional); cv, cp, P, rho of that call are never read by the path
template <int NS>
__device__ __forceinline__ void fluid_props(const Params<NS>& p, const double* rhoi, double T, double& R, double& c2) {
  const double s_ref = (p.ref_velocity * p.ref_velocity / p.ref_temperature);
  double rhoiDim[NS];
#pragma unroll
  for (int i = 0; i < NS; i++) rhoiDim[i] = rhoi[i] * p.ref_density;
  const double Tdim = T * p.ref_temperature;
  double RDim = 0.0, rhoDim = 0.0, cpDim = 0.0, cvDim = 0.0;
#pragma unroll
  for (int i = 0; i < NS; i++) rhoDim += rhoiDim[i];
#pragma unroll
  for (int i = 0; i < NS; i++) RDim += rhoiDim[i] * p.Rs[i];
  RDim /= rhoDim;
  const int rng = thermo_range(Tdim);
#pragma unroll
  for (int i = 0; i < NS; i++) {
    const double X = rhoiDim[i] / rhoDim;
    const double cpi = sp_cp(p, i, rng, Tdim);
    cpDim += cpi * X;
    cvDim += X * (cpi - p.Rs[i]);
  }
  const double g = cpDim / cvDim;
  const double c2Dim = g * RDim * Tdim;
  R = RDim / s_ref;
  c2 = c2Dim / (p.ref_velocity * p.ref_velocity);
}

// GetTotalEnthalpy (compressibleFR.tcc:1249-1273); hr = h rho and ke = rho |v|^2 / 2 are also what GetTotalEnergy
// (:1219-1246) is made of: E = (hr - P) + ke
template <int NS>
__device__ __forceinline__ double total_enthalpy(const Params<NS>& p, const double* Q, double& hr, double& ke) {
  double X[NS];
  const double T = Q[NS + 3], rho = Q[NS + 5], u = Q[NS], v = Q[NS + 1], w = Q[NS + 2];
  const double v2 = u * u + v * v + w * w;
#pragma unroll
  for (int i = 0; i < NS; i++) X[i] = Q[i] / rho;
  const double h = chem_h(p, X, T * p.ref_temperature) / p.ref_specific_enthalpy;
  hr = h * rho;
  ke = 0.5 * rho * v2;
  return hr + ke;
}
template <int NS>
__device__ __forceinline__ double total_enthalpy(const Params<NS>& p, const double* Q) {
  double hr, ke;
  return total_enthalpy(p, Q, hr, ke);
}
// GetTotalEnergy (compressibleFR.tcc:1219-1246)
template <int NS>
__device__ __forceinline__ double total_energy(const Params<NS>& p, const double* Q) {
  double X[NS];
  const double T = Q[NS + 3], P = Q[NS + 4], rho = Q[NS + 5], u = Q[NS], v = Q[NS + 1], w = Q[NS + 2];
  const double v2 = u * u + v * v + w * w;
#pragma unroll
  for (int i = 0; i < NS; i++) X[i] = Q[i] / rho;
  const double h = chem_h(p, X, T * p.ref_temperature) / p.ref_specific_enthalpy;
  const double E = h * rho - P;
  return E + 0.5 * rho * v2;
}
// GetTheta (compressibleFR.tcc:1640-1644)
template <int NS>
__device__ __forceinline__ double theta_of(const double* Q, const double* n, double vdotn) {
  return (Q[NS] * n[0] + Q[NS + 1] * n[1] + Q[NS + 2] * n[2] + vdotn);
}

// HLLCFlux (compressibleFR.tcc:301-548; RoeFlux :293-298 forwards here), cut into the pieces the finite-difference
// Jacobian can share between its perturbed evaluations (kfr_jac_edges): what depends on ONE state only
// (hllc_side_thermo), what depends on the densities and temperatures of both (hllc_roe_c2), and the rest (hllc_assemble).
// numerical_flux is their composition; every piece does the reference's operations in the reference's order.

// the speed of sound squared (GetFluidProperties) and h rho (GetTotalEnthalpy) of one state; Q needs rho at [NS+5]
template <int NS>
__device__ __forceinline__ void hllc_side_thermo(const Params<NS>& p, const double* Q, double& c2, double& hr) {
  double R, X[NS];
  fluid_props(p, Q, Q[NS + 3], R, c2);
  const double rho = Q[NS + 5];
#pragma unroll
  for (int i = 0; i < NS; i++) X[i] = Q[i] / rho;
  const double h = chem_h(p, X, Q[NS + 3] * p.ref_temperature) / p.ref_specific_enthalpy;
  hr = h * rho;
}
// the speed of sound squared of the Roe-averaged state: a function of rho_i, T and rho of both sides, not of the velocities
template <int NS>
__device__ __forceinline__ double hllc_roe_c2(const Params<NS>& p, const double* QL, const double* QR) {
  const double rhoL = QL[NS + 5], rhoR = QR[NS + 5];
  const double rho = sqrt(rhoL * rhoR);
  const double sigma = rho / (rhoL + rho);
  double roeQ[NS], Rm, c2;
#pragma unroll
  for (int i = 0; i < NS; i++) roeQ[i] = QL[i] + sigma * (QR[i] - QL[i]);
  const double T = QL[NS + 3] + sigma * (QR[NS + 3] - QL[NS + 3]);
  fluid_props(p, roeQ, T, Rm, c2);
  return c2;
}
// QL / QR need entries [0, NS+6).  The Roe-averaged state only feeds GetFluidProperties (rho_i, T), so its auxiliary
// variables are not formed.  The NaN kneecap of EqnSet::NumericalFlux (eqnset.tcc:73-88) is applied here.
template <int NS>
__device__ __forceinline__ void hllc_assemble(const Params<NS>& p, const double* QL, const double* QR, const double* av,
                                              double vdotn, double beta, double c2L, double hrL, double c2R, double hrR,
                                              double c2, double* flux, double* parts = nullptr) {
  const double uL = QL[NS], vL = QL[NS + 1], wL = QL[NS + 2], pL = QL[NS + 4];
  const double pgL = pL - p.Pref, rhoL = QL[NS + 5];
  const double keL = 0.5 * rhoL * (uL * uL + vL * vL + wL * wL);
  const double HTL = hrL + keL;
  const double ETL = HTL - pL;
  const double thetaL = theta_of<NS>(QL, av, vdotn);
  const double uR = QR[NS], vR = QR[NS + 1], wR = QR[NS + 2], pR = QR[NS + 4];
  const double pgR = pR - p.Pref, rhoR = QR[NS + 5];
  const double keR = 0.5 * rhoR * (uR * uR + vR * vR + wR * wR);
  const double HTR = hrR + keR;
  const double ETR = HTR - pR;
  if (parts) { parts[0] = hrL; parts[1] = keL; parts[2] = hrR; parts[3] = keR; }
  const double thetaR = theta_of<NS>(QR, av, vdotn);

  const double rho = sqrt(rhoL * rhoR);
  const double sigma = rho / (rhoL + rho);
  const double uRoe = uL + sigma * (uR - uL);
  const double vRoe = vL + sigma * (vR - vL);
  const double wRoe = wL + sigma * (wR - wL);
  double theta = (uRoe * av[0] + vRoe * av[1] + wRoe * av[2] + vdotn);

  const double oneMBeta = 1.0 - beta;
  const double thetaPrime = theta * (1.0 + beta) * 0.5;
  const double cPrime = 0.5 * sqrt(theta * theta * (oneMBeta * oneMBeta) + 4.0 * beta * c2);
  const double thetaLPrime = thetaL * (1.0 + beta) * 0.5;
  const double thetaRPrime = thetaR * (1.0 + beta) * 0.5;
  const double cLPrime = 0.5 * sqrt(thetaL * thetaL * (oneMBeta * oneMBeta) + 4.0 * beta * c2L);
  const double cRPrime = 0.5 * sqrt(thetaR * thetaR * (oneMBeta * oneMBeta) + 4.0 * beta * c2R);
  const double eig5L = thetaLPrime - cLPrime;
  const double eig4R = thetaRPrime + cRPrime;
  const double eig4 = thetaPrime + cPrime;
  const double eig5 = thetaPrime - cPrime;
  const double SL = mind(eig5L, eig5);
  const double SR = maxd(eig4R, eig4);
  double SM = (pgR - pgL + rhoL * thetaL * (SL - thetaL) - rhoR * thetaR * (SR - thetaR)) /
              (rhoL * (SL - thetaL) - rhoR * (SR - thetaR));

  double Qf[NS + 4], pStar = 0.0;
  if (SL >= 0.0) {
#pragma unroll
    for (int i = 0; i < NS; i++) Qf[i] = QL[i];
    Qf[NS] = rhoL * uL; Qf[NS + 1] = rhoL * vL; Qf[NS + 2] = rhoL * wL; Qf[NS + 3] = ETL;
    pStar = pgL;
    SM = thetaL;
  } else if (SR <= 0.0) {
#pragma unroll
    for (int i = 0; i < NS; i++) Qf[i] = QR[i];
    Qf[NS] = rhoR * uR; Qf[NS + 1] = rhoR * vR; Qf[NS + 2] = rhoR * wR; Qf[NS + 3] = ETR;
    pStar = pgR;
    SM = thetaR;
  } else if ((SL <= 0.0) && (SM >= 0.0)) {
    pStar = pgL + rhoL * (thetaL - SL) * (thetaL - SM);
    const double omega = 1.0 / (SL - SM);
    const double const1 = SL - thetaL;
    const double const2 = pStar - pgL;
#pragma unroll
    for (int i = 0; i < NS; i++) Qf[i] = omega * const1 * QL[i];
    Qf[NS] = omega * (const1 * rhoL * uL + const2 * av[0]);
    Qf[NS + 1] = omega * (const1 * rhoL * vL + const2 * av[1]);
    Qf[NS + 2] = omega * (const1 * rhoL * wL + const2 * av[2]);
    Qf[NS + 3] = omega * (const1 * ETL - pgL * thetaL + (pStar * SM)) + p.Pref * (SM - thetaL) * omega;
  } else if ((SM <= 0.0) && (SR >= 0.0)) {
    pStar = pgR + rhoR * (thetaR - SR) * (thetaR - SM);
    const double omega = 1.0 / (SR - SM);
    const double const1 = SR - thetaR;
    const double const2 = pStar - pgR;
#pragma unroll
    for (int i = 0; i < NS; i++) Qf[i] = omega * const1 * QR[i];
    Qf[NS] = omega * (const1 * rhoR * uR + const2 * av[0]);
    Qf[NS + 1] = omega * (const1 * rhoR * vR + const2 * av[1]);
    Qf[NS + 2] = omega * (const1 * rhoR * wR + const2 * av[2]);
    Qf[NS + 3] = omega * (const1 * ETR - pgR * thetaR + (pStar * SM)) + p.Pref * (SM - thetaR) * omega;
  } else {   // NaN wave speeds ("HLLC: Should never be here"): every component ends up kneecapped
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int i = 0; i < NS + 4; i++) Qf[i] = qnan;
  }
  const double area = av[3];
  const double Et = Qf[NS + 3];
  theta = SM;
  const double thetabar = theta - vdotn;
#pragma unroll
  for (int i = 0; i < NS; i++) flux[i] = area * Qf[i] * theta;
  flux[NS] = area * (Qf[NS] * theta + pStar * av[0]);
  flux[NS + 1] = area * (Qf[NS + 1] * theta + pStar * av[1]);
  flux[NS + 2] = area * (Qf[NS + 2] * theta + pStar * av[2]);
  flux[NS + 3] = area * (Et * theta + pStar * thetabar) + p.Pref * thetabar * area;
#pragma unroll
  for (int i = 0; i < NS + 4; i++) if (isnan(flux[i])) flux[i] = 0.0;
}
template <int NS>
__device__ __forceinline__ void numerical_flux(const Params<NS>& p, const double* QL, const double* QR, const double* av,
                                               double vdotn, double beta, double* flux, double* parts = nullptr) {
  double c2L, hrL, c2R, hrR;
  hllc_side_thermo(p, QL, c2L, hrL);
  hllc_side_thermo(p, QR, c2R, hrR);
  const double c2 = hllc_roe_c2(p, QL, QR);
  hllc_assemble(p, QL, QR, av, vdotn, beta, c2L, hrL, c2R, hrR, c2, flux, parts);
}

// MaxEigenvalue (compressibleFR.tcc:1603-1637)
template <int NS>
__device__ __forceinline__ double max_eigenvalue(const Params<NS>& p, const double* Q, const double* av, double vdotn,
                                                 double beta) {
  double R, c2;
  fluid_props(p, Q, Q[NS + 3], R, c2);
  const double theta = theta_of<NS>(Q, av, vdotn);
  const double oneMBeta = 1.0 - beta;
  const double thetaPrime = theta * (1.0 + beta) * 0.5;
  const double cPrime = 0.5 * sqrt(theta * theta * (oneMBeta * oneMBeta) + 4.0 * beta * c2);
  const double eig4 = thetaPrime + cPrime;
  const double eig5 = thetaPrime - cPrime;
  return maxd(fabs(eig4), fabs(eig5));
}

// ExtrapolateVariables (compressibleFR.tcc:734-752) with ExtrapolateCorrection (eqnset.h:231-238); g = first NEQ
// gradient rows of the node
template <int NS>
__device__ __forceinline__ void extrapolate(double chi, double* Qho, const double* Q, const double* dQedge,
                                            const double* g, const double* dx, const double* lim) {
#pragma unroll
  for (int i = 0; i < NS + 4; i++) {
    const double corr = (0.5 * chi * dQedge[i] + (1.0 - chi) * (g[i * 3] * dx[0] + g[i * 3 + 1] * dx[1] + g[i * 3 + 2] * dx[2]));
    Qho[i] = Q[i] + corr * lim[i];
  }
}

// BadExtrapolation (compressibleFR.tcc:714-731); Q carries valid P and rho
template <int NS>
__device__ __forceinline__ bool bad_extrapolation(const Params<NS>& p, const double* Q) {
#pragma unroll
  for (int i = 0; i < NS; i++) if (Q[i] < 0.0) return true;
  if (total_energy(p, Q) <= 0.0) return true;
  if (Q[NS + 4] < 1.0e-10) return true;
  if (Q[NS + 3] < 1.0e-10) return true;
  return false;
}

// the same test with h rho and rho |v|^2 / 2 of the state already at hand (from the flux evaluation)
template <int NS>
__device__ __forceinline__ bool bad_extrapolation(const Params<NS>& p, const double* Q, double hr, double ke) {
#pragma unroll
  for (int i = 0; i < NS; i++) if (Q[i] < 0.0) return true;
  const double E = hr - Q[NS + 4];
  if (E + ke <= 0.0) return true;
  if (Q[NS + 4] < 1.0e-10) return true;
  if (Q[NS + 3] < 1.0e-10) return true;
  return false;
}

// ---------------------------------------------------------------- small dense algebra (matrix.h)
// MatVecMult (matrix.h:63-74)
template <int N>
__device__ __forceinline__ void matvec(const double* a, const double* v, double* out) {
  for (int i = 0; i < N; i++) {
    double s = a[i * N + 0] * v[0];
    for (int j = 1; j < N; j++) s += a[i * N + j] * v[j];
    out[i] = s;
  }
}
// LU (matrix.h:110-190): partial pivoting through the permutation p, rows not swapped
template <int N>
__device__ __forceinline__ void lu(double* a, int* p) {
  for (int i = 0; i < N; i++) p[i] = i;
  int row = 0;
  for (int i = 0; i < N; i++) {
    double large = 0.0, lmag = 0.0;   // |large| carried explicitly: see k_lu_diag_lanes (nvcc 12.9 drops the abs otherwise)
    for (int j = i; j < N; j++) {
      const double v = a[p[j] * N + i], vmag = fabs(v);
      if (vmag > lmag) { large = v; lmag = vmag; row = j; }
    }
    const int t = p[i]; p[i] = p[row]; p[row] = t;
    large = 1.0 / large;
    for (int j = i + 1; j < N; j++) a[p[j] * N + i] *= large;
    for (int j = i + 1; j < N; j++) {
      for (int k = i + 1; k < N; k++) a[p[j] * N + k] -= a[p[j] * N + i] * a[p[i] * N + k];
    }
  }
}
// LuSolve (matrix.h:237-264): the solution overwrites b, x is scratch
template <int N>
__device__ __forceinline__ void lu_solve(const double* a, double* b, const int* p, double* x) {
  for (int i = 0; i < N; i++) {
    double sum = 0.0;
    for (int j = 0; j < i; j++) sum += a[p[i] * N + j] * x[j];
    x[i] = b[p[i]] - sum;
  }
  for (int i = N - 1; i >= 0; i--) {
    double sum = 0.0;
    for (int j = N - 1; j > i; j--) sum += a[p[i] * N + j] * b[j];
    b[i] = (x[i] - sum) / a[p[i] * N + i];
  }
}

// ---------------------------------------------------------------- boundary conditions
// PerpVectors (geometry.h:101-129)
__device__ __forceinline__ void perp_vectors(const double* n, double* v1, double* v2) {
  double dot;
  v1[0] = v1[1] = v1[2] = 0.0;
  if (fabs(dot = n[0]) < 0.95) v1[0] = 1.0;
  else if (fabs(dot = n[1]) < 0.95) v1[1] = 1.0;
  else { dot = n[2]; v1[2] = 1.0; }
  v1[0] -= dot * n[0];
  v1[1] -= dot * n[1];
  v1[2] -= dot * n[2];
  double mag = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
  v1[0] = v1[0] / mag; v1[1] = v1[1] / mag; v1[2] = v1[2] / mag;
  v2[0] = n[1] * v1[2] - v1[1] * n[2];
  v2[1] = n[2] * v1[0] - v1[2] * n[0];
  v2[2] = n[0] * v1[1] - v1[0] * n[1];
  mag = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
  v2[0] = v2[0] / mag; v2[1] = v2[1] / mag; v2[2] = v2[2] / mag;
}

// NOTE on __noinline__ (farfield_bc, inviscid_wall_bc, boundary_variables): with these inlined into the
// boundary-Jacobian kernel, nvcc 12.9 (-O3, sm_100a) overlaid the caller's perturbed left state with a temporary of
// the inlined wall BC (the state came back with its v-component overwritten by the BC solution; the same source built
// for the host is clean under ASan/UBSan).  As real calls each gets its own frame; they run per boundary half-edge only.

// Eigensystem (compressibleFR.tcc:150-290) in closed form.  The reference fills two dense N x N arrays T, Tinv; both
// are sparse (an identity block plus the velocity / temperature columns), so here only the scalars they are built from
// are kept and rows are applied on the fly.  Skipping the structurally-zero entries of a row sum is exact: every
// product with a stored 0.0 is a signed zero, and s + (+-0) == s.  The per-species c2i are overwritten by the bulk c2
// in the reference (:176-180).
template <int NS>
struct Eigen {
  double theta, c2, cPrime, thetaPrime, Xp, Xm, rho, bm1;
  double l[3], m[3], n[3];
};

template <int NS>
__device__ __forceinline__ void eigen_setup(const Params<NS>& p, const double* Q, const double* av, double vdotn, double beta,
                                            Eigen<NS>& E) {
  const double bp1 = beta + 1.0, oneMBeta = 1.0 - beta;
  E.bm1 = beta - 1.0;
  E.n[0] = av[0]; E.n[1] = av[1]; E.n[2] = av[2];
  E.theta = theta_of<NS>(Q, av, vdotn);
  E.rho = Q[NS + 5];
  double R;
  fluid_props(p, Q, Q[NS + 3], R, E.c2);
  E.thetaPrime = 0.5 * bp1 * E.theta;
  E.cPrime = 0.5 * sqrt(E.theta * E.theta * (oneMBeta * oneMBeta) + 4.0 * beta * E.c2);
  perp_vectors(av, E.l, E.m);
  const double betam = oneMBeta * 0.5;
  E.Xp = E.theta * betam + E.cPrime;
  E.Xm = E.theta * betam - E.cPrime;
}
// eigenvalue i (compressibleFR.tcc:199-203)
template <int NS>
__device__ __forceinline__ double eigen_value(const Eigen<NS>& E, int i) {
  if (i < NS + 2) return E.theta;
  return (i == NS + 2) ? (E.thetaPrime + E.cPrime) : (E.thetaPrime - E.cPrime);
}
// the non-zero entries of row i of Tinv (compressibleFR.tcc:240-273): columns u, v, w, t (+ the unit diagonal for a
// species row); returns the number of leading species columns that are non-zero (1 for species rows, else 0)
template <int NS>
__device__ __forceinline__ void tinv_row(const Eigen<NS>& E, const double* Q, int i, double& du, double& dv, double& dw,
                                         double& dt) {
  const double lx = E.l[0], ly = E.l[1], lz = E.l[2], mx = E.m[0], my = E.m[1], mz = E.m[2];
  const double nx = E.n[0], ny = E.n[1], nz = E.n[2];
  const double c2 = E.c2, Xm = E.Xm, Xp = E.Xp, bm1 = E.bm1, theta = E.theta, rho = E.rho, cPrime = E.cPrime;
  if (i < NS) {
    const double KK = -Q[i] * (c2 * (Xm + Xp) - bm1 * Xm * Xp * theta + bm1 * c2 * (Xm + Xp));
    du = -((ly * mz - lz * my) * KK) / (c2 * Xm * Xp);
    dv = ((lx * mz - lz * mx) * KK) / (c2 * Xm * Xp);
    dw = -((lx * my - ly * mx) * KK) / (c2 * Xm * Xp);
    dt = (Q[i] * (c2 + bm1 * c2)) / (rho * c2 * Xm * Xp);
  } else if (i == NS) {
    du = my * nz - mz * ny;
    dv = -(mx * nz - mz * nx);
    dw = mx * ny - my * nx;
    dt = 0.0;
  } else if (i == NS + 1) {
    du = -(ly * nz - lz * ny);
    dv = lx * nz - lz * nx;
    dw = -(lx * ny - ly * nx);
    dt = 0.0;
  } else if (i == NS + 2) {
    du = ((Xp) * (ly * mz - lz * my)) / (2.0 * cPrime);
    dv = -((Xp) * (lx * mz - lz * mx)) / (2.0 * cPrime);
    dw = ((Xp) * (lx * my - ly * mx)) / (2.0 * cPrime);
    dt = 1.0 / (2.0 * rho * cPrime);
  } else {
    du = ((Xm) * (ly * mz - lz * my)) / (2.0 * cPrime);
    dv = -((Xm) * (lx * mz - lz * mx)) / (2.0 * cPrime);
    dw = ((Xm) * (lx * my - ly * mx)) / (2.0 * cPrime);
    dt = 1.0 / (2.0 * rho * cPrime);
  }
}
// rhs[i] = sum_j Tinv[i][j] * val[j], j ascending from 0.0 (compressibleFR.tcc:988-993, 1090-1095)
template <int NS>
__device__ __forceinline__ double tinv_row_dot(const Eigen<NS>& E, const double* Q, int i, const double* val) {
  double du, dv, dw, dt;
  tinv_row(E, Q, i, du, dv, dw, dt);
  double s = 0.0;
  if (i < NS) s += 1.0 * val[i];
  s += du * val[NS];
  s += dv * val[NS + 1];
  s += dw * val[NS + 2];
  if (i != NS && i != NS + 1) s += dt * val[NS + 3];
  return s;
}
// (T v)[i] in MatVecMult order (matrix.h:63-74; T: compressibleFR.tcc:213-237)
template <int NS>
__device__ __forceinline__ double t_row_dot(const Eigen<NS>& E, const double* Q, int i, const double* v) {
  const double c2 = E.c2, Xm = E.Xm, Xp = E.Xp, bm1 = E.bm1, theta = E.theta, rho = E.rho;
  if (i < NS) {
    const double tw = -(Q[i] * (c2 + bm1 * c2 - theta * bm1 * Xm)) / (c2 * Xm);
    const double tt = (Q[i] * (c2 + bm1 * c2 - theta * bm1 * Xp)) / (c2 * Xp);
    double s = 1.0 * v[i];
    s += tw * v[NS + 2];
    s += tt * v[NS + 3];
    return s;
  }
  if (i < NS + 3) {
    const int k = i - NS;
    double s = E.l[k] * v[NS];
    s += E.m[k] * v[NS + 1];
    s += E.n[k] * v[NS + 2];
    s += -E.n[k] * v[NS + 3];
    return s;
  }
  double s = -rho * Xm * v[NS + 2];
  s += rho * Xp * v[NS + 3];
  return s;
}

// NewtonFindTGivenP (compressibleFR.tcc:2343-2378)
template <int NS>
__device__ double newton_T_given_P(const Params<NS>& p, const double* rhoi, double Pgoal, double Tinit) {
  double TDim = Tinit * p.ref_temperature;
  const double PgoalDim = Pgoal * p.ref_pressure;
  double rhoiDim[NS];
  for (int i = 0; i < NS; i++) rhoiDim[i] = rhoi[i] * p.ref_density;
  for (int j = 0; j < 28; j++) {
    const double TpDim = TDim + 1.0e-8;
    const double PDim = chem_P(p, rhoiDim, TDim);
    const double PpDim = chem_P(p, rhoiDim, TpDim);
    const double zpoint = PgoalDim - PDim;
    const double zpointp = PgoalDim - PpDim;
    const double dzdT = (zpointp - zpoint) / (TpDim - TDim);
    const double dT = -zpoint / dzdT;
    if (fabs(dT / p.ref_temperature) < 1.0e-15) break;
    else TDim += dT;
  }
  return TDim / p.ref_temperature;
}

constexpr int N_SUBIT = 10;   // compressibleFR.tcc:32

// GetFarfieldBoundaryVariables (compressibleFR.tcc:940-1039)
template <int NS>
__device__ __noinline__ void farfield_bc(const Params<NS>& p, const double* QL, double* QR, const double* av, double vdotn,
                                         double beta, const double* Qinf) {
  // Qinf: the free stream handed over by the caller (p.qinf, or the power-law scaled copy of the viscous far field)
  constexpr int N = NS + 4, NV = 3 * NS + 6;
  double qavg[NS + 6], rhs[N], ql[N], qinf[N];
  Eigen<NS> E;
  for (int subit = 0; subit < N_SUBIT; subit++) {
    for (int i = 0; i < N; i++) qavg[i] = 0.5 * (QL[i] + QR[i]);
    aux_pr(p, qavg);
    eigen_setup(p, qavg, av, vdotn, beta, E);
    if (p.no_cvbc) {
      if (E.theta >= 0.0) { for (int i = 0; i < NV; i++) QR[i] = QL[i]; }
      else { for (int i = 0; i < NV; i++) QR[i] = Qinf[i]; }
    } else {
      for (int i = 0; i < N; i++) { ql[i] = QL[i]; qinf[i] = Qinf[i]; }
      const double Tguess = ql[N - 1];
      ql[N - 1] = QL[NS + 4];
      qinf[N - 1] = Qinf[NS + 4];
#pragma unroll
      for (int i = 0; i < N; i++) rhs[i] = tinv_row_dot(E, qavg, i, (eigen_value(E, i) >= 0.0) ? ql : qinf);
#pragma unroll
      for (int i = 0; i < N; i++) QR[i] = t_row_dot(E, qavg, i, rhs);
      for (int i = 0; i < NS; i++) if (QR[i] < 0.0) QR[i] = 0.0;
      const double pgoal = QR[N - 1];
      QR[N - 1] = newton_T_given_P(p, QR, pgoal, Tguess);
    }
  }
}

// The characteristic solve of the slip wall: LU (matrix.h:110-190) + LuSolve (:237-264) of Tinv with its last row
// replaced by [0.., n, 0].  The species columns of that matrix hold a unit diagonal and nothing else, so the partial
// pivoting of the reference picks row i for every species column i, its multipliers are exact zeros and the
// elimination leaves the rest untouched: the N x N factorisation IS the pivoted LU of the trailing 4 x 4 block
// (rows / columns u, v, w, T), and the species unknowns follow by back-substitution over the columns T, w, v, u (the
// reference's descending order).  Same operations on the same operands, 16 matrix entries instead of (NS+4)^2.
template <int NS>
__device__ __forceinline__ void wall_solve(const Eigen<NS>& E, const double* Q, const double* av, double* rhs) {
  constexpr int N = NS + 4;
  double M[4][4];
  int p[4] = {0, 1, 2, 3};
#pragma unroll
  for (int r = 0; r < 3; r++) tinv_row(E, Q, NS + r, M[r][0], M[r][1], M[r][2], M[r][3]);
  M[3][0] = av[0]; M[3][1] = av[1]; M[3][2] = av[2]; M[3][3] = 0.0;
  int row = 0;   // (only read if a pivot column is entirely zero, i.e. for a singular system)
#pragma unroll
  for (int i = 0; i < 4; i++) {
    // |large| is carried explicitly: with `fabs(v) > fabs(large)` nvcc 12.9 emitted the second comparison of every pivot
    // step as DSETP.GT |v|, large -- abs on large dropped (see k_lu_diag_lanes, pcfd_internal.cuh)
    double large = 0.0, lmag = 0.0;
#pragma unroll
    for (int j = i; j < 4; j++) {
      const double v = M[p[j]][i], vmag = fabs(v);
      if (vmag > lmag) { large = v; lmag = vmag; row = j; }
    }
    const int t = p[i]; p[i] = p[row]; p[row] = t;
    large = 1.0 / large;
#pragma unroll
    for (int j = i + 1; j < 4; j++) M[p[j]][i] *= large;
#pragma unroll
    for (int j = i + 1; j < 4; j++)
#pragma unroll
      for (int k = i + 1; k < 4; k++) M[p[j]][k] -= M[p[j]][i] * M[p[i]][k];
  }
  double x[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = rhs[NS + i];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < i; j++) sum += M[p[i]][j] * x[j];
    x[i] = b[p[i]] - sum;
  }
#pragma unroll
  for (int i = 3; i >= 0; i--) {
    double sum = 0.0;
#pragma unroll
    for (int j = 3; j > i; j--) sum += M[p[i]][j] * b[j];
    b[i] = (x[i] - sum) / M[p[i]][i];
  }
#pragma unroll
  for (int i = 0; i < 4; i++) rhs[NS + i] = b[i];
#pragma unroll
  for (int i = NS - 1; i >= 0; i--) {
    double du, dv, dw, dt;
    tinv_row(E, Q, i, du, dv, dw, dt);
    double sum = 0.0;
    sum += dt * b[3];
    sum += dw * b[2];
    sum += dv * b[1];
    sum += du * b[0];
    rhs[i] = (rhs[i] - sum) / 1.0;
  }
  (void)N;
}

// GetInviscidWallBoundaryVariables (compressibleFR.tcc:1042-1134)
template <int NS>
__device__ __noinline__ void inviscid_wall_bc(const Params<NS>& p, const double* QL, double* QR, const double* av,
                                              double vdotn, double beta) {
  constexpr int N = NS + 4, NV = 3 * NS + 6;
  if (!p.no_cvbc) {
    double qavg[NS + 6], rhs[N], ql[N];
    Eigen<NS> E;
    for (int subit = 0; subit < N_SUBIT; subit++) {
      for (int i = 0; i < N; i++) qavg[i] = 0.5 * (QL[i] + QR[i]);
      aux_pr(p, qavg);
      eigen_setup(p, qavg, av, vdotn, beta, E);
      for (int i = 0; i < N; i++) ql[i] = QL[i];
      const double Tguess = ql[N - 1];
      ql[N - 1] = QL[NS + 4];
#pragma unroll
      for (int i = 0; i < N; i++) rhs[i] = tinv_row_dot(E, qavg, i, ql);
      rhs[N - 1] = vdotn;
      wall_solve(E, qavg, av, rhs);
      for (int i = 0; i < N; i++) QR[i] = rhs[i];
      const double pgoal = QR[N - 1];
      QR[N - 1] = newton_T_given_P(p, QR, pgoal, Tguess);
    }
  } else {
    double QLmod[NV];
    for (int i = 0; i < NV; i++) QLmod[i] = QL[i];
    QLmod[NS] += vdotn * av[0];
    QLmod[NS + 1] += vdotn * av[1];
    QLmod[NS + 2] += vdotn * av[2];
    for (int i = 0; i < N; i++) QR[i] = QLmod[i];
    const double dot = 2.0 * (QLmod[NS] * av[0] + QLmod[NS + 1] * av[1] + QLmod[NS + 2] * av[2]);   // MirrorVector
    QR[NS] = QLmod[NS] - dot * av[0];
    QR[NS + 1] = QLmod[NS + 1] - dot * av[1];
    QR[NS + 2] = QLmod[NS + 2] - dot * av[2];
  }
}

// CalculateBoundaryVariables (bc.tcc:1058-1397) for the BC types of the reacting configs; QL and QR are full rows
// Proteus_FarFieldViscous (bc.tcc:1092-1108): see eq::boundary_variables; GetMomentumLocation() == nspecies here
template <int NS>
__device__ __noinline__ void boundary_variables(const Params<NS>& p, double* QL, double* QR, const double* av, int bctype,
                                                double betaL, double ubar = 1.0, double* qref = nullptr) {
  constexpr int N = NS + 4, NV = 3 * NS + 6;
  const double vdotn = 0.0;   // static mesh
  switch (bctype) {
    case PCFD_BC_PARALLEL: return;
    case PCFD_BC_FARFIELD_VISCOUS: {
      double fresh[NV];
      double* Qinf = qref ? qref : fresh;
      if (!qref) for (int i = 0; i < NV; i++) fresh[i] = p.qinf[i];
      if (ubar < 1.0) for (int i = 0; i < 3; i++) Qinf[NS + i] = ubar * Qinf[NS + i];
      farfield_bc(p, QL, QR, av, vdotn, betaL, Qinf);
      break;
    }
    case PCFD_BC_SONIC_OUTFLOW: case PCFD_BC_NEUMANN:
      for (int i = 0; i < N; i++) QR[i] = QL[i];
      break;
    case PCFD_BC_FARFIELD:
      farfield_bc(p, QL, QR, av, vdotn, betaL, p.qinf);
      break;
    case PCFD_BC_IMPERMEABLE_WALL: case PCFD_BC_SYMMETRY:
      inviscid_wall_bc(p, QL, QR, av, vdotn, betaL);
      break;
    default: break;
  }
  aux(p, QR);
  aux(p, QL);
}

// ---------------------------------------------------------------- update
// GetViscousWallBoundaryVariables (compressibleFR.tcc:2048-2070), static wall (vel = 0 after bc.tcc:1207-1254 with no
// movement, no bleed steps, no grid speed): species densities kept, wall velocity, wall temperature = Twall or, for an
// adiabatic wall (Twall < 0), the temperature of the most-normal neighbour.  Rewrites QL itself (hard-set wall node).
template <int NS>
__device__ __forceinline__ void viscous_wall_bc(double* QL, double* QR, double normalT, double Twall) {
#pragma unroll
  for (int i = 0; i < NS; i++) QR[i] = QL[i];
  if (Twall < 0.0) QR[NS + 3] = QL[NS + 3] = normalT;
  else QR[NS + 3] = QL[NS + 3] = Twall;
  QR[NS] = QL[NS] = 0.0;
  QR[NS + 1] = QL[NS + 1] = 0.0;
  QR[NS + 2] = QL[NS + 2] = 0.0;
}

// CalculateBoundaryVariables (bc.tcc:1058-1397) including the BC types that rewrite the interior state: nodes owning
// such a half-edge are walked sequentially (kfr_update_bcs_nodes, kfr_jac_bnodes)
template <int NS>
__device__ __noinline__ void boundary_variables_seq(const Params<NS>& p, double* QL, double* QR, const double* av, int bctype,
                                                    double betaL, double normalT, double Twall, double ubar = 1.0,
                                                    double* qref = nullptr) {
  constexpr int NV = 3 * NS + 6;
  switch (bctype) {
    case PCFD_BC_SONIC_INFLOW: case PCFD_BC_DIRICHLET:
      for (int i = 0; i < NV; i++) QR[i] = QL[i] = p.qinf[i];
      break;
    case PCFD_BC_NOSLIP:
      viscous_wall_bc<NS>(QL, QR, normalT, Twall);
      break;
    default:
      boundary_variables(p, QL, QR, av, bctype, betaL, ubar, qref);
      return;
  }
  aux(p, QR);
  aux(p, QL);
}

// ApplyDQ (compressibleFR.tcc:886-937)
template <int NS>
__device__ __forceinline__ void apply_dq(const Params<NS>& p, const double* dQ, double* Q) {
  for (int i = 0; i < NS; i++) {
    const double rho = Q[i] + dQ[i];
    if (!(rho < 0.0)) Q[i] += dQ[i];   // a negative projected density: the update is refused
  }
  const double projectedT = Q[NS + 3] + dQ[NS + 3];
  if (projectedT < 0.0) Q[NS + 3] = 1.0e-10;
  else Q[NS + 3] = projectedT;
  Q[NS] += dQ[NS];
  Q[NS + 1] += dQ[NS + 1];
  Q[NS + 2] += dQ[NS + 2];
  aux(p, Q);
}

// NativeToConservative (compressibleFR.tcc:2117-2130)
template <int NS>
__device__ __forceinline__ void native_to_conservative(const Params<NS>& p, double* Q) {
  const double rho = Q[NS + 5];
  const double Et = total_energy(p, Q);
  Q[NS] *= rho;
  Q[NS + 1] *= rho;
  Q[NS + 2] *= rho;
  Q[NS + 3] = Et;
}
// ConservativeToNative (compressibleFR.tcc:2133-2201); returns false if the Newton iteration did not converge
template <int NS>
__device__ bool conservative_to_native(const Params<NS>& p, double* Q) {
  double rho = 0.0, Y[NS], rhoiDim[NS], R = 0.0;
  for (int i = 0; i < NS; i++) {
    rho += Q[i];
    rhoiDim[i] = Q[i] * p.ref_density;
    R += rhoiDim[i] * p.Rs[i];
  }
  R /= p.ref_density * rho;
  for (int i = 0; i < NS; i++) Y[i] = Q[i] / rho;
  const double u = Q[NS] / rho, v = Q[NS + 1] / rho, w = Q[NS + 2] / rho;
  const double v2 = u * u + v * v + w * w;
  const double res = Q[NS + 3] - 0.5 * v2 * rho;
  const double P = Q[NS + 4];
  double T = ((P * p.ref_pressure) / ((rho * p.ref_density) * R)) / p.ref_temperature;   // IdealGasEOS::GetT
  int j = 0;
  for (; j < 20; j++) {
    const double Tp = T + 1.0e-8;
    const double H = rho * (chem_h(p, Y, T * p.ref_temperature) / p.ref_specific_enthalpy);
    const double Hp = rho * (chem_h(p, Y, Tp * p.ref_temperature) / p.ref_specific_enthalpy);
    const double Pn = chem_P(p, rhoiDim, T * (p.ref_temperature)) / p.ref_pressure;
    const double Pp = chem_P(p, rhoiDim, Tp * (p.ref_temperature)) / p.ref_pressure;
    const double E = H - Pn;
    const double Ep = Hp - Pp;
    const double zpoint = res - E;
    const double zpointp = res - Ep;
    const double dzdT = (zpointp - zpoint) / (Tp - T);
    const double dT = -zpoint / dzdT;
    T += dT;
    if (fabs(dT) < 1.0e-12) break;
  }
  Q[NS] = u; Q[NS + 1] = v; Q[NS + 2] = w;
  Q[NS + 3] = T;
  return j != 20;
}

// SourceTerm with the rate constants of the state's temperature handed in (chemdev::rate_constants); rxn_on assumed
template <int NS>
__device__ __forceinline__ void source_term_rates(const Params<NS>& p, const double* Q, double vol, const double* Kf,
                                                  const double* Kb, double* source) {
  double rhoi[PCFD_CHEM_MAX_SPECIES], wdot[PCFD_CHEM_MAX_SPECIES];
#pragma unroll
  for (int i = 0; i < NS; i++) rhoi[i] = Q[i] * p.ref_density;
  chemdev::mass_production_rates(p.chem, rhoi, Kf, Kb, wdot);
#pragma unroll
  for (int i = 0; i < NS; i++) {
    double w = wdot[i];
    w /= (p.ref_density / p.ref_time);
    source[i] = vol * w;
  }
}

// SourceTerm (compressibleFR.tcc:1276-1316), species rows only (the others are zero; gravity off)
template <int NS>
__device__ __forceinline__ void source_term(const Params<NS>& p, const double* Q, double vol, double* source) {
  if (!p.rxn_on) {
#pragma unroll
    for (int i = 0; i < NS; i++) source[i] = 0.0;
    return;
  }
  double rhoi[PCFD_CHEM_MAX_SPECIES], wdot[PCFD_CHEM_MAX_SPECIES];
  const double T = Q[NS + 3] * p.ref_temperature;
#pragma unroll
  for (int i = 0; i < NS; i++) rhoi[i] = Q[i] * p.ref_density;
  chemdev::mass_production(p.chem, rhoi, T, wdot);
#pragma unroll
  for (int i = 0; i < NS; i++) {
    double w = wdot[i];
    w /= (p.ref_density / p.ref_time);
    source[i] = vol * w;
  }
}

// ChemModel::dRmixdRhoi (chem.tcc:861-873)
template <int NS>
__device__ __forceinline__ double dRmixdRhoi(const Params<NS>& p, const double* rhoi, double rho, int i) {
  double d = p.Rs[i] * (rho - rhoi[i]) / (rho * rho);
  for (int j = 0; j < NS; j++) {
    if (j == i) continue;
    d -= p.Rs[j] * rhoi[j] / (rho * rho);
  }
  return d;
}

// ====================================================================== transport + viscous terms (compressibleNSFR)
// Species transport data as Species holds it (species.h:35-49): Sutherland law (White) up to the transition
// temperature, NASA RP-1311 fits [Tlo, Thi, A, B, C, D] above.  The two molecular-weight powers of Wilke's rule
// (chem.tcc:908-909) do not depend on the state: they are evaluated once on the host with the C library's pow.
template <int NS>
struct Transport {
  double mu_fit[NS][3][6], k_fit[NS][3][6];
  double mu_white[NS][4], k_white[NS][4];     // reference value, T0, S, transition temperature
  int nmu[NS], nk[NS];
  double pw25[NS][NS];                        // pow(MW_j / MW_i, 0.25)
  double pwm05[NS][NS];                       // pow(1 + MW_i / MW_j, -0.5)
  double phi_ii[NS];                          // Wilke's phi for j == i: visc_i / visc_i == 1, nothing of the state is left
  double sqrt8;
  double ref_viscosity, ref_k, Re, PrT;
  int white_uniform;                          // every species carries the same Sutherland rows (the reference's do,
                                              // species.tcc:13-22) and mu / k share T0: one pow serves all ten fits
};

// Species::GetViscosity / GetThermalConductivity (species.tcc:393-479); the range search keeps the last match.
// logT = log(T) is passed in: the same argument gives the same value, so it is formed once per state.
__device__ __forceinline__ double sp_transport(const double* white, const double (*fit)[6], int nfit, double T, double logT,
                                               double conv) {
  if (T <= white[3]) {
    const double v0 = white[0], T0 = white[1], S = white[2];
    return v0 * (pow(T / T0, 1.5)) * ((T0 + S) / (T + S));
  }
  int range = -1;
  for (int i = 0; i < nfit; i++)
    if (T >= fit[i][0] && T <= fit[i][1]) range = i;
  if (range == -1) return nan("");   // the reference aborts
  const double A = fit[range][2], B = fit[range][3], C = fit[range][4], D = fit[range][5];
  const double logv = A * logT + B / T + C / (T * T) + D;
  return exp(logv) * conv;
}

// ChemModel::GetViscosity / GetThermalConductivity -> WilkesMixtureRule (chem.tcc:876-938) with
// MassFractionToMoleFraction (:941-958), through CompressibleFREqnSet::GetMolecularViscosity / GetThermalConductivity
// (compressibleFR.tcc:1580-1599).  Both properties are mixed with the same weights (the species viscosities), so the
// weights are formed once; rhoi and T non-dimensional, results non-dimensional.
template <int NS>
__device__ __forceinline__ void mixture_transport(const Params<NS>& p, const Transport<NS>& t, const double* rhoi, double T,
                                                  double& mu, double& k) {
  const double Td = T * p.ref_temperature;
  double rho = 0.0, summ = 0.0, mf[NS], visc[NS], cond[NS];
#pragma unroll
  for (int i = 0; i < NS; i++) { mf[i] = rhoi[i] * p.ref_density; rho += mf[i]; }
#pragma unroll
  for (int i = 0; i < NS; i++) {
    const double massfrac = mf[i] / rho;
    mf[i] = (massfrac) / p.mw[i];
    summ += mf[i];
  }
#pragma unroll
  for (int i = 0; i < NS; i++) mf[i] /= summ;
  summ = 0.0;
#pragma unroll
  for (int i = 0; i < NS - 1; i++) summ += mf[i];
  mf[NS - 1] = 1.0 - summ;
  if (t.white_uniform && Td <= t.mu_white[0][3]) {
    // Sutherland branch with identical rows for every species: identical operands, identical results -- one pow
    const double p15 = pow(Td / t.mu_white[0][1], 1.5);
    const double vm = t.mu_white[0][0] * (p15) * ((t.mu_white[0][1] + t.mu_white[0][2]) / (Td + t.mu_white[0][2]));
    const double vk = t.k_white[0][0] * (p15) * ((t.k_white[0][1] + t.k_white[0][2]) / (Td + t.k_white[0][2]));
#pragma unroll
    for (int i = 0; i < NS; i++) { visc[i] = vm; cond[i] = vk; }
  } else {
    const double logT = log(Td);
#pragma unroll
    for (int i = 0; i < NS; i++) {
      visc[i] = sp_transport(t.mu_white[i], t.mu_fit[i], t.nmu[i], Td, logT, 1.0e-7);
      cond[i] = sp_transport(t.k_white[i], t.k_fit[i], t.nk[i], Td, logT, 0.0001);
    }
  }
  double mmu = 0.0, mk = 0.0;
#pragma unroll
  for (int i = 0; i < NS; i++) {
    double wi = 0.0;
#pragma unroll
    for (int j = 0; j < NS; j++) {
      double phi;
      if (j == i) {
        phi = t.phi_ii[i];     // visc_i / visc_i == 1 exactly (a NaN viscosity still poisons wi through the other terms)
      } else {
        // equal viscosities (always, on the Sutherland branch): the ratio is exactly 1 and so is its square root
        const double rt = (visc[i] == visc[j]) ? 1.0 : sqrt(visc[i] / visc[j]);
        const double temp = (1.0 + rt * t.pw25[i][j]);
        phi = t.pwm05[i][j] * temp * temp / t.sqrt8;
      }
      wi += mf[j] * phi;
    }
    mmu += (mf[i] / wi) * visc[i];
    mk += (mf[i] / wi) * cond[i];
  }
  mu = mmu / t.ref_viscosity;
  k = mk / t.ref_k;
}

// GetFluidProperties as above, returning the mixture cv, cp and R (non-dimensional) the viscous terms read
template <int NS>
__device__ __forceinline__ void fluid_props_cvcp(const Params<NS>& p, const double* rhoi, double T, double& cv, double& cp,
                                                 double& R) {
  const double s_ref = (p.ref_velocity * p.ref_velocity / p.ref_temperature);
  double rhoiDim[NS];
#pragma unroll
  for (int i = 0; i < NS; i++) rhoiDim[i] = rhoi[i] * p.ref_density;
  const double Tdim = T * p.ref_temperature;
  double RDim = 0.0, rhoDim = 0.0, cpDim = 0.0, cvDim = 0.0;
#pragma unroll
  for (int i = 0; i < NS; i++) rhoDim += rhoiDim[i];
#pragma unroll
  for (int i = 0; i < NS; i++) RDim += rhoiDim[i] * p.Rs[i];
  RDim /= rhoDim;
  const int rng = thermo_range(Tdim);
#pragma unroll
  for (int i = 0; i < NS; i++) {
    const double X = rhoiDim[i] / rhoDim;
    const double cpi = sp_cp(p, i, rng, Tdim);
    cpDim += cpi * X;
    cvDim += X * (cpi - p.Rs[i]);
  }
  cv = cvDim / s_ref;
  cp = cpDim / s_ref;
  R = RDim / s_ref;
}

// CompressibleFREqnSet::ViscousFlux (compressibleFR.tcc:551-637): Q = [rho_i, u, v, w, T] of the face, g = gradients of
// u, v, w, T (12 doubles); f = the momentum and energy rows (the species rows are zero)
template <int NS>
__device__ __forceinline__ void viscous_flux(const Params<NS>& p, const Transport<NS>& t, const double* Q, const double* g,
                                             const double* av, double mut, double* f) {
  const double ux = g[0], uy = g[1], uz = g[2], vx = g[3], vy = g[4], vz = g[5], wx = g[6], wy = g[7], wz = g[8];
  const double Tx = g[9], Ty = g[10], Tz = g[11];
  const double u = Q[NS], v = Q[NS + 1], w = Q[NS + 2], T = Q[NS + 3];
  double cv, cp, R, mu, kc;
  fluid_props_cvcp(p, Q, T, cv, cp, R);
  mixture_transport(p, t, Q, T, mu, kc);
  const double tmut = (mu + mut);
  const double fact = 2.0 / 3.0;
  const double tauxx = 2.0 * fact * ux - fact * vy - fact * wz;
  const double tauyy = 2.0 * fact * vy - fact * ux - fact * wz;
  const double tauzz = 2.0 * fact * wz - fact * ux - fact * vy;
  const double tauxy = uy + vx;
  const double tauxz = uz + wx;
  const double tauyz = vz + wy;
  const double RK = av[3] / t.Re;
  const double RKT = RK * tmut;
  double k = -kc;
  const double Tn = Tx * av[0] + Ty * av[1] + Tz * av[2];
  const double kT = (cp * mut) / t.PrT;
  k -= kT;
  const double tauxn = -tauxx * av[0] - tauxy * av[1] - tauxz * av[2];
  const double tauyn = -tauxy * av[0] - tauyy * av[1] - tauyz * av[2];
  const double tauzn = -tauxz * av[0] - tauyz * av[1] - tauzz * av[2];
  f[0] = RKT * (tauxn);
  f[1] = RKT * (tauyn);
  f[2] = RKT * (tauzn);
  f[3] = RKT * (tauxn * u + tauyn * v + tauzn * w) + RK * k * Tn;
}

// the part of CompressibleFREqnSet::ViscousJacobian (compressibleFR.tcc:1713-2040) shared by both sides: averaged state,
// its viscosity / conductivity / cp
template <int NS>
struct ViscJacCommon {
  double u, v, w, RK, RKT, c1;
};
template <int NS>
__device__ __forceinline__ void viscous_jac_common(const Params<NS>& p, const Transport<NS>& t, const double* QL,
                                                   const double* QR, const double* av, double mut, ViscJacCommon<NS>& C) {
  double Q[NS + 4];
#pragma unroll
  for (int i = 0; i < NS + 4; i++) Q[i] = 0.5 * (QL[i] + QR[i]);
  const double T = Q[NS + 3];
  double mu, k, cv, cp, R;
  mixture_transport(p, t, Q, T, mu, k);
  const double tmut = (mu + mut);
  C.RK = av[3] / t.Re;
  C.RKT = C.RK * tmut;
  fluid_props_cvcp(p, Q, T, cv, cp, R);
  C.u = 0.5 * (QL[NS] + QR[NS]);
  C.v = 0.5 * (QL[NS + 1] + QR[NS + 1]);
  C.w = 0.5 * (QL[NS + 2] + QR[NS + 2]);
  const double kT = mut / t.PrT * cp;
  C.c1 = -(k + kT);
}

// one side (:1796-1960 right with D = dx/s2, :1963-2010 left with D = -dx/s2): rows NS..NS+3 of the block, a[4][NS+4]
// (the species rows are zero); Qs holds [0, NS+6) of that side's node (stored aux values P and rho)
template <int NS>
__device__ __forceinline__ void viscous_jac_side(const Params<NS>& p, const ViscJacCommon<NS>& C, const double* D,
                                                 const double* Qs, const double* av, double* a) {
  constexpr int NEQ = NS + 4;
  const double u = C.u, v = C.v, w = C.w, RK = C.RK, RKT = C.RKT;
  const double rhos = Qs[NS + 5], us = Qs[NS], vs = Qs[NS + 1], ws = Qs[NS + 2], Ps = Qs[NS + 4], Ts = Qs[NS + 3];
  double cvs, cps, Rs;
  fluid_props_cvcp(p, Qs, Ts, cvs, cps, Rs);
  double Dr[3];
#pragma unroll
  for (int i = 0; i < 3; i++) Dr[i] = D[i] / rhos;
  const double dux = -u * Dr[0], duy = -u * Dr[1], duz = -u * Dr[2];
  const double dvx = -v * Dr[0], dvy = -v * Dr[1], dvz = -v * Dr[2];
  const double dwx = -w * Dr[0], dwy = -w * Dr[1], dwz = -w * Dr[2];
  const double dfact = -2.0 / 3.0 * (dux + dvy + dwz);
  const double dtauxx = (2.0 * dux + dfact);
  const double dtauyy = (2.0 * dvy + dfact);
  const double dtauzz = (2.0 * dwz + dfact);
  const double dtauxy = duy + dvx;
  const double dtauxz = duz + dwx;
  const double dtauyz = dvz + dwy;
  const double dtauxn = dtauxx * av[0] + dtauxy * av[1] + dtauxz * av[2];
  const double dtauyn = dtauxy * av[0] + dtauyy * av[1] + dtauyz * av[2];
  const double dtauzn = dtauxz * av[0] + dtauyz * av[1] + dtauzz * av[2];
  const double c43 = 4.0 / 3.0, mc23 = -2.0 / 3.0;
  double* row = a;
  const double dR2_drhou = (c43 * Dr[0] * av[0] + Dr[1] * av[1] + Dr[2] * av[2]);
  const double dR2_drhov = (mc23 * Dr[1] * av[0] + Dr[0] * av[1]);
  const double dR2_drhow = (mc23 * Dr[2] * av[0] + Dr[0] * av[2]);
#pragma unroll
  for (int i = 0; i < NS; i++) row[i] = -RKT * (dtauxn);
  row[NS + 0] = -RKT * dR2_drhou;
  row[NS + 1] = -RKT * dR2_drhov;
  row[NS + 2] = -RKT * dR2_drhow;
  row[NS + 3] = 0.0;
  row = a + NEQ;
  const double dR3_drhou = (mc23 * Dr[0] * av[1] + Dr[1] * av[0]);
  const double dR3_drhov = (Dr[0] * av[0] + c43 * Dr[1] * av[1] + Dr[2] * av[2]);
  const double dR3_drhow = (mc23 * Dr[2] * av[1] + Dr[1] * av[2]);
#pragma unroll
  for (int i = 0; i < NS; i++) row[i] = -RKT * (dtauyn);
  row[NS + 0] = -RKT * dR3_drhou;
  row[NS + 1] = -RKT * dR3_drhov;
  row[NS + 2] = -RKT * dR3_drhow;
  row[NS + 3] = 0.0;
  row = a + 2 * NEQ;
  const double dR4_drhou = (mc23 * Dr[0] * av[2] + Dr[2] * av[0]);
  const double dR4_drhov = (mc23 * Dr[1] * av[2] + Dr[2] * av[1]);
  const double dR4_drhow = (Dr[0] * av[0] + Dr[1] * av[1] + c43 * Dr[2] * av[2]);
#pragma unroll
  for (int i = 0; i < NS; i++) row[i] = -RKT * (dtauzn);
  row[NS + 0] = -RKT * dR4_drhou;
  row[NS + 1] = -RKT * dR4_drhov;
  row[NS + 2] = -RKT * dR4_drhow;
  row[NS + 3] = 0.0;
  // IdealGasEOS::GetdT_dP (EOS.tcc:42-45); ChemModel::dTdRhoi (chem.tcc:687-704) on dimensional densities / pressure
  const double dT_dP = (1.0 / (rhos * Rs));
  double rhoidim[NS], dTdRhoi[NS];
  {
    double rho = 0.0, Rmix = 0.0;
    const double Pdim = Ps * p.ref_pressure;
#pragma unroll
    for (int i = 0; i < NS; i++) rhoidim[i] = Qs[i] * p.ref_density;
#pragma unroll
    for (int i = 0; i < NS; i++) { rho += rhoidim[i]; Rmix += rhoidim[i] * p.Rs[i]; }
    Rmix /= rho;
    const double dTdRho = (-Pdim / (Rmix * rho * rho));
    const double dTdR = (-Pdim / (rho * Rmix * Rmix));
#pragma unroll
    for (int i = 0; i < NS; i++) dTdRhoi[i] = dTdR * dRmixdRhoi(p, rhoidim, rho, i) + dTdRho;
  }
  const double dP_dru = Rs / cvs * us;
  const double dP_drv = Rs / cvs * vs;
  const double dP_drw = Rs / cvs * ws;
  const double dP_dret = Rs / cvs;
  const double Tn = (D[0] * av[0] + D[1] * av[1] + D[2] * av[2]) * C.c1;
  row = a + 3 * NEQ;
#pragma unroll
  for (int i = 0; i < NS; i++) {
    const double dT_drho = dTdRhoi[i] / (p.ref_temperature / p.ref_density);
    row[i] = -RKT * (dtauxn * u + dtauyn * v + dtauzn * w) + RK * Tn * dT_drho;
  }
  row[NS + 0] = -RKT * (dR2_drhou * u + dR3_drhou * v + dR4_drhou * w) + RK * Tn * dT_dP * dP_dru;
  row[NS + 1] = -RKT * (dR2_drhov * u + dR3_drhov * v + dR4_drhov * w) + RK * Tn * dT_dP * dP_drv;
  row[NS + 2] = -RKT * (dR2_drhow * u + dR3_drhow * v + dR4_drhow * w) + RK * Tn * dT_dP * dP_drw;
  row[NS + 3] = RK * Tn * dT_dP * dP_dret;
}

// ContributeTemporalTerms (compressibleFR.tcc:1319-1463) with ChemModel::dEtdP_dEtdRhoi (chem.tcc:829-858) and the
// IdealGasEOS derivatives (EOS.tcc:42-76); A = the node's diagonal block, row-major N x N, accumulated in place
template <int NS>
__device__ void temporal_terms(const Params<NS>& p, const double* Q, double vol, double cnp1, double dtau, double* A,
                               double beta) {
  constexpr int N = NS + 4, uloc = NS, vloc = NS + 1, wloc = NS + 2, tloc = NS + 3;
  double dEtdRhoi[NS], rhoiDim[NS], Yi[NS], thetaOBetai[NS];
  double vOverDt;
  if (p.use_local_dt) vOverDt = cnp1 * vol / p.dt + vol / dtau;
  else vOverDt = cnp1 * vol / dtau;
  const double rho = Q[NS + 5], T = Q[tloc], P = Q[NS + 4], u = Q[uloc], v = Q[vloc], w = Q[wloc];
  const double rvOverDt = rho * vOverDt;
  double v2 = u * u + v * v + w * w;
  double R, c2;
  fluid_props(p, Q, T, R, c2);
  const double s_ref = (p.ref_velocity * p.ref_velocity / p.ref_temperature);
  const double rhoDim = rho * p.ref_density;
  const double TDim = T * p.ref_temperature;
  const double PDim = P * p.ref_pressure;
  (void)PDim;
  v2 *= p.ref_velocity * p.ref_velocity;
  for (int i = 0; i < NS; i++) {
    rhoiDim[i] = Q[i] * p.ref_density;
    Yi[i] = Q[i] / rho;
  }
  // dEtdP_dEtdRhoi
  double dEtdP = 0.0;
  {
    const double hv2 = 0.5 * v2;
    double rhomix = 0.0, Rmix = 0.0;
    for (int i = 0; i < NS; i++) {
      rhomix += rhoiDim[i];
      Rmix += rhoiDim[i] * p.Rs[i];
    }
    Rmix /= rhomix;
    const double dPdrhomix = (Rmix * TDim);
    const double dPdRmix = (rhomix * TDim);
    const int rng = thermo_range(TDim);
    for (int i = 0; i < NS; i++)
      dEtdRhoi[i] = hv2 + sp_h(p, i, rng, TDim) - (dPdrhomix + dPdRmix * dRmixdRhoi(p, rhoiDim, rhomix, i));
    const double dTdP = (1.0 / (rhomix * Rmix));
    for (int i = 0; i < NS; i++) dEtdP += rhoiDim[i] * (sp_cp(p, i, rng, TDim)) * dTdP;
    dEtdP -= 1.0;
  }
  const double ref_detdrho = p.ref_specific_enthalpy * p.ref_density / p.ref_density;
  for (int i = 0; i < NS; i++) dEtdRhoi[i] /= ref_detdrho;
  const double ref_detdP = p.ref_specific_enthalpy * p.ref_density / p.ref_pressure;
  dEtdP /= ref_detdP;
  double dPdT = (rhoDim * (R * s_ref));
  dPdT /= (p.ref_pressure / p.ref_temperature);
  const double oneOBeta = 1.0 / beta;
  const double oneMbeta = 1.0 - beta;
  for (int i = 0; i < NS; i++) thetaOBetai[i] = (Yi[i] * oneMbeta / c2) * oneOBeta;
  for (int i = 0; i < NS; i++) {
    A[N * i + i] += vOverDt;
    A[N * i + tloc] += thetaOBetai[i] * vOverDt * dPdT;
  }
  double precond = 0.0;
  for (int j = 0; j < NS; j++) { A[N * uloc + j] += u * vOverDt; precond += thetaOBetai[j] * u; }
  A[N * uloc + uloc] += rvOverDt;
  A[N * uloc + tloc] += precond * vOverDt * dPdT;
  precond = 0.0;
  for (int j = 0; j < NS; j++) { A[N * vloc + j] += v * vOverDt; precond += thetaOBetai[j] * v; }
  A[N * vloc + vloc] += rvOverDt;
  A[N * vloc + tloc] += precond * vOverDt * dPdT;
  precond = 0.0;
  for (int j = 0; j < NS; j++) { A[N * wloc + j] += w * vOverDt; precond += thetaOBetai[j] * w; }
  A[N * wloc + wloc] += rvOverDt;
  A[N * wloc + tloc] += precond * vOverDt * dPdT;
  precond = 0.0;
  for (int j = 0; j < NS; j++) { A[N * tloc + j] += dEtdRhoi[j] * vOverDt; precond += dEtdRhoi[j] * thetaOBetai[j]; }
  A[N * tloc + uloc] += u * rvOverDt;
  A[N * tloc + vloc] += v * rvOverDt;
  A[N * tloc + wloc] += w * rvOverDt;
  precond += dEtdP * (oneOBeta);
  A[N * tloc + tloc] += precond * vOverDt * dPdT;
}

}  // namespace fr
