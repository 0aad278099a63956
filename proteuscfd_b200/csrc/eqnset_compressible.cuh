// eqnset_compressible.cuh -- device-side perfect-gas eqnset (EqnSet plugin surface,
// reference ucs/eqnset.h:44-299 as implemented by ucs/compressible.tcc).
//
// Conservative variables Q[0..4] = [rho, rho u, rho v, rho w, rho E], auxiliary
// Q[5..9] = [T, P, u, v, w] (compressible.tcc:27-28).
//
// PARITY CONTRACT: this translation unit is compiled with --fmad=false and every
// expression keeps the reference's operand order, so each double produced here
// is bit-identical to the reference built with plain x86-64 SSE2 arithmetic.
// (The FD Jacobian divides flux differences by h = 1e-8, so anything weaker than
// bit-identical fluxes would show up at 1e-8 relative in the matrix.)
#pragma once

#define PCFD_NEQN 5
#define PCFD_NVARS 10
#define PCFD_NTERMS 9

namespace eq {

// macros.h:32-42 (comparison form, not fmax/fmin: NaN and signed-zero behaviour differ)
__device__ __forceinline__ double maxd(double x, double y) { return (x > y) ? x : y; }
__device__ __forceinline__ double mind(double x, double y) { return (x < y) ? x : y; }

// compressible.tcc:1088-1101 ComputePressure
__device__ __forceinline__ double pressure(const double* Q, double gamma) {
  const double r = Q[0];
  const double u = Q[1] / r, v = Q[2] / r, w = Q[3] / r;
  const double v2h = 0.5 * (u * u + v * v + w * w);
  return (gamma - 1.0) * (Q[4] - r * v2h);
}

// compressible.tcc:1230-1243 ComputeAuxiliaryVariables (+ ComputeTemperature :1104-1110)
__device__ __forceinline__ void aux(double* Q, double gamma) {
  const double gm1 = gamma - 1.0;
  const double u = Q[1] / Q[0], v = Q[2] / Q[0], w = Q[3] / Q[0];
  const double V2 = u * u + v * v + w * w;
  Q[6] = gm1 * (Q[4] - 0.5 * Q[0] * V2);
  Q[5] = gamma * pressure(Q, gamma) / Q[0];
  Q[7] = u;
  Q[8] = v;
  Q[9] = w;
}

// compressible.tcc:996-1007 GetTheta
__device__ __forceinline__ double theta(const double* Q, const double* n, double vdotn) {
  const double u = Q[1] / Q[0], v = Q[2] / Q[0], w = Q[3] / Q[0];
  return u * n[0] + v * n[1] + w * n[2] + vdotn;
}

// eqnset.h:231-238 ExtrapolateCorrection, compressible.tcc:1039-1052 ExtrapolateVariables
__device__ __forceinline__ void extrapolate(double chi, double* Qho, const double* q, const double* dQedge,
                                            const double* gradQ, const double* dx, const double* lim) {
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const double corr =
        0.5 * chi * dQedge[i] + (1.0 - chi) * (gradQ[i * 3] * dx[0] + gradQ[i * 3 + 1] * dx[1] + gradQ[i * 3 + 2] * dx[2]);
    Qho[i] = q[i] + corr * lim[i];
  }
}

// compressible.tcc:1056-1079 BadExtrapolation
__device__ __forceinline__ bool bad_extrapolation(const double* Q, double gamma) {
  const double r = Q[0];
  const double u = Q[1] / r, v = Q[2] / r, w = Q[3] / r;
  const double E = Q[4];
  const double v2h = 0.5 * (u * u + v * v + w * w);
  const double p = (gamma - 1.0) * (E - r * v2h);
  return (p < 1.0e-10) || (r < 0.0) || (E < 1.0e-10);
}

// compressible.tcc:534-578 RoeVariables
__device__ __forceinline__ void roe_variables(const double* QL, const double* QR, double gamma, double* Qroe) {
  const double gm1 = gamma - 1.0;
  const double rhoL = QL[0], rhoR = QR[0];
  const double uL = QL[1] / QL[0], uR = QR[1] / QR[0];
  const double vL = QL[2] / QL[0], vR = QR[2] / QR[0];
  const double wL = QL[3] / QL[0], wR = QR[3] / QR[0];
  const double EL = QL[4], ER = QR[4];
  const double v2L = uL * uL + vL * vL + wL * wL;
  const double v2R = uR * uR + vR * vR + wR * wR;
  const double PL = gm1 * (EL - 0.5 * rhoL * v2L);
  const double PR = gm1 * (ER - 0.5 * rhoR * v2R);
  const double hL = (EL + PL) / rhoL;
  const double hR = (ER + PR) / rhoR;
  const double rho = sqrt(rhoL * rhoR);
  const double sigma = rho / (rhoL + rho);
  const double u = uL + sigma * (uR - uL);
  const double v = vL + sigma * (vR - vL);
  const double w = wL + sigma * (wR - wL);
  const double h = hL + sigma * (hR - hL);
  const double v2h = 0.5 * (u * u + v * v + w * w);
  Qroe[0] = rho;
  Qroe[1] = rho * u;
  Qroe[2] = rho * v;
  Qroe[3] = rho * w;
  Qroe[4] = rho / gamma * (h + gm1 * v2h);
}

// compressible.tcc:687-710 Flux
__device__ __forceinline__ void phys_flux(const double* Q, const double* n, double vdotn, double gamma, double* f,
                                          double* Pout = nullptr) {
  const double rho = Q[0];
  const double u = Q[1] / rho, v = Q[2] / rho, w = Q[3] / rho;
  const double rEt = Q[4];
  const double v2h = 0.5 * (u * u + v * v + w * w);
  const double P = (gamma - 1.0) * (rEt - rho * v2h);
  const double ht = (rEt + P) / rho;
  const double rhotheta = rho * (n[0] * u + n[1] * v + n[2] * w + vdotn);
  f[0] = rhotheta;
  f[1] = (u * rhotheta + P * n[0]);
  f[2] = (v * rhotheta + P * n[1]);
  f[3] = (w * rhotheta + P * n[2]);
  f[4] = (ht * rhotheta - vdotn * P);
  if (Pout) *Pout = P;   // == ComputePressure(Q): the quantity BadExtrapolation tests
}

// Harten-Hyman entropy fix #2 on one wave (compressible.tcc:150-196)
__device__ __forceinline__ double entropy_fix(double eig, double eigL, double eigR) {
  double eps = maxd((eig - eigL), (eigR - eig));
  eps = maxd(0.0, eps);
  if (fabs(eig) < eps) return 0.5 * (eig * eig / eps + eps);
  return fabs(eig);
}

// matrix.h:63-74 MatVecMult row: a0*v0 first, then += in column order
__device__ __forceinline__ double row5(double a0, double a1, double a2, double a3, double a4, const double* v) {
  double s = a0 * v[0];
  s += a1 * v[1];
  s += a2 * v[2];
  s += a3 * v[3];
  s += a4 * v[4];
  return s;
}

// compressible.tcc:93-230 RoeFlux, with Eigensystem (:581-684) evaluated row by row
// instead of through stored 5x5 T / Tinv arrays; n[0..2] unit normal, n[3] area.
// `bad` (optional) receives BadExtrapolation(QL) || BadExtrapolation(QR) || BadExtrapolation(Qroe)
// (compressible.tcc:1056-1079, the test of Kernel_PressureClip, limiters.tcc:777-813) from the pressures this
// routine forms anyway: the same expressions, so the same bits, at no extra divisions.
__device__ __forceinline__ void roe_flux(const double* QL, const double* QR, const double* n, double vdotn,
                                         double gamma, double* flux, bool* bad = nullptr) {
  double Qroe[5];
  roe_variables(QL, QR, gamma, Qroe);
  const double gm1 = gamma - 1.0;
  const double area = n[3];
  const double nx = n[0], ny = n[1], nz = n[2];

  // --- Roe-state eigensystem quantities
  const double rho = Qroe[0];
  const double u = Qroe[1] / rho, v = Qroe[2] / rho, w = Qroe[3] / rho;
  const double thetaf = u * nx + v * ny + w * nz;
  const double th = thetaf + vdotn;
  const double v2h = 0.5 * (u * u + v * v + w * w);
  const double P = gm1 * (Qroe[4] - rho * v2h);
  const double c2 = gamma * P / rho;
  const double c = sqrt(c2);

  // --- left/right wave speeds for the entropy fix
  double thetaL, thetaR, cL, cR;
  {
    const double rhoL = QL[0];
    const double uL = QL[1] / rhoL, vL = QL[2] / rhoL, wL = QL[3] / rhoL;
    const double PL = gm1 * (QL[4] - 0.5 * rhoL * (uL * uL + vL * vL + wL * wL));
    const double rhoR = QR[0];
    const double uR = QR[1] / rhoR, vR = QR[2] / rhoR, wR = QR[3] / rhoR;
    const double PR = gm1 * (QR[4] - 0.5 * rhoR * (uR * uR + vR * vR + wR * wR));
    thetaL = uL * nx + vL * ny + wL * nz + vdotn;
    thetaR = uR * nx + vR * ny + wR * nz + vdotn;
    cR = sqrt(gamma * PR / rhoR);
    cL = sqrt(gamma * PL / rhoL);
  }
  double lam[5];
  lam[0] = lam[1] = lam[2] = entropy_fix(th, thetaL, thetaR);
  lam[3] = entropy_fix(th + c, thetaL + cL, thetaR + cR);
  lam[4] = entropy_fix(th - c, thetaL - cL, thetaR - cR);

  double dQ[5], dv[5];
#pragma unroll
  for (int i = 0; i < 5; i++) dQ[i] = QR[i] - QL[i];

  // dv = Tinv * dQ (rows of Tinv, compressible.tcc:640-673)
  dv[0] = row5(nx - nz * v / rho + ny * w / rho - nx / c2 * v2h * gm1, nx / c2 * u * gm1, nz / rho + nx / c2 * v * gm1,
               -ny / rho + nx / c2 * w * gm1, -nx / c2 * gm1, dQ);
  dv[1] = row5(ny + nz * u / rho - nx * w / rho - ny / c2 * v2h * gm1, -nz / rho + ny / c2 * u * gm1, ny / c2 * v * gm1,
               nx / rho + ny / c2 * w * gm1, -ny / c2 * gm1, dQ);
  dv[2] = row5(nz - ny * u / rho + nx * v / rho - nz / c2 * v2h * gm1, ny / rho + nz / c2 * u * gm1,
               -nx / rho + nz / c2 * v * gm1, nz / c2 * w * gm1, -nz / c2 * gm1, dQ);
  dv[3] = row5(-0.5 / rho * (thetaf - gm1 * v2h / c), 0.5 / rho * (nx - gm1 * u / c), 0.5 / rho * (ny - gm1 * v / c),
               0.5 / rho * (nz - gm1 * w / c), 0.5 / rho * (gm1 / c), dQ);
  dv[4] = row5(0.5 / rho * (thetaf + gm1 * v2h / c), -0.5 / rho * (nx + gm1 * u / c), -0.5 / rho * (ny + gm1 * v / c),
               -0.5 / rho * (nz + gm1 * w / c), +0.5 / rho * (gm1 / c), dQ);
#pragma unroll
  for (int i = 0; i < 5; i++) dv[i] *= fabs(lam[i]);

  // dr = T * dv (rows of T, compressible.tcc:606-637)
  double dr[5];
  const double rc = rho / c;
  dr[0] = row5(nx, ny, nz, rc, rc, dv);
  dr[1] = row5(u * nx, u * ny - rho * nz, u * nz + rho * ny, rho * (u / c + nx), rho * (u / c - nx), dv);
  dr[2] = row5(v * nx + rho * nz, v * ny, v * nz - rho * nx, rho * (v / c + ny), rho * (v / c - ny), dv);
  dr[3] = row5(w * nx - rho * ny, w * ny + rho * nx, w * nz, rho * (w / c + nz), rho * (w / c - nz), dv);
  dr[4] = row5(v2h * nx + rho * (v * nz - w * ny), v2h * ny + rho * (w * nx - u * nz), v2h * nz + rho * (u * ny - v * nx),
               rho * (v2h / c + thetaf + c / gm1), rho * (v2h / c - thetaf + c / gm1), dv);

  double fL[5], fR[5], pL, pR;
  phys_flux(QL, n, vdotn, gamma, fL, &pL);
  phys_flux(QR, n, vdotn, gamma, fR, &pR);
#pragma unroll
  for (int i = 0; i < 5; i++) flux[i] = 0.5 * area * (fL[i] + fR[i] - dr[i]);
  if (bad)
    *bad = (pL < 1.0e-10) || (QL[0] < 0.0) || (QL[4] < 1.0e-10) || (pR < 1.0e-10) || (QR[0] < 0.0) || (QR[4] < 1.0e-10) ||
           (P < 1.0e-10) || (rho < 0.0) || (Qroe[4] < 1.0e-10);
}

// EqnSet::NumericalFlux (eqnset.tcc:55-90): Roe + the NaN kneecap
__device__ __forceinline__ void numerical_flux(const double* QL, const double* QR, const double* n, double vdotn,
                                               double gamma, double* flux, bool* bad = nullptr) {
  roe_flux(QL, QR, n, vdotn, gamma, flux, bad);
#pragma unroll
  for (int i = 0; i < 5; i++)
    if (isnan(flux[i])) flux[i] = 0.0;
}

// compressible.tcc:799-821 MaxEigenvalue
__device__ __forceinline__ double max_eigenvalue(const double* Q, const double* n, double vdotn, double gamma) {
  const double rho = Q[0];
  const double u = Q[1] / rho, v = Q[2] / rho, w = Q[3] / rho;
  const double v2h = 0.5 * (u * u + v * v + w * w);
  const double P = (gamma - 1.0) * (Q[4] - rho * v2h);
  const double c = sqrt(gamma * P / rho);
  const double th = theta(Q, n, vdotn);
  return maxd(fabs(th + c), fabs(th - c));
}

// compressible.tcc:929-993 ApplyDQ
__device__ __forceinline__ void apply_dq(const double* dQ, double* Q, double gamma) {
  const double gm1 = gamma - 1.0;
  const double minP = 1.0e-10, minRho = 1.0e-10, minE = 1.0e-10;
  if (Q[0] + dQ[0] < 0.0) Q[0] = minRho; else Q[0] += dQ[0];
  if (Q[4] + dQ[4] < 0.0) Q[4] = minE; else Q[4] += dQ[4];
  const double rho = Q[0];
  double u = Q[1] / rho, v = Q[2] / rho, w = Q[3] / rho;
  const double E = Q[4];
  const double v2 = u * u + v * v + w * w;
  if (E < 0.5 * rho * v2) {
    const double v2mod = 2.0 * (E - minP / gm1);
    double frac = 0.0;
    if (v2mod > 0.0) frac = sqrt(v2mod / v2);
    u *= frac; v *= frac; w *= frac;
    Q[1] = rho * u; Q[2] = rho * v; Q[3] = rho * w;
  } else {
    Q[1] += dQ[1]; Q[2] += dQ[2]; Q[3] += dQ[3];
  }
  aux(Q, gamma);
}

// ------------------------------------------------------------------ viscous terms

struct ViscParams {
  double gamma, Re, Pr, PrT, tref, mach;
};

// T^1.5 for Sutherland's law.  The reference calls libm pow(T, 1.5) (eqnset.h:264); glibc's pow is
// accurate to ~0.52 ulp, not reproducible bit for bit without its tables.  T*sqrt(T) is evaluated
// here in double-double and rounded once, i.e. to within 0.5 ulp (+2^-100) of the exact value: it
// equals glibc's result except where glibc itself misses the correctly rounded value (then 1 ulp).
__device__ __forceinline__ double pow15(double T) {
  const double s = sqrt(T);
  const double e = __fma_rn(-s, s, T);   // T - s*s, exact
  const double sl = e / (2.0 * s);       // sqrt(T) = s + sl
  const double p = T * s;
  const double pe = __fma_rn(T, s, -p);  // T*s - p, exact
  return p + (pe + T * sl);
}

// eqnset.h:251-268 ComputeViscosity (Sutherland, non-dimensional); T = Q[5]
__device__ __forceinline__ double viscosity(const ViscParams& vp, double T) {
  const double S = 110.4 / vp.tref;
  return (1.0 + S) * pow15(T) / (T + S);
}

// compressible.tcc:713-795 ViscousFlux.  Q[0..3] and T of the edge-averaged state; g = gradient rows
// [T, u, v, w] (terms 5..8 of qgrad, 12 doubles); flux[0] is identically 0 and not returned.
__device__ __forceinline__ void viscous_flux(const ViscParams& vp, const double* Q, double T, const double* g,
                                             const double* n, double mut, double* f14) {
  const double* gT = g;
  const double* gu = g + 3;
  const double* gv = g + 6;
  const double* gw = g + 9;
  const double rho = Q[0];
  const double u = Q[1] / rho, v = Q[2] / rho, w = Q[3] / rho;
  const double mu = viscosity(vp, T);
  const double tmut = (mu + mut);
  const double fact = -2.0 / 3.0 * (gu[0] + gv[1] + gw[2]);
  const double tauxx = 2.0 * gu[0] + fact, tauyy = 2.0 * gv[1] + fact, tauzz = 2.0 * gw[2] + fact;
  const double tauxy = gu[1] + gv[0], tauxz = gu[2] + gw[0], tauyz = gv[2] + gw[1];
  const double ReTilde = vp.Re / vp.mach;
  const double RK = n[3] / ReTilde;
  const double RKT = RK * tmut;
  const double cp = 1.0 / (vp.gamma - 1.0);
  const double k = mu / vp.Pr * cp;
  const double kT = mut / vp.PrT * cp;
  const double c1 = -(k + kT);
  const double Tn = gT[0] * n[0] + gT[1] * n[1] + gT[2] * n[2];
  const double tauxn = tauxx * n[0] + tauxy * n[1] + tauxz * n[2];
  const double tauyn = tauxy * n[0] + tauyy * n[1] + tauyz * n[2];
  const double tauzn = tauxz * n[0] + tauyz * n[1] + tauzz * n[2];
  f14[0] = -RKT * (tauxn);
  f14[1] = -RKT * (tauyn);
  f14[2] = -RKT * (tauzn);
  f14[3] = -RKT * (tauxn * u + tauyn * v + tauzn * w) + RK * c1 * Tn;
}

// one side of ViscousJacobian (compressible.tcc:1633-1893): D = -/+ dx/(rho_side*s2); (rs,us,vs,ws,Ps)
// the side's own state, (u,v,w) the edge-averaged velocity.  Adds sign*a to the 5x5 block at dst.
__device__ __forceinline__ void viscous_jac_side(const double* D, double rs, double us, double vs, double ws, double Ps,
                                                 double u, double v, double w, const double* n, double RK, double RKT,
                                                 double c1, double gamma, bool negate, double* dst) {
  const double c43 = 4.0 / 3.0, mc23 = -2.0 / 3.0;
  const double gm1 = (gamma - 1.0);
  const double dux = -u * D[0], duy = -u * D[1], duz = -u * D[2];
  const double dvx = -v * D[0], dvy = -v * D[1], dvz = -v * D[2];
  const double dwx = -w * D[0], dwy = -w * D[1], dwz = -w * D[2];
  const double dfact = -2.0 / 3.0 * (dux + dvy + dwz);
  const double dtauxx = (2.0 * dux + dfact), dtauyy = (2.0 * dvy + dfact), dtauzz = (2.0 * dwz + dfact);
  const double dtauxy = duy + dvx, dtauxz = duz + dwx, dtauyz = dvz + dwy;
  const double dtauxn = dtauxx * n[0] + dtauxy * n[1] + dtauxz * n[2];
  const double dtauyn = dtauxy * n[0] + dtauyy * n[1] + dtauyz * n[2];
  const double dtauzn = dtauxz * n[0] + dtauyz * n[1] + dtauzz * n[2];
  const double dR2u = (c43 * D[0] * n[0] + D[1] * n[1] + D[2] * n[2]);
  const double dR2v = (mc23 * D[1] * n[0] + D[0] * n[1]);
  const double dR2w = (mc23 * D[2] * n[0] + D[0] * n[2]);
  const double dR3u = (mc23 * D[0] * n[1] + D[1] * n[0]);
  const double dR3v = (D[0] * n[0] + c43 * D[1] * n[1] + D[2] * n[2]);
  const double dR3w = (mc23 * D[2] * n[1] + D[1] * n[2]);
  const double dR4u = (mc23 * D[0] * n[2] + D[2] * n[0]);
  const double dR4v = (mc23 * D[1] * n[2] + D[2] * n[1]);
  const double dR4w = (D[0] * n[0] + D[1] * n[1] + c43 * D[2] * n[2]);
  const double v2 = (us * us + vs * vs + ws * ws);
  const double dT_dP = gamma / rs;
  const double dP_dr = +gm1 * 0.5 * v2, dP_dru = -gm1 * us, dP_drv = -gm1 * vs, dP_drw = -gm1 * ws, dP_dret = +gm1;
  const double Tn = (D[0] * n[0] + D[1] * n[1] + D[2] * n[2]) * dT_dP * c1;
  double a[20];   // rows 1..4 (row 0 is zero)
  a[0] = -RKT * (dtauxn); a[1] = -RKT * dR2u; a[2] = -RKT * dR2v; a[3] = -RKT * dR2w; a[4] = 0.0;
  a[5] = -RKT * (dtauyn); a[6] = -RKT * dR3u; a[7] = -RKT * dR3v; a[8] = -RKT * dR3w; a[9] = 0.0;
  a[10] = -RKT * (dtauzn); a[11] = -RKT * dR4u; a[12] = -RKT * dR4v; a[13] = -RKT * dR4w; a[14] = 0.0;
  a[15] = -RKT * (dtauxn * u + dtauyn * v + dtauzn * w) + RK * Tn * (dP_dr - Ps / rs);
  a[16] = -RKT * (dR2u * u + dR3u * v + dR4u * w) + RK * Tn * dP_dru;
  a[17] = -RKT * (dR2v * u + dR3v * v + dR4v * w) + RK * Tn * dP_drv;
  a[18] = -RKT * (dR2w * u + dR3w * v + dR4w * w) + RK * Tn * dP_drw;
  a[19] = RK * Tn * dP_dret;
  // row 0 receives +/-0.0, which leaves every value (up to the sign of a zero) unchanged
#pragma unroll
  for (int k = 0; k < 20; k++) dst[5 + k] += negate ? -a[k] : a[k];
}

// Kernel_Viscous_Jac (jacobian.tcc:768-800) for one edge: A(l,r) += aR, A(r,l) += aL (aL sign-flipped)
__device__ __forceinline__ void viscous_jacobian(const ViscParams& vp, const double* QL, const double* QR, const double* dx,
                                                 double s2, const double* n, double mut, double* pRL, double* pLR) {
  double Qavg[5];
#pragma unroll
  for (int i = 0; i < 5; i++) Qavg[i] = 0.5 * (QL[i] + QR[i]);
  const double Tavg = vp.gamma * pressure(Qavg, vp.gamma) / Qavg[0];   // aux(): Q[5]
  const double mu = viscosity(vp, Tavg);
  const double tmut = (mu + mut);
  const double ReTilde = vp.Re / vp.mach;
  const double RK = n[3] / ReTilde;
  const double RKT = RK * tmut;
  const double rhoL = QL[0], uL = QL[1] / rhoL, vL = QL[2] / rhoL, wL = QL[3] / rhoL, PL = QL[6];
  const double rhoR = QR[0], uR = QR[1] / rhoR, vR = QR[2] / rhoR, wR = QR[3] / rhoR, PR = QR[6];
  const double u = 0.5 * (uL + uR), v = 0.5 * (vL + vR), w = 0.5 * (wL + wR);
  double DxL[3], DxR[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    DxL[i] = -dx[i] / (rhoL * s2);
    DxR[i] = dx[i] / (rhoR * s2);
  }
  const double cp = 1.0 / (vp.gamma - 1.0);
  const double k = mu / vp.Pr * cp;
  const double kT = mut / vp.PrT * cp;
  const double c1 = -(k + kT);
  viscous_jac_side(DxR, rhoR, uR, vR, wR, PR, u, v, w, n, RK, RKT, c1, vp.gamma, false, pLR);
  viscous_jac_side(DxL, rhoL, uL, vL, wL, PL, u, v, w, n, RK, RKT, c1, vp.gamma, true, pRL);
}

// ---------------------------------------------------------------- boundary states

struct BcParams {
  double gamma;
  int no_cvbc;
  double qinf[PCFD_NVARS];
};

// compressible.tcc:1544-1574 GetViscousWallBoundaryVariables on a static wall (vel = 0: bc.tcc:1207-1254
// with movement == 0, bleedSteps == 0, velw == 0); normalQ = state of the most-normal neighbour
__device__ __forceinline__ void viscous_wall_bc(double gamma, double* QL, double* QR, const double* normalQ, double Twall) {
  const double vel[3] = {0.0, 0.0, 0.0};
  if (Twall < 0.0) {
    QR[0] = QL[0] = normalQ[0];
    QR[4] = QL[4] = normalQ[4];
  } else {
    const double v2 = vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2];
    const double gm1 = gamma - 1.0;
    const double rhoEt = (Twall * QL[0] / (gamma * gm1) + 0.5 * QR[0] * v2);
    QR[0] = QL[0];
    QR[4] = QL[4] = rhoEt;
  }
  QR[1] = QL[1] = QL[0] * vel[0];
  QR[2] = QL[2] = QL[0] * vel[1];
  QR[3] = QL[3] = QL[0] * vel[2];
}

// compressible.tcc:1246-1372 characteristic far field (static mesh, 10 sub-iterations)
// Qinf: the free stream the caller hands over (p.qinf, or the power-law scaled copy of the viscous far field)
__device__ __forceinline__ void farfield_bc(const BcParams& p, const double* QL, double* QR, const double* n, double vdotn,
                                            const double* Qinf) {
  const double gamma = p.gamma;
  for (int subit = 0; subit < 10; subit++) {
    const double u = vdotn * n[0], v = vdotn * n[1], w = vdotn * n[2];
    double avg[5];
#pragma unroll
    for (int i = 0; i < 5; i++) avg[i] = (QL[i] + QR[i]) / 2.0;
    const double th = theta(avg, n, vdotn);
    const double rhoi = QL[0];
    const double ui = QL[1] / QL[0] + u, vi = QL[2] / QL[0] + v, wi = QL[3] / QL[0] + w;
    const double pi = QL[6];
    const double rhoinf = Qinf[0];
    const double uinf = Qinf[1] / Qinf[0] + u, vinf = Qinf[2] / Qinf[0] + v, winf = Qinf[3] / Qinf[0] + w;
    const double pinf = Qinf[6];
    const double pavg = pressure(avg, gamma);
    const double rhoavg = avg[0];
    const double c2avg = gamma * (pavg / rhoavg);
    const double cavg = sqrt(c2avg);
    const double nx = n[0], ny = n[1], nz = n[2];
    if (th == 0.0) {
      return;
    } else if (th > 0.0 && fabs(th / cavg) >= 1.0) {
#pragma unroll
      for (int i = 0; i < 5; i++) QR[i] = QL[i];
    } else if (th < 0.0 && fabs(th / cavg) >= 1.0) {
#pragma unroll
      for (int i = 0; i < 5; i++) QR[i] = Qinf[i];
    } else if (th > 0.0 && fabs(th / cavg) < 1.0) {
      const double pb = pinf;
      const double rhob = rhoi + (pb - pi) / c2avg;
      const double t = (pb - pi) / (rhoavg * cavg);
      const double ub = ui - nx * t, vb = vi - ny * t, wb = wi - nz * t;
      QR[0] = rhob; QR[1] = rhob * ub; QR[2] = rhob * vb; QR[3] = rhob * wb;
      QR[4] = pb / (gamma - 1.0) + 0.5 * rhob * (ub * ub + vb * vb + wb * wb);
    } else if (th < 0.0 && fabs(th / cavg) < 1.0) {
      const double pb = 0.5 * (pinf + pi + rhoavg * cavg * (nx * (uinf - ui) + ny * (vinf - vi) + nz * (winf - wi)));
      const double rhob = rhoinf + (pb - pinf) / c2avg;
      const double t = (pb - pinf) / (rhoavg * cavg);
      const double ub = uinf + nx * t, vb = vinf + ny * t, wb = winf + nz * t;
      QR[0] = rhob; QR[1] = rhob * ub; QR[2] = rhob * vb; QR[3] = rhob * wb;
      QR[4] = pb / (gamma - 1.0) + 0.5 * rhob * (ub * ub + vb * vb + wb * wb);
    } else {
      return;   // NaN theta
    }
  }
}

// compressible.tcc:1374-1472 slip wall / symmetry
__device__ __forceinline__ void inviscid_wall_bc(const BcParams& p, const double* QL, double* QR, const double* n,
                                                 double vdotn) {
  const double gamma = p.gamma;
  for (int subit = 0; subit < 10; subit++) {
    double QLmod[5], avg[5];
#pragma unroll
    for (int i = 0; i < 5; i++) QLmod[i] = QL[i];
    double rhoi = QLmod[0];
    const double u = vdotn * n[0], v = vdotn * n[1], w = vdotn * n[2];
    const double ru = u * rhoi, rv = v * rhoi, rw = w * rhoi;
    QLmod[1] += ru; QLmod[2] += rv; QLmod[3] += rw;
#pragma unroll
    for (int i = 0; i < 5; i++) avg[i] = (QL[i] + QR[i]) / 2.0;
    rhoi = QL[0];
    const double ui = QL[1] / QL[0] + ru, vi = QL[2] / QL[0] + rv, wi = QL[3] / QL[0] + rw;
    const double pi = QL[6];
    const double rhoavg = avg[0];
    avg[1] += rhoavg * u; avg[2] += rhoavg * v; avg[3] += rhoavg * w;
    const double pavg = pressure(avg, gamma);
    const double c2avg = gamma * (pavg / rhoavg);
    const double cavg = sqrt(c2avg);
    if (!p.no_cvbc) {
      const double nx = n[0], ny = n[1], nz = n[2];
      const double pb = pi + rhoavg * cavg * (theta(QL, n, vdotn));
      const double rhob = rhoi + (pb - pi) / c2avg;
      const double t = (pb - pi) / (rhoavg * cavg);
      const double ub = ui - nx * t, vb = vi - ny * t, wb = wi - nz * t;
      QR[0] = rhob; QR[1] = rhob * ub; QR[2] = rhob * vb; QR[3] = rhob * wb;
      QR[4] = pb / (gamma - 1.0) + 0.5 * rhob * (ub * ub + vb * vb + wb * wb);
    } else {
      // MirrorVector (geometry.h): v - 2 (v.n) n
#pragma unroll
      for (int i = 0; i < 5; i++) QR[i] = QLmod[i];
      const double dot = QLmod[1] * n[0] + QLmod[2] * n[1] + QLmod[3] * n[2];
      QR[1] = QLmod[1] - 2.0 * dot * n[0];
      QR[2] = QLmod[2] - 2.0 * dot * n[1];
      QR[3] = QLmod[3] - 2.0 * dot * n[2];
    }
  }
}

// bc.tcc:1058-1120 + :1392-1396 CalculateBoundaryVariables for the BC types of the
// hot-path configs; QL and QR are full nvars states.
// Proteus_FarFieldViscous (bc.tcc:1092-1108): the momentum of the free stream is scaled by ubar = PowerLawU(1, wall
// distance of the left node, Re) (powerLaw.h:11-29; formed on the host with the C library's pow, pcfd_set_field of
// PCFD_F_WALLDIST) when ubar < 1.  qref: the caller's copy of Qinf, scaled IN PLACE as the reference does -- BC_Kernel
// takes a fresh copy per call (bc.tcc:741-744: pass NULL), Bkernel_NumJac hands ONE copy to all its re-evaluations
// (jacobian.tcc:485-506), so the scaling compounds from perturbation to perturbation there.
__device__ __forceinline__ void boundary_variables(const BcParams& p, double* QL, double* QR, const double* n, int bctype,
                                                   const double* normalQ = nullptr, double twall = 0.0, double ubar = 1.0,
                                                   double* qref = nullptr) {
  const double vdotn = 0.0;   // static mesh: driver.tcc:97-113 with Mesh::nv == 0
  switch (bctype) {
    case PCFD_BC_FARFIELD_VISCOUS: {
      double fresh[PCFD_NVARS];
      double* Qinf = qref ? qref : fresh;
      if (!qref) {
#pragma unroll
        for (int i = 0; i < PCFD_NVARS; i++) fresh[i] = p.qinf[i];
      }
      if (ubar < 1.0) {
#pragma unroll
        for (int i = 0; i < 3; i++) Qinf[1 + i] = ubar * Qinf[1 + i];   // GetMomentumLocation() == 1
      }
      farfield_bc(p, QL, QR, n, vdotn, Qinf);
      break;
    }
    case PCFD_BC_PARALLEL:
      return;
    case PCFD_BC_SONIC_INFLOW:
    case PCFD_BC_DIRICHLET:
#pragma unroll
      for (int i = 0; i < PCFD_NVARS; i++) QR[i] = QL[i] = p.qinf[i];
      break;
    case PCFD_BC_SONIC_OUTFLOW:
    case PCFD_BC_NEUMANN:
#pragma unroll
      for (int i = 0; i < PCFD_NEQN; i++) QR[i] = QL[i];
      break;
    case PCFD_BC_FARFIELD:
      farfield_bc(p, QL, QR, n, vdotn, p.qinf);
      break;
    case PCFD_BC_IMPERMEABLE_WALL:
    case PCFD_BC_SYMMETRY:
      inviscid_wall_bc(p, QL, QR, n, vdotn);
      break;
    case PCFD_BC_NOSLIP:   // bc.tcc:1182-1291
      viscous_wall_bc(p.gamma, QL, QR, normalQ, twall);
      break;
    default:
      break;
  }
  aux(QR, p.gamma);
  aux(QL, p.gamma);
}

}  // namespace eq
