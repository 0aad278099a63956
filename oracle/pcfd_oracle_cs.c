/* pcfd_oracle_cs.c -- TEST INFRASTRUCTURE (see pcfd_oracle.c): the complex-step field Jacobian of the reference,
   Kernel_NumJac_Complex (ucs/jacobian.tcc:370-433), selected by jacobianFieldType = 2 (:150-176).

   The reference instantiates its eqnset on std::complex<double> (RCmplx) and evaluates the Roe flux with a state
   perturbed by i*1e-11; the Jacobian column is imag(flux)/1e-11.  This file is the Roe flux of pcfd_oracle.c
   (compressible.tcc:93-230, 534-710) on C99 `double complex`, whose product, quotient and square root are the same
   libgcc / glibc routines libstdc++'s std::complex uses (__muldc3, __divdc3, csqrt), with the reference's complex
   overloads: CAbs flips the sign by the real part, MAX / comparisons look at the real part (macros.h:32-68).  Every
   quantity the reference holds as `Type` (gamma, the area vector, vdotn) is complex here too, literals stay real.
   The auxiliary variables the reference recomputes for the perturbed state (ComputeAuxiliaryVariables) are not read by
   the Roe flux and are left out.  Pinned bit for bit by tests/golden/box6_implicit_complex.npz. */
#include "pcfd_oracle.h"

#include <complex.h>
#include <math.h>
#include <string.h>

typedef double complex cplx;

static cplx cs_abs(cplx x){ cplx b = x; if(creal(x) < 0.0) b = -x; return b; }
static cplx cs_max(cplx x, cplx y){ return (creal(x) > creal(y)) ? x : y; }

static void cs_roe_variables(const cplx* QL, const cplx* QR, cplx gamma, cplx* Qroe)
{
  cplx gm1 = gamma - 1.0;
  cplx rhoL = QL[0], rhoR = QR[0];
  cplx uL = QL[1]/QL[0], uR = QR[1]/QR[0];
  cplx vL = QL[2]/QL[0], vR = QR[2]/QR[0];
  cplx wL = QL[3]/QL[0], wR = QR[3]/QR[0];
  cplx EL = QL[4], ER = QR[4];
  cplx v2L = uL*uL + vL*vL + wL*wL;
  cplx v2R = uR*uR + vR*vR + wR*wR;
  cplx PL = gm1*(EL - 0.5*rhoL*v2L);
  cplx PR = gm1*(ER - 0.5*rhoR*v2R);
  cplx hL = (EL + PL)/rhoL;
  cplx hR = (ER + PR)/rhoR;
  cplx rho = csqrt(rhoL*rhoR);
  cplx sigma = rho/(rhoL + rho);
  cplx u = uL + sigma*(uR - uL);
  cplx v = vL + sigma*(vR - vL);
  cplx w = wL + sigma*(wR - wL);
  cplx h = hL + sigma*(hR - hL);
  cplx v2h = 0.5*(u*u + v*v + w*w);
  Qroe[0] = rho;
  Qroe[1] = rho*u;
  Qroe[2] = rho*v;
  Qroe[3] = rho*w;
  Qroe[4] = rho/gamma*(h + gm1*v2h);
}

static void cs_eigensystem(const cplx* Q, const cplx* avec, cplx vdotn, cplx gamma,
			cplx* eigenvalues, cplx* T, cplx* Tinv)
{
  cplx nx = avec[0], ny = avec[1], nz = avec[2];
  cplx rho = Q[0], ru = Q[1], rv = Q[2], rw = Q[3], rE = Q[4];
  cplx u = ru/rho, v = rv/rho, w = rw/rho;
  cplx gm1 = gamma - 1.0;
  cplx thetaf = u*nx + v*ny + w*nz;
  cplx theta = thetaf + vdotn;
  cplx v2h = 0.5*(u*u + v*v + w*w);
  cplx P = gm1*(rE - rho*v2h);
  cplx c2 = gamma*P/rho;
  cplx c = csqrt(c2);

  T[0] = nx;
  T[5] = u*nx;
  T[10] = v*nx + rho*nz;
  T[15] = w*nx - rho*ny;
  T[20] = v2h*nx + rho*(v*nz - w*ny);

  T[1] = ny;
  T[6] = u*ny - rho*nz;
  T[11] = v*ny;
  T[16] = w*ny + rho*nx;
  T[21] = v2h*ny + rho*(w*nx - u*nz);

  T[2] = nz;
  T[7] = u*nz + rho*ny;
  T[12] = v*nz - rho*nx;
  T[17] = w*nz;
  T[22] = v2h*nz + rho*(u*ny - v*nx);

  T[3] = rho/c;
  T[8] = rho*(u/c + nx);
  T[13] = rho*(v/c + ny);
  T[18] = rho*(w/c + nz);
  T[23] = rho*(v2h/c + thetaf + c/gm1);

  T[4] = rho/c;
  T[9] = rho*(u/c - nx);
  T[14] = rho*(v/c - ny);
  T[19] = rho*(w/c - nz);
  T[24] = rho*(v2h/c - thetaf + c/gm1);

  Tinv[0]  = nx - nz*v/rho + ny*w/rho - nx/c2*v2h*gm1;
  Tinv[1]  = nx/c2*u*gm1;
  Tinv[2]  = nz/rho + nx/c2*v*gm1;
  Tinv[3]  = -ny/rho + nx/c2*w*gm1;
  Tinv[4]  = -nx/c2*  gm1;

  Tinv[5]  = ny + nz*u/rho - nx*w/rho - ny/c2*v2h*gm1;
  Tinv[6]  = -nz/rho + ny/c2*u*gm1;
  Tinv[7]  = ny/c2*v*gm1;
  Tinv[8]  = nx/rho + ny/c2*w*gm1;
  Tinv[9]  = -ny/c2*gm1;

  Tinv[10] = nz - ny*u/rho + nx*v/rho - nz/c2*v2h*gm1;
  Tinv[11] = ny/rho + nz/c2*u*gm1;
  Tinv[12] = -nx/rho + nz/c2*v*gm1;
  Tinv[13] = nz/c2*w*gm1;
  Tinv[14] = -nz/c2*  gm1;

  Tinv[15] = -0.5/rho*(thetaf - gm1*v2h/c);
  Tinv[16] = 0.5/rho*(nx - gm1*u/c);
  Tinv[17] = 0.5/rho*(ny - gm1*v/c);
  Tinv[18] = 0.5/rho*(nz - gm1*w/c);
  Tinv[19] = 0.5/rho*(gm1 /c);

  Tinv[20] = 0.5/rho*(thetaf + gm1*v2h/c);
  Tinv[21] = -0.5/rho*(nx + gm1*u/c);
  Tinv[22] = -0.5/rho*(ny + gm1*v/c);
  Tinv[23] = -0.5/rho*(nz + gm1*w/c);
  Tinv[24] = +0.5/rho*(gm1 /c);

  eigenvalues[0] = theta;
  eigenvalues[1] = theta;
  eigenvalues[2] = theta;
  eigenvalues[3] = theta + c;
  eigenvalues[4] = theta - c;
}

static void cs_phys_flux(const cplx* Q, const cplx* avec, cplx vdotn, cplx gamma, cplx* flux)
{
  cplx rho = Q[0], ru = Q[1], rv = Q[2], rw = Q[3], rEt = Q[4];
  cplx u = ru/rho, v = rv/rho, w = rw/rho;
  cplx v2h = 0.5*(u*u + v*v + w*w);
  cplx gm1 = gamma - 1.0;
  cplx P = gm1*(rEt - rho*v2h);
  cplx ht = (rEt + P)/rho;
  cplx rhotheta = rho*(avec[0]*u + avec[1]*v + avec[2]*w + vdotn);
  flux[0] = rhotheta;
  flux[1] = (u*rhotheta + P*avec[0]);
  flux[2] = (v*rhotheta + P*avec[1]);
  flux[3] = (w*rhotheta + P*avec[2]);
  flux[4] = (ht*rhotheta - vdotn*P);
}

static void cs_matvec(const cplx* a, const cplx* v, cplx* vout, int n)
{
  int i, j;
  for(i = 0; i < n; i++){
    vout[i] = a[i*n + 0]*v[0];
    for(j = 1; j < n; j++) vout[i] += a[i*n + j]*v[j];
  }
}

static void cs_roe_flux(const cplx* QL, const cplx* QR, const cplx* avec, cplx vdotn, cplx gamma,
		  cplx* flux)
{
  int i;
  cplx Qroe[5], T[25], Tinv[25], eigenvalues[5], fluxL[5], fluxR[5], dQ[5], dv[5], dr[5];
  cplx area = avec[3];
  cplx gm1, thetaR, thetaL, eigL, eigR, eps, cR, cL, eig;

  cs_roe_variables(QL, QR, gamma, Qroe);
  cs_eigensystem(Qroe, avec, vdotn, gamma, eigenvalues, T, Tinv);

  gm1 = gamma - 1.0;
  {
    const cplx rhoL = QL[0];
    const cplx uL = QL[1]/rhoL, vL = QL[2]/rhoL, wL = QL[3]/rhoL;
    const cplx EL = QL[4];
    const cplx vmag2L = uL*uL + vL*vL + wL*wL;
    const cplx PL = gm1*(EL - 0.5*rhoL*vmag2L);
    const cplx rhoR = QR[0];
    const cplx uR = QR[1]/rhoR, vR = QR[2]/rhoR, wR = QR[3]/rhoR;
    const cplx ER = QR[4];
    const cplx vmag2R = uR*uR + vR*vR + wR*wR;
    const cplx PR = gm1*(ER - 0.5*rhoR*vmag2R);
    thetaL = uL*avec[0] + vL*avec[1] + wL*avec[2] + vdotn;
    thetaR = uR*avec[0] + vR*avec[1] + wR*avec[2] + vdotn;
    cR = csqrt(gamma*PR/rhoR);
    cL = csqrt(gamma*PL/rhoL);
  }

  eigL = thetaL; eigR = thetaR; eig = eigenvalues[0];
  eps = cs_max((eig - eigL), (eigR - eig));
  eps = cs_max(0.0, eps);
  if(creal(cs_abs(eigenvalues[0])) < creal(eps)){
    eigenvalues[0] = 0.5*(eigenvalues[0]*eigenvalues[0]/eps + eps);
    eigenvalues[1] = eigenvalues[0];
    eigenvalues[2] = eigenvalues[0];
  }
  else{
    eigenvalues[0] = eigenvalues[1] = eigenvalues[2] = cs_abs(eigenvalues[0]);
  }

  eigL = thetaL + cL; eigR = thetaR + cR; eig = eigenvalues[3];
  eps = cs_max((eig - eigL), (eigR - eig));
  eps = cs_max(0.0, eps);
  if(creal(cs_abs(eigenvalues[3])) < creal(eps)) eigenvalues[3] = 0.5*(eigenvalues[3]*eigenvalues[3]/eps + eps);
  else eigenvalues[3] = cs_abs(eigenvalues[3]);

  eigL = thetaL - cL; eigR = thetaR - cR; eig = eigenvalues[4];
  eps = cs_max((eig - eigL), (eigR - eig));
  eps = cs_max(0.0, eps);
  if(creal(cs_abs(eigenvalues[4])) < creal(eps)) eigenvalues[4] = 0.5*(eigenvalues[4]*eigenvalues[4]/eps + eps);
  else eigenvalues[4] = cs_abs(eigenvalues[4]);

  for(i = 0; i < 5; i++) dQ[i] = QR[i] - QL[i];
  cs_matvec(Tinv, dQ, dv, 5);
  for(i = 0; i < 5; i++) dv[i] *= cs_abs(eigenvalues[i]);
  cs_matvec(T, dv, dr, 5);
  cs_phys_flux(QL, avec, vdotn, gamma, fluxL);
  cs_phys_flux(QR, avec, vdotn, gamma, fluxR);
  for(i = 0; i < 5; i++) flux[i] = 0.5*area*(fluxL[i] + fluxR[i] - dr[i]);
}

static double* cs_block(const int* ia, const int* ja, double* A, int row, int col)
{
  int k;
  for(k = ia[row]; k < ia[row+1]; k++) if(ja[k] == col) return &A[(size_t)k*25];
  return NULL;
}

/* Kernel_NumJac_Complex over the interior edges, blocks added into A like the driver's scatter */
void orc_jac_edges_complex(int nedge, const int* edges_n, const double* edges_a, const double* q, int nvars, double gamma_r,
			   const int* ia, const int* ja, double* A)
{
  int e, i, j, k;
  const cplx h = 1.0e-11*_Complex_I;
  const cplx gamma = gamma_r;
  for(e = 0; e < nedge; e++){
    int l = edges_n[2*e], r = edges_n[2*e+1];
    cplx avec[4], QL[5], QR[5], QPL[5], QPR[5], fluxL[5], fluxR[5];
    double tempL[25], tempR[25], *pR, *pL;
    for(j = 0; j < 4; j++) avec[j] = edges_a[4*e + j];
    for(j = 0; j < 5; j++){ QL[j] = q[(size_t)l*nvars + j]; QR[j] = q[(size_t)r*nvars + j]; }
    for(i = 0; i < 5; i++){
      memcpy(QPL, QL, sizeof(QL));
      memcpy(QPR, QR, sizeof(QR));
      QPL[i] += h;
      QPR[i] += h;
      cs_roe_flux(QPL, QR, avec, 0.0, gamma, fluxL);
      cs_roe_flux(QL, QPR, avec, 0.0, gamma, fluxR);
      for(j = 0; j < 5; j++){      /* EqnSet::NumericalFlux: a NaN real part is kneecapped (eqnset.tcc:73-88) */
	if(isnan(creal(fluxL[j]))) fluxL[j] = 0.0;
	if(isnan(creal(fluxR[j]))) fluxR[j] = 0.0;
      }
      for(j = 0; j < 5; j++) tempL[j*5 + i] = -cimag(fluxL[j])/cimag(h);
      for(j = 0; j < 5; j++) tempR[j*5 + i] = cimag(fluxR[j])/cimag(h);
    }
    pR = cs_block(ia, ja, A, l, r);
    pL = cs_block(ia, ja, A, r, l);
    for(k = 0; k < 25; k++) pR[k] += tempR[k];
    for(k = 0; k < 25; k++) pL[k] += tempL[k];
  }
}
