// eqnset_compressible_cs.cuh -- the perfect-gas Roe flux on a complex state, for the complex-step field Jacobian
// (Kernel_NumJac_Complex, ucs/jacobian.tcc:370-433; Param::fieldJacType == 2).
//
// The reference instantiates its eqnset on std::complex<double> and perturbs one conservative variable by i*1e-11; the
// Jacobian column is imag(flux) / 1e-11.  The functions below are the flux functions of eqnset_compressible.cuh
// (roe_variables, phys_flux, entropy_fix, row5, roe_flux) as templates over the state type -- the SAME statements,
// generated from that text, so that instantiated on double they are eq::roe_flux bit for bit
// (tests/test_host_emulation.py checks exactly that) -- and `cplx` carries the arithmetic std::complex<double> has
// in the reference build: product (ac - bd, ad + bc), quotient by Smith's formula in the order libgcc's __divdc3
// evaluates it, square root of a number with positive real part as glibc's csqrt forms it, and the reference's own
// overloads for complex numbers (macros.h:32-68: CAbs flips the sign by the real part, MAX and every comparison look at
// the real part).  Area vector, gamma and vdotn are real here: the reference holds them as complex numbers with a zero
// imaginary part, whose products and quotients reduce to the real-scalar forms below.
#pragma once

namespace eqcs {

struct cplx {
  double re, im;
  __device__ __forceinline__ cplx() {}
  __device__ __forceinline__ cplx(double r) : re(r), im(0.0) {}
  __device__ __forceinline__ cplx(double r, double i) : re(r), im(i) {}
};
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return cplx(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return cplx(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx operator-(cplx a) { return cplx(-a.re, -a.im); }
__device__ __forceinline__ cplx operator+(cplx a, double b) { return cplx(a.re + b, a.im); }
__device__ __forceinline__ cplx operator+(double a, cplx b) { return cplx(a + b.re, b.im); }
__device__ __forceinline__ cplx operator-(cplx a, double b) { return cplx(a.re - b, a.im); }
__device__ __forceinline__ cplx operator-(double a, cplx b) { return cplx(a - b.re, -b.im); }
__device__ __forceinline__ cplx operator*(cplx a, cplx b) { return cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ cplx operator*(double a, cplx b) { return cplx(a * b.re, a * b.im); }
__device__ __forceinline__ cplx operator*(cplx a, double b) { return cplx(a.re * b, a.im * b); }
__device__ __forceinline__ cplx operator/(cplx a, double b) { return cplx(a.re / b, a.im / b); }
// __divdc3 (libgcc): Smith's formula, the branch taken by the larger denominator part
__device__ __forceinline__ cplx operator/(cplx n, cplx d) {
  if (fabs(d.re) < fabs(d.im)) {
    const double ratio = d.re / d.im, denom = d.re * ratio + d.im;
    return cplx((n.re * ratio + n.im) / denom, (n.im * ratio - n.re) / denom);
  }
  const double ratio = d.im / d.re, denom = d.im * ratio + d.re;
  return cplx((n.im * ratio + n.re) / denom, (n.im - n.re * ratio) / denom);
}
__device__ __forceinline__ cplx operator/(double a, cplx d) { return cplx(a, 0.0) / d; }
__device__ __forceinline__ cplx& operator+=(cplx& a, cplx b) { a = a + b; return a; }
__device__ __forceinline__ cplx& operator*=(cplx& a, cplx b) { a = a * b; return a; }
__device__ __forceinline__ bool operator<(cplx a, cplx b) { return a.re < b.re; }
__device__ __forceinline__ bool operator<(cplx a, double b) { return a.re < b; }
__device__ __forceinline__ cplx fabs(cplx a) { return (a.re < 0.0) ? -a : a; }                  // CAbs, macros.h:62-68
// csqrt (glibc) for a positive real part: r = sqrt((|z| + x) / 2) with |z| = x to the last bit for |y| << x,
// imaginary part y / r / 2; an exactly real number keeps an exactly zero imaginary part
__device__ __forceinline__ cplx sqrt(cplx a) {
  if (a.im == 0.0) return cplx(::sqrt(a.re), a.im);
  // |z|: for |y| < 2^-27 |x| the correctly rounded modulus IS |x| (1 + y^2 / 2x^2 < 1 + 2^-55); keeps the complex-step
  // regime independent of the device library's hypot, which is allowed an ulp
  const double ar = ::fabs(a.re), ai = ::fabs(a.im);
  const double mod = (ai < ar * 7.450580596923828e-09) ? ar : ::hypot(a.re, a.im);
  const double r = ::sqrt(0.5 * (mod + a.re));
  return cplx(r, 0.5 * (a.im / r));
}
__device__ __forceinline__ double fabs(double a) { return ::fabs(a); }
__device__ __forceinline__ double sqrt(double a) { return ::sqrt(a); }
template <class T>
__device__ __forceinline__ T maxd(T x, T y) { return (y < x) ? x : y; }                          // MAX, macros.h:32-36

template <class T>
__device__ __forceinline__ void roe_variables(const T* QL, const T* QR, double gamma, T* Qroe) {
  const double gm1 = gamma - 1.0;
  const T rhoL = QL[0], rhoR = QR[0];
  const T uL = QL[1] / QL[0], uR = QR[1] / QR[0];
  const T vL = QL[2] / QL[0], vR = QR[2] / QR[0];
  const T wL = QL[3] / QL[0], wR = QR[3] / QR[0];
  const T EL = QL[4], ER = QR[4];
  const T v2L = uL * uL + vL * vL + wL * wL;
  const T v2R = uR * uR + vR * vR + wR * wR;
  const T PL = gm1 * (EL - 0.5 * rhoL * v2L);
  const T PR = gm1 * (ER - 0.5 * rhoR * v2R);
  const T hL = (EL + PL) / rhoL;
  const T hR = (ER + PR) / rhoR;
  const T rho = sqrt(rhoL * rhoR);
  const T sigma = rho / (rhoL + rho);
  const T u = uL + sigma * (uR - uL);
  const T v = vL + sigma * (vR - vL);
  const T w = wL + sigma * (wR - wL);
  const T h = hL + sigma * (hR - hL);
  const T v2h = 0.5 * (u * u + v * v + w * w);
  Qroe[0] = rho;
  Qroe[1] = rho * u;
  Qroe[2] = rho * v;
  Qroe[3] = rho * w;
  Qroe[4] = rho / gamma * (h + gm1 * v2h);
}

template <class T>
__device__ __forceinline__ void phys_flux(const T* Q, const double* n, double vdotn, double gamma, T* f,
                                          T* Pout = nullptr) {
  const T rho = Q[0];
  const T u = Q[1] / rho, v = Q[2] / rho, w = Q[3] / rho;
  const T rEt = Q[4];
  const T v2h = 0.5 * (u * u + v * v + w * w);
  const T P = (gamma - 1.0) * (rEt - rho * v2h);
  const T ht = (rEt + P) / rho;
  const T rhotheta = rho * (n[0] * u + n[1] * v + n[2] * w + vdotn);
  f[0] = rhotheta;
  f[1] = (u * rhotheta + P * n[0]);
  f[2] = (v * rhotheta + P * n[1]);
  f[3] = (w * rhotheta + P * n[2]);
  f[4] = (ht * rhotheta - vdotn * P);
  if (Pout) *Pout = P;   // == ComputePressure(Q): the quantity BadExtrapolation tests
}

template <class T>
__device__ __forceinline__ T entropy_fix(T eig, T eigL, T eigR) {
  T eps = maxd((eig - eigL), (eigR - eig));
  eps = maxd(T(0.0), eps);
  if (fabs(eig) < eps) return 0.5 * (eig * eig / eps + eps);
  return fabs(eig);
}

template <class T>
__device__ __forceinline__ T row5(T a0, T a1, T a2, T a3, T a4, const T* v) {
  T s = a0 * v[0];
  s += a1 * v[1];
  s += a2 * v[2];
  s += a3 * v[3];
  s += a4 * v[4];
  return s;
}

template <class T>
__device__ __forceinline__ void roe_flux(const T* QL, const T* QR, const double* n, double vdotn,
                                         double gamma, T* flux, bool* bad = nullptr) {
  T Qroe[5];
  roe_variables(QL, QR, gamma, Qroe);
  const double gm1 = gamma - 1.0;
  const double area = n[3];
  const double nx = n[0], ny = n[1], nz = n[2];

  // --- Roe-state eigensystem quantities
  const T rho = Qroe[0];
  const T u = Qroe[1] / rho, v = Qroe[2] / rho, w = Qroe[3] / rho;
  const T thetaf = u * nx + v * ny + w * nz;
  const T th = thetaf + vdotn;
  const T v2h = 0.5 * (u * u + v * v + w * w);
  const T P = gm1 * (Qroe[4] - rho * v2h);
  const T c2 = gamma * P / rho;
  const T c = sqrt(c2);

  // --- left/right wave speeds for the entropy fix
  T thetaL, thetaR, cL, cR;
  {
    const T rhoL = QL[0];
    const T uL = QL[1] / rhoL, vL = QL[2] / rhoL, wL = QL[3] / rhoL;
    const T PL = gm1 * (QL[4] - 0.5 * rhoL * (uL * uL + vL * vL + wL * wL));
    const T rhoR = QR[0];
    const T uR = QR[1] / rhoR, vR = QR[2] / rhoR, wR = QR[3] / rhoR;
    const T PR = gm1 * (QR[4] - 0.5 * rhoR * (uR * uR + vR * vR + wR * wR));
    thetaL = uL * nx + vL * ny + wL * nz + vdotn;
    thetaR = uR * nx + vR * ny + wR * nz + vdotn;
    cR = sqrt(gamma * PR / rhoR);
    cL = sqrt(gamma * PL / rhoL);
  }
  T lam[5];
  lam[0] = lam[1] = lam[2] = entropy_fix(th, thetaL, thetaR);
  lam[3] = entropy_fix(th + c, thetaL + cL, thetaR + cR);
  lam[4] = entropy_fix(th - c, thetaL - cL, thetaR - cR);

  T dQ[5], dv[5];
#pragma unroll
  for (int i = 0; i < 5; i++) dQ[i] = QR[i] - QL[i];

  // dv = Tinv * dQ (rows of Tinv, compressible.tcc:640-673)
  dv[0] = row5<T>(nx - nz * v / rho + ny * w / rho - nx / c2 * v2h * gm1, nx / c2 * u * gm1, nz / rho + nx / c2 * v * gm1,
               -ny / rho + nx / c2 * w * gm1, -nx / c2 * gm1, dQ);
  dv[1] = row5<T>(ny + nz * u / rho - nx * w / rho - ny / c2 * v2h * gm1, -nz / rho + ny / c2 * u * gm1, ny / c2 * v * gm1,
               nx / rho + ny / c2 * w * gm1, -ny / c2 * gm1, dQ);
  dv[2] = row5<T>(nz - ny * u / rho + nx * v / rho - nz / c2 * v2h * gm1, ny / rho + nz / c2 * u * gm1,
               -nx / rho + nz / c2 * v * gm1, nz / c2 * w * gm1, -nz / c2 * gm1, dQ);
  dv[3] = row5<T>(-0.5 / rho * (thetaf - gm1 * v2h / c), 0.5 / rho * (nx - gm1 * u / c), 0.5 / rho * (ny - gm1 * v / c),
               0.5 / rho * (nz - gm1 * w / c), 0.5 / rho * (gm1 / c), dQ);
  dv[4] = row5<T>(0.5 / rho * (thetaf + gm1 * v2h / c), -0.5 / rho * (nx + gm1 * u / c), -0.5 / rho * (ny + gm1 * v / c),
               -0.5 / rho * (nz + gm1 * w / c), +0.5 / rho * (gm1 / c), dQ);
#pragma unroll
  for (int i = 0; i < 5; i++) dv[i] *= fabs(lam[i]);

  // dr = T * dv (rows of T, compressible.tcc:606-637)
  T dr[5];
  const T rc = rho / c;
  dr[0] = row5<T>(nx, ny, nz, rc, rc, dv);
  dr[1] = row5<T>(u * nx, u * ny - rho * nz, u * nz + rho * ny, rho * (u / c + nx), rho * (u / c - nx), dv);
  dr[2] = row5<T>(v * nx + rho * nz, v * ny, v * nz - rho * nx, rho * (v / c + ny), rho * (v / c - ny), dv);
  dr[3] = row5<T>(w * nx - rho * ny, w * ny + rho * nx, w * nz, rho * (w / c + nz), rho * (w / c - nz), dv);
  dr[4] = row5<T>(v2h * nx + rho * (v * nz - w * ny), v2h * ny + rho * (w * nx - u * nz), v2h * nz + rho * (u * ny - v * nx),
               rho * (v2h / c + thetaf + c / gm1), rho * (v2h / c - thetaf + c / gm1), dv);

  T fL[5], fR[5], pL, pR;
  phys_flux(QL, n, vdotn, gamma, fL, &pL);
  phys_flux(QR, n, vdotn, gamma, fR, &pR);
#pragma unroll
  for (int i = 0; i < 5; i++) flux[i] = 0.5 * area * (fL[i] + fR[i] - dr[i]);
  if (bad)
    *bad = (pL < 1.0e-10) || (QL[0] < 0.0) || (QL[4] < 1.0e-10) || (pR < 1.0e-10) || (QR[0] < 0.0) || (QR[4] < 1.0e-10) ||
           (P < 1.0e-10) || (rho < 0.0) || (Qroe[4] < 1.0e-10);
}

}  // namespace eqcs
