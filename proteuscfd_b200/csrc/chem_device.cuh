// chem_device.cuh -- device functions of the finite-rate chemistry model (ChemModel / Reaction / Species of the
// reference), shared by pcfd_chem.cu (stand-alone source-term entry points) and pcfd_fr.cu (reacting eqnset).
// See pcfd_chem.cu for the parity notes (operation order is the reference's; exp / log / pow are CUDA libm).
#pragma once

#include <cuda_runtime.h>

#include <cmath>

#include "../../include/pcfd.h"

namespace chemdev {

constexpr double UNIV_R = 8.31447215;   // chem_constants.h:5

// macros.h:24-30 isWholeNumber (truncation towards zero, as written there)
__device__ __forceinline__ bool is_whole(double x) { return (int)(x + 0.5) == (int)x; }

// std::pow(Type, Int) of the reference promotes the exponent to double (C++11): same value; the small integer
// powers that stoichiometric coefficients produce are formed exactly as pow does (x^0 = 1, x^1 = x)
__device__ __forceinline__ double pow_stoich(double x, double nu) {
  if (is_whole(nu)) {
    const int k = (int)nu;
    if (k == 0) return 1.0;
    if (k == 1) return x;
    return pow(x, (double)k);
  }
  return pow(x, nu);
}

// reaction.tcc:607-624, 705-760
__device__ __forceinline__ double rate_constant(int type, double A, double EA, double n, double T) {
  switch (type) {
    case 0: return A * exp(-EA / (UNIV_R * T));
    case 1: return A * pow(T, n) * exp(-EA / (UNIV_R * T));
    case 2: return A * pow(T, n) * exp(-EA / T);
    case 3: return A * pow(T, n);
    default: return -999.0;
  }
}

// Species::GetThermoCoeff (species.tcc:96-138): which NASA-7 range
__device__ __forceinline__ int thermo_range(double T) {
  if (T < 200.0) return 0;
  if (T > 6000.0) return 1;
  return (T > 1000.0) ? 1 : 0;
}

// Reaction::GetEquilibriumReactionRate (reaction.tcc:626-680)
static __device__ double equilibrium_constant(const pcfd_chem_model* __restrict__ m, int j, double T) {
  double nu = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0, d4 = 0.0, d5 = 0.0, d6 = 0.0, d7 = 0.0;
  const int rng = thermo_range(T);
  for (int i = 0; i < m->nsp[j]; i++) {
    const double* a = m->nasa7[m->species[j][i]][rng];
    const double dnu = m->nupp[j][i] - m->nup[j][i];
    nu += dnu;
    d1 += dnu * a[0]; d2 += dnu * a[1]; d3 += dnu * a[2]; d4 += dnu * a[3];
    d5 += dnu * a[4]; d6 += dnu * a[5]; d7 += dnu * a[6];
  }
  const double Kp = exp(d1 * (log(T) - 1.0) + T * (d2 / 2.0 + T * (d3 / 6.0 + T * (d4 / 12.0 + d5 / 20.0 * T))) - d6 / T + d7);
  double Kc = Kp;
  if (!(fabs(nu) < 1.0e-15)) {
    const double Pref = 101325.0;
    if (is_whole(nu)) Kc *= pow(Pref / (UNIV_R * T), (double)(int)nu);
    else Kc *= pow(Pref / (UNIV_R * T), nu);
  }
  return Kc;
}

// wdot[i] for all species of one state; rhoi dimensional [kg/m^3], T [K]
static __device__ void mass_production(const pcfd_chem_model* __restrict__ m, const double* rhoi, double T, double* wdot) {
  const int ns = m->nspecies, nr = m->nreactions;
  double X[PCFD_CHEM_MAX_SPECIES];
  for (int k = 0; k < ns; k++) X[k] = rhoi[k] / m->mw[k];
  for (int i = 0; i < ns; i++) wdot[i] = 0.0;
  // reactions outer here, species outer in the reference: every wdot[i] still receives its reaction terms in
  // reaction order, which is all its rounding depends on
  for (int j = 0; j < nr; j++) {
    const double Kf = rate_constant(m->rxn_type[j], m->A[j], m->EA[j], m->n[j], T);
    double Kb;
    if (m->backward_given[j]) Kb = rate_constant(m->rxn_type_b[j], m->Ab[j], m->EAb[j], m->nb[j], T);
    else Kb = Kf / equilibrium_constant(m, j, T);
    double prod_form = 1.0, prod_dest = 1.0;
    for (int k = 0; k < m->nsp[j]; k++) {
      const double x = X[m->species[j][k]];
      prod_form *= pow_stoich(x, m->nup[j][k]);
      prod_dest *= pow_stoich(x, m->nupp[j][k]);
    }
    double Mconc = 1.0;
    if (m->third_body[j]) {
      Mconc = 0.0;
      for (int k = 0; k < ns; k++) Mconc += X[k];
      for (int k = 0; k < m->nsp[j]; k++) Mconc += (m->tbeff[j][k] - 1.0) * X[m->species[j][k]];
    }
    const double net = Kf * prod_form - Kb * prod_dest;
    for (int k = 0; k < m->nsp[j]; k++) {
      const double nu_ir = m->nupp[j][k] - m->nup[j][k];
      if (nu_ir == 0.0) continue;   // AlmostEqualRelative(nu_ir, 0, 1e-15) (reaction.tcc:789)
      const int g = m->species[j][k];
      double w = nu_ir * Mconc * net;
      w *= m->mw[g];
      wdot[g] += w;
    }
  }
}

// The same sum with the rate constants handed in: the finite-difference source Jacobian (kfr_jac_node) perturbs one
// density at a time at a fixed temperature, so Kf / Kb (all the exp / pow / log of the model) are evaluated once per
// temperature instead of once per perturbation.  Same operations on the same values as mass_production.
static __device__ void rate_constants(const pcfd_chem_model* __restrict__ m, double T, double* Kf, double* Kb) {
  const int nr = m->nreactions;
  for (int j = 0; j < nr; j++) {
    Kf[j] = rate_constant(m->rxn_type[j], m->A[j], m->EA[j], m->n[j], T);
    if (m->backward_given[j]) Kb[j] = rate_constant(m->rxn_type_b[j], m->Ab[j], m->EAb[j], m->nb[j], T);
    else Kb[j] = Kf[j] / equilibrium_constant(m, j, T);
  }
}
static __device__ void mass_production_rates(const pcfd_chem_model* __restrict__ m, const double* rhoi, const double* Kf,
                                             const double* Kb, double* wdot) {
  const int ns = m->nspecies, nr = m->nreactions;
  double X[PCFD_CHEM_MAX_SPECIES];
  for (int k = 0; k < ns; k++) X[k] = rhoi[k] / m->mw[k];
  for (int i = 0; i < ns; i++) wdot[i] = 0.0;
  for (int j = 0; j < nr; j++) {
    double prod_form = 1.0, prod_dest = 1.0;
    for (int k = 0; k < m->nsp[j]; k++) {
      const double x = X[m->species[j][k]];
      prod_form *= pow_stoich(x, m->nup[j][k]);
      prod_dest *= pow_stoich(x, m->nupp[j][k]);
    }
    double Mconc = 1.0;
    if (m->third_body[j]) {
      Mconc = 0.0;
      for (int k = 0; k < ns; k++) Mconc += X[k];
      for (int k = 0; k < m->nsp[j]; k++) Mconc += (m->tbeff[j][k] - 1.0) * X[m->species[j][k]];
    }
    const double net = Kf[j] * prod_form - Kb[j] * prod_dest;
    for (int k = 0; k < m->nsp[j]; k++) {
      const double nu_ir = m->nupp[j][k] - m->nup[j][k];
      if (nu_ir == 0.0) continue;
      const int g = m->species[j][k];
      double w = nu_ir * Mconc * net;
      w *= m->mw[g];
      wdot[g] += w;
    }
  }
}

}  // namespace chemdev
