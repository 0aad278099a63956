// pcfd_chem.cu -- finite-rate chemistry source term of the reacting eqnset (compressibleFR) on sm_100a.
//
// One thread per node.  The reference evaluates ChemModel::GetMassProductionRates (chem.tcc:575-583) as a double loop
// species x reactions and recomputes the rate constants k_f, k_b of a reaction for EVERY species (30 x per node for
// 5-species air, plus a std::vector allocation per call, reaction.tcc:772); they depend on T only, so here each
// reaction is evaluated once per node and its net rate is then distributed to the species in the reference's
// accumulation order (species outer, reactions inner), which keeps every partial sum identical.
//
// Parity: +, -, *, / and the operand order are the reference's (compiled with --fmad=false); exp, log and pow are
// CUDA libm (<= 1-2 ulp) where the reference calls glibc, so results agree to rounding, not bit for bit, and
// net = k_f prod_f - k_b prod_b cancels near equilibrium: tests compare against the scale of the two terms.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>

#include "../../include/pcfd.h"
#include "chem_device.cuh"

namespace {

using namespace chemdev;

__global__ void __launch_bounds__(128) k_chem_wdot(const pcfd_chem_model* __restrict__ m, int n, const double* __restrict__ rhoi,
                                                    const double* __restrict__ T, double* __restrict__ wdot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ns = m->nspecies;
  double r[PCFD_CHEM_MAX_SPECIES], w[PCFD_CHEM_MAX_SPECIES];
  for (int k = 0; k < ns; k++) r[k] = rhoi[(size_t)i * ns + k];
  mass_production(m, r, T[i], w);
  for (int k = 0; k < ns; k++) wdot[(size_t)i * ns + k] = w[k];
}

// CompressibleFREqnSet::SourceTerm (compressibleFR.tcc:1276-1316)
__global__ void __launch_bounds__(128) k_chem_source(const pcfd_chem_model* __restrict__ m, int n, int stride,
                                                      const double* __restrict__ Q, const double* __restrict__ vol,
                                                      double ref_density, double ref_time, double ref_temperature,
                                                      double* __restrict__ source) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ns = m->nspecies;
  const double* q = Q + (size_t)i * stride;
  double r[PCFD_CHEM_MAX_SPECIES], w[PCFD_CHEM_MAX_SPECIES];
  const double T = q[ns + 3] * ref_temperature;
  for (int k = 0; k < ns; k++) r[k] = q[k] * ref_density;
  mass_production(m, r, T, w);
  double* s = source + (size_t)i * (ns + 4);
  const double v = vol[i];
  for (int k = 0; k < ns; k++) {
    double wk = w[k];
    wk /= (ref_density / ref_time);
    s[k] = v * wk;
  }
  for (int k = ns; k < ns + 4; k++) s[k] = 0.0;
}

std::string g_chem_err;

}  // namespace

struct pcfd_chem {
  int device = 0;
  pcfd_chem_model host{};
  pcfd_chem_model* dev = nullptr;
  std::string err;
};

namespace {
int cfail(pcfd_chem* c, const std::string& msg) {
  if (c) c->err = msg; else g_chem_err = msg;
  return 1;
}
#define CCK(call)                                                                                     \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) return cfail(c, std::string(#call) + ": " + cudaGetErrorString(e_));       \
  } while (0)
}  // namespace

extern "C" {

const char* pcfd_chem_last_error(const pcfd_chem* c) { return c ? c->err.c_str() : g_chem_err.c_str(); }

int pcfd_chem_destroy(pcfd_chem* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->dev) cudaFree(c->dev);
  delete c;
  return 0;
}

int pcfd_chem_create(const pcfd_chem_model* model, int device, pcfd_chem** out) {
  pcfd_chem* c = nullptr;
  if (!model || !out) return cfail(c, "pcfd_chem_create: null argument");
  *out = nullptr;
  if (model->nspecies < 1 || model->nspecies > PCFD_CHEM_MAX_SPECIES || model->nreactions < 0 ||
      model->nreactions > PCFD_CHEM_MAX_REACTIONS)
    return cfail(c, "pcfd_chem_create: species / reaction count out of range");
  for (int j = 0; j < model->nreactions; j++) {
    if (model->nsp[j] < 1 || model->nsp[j] > model->nspecies) return cfail(c, "pcfd_chem_create: bad species count in a reaction");
    for (int k = 0; k < model->nsp[j]; k++)
      if (model->species[j][k] < 0 || model->species[j][k] >= model->nspecies)
        return cfail(c, "pcfd_chem_create: reaction references an unknown species");
    if (model->rxn_type[j] < 0 || model->rxn_type[j] > 3) return cfail(c, "pcfd_chem_create: unknown reaction type");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return cfail(c, "pcfd_chem_create: no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return cfail(c, "pcfd_chem_create: bad device ordinal");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return cfail(c, "pcfd_chem_create: built for sm_100a only");
  c = new pcfd_chem();
  c->device = device;
  c->host = *model;
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&c->dev), sizeof(pcfd_chem_model)) != cudaSuccess ||
      cudaMemcpy(c->dev, model, sizeof(pcfd_chem_model), cudaMemcpyHostToDevice) != cudaSuccess) {
    g_chem_err = "pcfd_chem_create: device allocation failed";
    pcfd_chem_destroy(c);
    return 1;
  }
  *out = c;
  return 0;
}

int pcfd_chem_mass_production(pcfd_chem* c, int n, const double* rhoi, const double* T, double* wdot) {
  if (!c) return 1;
  if (n < 0 || (n > 0 && (!rhoi || !T || !wdot))) return cfail(c, "pcfd_chem_mass_production: bad argument");
  if (n == 0) return 0;
  CCK(cudaSetDevice(c->device));
  const int ns = c->host.nspecies;
  double *dr = nullptr, *dT = nullptr, *dw = nullptr;
  CCK(cudaMalloc(reinterpret_cast<void**>(&dr), (size_t)n * ns * sizeof(double)));
  CCK(cudaMalloc(reinterpret_cast<void**>(&dT), (size_t)n * sizeof(double)));
  CCK(cudaMalloc(reinterpret_cast<void**>(&dw), (size_t)n * ns * sizeof(double)));
  CCK(cudaMemcpy(dr, rhoi, (size_t)n * ns * sizeof(double), cudaMemcpyHostToDevice));
  CCK(cudaMemcpy(dT, T, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  k_chem_wdot<<<(n + 127) / 128, 128>>>(c->dev, n, dr, dT, dw);
  CCK(cudaGetLastError());
  CCK(cudaMemcpy(wdot, dw, (size_t)n * ns * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(dr); cudaFree(dT); cudaFree(dw);
  return 0;
}

int pcfd_chem_source_term_device(pcfd_chem* c, int n, int stride, const void* dQ, const void* dvol, double ref_density,
                                 double ref_time, double ref_temperature, void* dsource, void* stream) {
  if (!c) return 1;
  if (n < 0 || stride < c->host.nspecies + 4 || (n > 0 && (!dQ || !dvol || !dsource)))
    return cfail(c, "pcfd_chem_source_term_device: bad argument");
  if (n == 0) return 0;
  CCK(cudaSetDevice(c->device));
  k_chem_source<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      c->dev, n, stride, static_cast<const double*>(dQ), static_cast<const double*>(dvol), ref_density, ref_time,
      ref_temperature, static_cast<double*>(dsource));
  CCK(cudaGetLastError());
  return 0;
}

int pcfd_chem_source_term(pcfd_chem* c, int n, int stride, const double* Q, const double* vol, double ref_density,
                          double ref_time, double ref_temperature, double* source) {
  if (!c) return 1;
  if (n < 0 || (n > 0 && (!Q || !vol || !source))) return cfail(c, "pcfd_chem_source_term: bad argument");
  if (n == 0) return 0;
  CCK(cudaSetDevice(c->device));
  const int neqn = c->host.nspecies + 4;
  double *dq = nullptr, *dv = nullptr, *ds = nullptr;
  CCK(cudaMalloc(reinterpret_cast<void**>(&dq), (size_t)n * stride * sizeof(double)));
  CCK(cudaMalloc(reinterpret_cast<void**>(&dv), (size_t)n * sizeof(double)));
  CCK(cudaMalloc(reinterpret_cast<void**>(&ds), (size_t)n * neqn * sizeof(double)));
  CCK(cudaMemcpy(dq, Q, (size_t)n * stride * sizeof(double), cudaMemcpyHostToDevice));
  CCK(cudaMemcpy(dv, vol, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  int rc = pcfd_chem_source_term_device(c, n, stride, dq, dv, ref_density, ref_time, ref_temperature, ds, nullptr);
  if (rc == 0) {
    cudaError_t e = cudaMemcpy(source, ds, (size_t)n * neqn * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cfail(c, cudaGetErrorString(e));
  }
  cudaFree(dq); cudaFree(dv); cudaFree(ds);
  return rc;
}

}  // extern "C"
