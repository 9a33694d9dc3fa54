// Per-campaign drivers of the variational path, shared by the CUDA kernels and the host emulation.
#pragma once
#include "variational.cuh"

namespace stp {

struct VarFitArgs {
  int lik, epochs, fresh;
  double lr_ngd, lr_adam;
  double hyp0[kMaxD + 4];        // c, ls x d, os, noise, rho at construction (models.py, App. A.2)
  VarHyperPrior prior;
  double gh_x[kGH], gh_w[kGH];   // Gauss-Hermite nodes / weights, fp32-rounded upstream (App. D-9)
  unsigned long long seed;
};

// Load X (transposed) and the standardised targets: ys = (y - mean) / (std_population + 1e-10)
// (optim.py:126-127 via utils.TorchStandardScaler).
template <class Team>
STP_HD void var_load(const Team& tm, const VarSmem& s, int n, int d, const double* X, const double* y,
                     int standardise) {
  for (int idx = tm.rank(); idx < n * d; idx += tm.size()) {
    const int i = idx / d, k = idx - i * d;
    s.Xt[k * s.cap + i] = X[idx];
  }
  double a = 0.0;
  for (int i = tm.rank(); i < n; i += tm.size()) a += y[i];
  const double mean = standardise ? tm.sum(a) / n : 0.0;
  double q = 0.0;
  for (int i = tm.rank(); i < n; i += tm.size()) { const double t = y[i] - mean; q += t * t; }
  const double sd = standardise ? sqrt(tm.sum(q) / n) + 1e-10 : 1.0;
  tm.sync();
  for (int i = tm.rank(); i < n; i += tm.size()) s.ys[i] = (y[i] - mean) / sd;
  tm.sync();
}

// Fresh model (runners.py:66-74 builds a new one every trial, App. D-7): Z = X, hyper-parameters at
// their construction values, q(u): theta1 = 1e-3 * noise, Theta2 = -I/2, Adam moments zero.
template <class Team>
STP_HD void var_init_state(const Team& tm, const VarModel& M, int n, int d, int cap, const double* X,
                           const double* init_noise, const VarFitArgs& a, unsigned long long stream) {
  for (int idx = tm.rank(); idx < n * d; idx += tm.size()) M.Z[idx] = X[idx];
  for (int k = tm.rank(); k < d + 4; k += tm.size()) M.hyp[k] = a.hyp0[k];
  for (int i = tm.rank(); i < n; i += tm.size()) {
    double e;
    if (init_noise) e = init_noise[i];
    else { double z1; philox_normal2(a.seed ^ 0x5bd1e995u, stream, (unsigned long long)i, e, z1); }
    M.nat1[i] = 1e-3 * e;
  }
  for (int idx = tm.rank(); idx < n * n; idx += tm.size()) {
    const int i = idx / n, k = idx - i * n;
    M.nat2[(size_t)i * cap + k] = (i == k) ? -0.5 : 0.0;
  }
  for (int idx = tm.rank(); idx < cap * d + d + 4; idx += tm.size()) { M.adam_m[idx] = 0.0; M.adam_v[idx] = 0.0; }
  tm.sync();
}

// train_natural_variational_model (optim.py:105-181): `epochs` full-batch steps.  Returns the
// status of the first failing epoch (0 = all fine) and leaves the last loss in *last_loss.
template <class Team>
STP_HD int var_fit_run(const Team& tm, const VarSmem& s, const VarWork& w, const VarModel& M, int n, int d,
                       const VarFitArgs& a, double* loss_hist, double* last_loss) {
  int status = 0;
  double last = NAN;
  for (int e = 0; e < a.epochs; ++e) {
    const VarEpochOut r = var_epoch(tm, s, w, M, n, d, a.lik, a.prior, a.gh_x, a.gh_w, e + 1, a.lr_ngd, a.lr_adam,
                                    1, (double*)nullptr);
    if (r.status) { status = r.status; break; }
    last = r.loss;
    if (loss_hist && tm.rank() == 0) loss_hist[e] = r.loss;
  }
  *last_loss = last;
  return status;
}

}  // namespace stp
