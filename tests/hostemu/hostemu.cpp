// Host emulation of the per-campaign device routines -- TEST SCAFFOLDING, not a CPU path.
//
// The CUDA kernels of stpbenchmark_b200/csrc are written against a "team" abstraction
// (team.cuh).  This file compiles the very same headers with g++ and a one-thread team so
// that the algebra (sweep, Cholesky, MLL gradient, L-BFGS-B, LogEI, ELBO backward, ...) can be
// checked against the oracle on the GPU-less build box.  It is built on demand by
// tests/hostemu/__init__.py, is never imported by the stpbenchmark_b200 package and is not
// part of libstp_b200.so.  Pointers are HOST pointers here.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../stpbenchmark_b200/csrc/exact_campaign.cuh"
#include "../../stpbenchmark_b200/csrc/var_campaign.cuh"

using namespace stp;

namespace {
int g_kern = 0, g_acq = 0;   // emu_set_model: kernel / acquisition variant of the next emu_exact_* calls
double g_beta = 0.0;
int g_large = 0;             // emu_set_large: carve the campaign as the stp_exact_*_large entry points do
struct Scratch {
  std::vector<double> buf, square;
  ExactSmem s;
  Scratch(int cap, int d) {
    if (g_large) {   // stp_exact_*_large: the working square in its own ("global") buffer
      buf.assign(exact_large_smem_doubles(cap, d, 1) + 8, 0.0);
      square.assign(exact_large_work_doubles(cap), 0.0);
      s = exact_smem_carve_large(buf.data(), square.data(), cap, d, 1);
    } else {
      buf.assign(exact_smem_doubles(cap, d, 1, 1) + 8, 0.0);
      s = exact_smem_carve(buf.data(), cap, d, 1, 1);
    }
    s.kern = g_kern; s.acq_kind = g_acq; s.acq_beta = g_beta;
  }
};
ExactPrior prior_from(const double* p) { return ExactPrior{p[0], p[1], p[2], p[3], p[4], p[5]}; }
}  // namespace

extern "C" {

void emu_set_large(int on) { g_large = on; }

int emu_exact_mll_fg(int B, int n_max, int d, const double* X, const double* y, const int* n_arr,
                     const double* theta, const double* prior, double* f, double* g, int* status) {
  SoloTeam tm;
  const ExactPrior pr = prior_from(prior);
  for (int b = 0; b < B; ++b) {
    const int n = n_arr ? n_arr[b] : n_max;
    Scratch sc(n_max, d);
    exact_load(tm, sc.s, n, d, X + (size_t)b * n_max * d, y + (size_t)b * n_max);
    double* th = sc.s.small;
    for (int k = 0; k < d + 3; ++k) th[k] = theta[b * (d + 3) + k];
    status[b] = exact_mll_fg(tm, sc.s, n, d, th, pr, &f[b], sc.s.small + 16);
    for (int k = 0; k < d + 3; ++k) g[b * (d + 3) + k] = sc.s.small[16 + k];
  }
  return 0;
}

int emu_exact_fit(int B, int n_max, int d, const double* X, const double* y, const int* n_arr, double* theta,
                  const double* lb, const double* ub, const double* prior, int m, double factr, double pgtol,
                  int maxiter, int maxfun, int maxls, double* f_out, int* nit, int* nfev, int* status) {
  SoloTeam tm;
  const ExactPrior pr = prior_from(prior);
  const LbfgsbOpts o{m, maxiter, maxfun, maxls, factr, pgtol};
  for (int b = 0; b < B; ++b) {
    const int n = n_arr ? n_arr[b] : n_max;
    Scratch sc(n_max, d);
    exact_load(tm, sc.s, n, d, X + (size_t)b * n_max * d, y + (size_t)b * n_max);
    LbfgsbState S;
    int flag = 0;
    double* th = sc.s.small;
    for (int k = 0; k < d + 3; ++k) th[k] = theta[b * (d + 3) + k];
    const ExactFitResult r = exact_fit_run(tm, sc.s, &S, &flag, n, d, th, lb, ub, pr, o);
    for (int k = 0; k < d + 3; ++k) theta[b * (d + 3) + k] = th[k];
    f_out[b] = r.f; nit[b] = r.nit; nfev[b] = r.nfev; status[b] = r.status;
  }
  return 0;
}

int emu_exact_acq(int B, int n_max, int d, int N, const double* pool, const int* pool_id, const uint8_t* mask,
                  const double* X, const double* y, const int* n_arr, const double* theta, const double* best_f,
                  int* out_idx, double* out_acq, double* mean, double* var, double* acq, int* status) {
  SoloTeam tm;
  for (int b = 0; b < B; ++b) {
    const int n = n_arr ? n_arr[b] : n_max;
    Scratch sc(n_max, d);
    exact_load(tm, sc.s, n, d, X + (size_t)b * n_max * d, y + (size_t)b * n_max);
    double* th = sc.s.small;
    for (int k = 0; k < d + 3; ++k) th[k] = theta[b * (d + 3) + k];
    status[b] = exact_factorize(tm, sc.s, n, d, th);
    if (status[b]) { out_idx[b] = -1; out_acq[b] = NAN; continue; }
    const double* P = pool + (size_t)(pool_id ? pool_id[b] : 0) * N * d;
    exact_acquire(tm, sc.s, n, d, th, P, N, mask ? mask + (size_t)b * N : nullptr, best_f ? best_f[b] : NAN,
                  &out_idx[b], &out_acq[b], mean ? mean + (size_t)b * N : nullptr,
                  var ? var + (size_t)b * N : nullptr, acq ? acq + (size_t)b * N : nullptr);
  }
  return 0;
}

// ---- generic ask/tell L-BFGS-B handle (tests drive it with arbitrary Python objectives) ----
void* emu_lbfgsb_new(int np, const double* x0, const double* lb, const double* ub, int m, double factr,
                     double pgtol, int maxiter, int maxfun, int maxls) {
  LbfgsbState* S = (LbfgsbState*)calloc(1, sizeof(LbfgsbState));
  const LbfgsbOpts o{m, maxiter, maxfun, maxls, factr, pgtol};
  lbfgsb_init(*S, np, x0, lb, ub, o);
  return S;
}
void emu_lbfgsb_x(void* h, double* x) { LbfgsbState* S = (LbfgsbState*)h; memcpy(x, S->x, sizeof(double) * S->np); }
int emu_lbfgsb_advance(void* h, double f, const double* g) { return lbfgsb_advance(*(LbfgsbState*)h, f, g); }
void emu_lbfgsb_result(void* h, double* f, int* nit, int* nfev, int* status) {
  LbfgsbState* S = (LbfgsbState*)h; *f = S->f; *nit = S->iter; *nfev = S->nfev; *status = S->status;
}
void emu_lbfgsb_free(void* h) { free(h); }

void emu_set_model(int kern, int acq, double beta) { g_kern = kern; g_acq = acq; g_beta = beta; }
double emu_acq_value(int kind, double beta, double mean, double var, double best_f) {
  return acq_value(kind, beta, mean, var, best_f);
}
double emu_log_ei_t(double mean, double sd, double df, double best_f) { return log_ei_t(mean, sd, df, best_f); }
double emu_log_h(double u) { return log_h(u); }
double emu_log_h_mc(double u) { return log_h_mc(u); }
double emu_log_ei(double mean, double var, double best_f) { return log_ei(mean, var, best_f); }

#include "hostemu_var.inc"

}  // extern "C"
