// Per-campaign drivers of the exact-GP path: what ONE team (CTA) does for ONE campaign.
// Shared verbatim by the CUDA kernels (stp_kernels.cu) and the host emulation used by the
// CPU tests (tests/hostemu/hostemu.cpp).
#pragma once
#include "exact.cuh"
#include "lbfgsb.cuh"

namespace stp {

struct ExactFitResult {
  double f;
  int nit, nfev, status;  // status: 0 converged, 1 cap reached, 2 abnormal line search, 3 non-finite / not PD
};

// Fit theta by L-BFGS-B on the exact MLL (train_exact_model_botorch, optim.py:11-32).
// `theta` (d+3 doubles, shared memory) holds the start on entry and the optimum on exit.
// `S` lives in shared memory; `flag` is one shared int.
template <class Team>
STP_HD ExactFitResult exact_fit_run(const Team& tm, const ExactSmem& s, LbfgsbState* S, int* flag, int n, int d,
                                    double* theta, const double* lb, const double* ub, const ExactPrior& pr,
                                    const LbfgsbOpts& opts) {
  const int np = d + 3;
  double* g = s.small + 16;
  const typename Team::Lanes L = tm.lanes();
  tm.sync();
  if (tm.warp() == 0) {   // the optimiser is run by the first warp, lane i <-> parameter i
    lbfgsb_init(L, *S, np, theta, lb, ub, opts);
    if (L.lane() == 0) *flag = 1;
  }
  tm.sync();
  while (*flag) {
    double f;
    const int st = exact_mll_fg(tm, s, n, d, S->x, pr, &f, g);
    tm.sync();
    STP_PHASE(-1);
    if (tm.warp() == 0) {
      if (st) {
        f = NAN;
        STP_OWNED(L, k, np) g[k] = NAN;
        L.sync();
      }
      const int more = lbfgsb_advance(L, *S, f, g);
      if (L.lane() == 0) *flag = more;
    }
    tm.sync();
    STP_PHASE(7);
  }
  STP_OWNED(tm, k, np) theta[k] = S->x[k];
  tm.sync();
  ExactFitResult r;
  r.f = S->f; r.nit = S->iter; r.nfev = S->nfev; r.status = S->status;
  return r;
}

}  // namespace stp
