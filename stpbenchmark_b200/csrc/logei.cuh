// LogExpectedImprovement of botorch's analytic acquisition (what runners.py:88-90 calls),
// restated for the device: log_h(u) = log(phi(u) + u Phi(u)) in three branches
// (SURVEY.md App. A.5), and LogEI = log_h((mean - best_f) / sigma) + log sigma with
// sigma = sqrt(max(var, 1e-12)).
#pragma once
#include "team.cuh"

namespace stp {

#if defined(__CUDA_ARCH__)
STP_HD double erfcx_pos(double x) { return erfcx(x); }
#else
// Host emulation only (glibc has no erfcx); argument is >= 1/sqrt(2) on every call site.
inline double erfcx_pos(double x) {
  if (x < 25.0) return exp(x * x) * erfc(x);
  const double r = 1.0 / (x * x);
  return (1.0 / (x * 1.7724538509055160273)) *
         (1.0 + r * (-0.5 + r * (0.75 + r * (-1.875 + r * (6.5625 - 29.53125 * r)))));
}
#endif

STP_HD double log1mexp(double x) {
  // log(1 - exp(x)), x < 0
  return (x > -0.69314718055994530942) ? log(-expm1(x)) : log1p(-exp(x));
}

STP_HD double log_h(double u) {
  const double kInvSqrt2 = 0.70710678118654752440;
  const double kInvSqrt2Pi = 0.39894228040143267794;
  const double kLog2Pi = 1.83787706640934548356;
  const double kLogSqrtPiOver2 = 0.22579135264472743236;  // log(sqrt(pi / 2))
  if (u > -1.0) {
    const double pdf = kInvSqrt2Pi * exp(-0.5 * u * u);
    const double cdf = 0.5 * erfc(-kInvSqrt2 * u);
    return log(pdf + u * cdf);
  }
  if (u != u) return u;
  const double head = -0.5 * (u * u + kLog2Pi);
  if (u > -1.0e6) {
    const double w = log(erfcx_pos(-kInvSqrt2 * u) * fabs(u)) + kLogSqrtPiOver2;
    return head + log1mexp(w);
  }
  return head - 2.0 * log(fabs(u));
}

STP_HD double log_ei(double mean, double var, double best_f) {
  const double sigma = sqrt(var > 1e-12 ? var : (var != var ? var : 1e-12));
  const double u = (mean - best_f) / sigma;
  return log_h(u) + log(sigma);
}

// LogEI when sigma (and log sigma) are fixed across the call, as for the MC models.
STP_HD double log_ei_fixed_sigma(double f, double best_f, double inv_sigma, double log_sigma) {
  return log_h((f - best_f) * inv_sigma) + log_sigma;
}

}  // namespace stp
