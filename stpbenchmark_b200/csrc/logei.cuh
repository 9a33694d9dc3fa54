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
// Host emulation only (glibc has no erfcx); argument is >= 0 on every call site.
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

// The same function for the Monte-Carlo acquisition (2048 x Nc evaluations per campaign and iteration: 97 % of
// k_var_acq).  In the middle branch botorch writes log(1 - t), t = erfcx(-u / sqrt 2) |u| sqrt(pi / 2), as
// log1mexp(log t) for the sake of autograd; the value is log1p(-t), equally conditioned (both lose the digits 1 - t
// loses for |u| >> 1) and two transcendentals cheaper.  The analytic paths keep log_h itself, so that nothing moves in
// the index-exact ExactGP / VarGP traces.
STP_HD double log_h_mc(double u) {
  const double kInvSqrt2 = 0.70710678118654752440;
  const double kInvSqrt2Pi = 0.39894228040143267794;
  const double kLog2Pi = 1.83787706640934548356;
  const double kSqrtPiOver2 = 1.25331413731550025121;
  if (u >= 0.0) {
    const double pdf = kInvSqrt2Pi * exp(-0.5 * u * u);
    const double cdf = 0.5 * erfc(-kInvSqrt2 * u);
    return log(pdf + u * cdf);
  }
  if (u != u) return u;
  // phi(u) + u Phi(u) = phi(u) (1 - t) for every u < 0 (botorch switches to this form at u = -1 only)
  const double head = -0.5 * (u * u + kLog2Pi);
  if (u > -1.0e6) return head + log1p(-(erfcx_pos(-kInvSqrt2 * u) * fabs(u) * kSqrtPiOver2));
  return head - 2.0 * log(fabs(u));
}

STP_HD double log_ei(double mean, double var, double best_f) {
  const double sigma = sqrt(var > 1e-12 ? var : (var != var ? var : 1e-12));
  const double u = (mean - best_f) / sigma;
  return log_h(u) + log(sigma);
}

// ---------------------------------------------------------------------------------------
// The other acquisitions over the same fused posterior (north_star names UCB and EI):
//   UCB  botorch UpperConfidenceBound:  mean + sqrt(beta) sigma
//   EI   botorch ExpectedImprovement / utils.EI (utils.py:299-309):  sigma (phi(u) + u Phi(u)), u = (mean - best_f) / sigma
// with sigma = sqrt(max(var, 1e-12)) as in botorch's _mean_and_sigma.
// ---------------------------------------------------------------------------------------
enum { kAcqLogEI = 0, kAcqUCB = 1, kAcqEI = 2 };

STP_HD double ei_h(double u) {      // phi(u) + u Phi(u)
  const double kInvSqrt2 = 0.70710678118654752440;
  const double kInvSqrt2Pi = 0.39894228040143267794;
  return kInvSqrt2Pi * exp(-0.5 * u * u) + u * (0.5 * erfc(-kInvSqrt2 * u));
}

STP_HD double acq_value(int kind, double beta, double mean, double var, double best_f) {
  if (kind == kAcqLogEI) return log_ei(mean, var, best_f);
  const double sigma = sqrt(var > 1e-12 ? var : (var != var ? var : 1e-12));
  if (kind == kAcqUCB) return mean + sqrt(beta) * sigma;
  return sigma * ei_h((mean - best_f) / sigma);
}

// One Monte-Carlo sample f of the latent function under a likelihood of constant scale sigma_l (VarTGP / VarLGP).
// (LogEI: log_h only -- the caller adds log sigma_l once after averaging, as the reference's mean over samples does.)
STP_HD double acq_sample(int kind, double sqrt_beta, double f, double best_f, double sigma_l) {
  if (kind == kAcqLogEI) return log_h_mc((f - best_f) / sigma_l);
  if (kind == kAcqUCB) return f + sqrt_beta * sigma_l;
  return sigma_l * ei_h((f - best_f) / sigma_l);
}

// ---------------------------------------------------------------------------------------
// Student-t LogEI of acqf.log_expected_improvement_t (acqf.py:7-52), which the reference evaluates with
// scipy.stats.t on the host:  Z = (mean - best_f) / std;  log(mean - best_f) + logcdf_t(Z) where the improvement is
// positive, logpdf_t(Z) + log(std) elsewhere; -inf where std <= 1e-9.
// t CDF through the regularised incomplete beta function I_x(nu/2, 1/2), x = nu / (nu + Z^2) (continued fraction,
// modified Lentz; the symmetric form is used where it converges faster).
// ---------------------------------------------------------------------------------------
STP_HD double betacf(double a, double b, double x) {
  const double tiny = 1e-300;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, dd = 1.0 - qab * x / qap;
  if (fabs(dd) < tiny) dd = tiny;
  dd = 1.0 / dd;
  double h = dd;
  for (int m = 1; m <= 500; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    dd = 1.0 + aa * dd; if (fabs(dd) < tiny) dd = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    dd = 1.0 / dd;
    h *= dd * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    dd = 1.0 + aa * dd; if (fabs(dd) < tiny) dd = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    dd = 1.0 / dd;
    const double del = dd * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-16) break;
  }
  return h;
}

// log of the regularised incomplete beta function I_x(a, b), 0 < x < 1
STP_HD double log_betainc(double a, double b, double x) {
  const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
  if (x < (a + 1.0) / (a + b + 2.0))
    return a * log(x) + b * log1p(-x) - lbeta + log(betacf(a, b, x)) - log(a);
  const double other = exp(b * log1p(-x) + a * log(x) - lbeta + log(betacf(b, a, 1.0 - x)) - log(b));
  return log1p(-other);
}

STP_HD double student_t_logpdf(double z, double nu) {
  return lgamma(0.5 * (nu + 1.0)) - lgamma(0.5 * nu) - 0.5 * log(nu * 3.14159265358979323846) -
         0.5 * (nu + 1.0) * log1p(z * z / nu);
}

STP_HD double student_t_logcdf(double z, double nu) {
  if (z == 0.0) return -0.69314718055994530942;
  const double x = nu / (nu + z * z);
  const double lhalf = log_betainc(0.5 * nu, 0.5, x) - 0.69314718055994530942;   // log of the smaller tail
  return (z < 0.0) ? lhalf : log1p(-exp(lhalf));
}

STP_HD double log_ei_t(double mean, double sd, double nu, double best_f) {
  if (!(sd > 1e-9)) return (sd != sd) ? sd : -INFINITY;
  const double gain = mean - best_f, z = gain / sd;
  return (gain > 0.0) ? log(gain) + student_t_logcdf(z, nu) : student_t_logpdf(z, nu) + log(sd);
}

// LogEI when sigma (and log sigma) are fixed across the call, as for the MC models.
STP_HD double log_ei_fixed_sigma(double f, double best_f, double inv_sigma, double log_sigma) {
  return log_h((f - best_f) * inv_sigma) + log_sigma;
}

}  // namespace stp
