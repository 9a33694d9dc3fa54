/* stp_b200.h -- C ABI of libstp_b200.so, the sm_100a implementation of the pool-based BO hot path
 * (fit -> predict -> acquire, batched over independent campaigns) of standw7/STPBenchmark.
 *
 * The reference is pure Python and has no FFI of its own; the seam this library sits behind is
 * the arithmetic the reference reaches through gpytorch / botorch / scipy.  Each entry point
 * cites the reference call site whose arithmetic it replaces (paths relative to the reference
 * checkout).  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory (e.g. torch.Tensor.data_ptr()),
 *     except `opts`, `prior`, `lower`, `upper`, which are small HOST arrays read at call time;
 *   - tensors are contiguous, row-major, batch-major; fp64 values, int32 indices, uint8 masks;
 *   - B = number of campaigns in the batch; one CTA owns one campaign;
 *   - `n` (int32[B], may be NULL = all n_cap) is the number of training points per campaign;
 *     X is [B, n_cap, d], y is [B, n_cap]; rows >= n[b] are ignored;
 *   - theta is [B, d+3] = (noise, constant, outputscale, lengthscale_1..d): the order botorch's
 *     fit hands the parameters of models.ExactGP to scipy (models.py:319-358);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     asynchronous and stream-ordered; nothing allocates, nothing synchronises;
 *   - return value: 0 = launched, negative = rejected (see stp_last_error_string());
 *   - status[b] (int32): 0 ok, 1 matrix not positive definite, 2 non-finite parameter or
 *     objective, 3 fit stopped on a non-finite objective, 4 iteration / evaluation cap reached,
 *     5 abnormal line-search termination, 6 campaign finished (turnover bank exhausted).  No exception crosses
 *     the boundary.
 *   - limits: 1 <= d <= STP_MAX_D, n_cap <= STP_MAX_N (stp_exact_*) or STP_MAX_N_VAR (stp_var_*).
 */
#ifndef STP_B200_H
#define STP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STP_MAX_D 8
#define STP_MAX_N 144   /* exact-GP entry points: (n_cap + 1)^2 doubles + the pool tile must fit in 227 KB of shared memory;
                           stp_exact_fit / stp_exact_bo_step also hold the optimiser state there: n_cap <= 142, beyond
                           that they return an error ("campaign does not fit in shared memory") */
#define STP_MAX_N_VAR 512   /* variational entry points: above 144 the working square moves to the global workspace; the
                               vectors and the sweep panels / pool tile still have to fit in 227 KB of shared memory
                               (n_cap = 480 fits for d <= 5, n_cap = 272 for d = 8), otherwise the call returns an error */
#define STP_MAX_N_LARGE 512 /* stp_exact_*_large: the working square lives in a caller-owned global workspace */
#define STP_ABI_VERSION 3

/* L-BFGS-B options; scipy.optimize.minimize(method="L-BFGS-B") defaults are
 * {10, 15000, 15000, 20, ftol / DBL_EPSILON with ftol = 2.220446049250313e-09, 1e-5}. */
typedef struct stp_lbfgsb_opts {
  int32_t m;
  int32_t maxiter;
  int32_t maxfun;
  int32_t maxls;
  double factr;
  double pgtol;
} stp_lbfgsb_opts;

/* Continuous batching for stp_exact_bo_step (all device pointers): a campaign whose training set reaches n_final
 * rows has finished; its trace row is appended to `sink` (row index from atomicAdd(sink_count, 1), dropped beyond
 * sink_capacity; sink may be NULL) and its slot restarts from the next initial design of its bank:
 * bank is int32[B, generations, n_init] pool indices, gen int32[B] the next generation to use (in/out).  When the
 * bank is exhausted the campaign stops with status 6.  n_final = 0 (or a NULL pointer) disables the feature. */
typedef struct stp_turnover {
  int32_t n_final;
  int32_t n_init;
  int32_t generations;
  int32_t sink_capacity;
  const int32_t* bank;
  int32_t* gen;
  int32_t* sink;
  int32_t* sink_count;
} stp_turnover;

/* Model / acquisition variants beyond what the reference instantiates (north_star names a
 * Matern kernel and UCB / EI acquisitions; the reference itself has RBF only (models.py:12) and LogEI in its loop).
 * Every entry point that takes a `const stp_model_opts* model` accepts NULL = {RBF, LogEI}: the reference's path.
 *   kernel       STP_KERNEL_RBF | STP_KERNEL_MATERN52 (gpytorch MaternKernel(nu=2.5), ARD): exact-GP entry points only
 *   acquisition  STP_ACQ_LOGEI (botorch LogExpectedImprovement, runners.py:88-90) | STP_ACQ_UCB (botorch
 *                UpperConfidenceBound: mean + sqrt(beta) sigma) | STP_ACQ_EI (botorch ExpectedImprovement = utils.EI,
 *                utils.py:299-309); sigma = sqrt(max(var, 1e-12)); for the Monte-Carlo models the mean over the samples. */
enum { STP_KERNEL_RBF = 0, STP_KERNEL_MATERN52 = 1 };
enum { STP_ACQ_LOGEI = 0, STP_ACQ_UCB = 1, STP_ACQ_EI = 2 };
typedef struct stp_model_opts {
  int32_t kernel;
  int32_t acquisition;
  double beta;
} stp_model_opts;

/* Likelihood of a variational model (models.py: VarGP :166, VarTGP :76, VarLGP :262). */
enum { STP_LIK_GAUSSIAN = 0, STP_LIK_STUDENT_T = 1, STP_LIK_LAPLACE = 2 };

int stp_version(void);
const char* stp_last_error_string(void);

/* Negative exact marginal log-likelihood (+ log-priors) / n and its gradient for B campaigns.
 * Replaces one closure call of botorch.fit.fit_gpytorch_mll -> ExactMarginalLogLikelihood
 * (optim.py:31-32) on models.ExactGP (models.py:300-395).  prior = {noise_loc, noise_scale,
 * os_loc, os_scale, ls_loc, ls_scale} of the LogNormal priors.  f: [B], g: [B, d+3]. */
int stp_exact_mll_fg(int B, int n_cap, int d, const double* X, const double* y, const int32_t* n,
                     const double* theta, const double* prior, double* f, double* g, int32_t* status,
                     const stp_model_opts* model, void* stream);

/* Device-resident L-BFGS-B fit of theta (in/out) for B campaigns: train_exact_model_botorch
 * (optim.py:11-32).  lower/upper: host arrays [d+3], +-HUGE_VAL = unbounded.
 * f_out [B], nit [B], nfev [B] may be NULL. */
int stp_exact_fit(int B, int n_cap, int d, const double* X, const double* y, const int32_t* n, double* theta,
                  const double* lower, const double* upper, const double* prior, const stp_lbfgsb_opts* opts,
                  double* f_out, int32_t* nit, int32_t* nfev, int32_t* status, const stp_model_opts* model, void* stream);

/* Fused predictive + LogEI + argmax over the pool at fixed theta: ExactGP.posterior
 * (models.py:387-395) + botorch LogExpectedImprovement.forward + argmax (runners.py:88-95).
 * pool: [G, N, d]; pool_id: int32[B] (NULL = pool 0); mask: uint8[B, N], non-zero = candidate
 * still available (NULL = all); best_f: [B] (NULL = predictive only).  out_idx [B] receives the
 * POOL index of the first maximal LogEI among available candidates (-1 if none / failure),
 * out_acq [B] its value.  mean / var / acq: optional [B, N] outputs for every pool point. */
int stp_exact_acq(int B, int n_cap, int d, int N, const double* pool, const int32_t* pool_id, const uint8_t* mask,
                  const double* X, const double* y, const int32_t* n, const double* theta, const double* best_f,
                  int32_t* out_idx, double* out_acq, double* mean, double* var, double* acq, int32_t* status,
                  const stp_model_opts* model, void* stream);

/* The same three operations for training sets beyond STP_MAX_N (the 80/20 fit-and-predict
 * study benchmark_predictions.py:19-75 fits ExactGP on up to 480 points and predicts the held-out 20 % through
 * model.posterior, :59-67).  The (n_cap + 1)^2 working square of each campaign lives in `work`, a device buffer of
 * stp_exact_large_workspace_bytes(B, n_cap) bytes owned by the caller (no hidden allocation; contents are scratch,
 * nothing is carried between calls); vectors, sweep panels, pool tile and optimiser state stay in shared memory.
 * Arguments, results, status codes and arithmetic are those of stp_exact_mll_fg / stp_exact_fit / stp_exact_acq;
 * any 1 <= n_cap <= STP_MAX_N_LARGE is accepted (small capacities run too, just slower than the shared-memory form). */
size_t stp_exact_large_workspace_bytes(int B, int n_cap);
int stp_exact_mll_fg_large(int B, int n_cap, int d, const double* X, const double* y, const int32_t* n,
                           const double* theta, const double* prior, double* f, double* g, int32_t* status,
                           const stp_model_opts* model, void* work, void* stream);
int stp_exact_fit_large(int B, int n_cap, int d, const double* X, const double* y, const int32_t* n, double* theta,
                        const double* lower, const double* upper, const double* prior, const stp_lbfgsb_opts* opts,
                        double* f_out, int32_t* nit, int32_t* nfev, int32_t* status, const stp_model_opts* model,
                        void* work, void* stream);
int stp_exact_acq_large(int B, int n_cap, int d, int N, const double* pool, const int32_t* pool_id, const uint8_t* mask,
                        const double* X, const double* y, const int32_t* n, const double* theta, const double* best_f,
                        int32_t* out_idx, double* out_acq, double* mean, double* var, double* acq, int32_t* status,
                        const stp_model_opts* model, void* work, void* stream);

/* One whole BO iteration (the loop body runners.py:64-109) for B campaigns in ONE launch:
 * fresh theta0 -> L-BFGS-B fit -> factorise -> pool pass -> argmax -> append the pick to
 * (X, y, n, trace) and clear its mask byte.  pool_y: [G, N].  trace: int32[B, n_cap] receives
 * the picked pool index at position n[b] (before the increment).  theta0: host [d+3] start
 * (the prior modes, models.py:326,334,347,357); theta_out [B, d+3] fitted values.
 * Campaigns whose status[b] is non-zero on entry are skipped (frozen, like the reference's
 * try/except at runners.py:113-115); n[b] >= n_cap is rejected by status 2.
 * order: optional int32[B] permutation (NULL = identity): CTA i works on campaign order[i].  CTAs are issued in
 * index order, so passing the campaigns sorted by decreasing n (longest fits first) shortens the tail of the
 * launch when the batch mixes training-set sizes.
 * turnover: optional continuous batching (see stp_turnover).  work_acc: optional double[B], += the algorithmic
 * flops of the iteration, F (n^3 + n^2 (3d+6)) + n^3/3 + (N-n)(n^2 + n(3d+8) + 60) with F the closure count;
 * iters_acc: optional int32[B], += 1 per completed iteration (both let a caller account a run without any
 * per-step device work of its own).
 * n_max: an upper bound of n[b] over the active campaigns of THIS launch if the caller knows one (campaigns that
 * advance in lockstep: n_initial + steps done), else 0.  Shared memory and the thread count of the launch are sized
 * for n_max instead of n_cap - 1, which fits 3 - 4 campaigns per SM instead of 2 while the training sets are
 * small; a campaign with n[b] > n_max stops with status 2.  Results do not depend on n_max. */
int stp_exact_bo_step(int B, int n_cap, int d, int N, const double* pool, const double* pool_y,
                      const int32_t* pool_id, uint8_t* mask, double* X, double* y, int32_t* n, int32_t* trace,
                      const double* theta0, const double* lower, const double* upper, const double* prior,
                      const stp_lbfgsb_opts* opts, double* theta_out, double* acq_out, int32_t* nit, int32_t* nfev,
                      int32_t* status, const int32_t* order, const stp_turnover* turnover, double* work_acc,
                      int32_t* iters_acc, int n_max, const stp_model_opts* model, void* stream);

/* stp_exact_bo_step for campaigns beyond the shared-memory capacity (n_cap up to STP_MAX_N_LARGE: a BO loop with more
 * than ~130 trials, runners.py:64-109 with a large n_trials): identical arguments and semantics, the working square
 * of every campaign in `work` (stp_exact_large_workspace_bytes(B, n_cap) bytes, scratch). */
int stp_exact_bo_step_large(int B, int n_cap, int d, int N, const double* pool, const double* pool_y,
                            const int32_t* pool_id, uint8_t* mask, double* X, double* y, int32_t* n, int32_t* trace,
                            const double* theta0, const double* lower, const double* upper, const double* prior,
                            const stp_lbfgsb_opts* opts, double* theta_out, double* acq_out, int32_t* nit, int32_t* nfev,
                            int32_t* status, const int32_t* order, const stp_turnover* turnover, double* work_acc,
                            int32_t* iters_acc, int n_max, const stp_model_opts* model, void* work, void* stream);

/* The same loop body as a PERSISTENT launch over a work queue: one CTA per SM slot (as many as are resident at the
 * launch's shared-memory footprint, at most max_ctas if max_ctas > 0) serves the B campaigns from a ticket ring; a
 * work item is one BO iteration of one campaign, and campaign b is advanced until iters_left[b] (int32[B], in/out,
 * decremented per completed iteration) reaches 0 or its status becomes non-zero.  Campaigns advance independently
 * -- nobody waits for the slowest fit of a lock-step batch -- so one call replaces iters x stp_exact_bo_step
 * (runners.py:64-109 inside `for trial in range(n_trials)`, runners.py:64) with identical per-campaign results.
 * sched: int32[B + 1] scratch (busy flags + ticket counter), zero it once after allocation.
 * n_max: upper bound of n[b] + iters_left[b] over the campaigns (what a training set can grow to in this launch),
 * 0 = n_cap - 1; other arguments as for stp_exact_bo_step (order = ring order of the campaigns). */
int stp_exact_bo_run(int B, int n_cap, int d, int N, const double* pool, const double* pool_y,
                     const int32_t* pool_id, uint8_t* mask, double* X, double* y, int32_t* n, int32_t* trace,
                     const double* theta0, const double* lower, const double* upper, const double* prior,
                     const stp_lbfgsb_opts* opts, double* theta_out, double* acq_out, int32_t* nit, int32_t* nfev,
                     int32_t* status, const int32_t* order, const stp_turnover* turnover, double* work_acc,
                     int32_t* iters_acc, int n_max, int32_t* iters_left, int32_t* sched, int max_ctas,
                     const stp_model_opts* model, void* stream);

/* ---- variational models: VarGP / VarTGP / VarLGP (models.py:114-202, 18-111, 205-297) ---------------
 * State arrays (device, caller-owned, batch-major): Z [B, n_cap, d] inducing locations, hyp [B, d+4] =
 * (constant, lengthscale_1..d, outputscale, noise, raw_deg_free), nat1 [B, n_cap], nat2 [B, n_cap, n_cap]
 * natural parameters of the whitened q(u), adam_m / adam_v [B, n_cap*d + d + 4] Adam moments laid out as
 * (Z | hyp).  work: device scratch of stp_var_workspace_bytes(B, n_cap) bytes.
 * Host tables: hyp0 [d+4] construction values; prior [8] = (ls_loc, ls_scale, os_loc, os_scale, noise_loc,
 * noise_scale, df_concentration, df_rate); gh [40] = 20 Gauss-Hermite nodes then 20 weights. */
size_t stp_var_workspace_bytes(int B, int n_cap);
/* Smallest capacity n_cap >= n that EVERY variational entry point can launch with for input dimension d, or -1.
 * The shared-memory budget has a gap just below 144 (the working square, the vectors and the pool tile of stp_var_acq
 * together exceed 227 KB from n_cap = 138 at d = 5) that ends where the square moves to the global workspace (145):
 * a caller that needs n = 143 rows (benchmark_predictions.py's 80 % split of the 178 P3HT points) allocates 145. */
int stp_var_capacity(int n, int d);

/* train_natural_variational_model (optim.py:105-181): `epochs` full-batch steps in ONE persistent launch --
 * ELBO forward (20-point Gauss-Hermite for Student-T / Laplace), hand-derived backward, natural-gradient step
 * lr_ngd * n on (nat1, nat2) (gpytorch.optim.NGD, optim.py:132-134) and Adam(lr_adam; optim.py:136-141) on
 * (Z, hyp).  y is standardised inside (optim.py:126-127).  fresh != 0 first builds a new model
 * (runners.py:66-74): Z = X, hyp = hyp0, nat1 = 1e-3 * init_noise (init_noise [B, n_cap]; NULL = Philox(seed)),
 * nat2 = -I/2, Adam moments 0; fresh == 0 continues from the given state with Adam step offset t0.
 * loss_hist: optional [B, epochs].  Campaigns with status != 0 on entry are skipped.  order: optional int32[B]
 * CTA -> campaign permutation (longest first), as for stp_exact_bo_step. */
int stp_var_fit(int B, int n_cap, int d, int lik, int epochs, double lr_ngd, double lr_adam, int fresh, int t0,
                const double* X, const double* y, const int32_t* n, const double* init_noise, unsigned long long seed,
                const double* hyp0, const double* prior, const double* gh, double* Z, double* hyp, double* nat1,
                double* nat2, double* adam_m, double* adam_v, double* loss_hist, int32_t* status, double* work,
                const int32_t* order, void* stream);

/* One evaluation of -ELBO/n (VariationalELBO, optim.py:143, 153-154) and its raw gradients at a given state,
 * nothing updated.  grads [B, G], G = n_cap + n_cap^2 + n_cap*d + d + 4, laid out as
 * (natural gradient of nat1 | of nat2 | dZ | d hyp).  standardise != 0 standardises y first. */
int stp_var_elbo_fg(int B, int n_cap, int d, int lik, int standardise, const double* X, const double* y,
                    const int32_t* n, const double* Z, const double* hyp, const double* nat1, const double* nat2,
                    const double* prior, const double* gh, double* loss, double* grads, int32_t* status, double* work,
                    void* stream);

/* Fused pool pass of a trained variational model: Var*.posterior (models.py:95-111, 187-202, 280-297) +
 * LogExpectedImprovement + mean over the S likelihood samples + argmax (runners.py:88-95).  mean / var
 * (optional [B, N]) receive the latent q(f) moments; acq (optional [B, N]) the acquisition.  Gaussian: analytic
 * LogEI on N(m, v + noise).  Student-T / Laplace: mean over S samples f = m + sqrt(v) eps of LogEI with the
 * constant likelihood sigma; eps [B, N, S] if non-NULL, else Philox4x32-10(seed, stream = campaign,
 * counter = candidate * ceil(S/2) + pair) -- stp_philox_normal reproduces that stream.  best_f [B] or NULL
 * (= max of y[b, :n], raw).  update != 0 additionally performs runners.py:98-109 (append the pick to X, y, n,
 * trace; clear its mask byte). */
int stp_var_acq(int B, int n_cap, int d, int N, int lik, const double* pool, const int32_t* pool_id, uint8_t* mask,
                const double* Z, const double* hyp, const double* nat1, const double* nat2, int32_t* n,
                const double* best_f, int S, const double* eps, unsigned long long seed, int32_t* out_idx,
                double* out_acq, double* mean, double* var, double* acq, int update, const double* pool_y, double* X,
                double* y, int32_t* trace, int32_t* status, double* work, const int32_t* order,
                const stp_model_opts* model, void* stream);

/* 2 * pairs standard normals of the in-kernel generator: out[2q], out[2q+1] for counters counter0 + q. */
int stp_philox_normal(unsigned long long seed, unsigned long long stream_id, unsigned long long counter0,
                      long long pairs, double* out, void* stream);

/* Elementwise analytic LogEI (botorch LogExpectedImprovement.forward, runners.py:88-90) of
 * `count` (mean, var) pairs against one best_f: out[i] = log_h((mean - best_f)/sigma) + log sigma,
 * sigma = sqrt(max(var, 1e-12)). */
int stp_log_ei(long long count, const double* mean, const double* var, double best_f, double* out, void* stream);

/* The same elementwise pass for any acquisition of stp_model_opts (kernel ignored): UCB and EI next to LogEI. */
int stp_acq_values(long long count, const double* mean, const double* var, double best_f, const stp_model_opts* model,
                   double* out, void* stream);

/* Elementwise Student-t LogEI, the device form of acqf.log_expected_improvement_t (acqf.py:7-52, scipy.stats.t on the
 * host in the reference): Z = (mean - best_f) / sd; log(mean - best_f) + log T_df(Z) where the improvement is positive,
 * log t_df(Z) + log sd elsewhere, -inf where sd <= 1e-9.  mean, sd, df: [count]. */
int stp_log_ei_t(long long count, const double* mean, const double* sd, const double* df, double best_f, double* out,
                 void* stream);

/* Sustained FP64 FMA throughput probe (roofline denominator; MEASURED_PEAKS.json has no FP64
 * figure).  Launches 148 * 8 CTAs x 256 threads, each running iters x 16 x 8 independent DFMAs
 * (= 2 flop each; the caller times the launch with CUDA events and divides).  `sink` (device,
 * one double) only keeps the loop alive: nothing meaningful is written to it. */
int stp_fp64_probe(int iters, double* sink, void* stream);

/* The same probe on the fp64 tensor-core path (mma.sync.m8n8k4.f64, DMMA): per launch 148 * 8 CTAs x 8 warps x
 * iters x 16 x 8 instructions of 256 MAC.  The sweep, the pool-pass product and the variational contractions
 * issue DMMA; on B200 its peak equals the DFMA peak. */
int stp_fp64_mma_probe(int iters, double* sink, void* stream);

/* Initial design, the O(Q Ne d) step of get_initial_samples_sobol (utils.py:272-279): out_idx[q] = index of the first
 * eligible pool point nearest to query point q (Euclidean; the queries are the scrambled-Sobol points already
 * scaled to the bounding box of the eligible points, which stay a host computation because they are defined by
 * torch's CPU generator, utils.py:262-271).  points [Q, d], eligible [Ne, d], out_idx int32 [Q]. */
int stp_design_snap(int Q, int Ne, int d, const double* points, const double* eligible, int32_t* out_idx, void* stream);

/* Diagnostic: per-phase cycle counters of stp_exact_bo_step (clock64 of thread 0 of every CTA, summed over CTAs;
 * ids in csrc/team.cuh).  Only in libraries built with -DSTP_PHASE_TIMERS (tools/gpu_phase_timers.py builds one
 * beside the shipped library); the shipped build returns an error.  out32: HOST array of 64 counters (0-31 exact path, 32-63 variational epoch) (may be
 * NULL); reset != 0 zeroes them.  Synchronises the device. */
int stp_phase_timers(unsigned long long* out32, int reset);

#ifdef __cplusplus
}
#endif
#endif /* STP_B200_H */
