// libstp_b200.so -- sm_100a kernels and C ABI (include/stp_b200.h) of the batched BO hot path.
// One CTA owns one campaign; every matrix of an exact-GP campaign stays in shared memory for the whole
// launch (43 KB at n <= 48, 65 KB at n <= 66, 106 KB at n <= 89: four, three or two campaigns resident per SM).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/stp_b200.h"
#include "exact_campaign.cuh"
#include "var_campaign.cuh"

using namespace stp;

namespace {

thread_local char g_err[256] = "";
int fail(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}
int fail_cuda(const char* where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return -2;
}

constexpr int kTC = 32;  // acquisition tile: one candidate per lane

// extra dynamic shared memory per CTA (bytes; diagnostic: lowers the number of resident campaigns per SM)
size_t smem_pad() {
  const char* env = getenv("STP_SMEM_PAD");
  return env ? (size_t)atol(env) : 0;
}

// launch bounds of the exact-GP kernels (overridable for A/B runs, tools/gpu_ab_exact.sh)
#ifndef STP_EXACT_MINB
#define STP_EXACT_MINB 2
#endif
#ifndef STP_EXACT_MAXT
#define STP_EXACT_MAXT 256
#endif

int threads_for(int n_cap) {
  const char* env = getenv("STP_THREADS");
  if (env) {
    const int t = atoi(env);
    if (t >= 32 && t <= STP_EXACT_MAXT && t % 32 == 0) return t;
  }
  // 128 registers per thread (launch bounds): 128 threads -> 4 CTAs/SM, 160 -> 3, 256 -> 2; the shared-memory
  // footprint of the carve-up follows the same steps (48 KB at cap 48, 74 KB at cap 66, 113 KB at cap 90)
  return n_cap <= 48 ? 128 : (n_cap <= 66 ? 160 : 256);
}

size_t exact_smem_bytes(int n_cap, int d, int threads, bool with_optimizer) {
  size_t b = exact_smem_doubles(n_cap, d, kTC, threads / 32) * sizeof(double);
  if (with_optimizer) b += sizeof(LbfgsbState) + 16;
  return b;
}

// Large capacities (n_cap > STP_MAX_N; row f3): square in the global workspace, kExactLargeThreads threads, one CTA per SM.
constexpr int kExactLargeThreads = 512;
int threads_large() {
  const char* env = getenv("STP_THREADS_LARGE");
  if (env) {
    const int t = atoi(env);
    if (t >= 32 && t <= kExactLargeThreads && t % 32 == 0) return t;
  }
  return kExactLargeThreads;
}
size_t exact_large_smem_bytes(int n_cap, int d, bool with_optimizer) {
  size_t b = exact_large_smem_doubles(n_cap, d, kTC) * sizeof(double);
  if (with_optimizer) b += sizeof(LbfgsbState) + 16;
  return b;
}

struct PriorArg { double v[6]; };
struct ModelArg { int kernel, acquisition; double beta; };

int fill_model(ModelArg& m, const stp_model_opts* o, bool kernel_allowed) {
  m.kernel = 0; m.acquisition = 0; m.beta = 0.0;
  if (!o) return 0;
  if (o->kernel != STP_KERNEL_RBF && (o->kernel != STP_KERNEL_MATERN52 || !kernel_allowed))
    return fail(kernel_allowed ? "unknown kernel" : "the variational entry points provide the RBF kernel only");
  if (o->acquisition < STP_ACQ_LOGEI || o->acquisition > STP_ACQ_EI) return fail("unknown acquisition");
  if (o->acquisition == STP_ACQ_UCB && !(o->beta >= 0.0)) return fail("UCB needs beta >= 0");
  m.kernel = o->kernel; m.acquisition = o->acquisition; m.beta = o->beta;
  return 0;
}
__device__ __forceinline__ ExactSmem with_model(ExactSmem s, const ModelArg& m) {
  s.kern = m.kernel; s.acq_kind = m.acquisition; s.acq_beta = m.beta;
  return s;
}
struct BoundsArg { double lb[kMaxP], ub[kMaxP], theta0[kMaxP]; };

__device__ __forceinline__ ExactPrior to_prior(const PriorArg& p) {
  return ExactPrior{p.v[0], p.v[1], p.v[2], p.v[3], p.v[4], p.v[5]};
}

__device__ __forceinline__ int fit_status_code(int s) {  // optimiser status -> ABI status
  return s == 0 ? 0 : (s == 1 ? 4 : (s == 2 ? 5 : 3));
}

// ------------------------------------------------------------------------------------------
// LARGE (all three model-level kernels): the working square of campaign b is work + b * exact_large_work_doubles(n_cap).
template <bool LARGE>
__device__ __forceinline__ ExactSmem exact_carve_for(double* smem, double* work, int b, int n_cap, int d, int nwarps) {
  if (LARGE) return exact_smem_carve_large(smem, work + (size_t)b * exact_large_work_doubles(n_cap), n_cap, d, kTC);
  return exact_smem_carve(smem, n_cap, d, kTC, nwarps);
}

template <bool LARGE>
__device__ __forceinline__ void exact_mll_fg_body(double* smem, int n_cap, int d, const double* __restrict__ X,
                                                  const double* __restrict__ y, const int32_t* __restrict__ n_arr,
                                                  const double* __restrict__ theta, const PriorArg& prior,
                                                  double* __restrict__ f, double* __restrict__ g,
                                                  int32_t* __restrict__ status, const ModelArg& model,
                                                  double* __restrict__ work) {
  const int b = blockIdx.x;
  const int n = n_arr ? n_arr[b] : n_cap;
  const ExactSmem s = with_model(exact_carve_for<LARGE>(smem, work, b, n_cap, d, blockDim.x >> 5), model);
  CtaTeam tm{s.red};
  const int np = d + 3;
  exact_load(tm, s, n, d, X + (size_t)b * n_cap * d, y + (size_t)b * n_cap);
  if (threadIdx.x < np) s.small[threadIdx.x] = theta[b * np + threadIdx.x];
  __syncthreads();
  double fv;
  const int st = exact_mll_fg(tm, s, n, d, s.small, to_prior(prior), &fv, s.small + 16);
  if (threadIdx.x == 0) { f[b] = fv; status[b] = st; }
  if (threadIdx.x < np) g[b * np + threadIdx.x] = st ? NAN : s.small[16 + threadIdx.x];
}
__global__ void k_exact_mll_fg(int n_cap, int d, const double* __restrict__ X, const double* __restrict__ y,
                               const int32_t* __restrict__ n_arr, const double* __restrict__ theta, PriorArg prior,
                               double* __restrict__ f, double* __restrict__ g, int32_t* __restrict__ status, ModelArg model) {
  extern __shared__ double smem[];
  exact_mll_fg_body<false>(smem, n_cap, d, X, y, n_arr, theta, prior, f, g, status, model, nullptr);
}
__global__ void __launch_bounds__(kExactLargeThreads, 1)
k_exact_mll_fg_large(int n_cap, int d, const double* __restrict__ X, const double* __restrict__ y,
                     const int32_t* __restrict__ n_arr, const double* __restrict__ theta, PriorArg prior,
                     double* __restrict__ f, double* __restrict__ g, int32_t* __restrict__ status, ModelArg model,
                     double* __restrict__ work) {
  extern __shared__ double smem[];
  exact_mll_fg_body<true>(smem, n_cap, d, X, y, n_arr, theta, prior, f, g, status, model, work);
}

template <bool LARGE>
__global__ void __launch_bounds__(LARGE ? kExactLargeThreads : STP_EXACT_MAXT, LARGE ? 1 : STP_EXACT_MINB)
k_exact_fit(int n_cap, int d, const double* __restrict__ X, const double* __restrict__ y,
            const int32_t* __restrict__ n_arr, double* __restrict__ theta, BoundsArg bnd,
            PriorArg prior, LbfgsbOpts opts, double* __restrict__ f_out, int32_t* __restrict__ nit,
            int32_t* __restrict__ nfev, int32_t* __restrict__ status, ModelArg model, double* __restrict__ work) {
  extern __shared__ double smem[];
  const int b = blockIdx.x;
  const int n = n_arr ? n_arr[b] : n_cap;
  const int nw = blockDim.x >> 5;
  const ExactSmem s = with_model(exact_carve_for<LARGE>(smem, work, b, n_cap, d, nw), model);
  LbfgsbState* S = reinterpret_cast<LbfgsbState*>(
      smem + (LARGE ? exact_large_smem_doubles(n_cap, d, kTC) : exact_smem_doubles(n_cap, d, kTC, nw)));
  int* flag = reinterpret_cast<int*>(reinterpret_cast<char*>(S) + sizeof(LbfgsbState));
  CtaTeam tm{s.red};
  const int np = d + 3;
  exact_load(tm, s, n, d, X + (size_t)b * n_cap * d, y + (size_t)b * n_cap);
  double* th = s.small;
  if (threadIdx.x < np) th[threadIdx.x] = theta[b * np + threadIdx.x];
  __syncthreads();
  const ExactFitResult r = exact_fit_run(tm, s, S, flag, n, d, th, bnd.lb, bnd.ub, to_prior(prior), opts);
  if (threadIdx.x < np) theta[b * np + threadIdx.x] = th[threadIdx.x];
  if (threadIdx.x == 0) {
    if (f_out) f_out[b] = r.f;
    if (nit) nit[b] = r.nit;
    if (nfev) nfev[b] = r.nfev;
    status[b] = fit_status_code(r.status);
  }
}

template <bool LARGE>
__device__ __forceinline__ void exact_acq_body(double* smem, int n_cap, int d, int N, const double* __restrict__ pool,
            const int32_t* __restrict__ pool_id, const uint8_t* __restrict__ mask,
            const double* __restrict__ X, const double* __restrict__ y,
            const int32_t* __restrict__ n_arr, const double* __restrict__ theta,
            const double* __restrict__ best_f, int32_t* __restrict__ out_idx,
            double* __restrict__ out_acq, double* __restrict__ mean, double* __restrict__ var,
            double* __restrict__ acq, int32_t* __restrict__ status, const ModelArg& model, double* __restrict__ work) {
  const int b = blockIdx.x;
  const int n = n_arr ? n_arr[b] : n_cap;
  const ExactSmem s = with_model(exact_carve_for<LARGE>(smem, work, b, n_cap, d, blockDim.x >> 5), model);
  CtaTeam tm{s.red};
  const int np = d + 3;
  exact_load(tm, s, n, d, X + (size_t)b * n_cap * d, y + (size_t)b * n_cap);
  if (threadIdx.x < np) s.small[threadIdx.x] = theta[b * np + threadIdx.x];
  __syncthreads();
  const int st = exact_factorize(tm, s, n, d, s.small);
  if (st) {
    if (threadIdx.x == 0) { status[b] = 1; if (out_idx) out_idx[b] = -1; if (out_acq) out_acq[b] = NAN; }
    return;
  }
  int bi; double bv;
  const double* P = pool + (size_t)(pool_id ? pool_id[b] : 0) * N * d;
  exact_acquire(tm, s, n, d, s.small, P, N, mask ? mask + (size_t)b * N : nullptr, best_f ? best_f[b] : NAN, &bi,
                &bv, mean ? mean + (size_t)b * N : nullptr, var ? var + (size_t)b * N : nullptr,
                acq ? acq + (size_t)b * N : nullptr);
  if (threadIdx.x == 0) { status[b] = 0; if (out_idx) out_idx[b] = bi; if (out_acq) out_acq[b] = bv; }
}
__global__ void k_exact_acq(int n_cap, int d, int N, const double* __restrict__ pool,
                            const int32_t* __restrict__ pool_id, const uint8_t* __restrict__ mask,
                            const double* __restrict__ X, const double* __restrict__ y,
                            const int32_t* __restrict__ n_arr, const double* __restrict__ theta,
                            const double* __restrict__ best_f, int32_t* __restrict__ out_idx,
                            double* __restrict__ out_acq, double* __restrict__ mean, double* __restrict__ var,
                            double* __restrict__ acq, int32_t* __restrict__ status, ModelArg model) {
  extern __shared__ double smem[];
  exact_acq_body<false>(smem, n_cap, d, N, pool, pool_id, mask, X, y, n_arr, theta, best_f, out_idx, out_acq, mean, var,
                        acq, status, model, nullptr);
}
__global__ void __launch_bounds__(kExactLargeThreads, 1)
k_exact_acq_large(int n_cap, int d, int N, const double* __restrict__ pool,
                  const int32_t* __restrict__ pool_id, const uint8_t* __restrict__ mask,
                  const double* __restrict__ X, const double* __restrict__ y,
                  const int32_t* __restrict__ n_arr, const double* __restrict__ theta,
                  const double* __restrict__ best_f, int32_t* __restrict__ out_idx,
                  double* __restrict__ out_acq, double* __restrict__ mean, double* __restrict__ var,
                  double* __restrict__ acq, int32_t* __restrict__ status, ModelArg model, double* __restrict__ work) {
  extern __shared__ double smem[];
  exact_acq_body<true>(smem, n_cap, d, N, pool, pool_id, mask, X, y, n_arr, theta, best_f, out_idx, out_acq, mean, var,
                       acq, status, model, work);
}

// Arguments of one BO iteration (the whole loop body runners.py:64-109) shared by the two exact-GP campaign kernels.
struct BoArgs {
  int n_cap, d, N, s_cap;
  const double* pool; const double* pool_y; const int32_t* pool_id;
  uint8_t* mask; double* X; double* y; int32_t* n_arr; int32_t* trace;
  BoundsArg bnd; PriorArg prior; LbfgsbOpts opts;
  double* theta_out; double* acq_out; int32_t* nit; int32_t* nfev; int32_t* status;
  stp_turnover turn; double* work_acc; int32_t* iters_acc;
  ModelArg model;
  double* work;   // LARGE only: the working squares (stp_exact_large_workspace_bytes)
};

// One BO iteration of campaign b by the whole CTA.  CG: the campaign state may have been written by another SM
// during this launch (persistent kernel), so it is read through the L2 (ld.global.cg), never through this SM's L1.
// Returns 0 when the iteration completed, non-zero when the campaign stopped (status[b] set).
template <bool CG, bool LARGE = false>
__device__ __forceinline__ int exact_bo_iteration(const BoArgs& a, double* smem, int b) {
  const int n_cap = a.n_cap, d = a.d, N = a.N, s_cap = a.s_cap;
  const int n = CG ? __ldcg(a.n_arr + b) : a.n_arr[b];
  // n_cap: row capacity of the global arrays; s_cap <= n_cap - 1: what the shared-memory carve-up of this launch holds
  if (n >= n_cap || n < 1 || n > s_cap) { if (threadIdx.x == 0) a.status[b] = 2; return 2; }
  const int nw = blockDim.x >> 5;
  const ExactSmem s = with_model(exact_carve_for<LARGE>(smem, a.work, b, s_cap, d, nw), a.model);
  LbfgsbState* S = reinterpret_cast<LbfgsbState*>(
      smem + (LARGE ? exact_large_smem_doubles(s_cap, d, kTC) : exact_smem_doubles(s_cap, d, kTC, nw)));
  int* flag = reinterpret_cast<int*>(reinterpret_cast<char*>(S) + sizeof(LbfgsbState));
  CtaTeam tm{s.red};
  const int np = d + 3;
  double* Xb = a.X + (size_t)b * n_cap * d;
  double* yb = a.y + (size_t)b * n_cap;
  STP_PHASE(-1); STP_COUNT(18);
  __syncthreads();   // the previous work item of this CTA is done with the shared memory
  if (CG) {
    for (int idx = threadIdx.x; idx < n * d; idx += blockDim.x) {
      const int i = idx / d, k = idx - i * d;
      s.Xt[k * s.cap + i] = __ldcg(Xb + idx);
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) s.y[i] = __ldcg(yb + i);
    __syncthreads();
  } else {
    exact_load(tm, s, n, d, Xb, yb);
  }
  double* th = s.small;
  if (threadIdx.x < np) th[threadIdx.x] = a.bnd.theta0[threadIdx.x];  // no warm start (App. D-7)
  __syncthreads();
  STP_PHASE(0);
  const ExactFitResult r = exact_fit_run(tm, s, S, flag, n, d, th, a.bnd.lb, a.bnd.ub, to_prior(a.prior), a.opts);
  if (threadIdx.x < np && a.theta_out) a.theta_out[b * np + threadIdx.x] = th[threadIdx.x];
  if (threadIdx.x == 0) { if (a.nit) a.nit[b] = r.nit; if (a.nfev) a.nfev[b] = r.nfev; }
  if (r.status == 3) { if (threadIdx.x == 0) a.status[b] = 3; return 3; }
  STP_PHASE(-1);
  if (exact_factorize(tm, s, n, d, th)) { if (threadIdx.x == 0) a.status[b] = 1; return 1; }
  STP_PHASE(8);
  double best = -INFINITY;                      // best_f = y_train.max(), raw (runners.py:89)
  STP_ROLLED for (int i = 0; i < n; ++i) best = fmax(best, s.y[i]);
  int bi; double bv;
  const int pid = a.pool_id ? a.pool_id[b] : 0;
  const double* P = a.pool + (size_t)pid * N * d;
  uint8_t* mb = a.mask + (size_t)b * N;
  exact_acquire(tm, s, n, d, th, P, N, mb, best, &bi, &bv, nullptr, nullptr, nullptr);
  STP_PHASE(9);
  if (bi < 0) { if (threadIdx.x == 0) a.status[b] = 2; return 2; }
  // runners.py:98-109: record the pick, append it to the training set, drop it from the pool
  if (threadIdx.x < d) Xb[(size_t)n * d + threadIdx.x] = P[(size_t)bi * d + threadIdx.x];
  if (threadIdx.x == 0) {
    yb[n] = a.pool_y[(size_t)pid * N + bi];
    a.trace[(size_t)b * n_cap + n] = bi;
    mb[bi] = 0;
    a.n_arr[b] = n + 1;
    if (a.acq_out) a.acq_out[b] = bv;
    if (a.iters_acc) a.iters_acc[b] = (CG ? __ldcg(a.iters_acc + b) : a.iters_acc[b]) + 1;
    if (a.work_acc) {   // algorithmic flops of this campaign-iteration, actual closure count
      const double nn = n, F = r.nfev;
      a.work_acc[b] = (CG ? __ldcg(a.work_acc + b) : a.work_acc[b]) +
                      F * (nn * nn * nn + nn * nn * (3 * d + 6)) + nn * nn * nn / 3.0 +
                      (N - nn) * (nn * nn + nn * (3 * d + 8) + 60.0);
    }
  }
  // Continuous batching: a campaign that has done its last trial hands its trace to the sink and its slot to the
  // next seed of its bank (fresh mask / trace / training set); nothing to do for the host between steps.
  const stp_turnover& turn = a.turn;
  if (turn.n_final > 0 && n + 1 >= turn.n_final) {
    __shared__ int s_slot, s_gen;
    __syncthreads();                               // the pick has been recorded in the trace row
    if (threadIdx.x == 0) {
      s_slot = turn.sink ? atomicAdd(turn.sink_count, 1) : -1;
      s_gen = CG ? __ldcg(turn.gen + b) : turn.gen[b];
    }
    __syncthreads();
    int32_t* tb = a.trace + (size_t)b * n_cap;
    if (s_slot >= 0 && s_slot < turn.sink_capacity)
      for (int i = threadIdx.x; i < n_cap; i += blockDim.x)
        turn.sink[(size_t)s_slot * n_cap + i] = CG ? __ldcg(tb + i) : tb[i];
    __syncthreads();
    if (s_gen >= turn.generations) { if (threadIdx.x == 0) a.status[b] = 6; return 6; }   // bank exhausted: finished
    const int32_t* init = turn.bank + ((size_t)b * turn.generations + s_gen) * turn.n_init;
    for (int i = threadIdx.x; i < N; i += blockDim.x) mb[i] = 1;
    for (int i = threadIdx.x; i < n_cap; i += blockDim.x) tb[i] = (i < turn.n_init) ? init[i] : -1;
    for (int i = threadIdx.x; i < n_cap * d; i += blockDim.x) {
      const int row = i / d, k = i - row * d;
      Xb[i] = (row < turn.n_init) ? P[(size_t)init[row] * d + k] : 0.0;
    }
    for (int i = threadIdx.x; i < n_cap; i += blockDim.x)
      yb[i] = (i < turn.n_init) ? a.pool_y[(size_t)pid * N + init[i]] : 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i < turn.n_init; i += blockDim.x) mb[init[i]] = 0;
    if (threadIdx.x == 0) { a.n_arr[b] = turn.n_init; turn.gen[b] = s_gen + 1; }
  }
  return 0;
}

// One BO iteration per campaign, one CTA per campaign (lock-step batch).
__global__ void __launch_bounds__(STP_EXACT_MAXT, STP_EXACT_MINB) k_exact_bo_step(BoArgs a, const int32_t* __restrict__ order) {
  extern __shared__ double smem[];
  const int b = order ? order[blockIdx.x] : blockIdx.x;   // longest-first scheduling: CTAs are issued in blockIdx order
  if (a.status[b] != 0) return;  // frozen campaign
  exact_bo_iteration<false>(a, smem, b);
}

// The same for campaigns beyond the shared-memory capacity (n_cap > STP_MAX_N): working squares in a.work.
__global__ void __launch_bounds__(kExactLargeThreads, 1) k_exact_bo_step_large(BoArgs a, const int32_t* __restrict__ order) {
  extern __shared__ double smem[];
  const int b = order ? order[blockIdx.x] : blockIdx.x;
  if (a.status[b] != 0) return;  // frozen campaign
  exact_bo_iteration<false, true>(a, smem, b);
}

// Persistent form: gridDim.x resident CTAs serve B campaigns from a ticket ring until every campaign has done its
// `iters_left[b]` iterations (or stopped).  A work item is ONE BO iteration of one campaign: campaigns advance
// independently, so no CTA ever waits for the slowest fit of a lock-step batch (the tail that limits
// k_exact_bo_step) -- the SM slots stay full until the whole job is done.  busy[b] serialises the iterations of a
// campaign; the release/acquire pair (threadfence + atomic) orders its state between SMs.
__global__ void __launch_bounds__(STP_EXACT_MAXT, STP_EXACT_MINB) k_exact_bo_run(BoArgs a, int B, const int32_t* __restrict__ order,
                                                                               int32_t* iters_left, int32_t* busy,
                                                                               unsigned int* ticket) {
  extern __shared__ double smem[];
  __shared__ int s_b;
  for (;;) {
    if (threadIdx.x == 0) {
      int found = -1;
      for (int tries = 0; tries < B; ++tries) {
        const unsigned int t = atomicAdd(ticket, 1u);
        const int slot = (int)(t % (unsigned int)B);
        const int b = order ? order[slot] : slot;
        if (__ldcg(a.status + b) != 0 || __ldcg(iters_left + b) <= 0) continue;
        if (atomicCAS(busy + b, 0, 1) != 0) continue;                 // another CTA is advancing this campaign
        __threadfence();                                               // acquire: see that CTA's writes
        if (__ldcg(a.status + b) != 0 || __ldcg(iters_left + b) <= 0) { atomicExch(busy + b, 0); continue; }
        found = b;
        break;
      }
      s_b = found;
    }
    __syncthreads();
    const int b = s_b;
    if (b < 0) return;              // a whole ring of tickets without work: the job is done (or in other hands)
    exact_bo_iteration<true>(a, smem, b);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();              // release: the campaign state is visible before the slot is handed back
      iters_left[b] = __ldcg(iters_left + b) - 1;
      __threadfence();
      atomicExch(busy + b, 0);
    }
  }
}


// ------------------------------------------------------------------------------------------
// Variational models.  State arrays are batch-major: Z [B, cap, d], hyp [B, d+4], nat1 [B, cap],
// nat2 [B, cap, cap], adam_m / adam_v [B, cap*d + d + 4]; work [B, var_work_doubles(cap)].
// ------------------------------------------------------------------------------------------
struct VarPtrs { double *Z, *hyp, *nat1, *nat2, *adam_m, *adam_v; };

__device__ __forceinline__ VarModel var_model_of(const VarPtrs& P, int b, int cap, int d) {
  const size_t na = (size_t)cap * d + d + 4;
  return VarModel{P.Z + (size_t)b * cap * d, P.hyp + (size_t)b * (d + 4), P.nat1 + (size_t)b * cap,
                  P.nat2 + (size_t)b * cap * cap, P.adam_m ? P.adam_m + b * na : nullptr,
                  P.adam_v ? P.adam_v + b * na : nullptr};
}

#ifndef STP_VAR_MINB
#define STP_VAR_MINB 2
#endif
// BIG: the working square M0 of a campaign lives in its global workspace (n_cap > kVarSmemSquareMax) instead of
// shared memory; a template parameter so that each instantiation addresses M0 in one known address space.
// BIG campaigns (M0 in the global workspace, 145 KB of shared memory: one CTA per SM) run with kVarBigThreads threads.
constexpr int kVarBigThreads = 512;
template <bool BIG>
__global__ void __launch_bounds__(BIG ? kVarBigThreads : 256, BIG ? 1 : STP_VAR_MINB) k_var_fit(int n_cap, int d, const double* __restrict__ X,
                                                 const double* __restrict__ y, const int32_t* __restrict__ n_arr,
                                                 const double* __restrict__ init_noise, VarFitArgs args, int t0,
                                                 VarPtrs P, double* __restrict__ loss_hist, int32_t* __restrict__ status,
                                                 double* __restrict__ work, const int32_t* __restrict__ order,
                                                 int with_panels) {
  extern __shared__ double smem[];
  const int b = order ? order[blockIdx.x] : blockIdx.x;   // longest-first scheduling (see stp_exact_bo_step)
  if (status[b] != 0) return;
  const int n = n_arr ? n_arr[b] : n_cap;
  if (n < 1 || n > n_cap) { if (threadIdx.x == 0) status[b] = 2; return; }
  VarSmem s = var_smem_carve(smem, n_cap, d, with_panels != 0, BIG ? 0 : 1);
  CtaTeam tm{s.red};
  const VarWork w = var_work_carve(work + (size_t)b * var_work_doubles(n_cap), n_cap);
  if (BIG) s.M0 = w.M0g;
  const VarModel M = var_model_of(P, b, n_cap, d);
  const double* Xb = X + (size_t)b * n_cap * d;
  var_load(tm, s, n, d, Xb, y + (size_t)b * n_cap, 1);
  if (args.fresh) var_init_state(tm, M, n, d, n_cap, Xb, init_noise ? init_noise + (size_t)b * n_cap : nullptr, args, b);
  int st = 0;
  for (int e = 0; e < args.epochs; ++e) {
    const VarEpochOut r = var_epoch<BIG>(tm, s, w, M, n, d, args.lik, args.prior, args.gh_x, args.gh_w, t0 + e + 1,
                                    args.lr_ngd, args.lr_adam, 1, (double*)nullptr);
    if (r.status) { st = r.status; break; }
    if (loss_hist && threadIdx.x == 0) loss_hist[(size_t)b * args.epochs + e] = r.loss;
  }
  if (threadIdx.x == 0) status[b] = st;
}

template <bool BIG>
__global__ void __launch_bounds__(256) k_var_elbo_fg(int n_cap, int d, int lik, int standardise,
                                                     const double* __restrict__ X, const double* __restrict__ y,
                                                     const int32_t* __restrict__ n_arr, VarFitArgs args, VarPtrs P,
                                                     double* __restrict__ loss, double* __restrict__ grads,
                                                     int32_t* __restrict__ status, double* __restrict__ work) {
  extern __shared__ double smem[];
  const int b = blockIdx.x;
  const int n = n_arr ? n_arr[b] : n_cap;
  VarSmem s = var_smem_carve(smem, n_cap, d, false, BIG ? 0 : 1);
  CtaTeam tm{s.red};
  const VarWork w = var_work_carve(work + (size_t)b * var_work_doubles(n_cap), n_cap);
  if (BIG) s.M0 = w.M0g;
  const VarModel M = var_model_of(P, b, n_cap, d);
  var_load(tm, s, n, d, X + (size_t)b * n_cap * d, y + (size_t)b * n_cap, standardise);
  const size_t G = (size_t)n_cap + (size_t)n_cap * n_cap + (size_t)n_cap * d + d + 4;
  const VarEpochOut r = var_epoch<BIG>(tm, s, w, M, n, d, lik, args.prior, args.gh_x, args.gh_w, 1, 0.0, 0.0, 0,
                                  grads + (size_t)b * G);
  if (threadIdx.x == 0) { loss[b] = r.loss; status[b] = r.status; }
}

// T512: 512-thread CTAs -- always for BIG, and for the smaller capacities whenever the shared-memory footprint leaves
// room for one CTA per SM only (n_cap > ~75): the Monte-Carlo acquisition is 97 % of this kernel and runs at the rate
// the fp64 pipe is kept fed, i.e. with the number of resident warps (ncu: 23 % warps active, 42 % of the samples on
// fixed-latency waits with 8 warps per SM).  Otherwise 256 threads with at most 85 registers: three CTAs per SM.
template <bool BIG, bool T512 = BIG>
__global__ void __launch_bounds__(T512 ? kVarBigThreads : 256, T512 ? 1 : 3) k_var_acq(int n_cap, int d, int N, int lik, const double* __restrict__ pool,
                                                 const int32_t* __restrict__ pool_id, uint8_t* __restrict__ mask,
                                                 VarPtrs P, int32_t* __restrict__ n_arr, const double* __restrict__ best_f,
                                                 int S, const double* __restrict__ eps, unsigned long long seed,
                                                 int32_t* __restrict__ out_idx, double* __restrict__ out_acq,
                                                 double* __restrict__ mean, double* __restrict__ var,
                                                 double* __restrict__ acq, int update, const double* __restrict__ pool_y,
                                                 double* __restrict__ X, double* __restrict__ y, int32_t* __restrict__ trace,
                                                 int32_t* __restrict__ status, double* __restrict__ work,
                                                 const int32_t* __restrict__ order, ModelArg model) {
  extern __shared__ double smem[];
  const int b = order ? order[blockIdx.x] : blockIdx.x;
  if (status[b] != 0) { if (threadIdx.x == 0 && out_idx) out_idx[b] = -1; return; }
  const int n = n_arr ? n_arr[b] : n_cap;
  if (n < 1 || n > n_cap || (update && n >= n_cap)) { if (threadIdx.x == 0) status[b] = 2; return; }
  VarSmem s = var_smem_carve(smem, n_cap, d, false, BIG ? 0 : 1);
  // BIG: the working square is global, so the stage region (busy only inside var_prepare) also hosts the candidate tile
  // (the tile starts on gZ, which only the training epoch uses)
  double* tile = BIG ? smem + var_smem_acq_vector_doubles(n_cap, d) : smem + var_smem_doubles(n_cap, d);
  double* part = tile + (size_t)(((n_cap + 15) & ~15) + kMaxD) * kTC;
#if !defined(STP_NO_DMMA)
  s.panels = tile;   // the candidate tile is idle during var_prepare and larger than the two sweep panels
#endif
  CtaTeam tm{s.red};
  const VarWork w = var_work_carve(work + (size_t)b * var_work_doubles(n_cap), n_cap);
  if (BIG) s.M0 = w.M0g;
  const VarModel M = var_model_of(P, b, n_cap, d);
  const int st = var_prepare<BIG>(tm, s, w, M, n, d, !BIG);
  if (st) { if (threadIdx.x == 0) { status[b] = st; if (out_idx) out_idx[b] = -1; if (out_acq) out_acq[b] = NAN; } return; }
  double best;
  if (best_f) best = best_f[b];
  else {                                             // best_f = y_train.max(), RAW (runners.py:89, App. D-2)
    best = -INFINITY;
    for (int i = 0; i < n; ++i) best = fmax(best, y[(size_t)b * n_cap + i]);
  }
  const int pid = pool_id ? pool_id[b] : 0;
  const double* Pl = pool + (size_t)pid * N * d;
  uint8_t* mb = mask ? mask + (size_t)b * N : nullptr;
  int bi; double bv;
  var_acquire(tm, s, w, n, d, lik, tile, part, Pl, N, mb, best, S, eps ? eps + (size_t)b * N * S : nullptr, seed,
              (unsigned long long)b, &bi, &bv, mean ? mean + (size_t)b * N : nullptr,
              var ? var + (size_t)b * N : nullptr, acq ? acq + (size_t)b * N : nullptr, !BIG, model.acquisition, model.beta);
  if (threadIdx.x == 0) { if (out_idx) out_idx[b] = bi; if (out_acq) out_acq[b] = bv; }
  if (!update) return;
  if (bi < 0) { if (threadIdx.x == 0) status[b] = 2; return; }
  if (threadIdx.x < d) X[((size_t)b * n_cap + n) * d + threadIdx.x] = Pl[(size_t)bi * d + threadIdx.x];
  if (threadIdx.x == 0) {
    y[(size_t)b * n_cap + n] = pool_y[(size_t)pid * N + bi];
    trace[(size_t)b * n_cap + n] = bi;
    mb[bi] = 0;
    n_arr[b] = n + 1;
  }
}

__global__ void k_philox_normal(unsigned long long seed, unsigned long long stream, unsigned long long counter0,
                                long long pairs, double* __restrict__ out) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < pairs; q += (long long)gridDim.x * blockDim.x)
    philox_normal2(seed, stream, counter0 + q, out[2 * q], out[2 * q + 1]);
}

// ------------------------------------------------------------------------------------------
// FP64 FMA throughput probe: 8 independent accumulator chains per thread.
__global__ void k_fp64_probe(int iters, double* sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 12345.678) sink[0] = r;  // keeps the loop alive; never true in practice
}

// The same for the fp64 tensor-core path: 8 independent m8n8k4 DMMA accumulator pairs per warp
// (256 MAC per instruction).
__global__ void k_fp64_mma_probe(int iters, double* sink) {
  double c[8][2];
#pragma unroll
  for (int u = 0; u < 8; ++u) { c[u][0] = threadIdx.x * 1e-9 + u; c[u][1] = c[u][0] + 0.5; }
  const double a = 1.0000001, b = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int u = 0; u < 8; ++u) dmma884(c[u][0], c[u][1], a, b);
    }
  }
  double r = 0.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) r += c[u][0] + c[u][1];
  if (r == 12345.678) sink[0] = r;
}

// Elementwise LogEI over `count` (mean, var) pairs; grid-stride.
__global__ void k_log_ei(long long count, const double* __restrict__ mean, const double* __restrict__ var,
                         double best_f, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = log_ei(mean[i], var[i], best_f);
}

// Initial design, step 2 (utils.py:272-279): every query point (a scaled Sobol point) is replaced by the FIRST nearest
// eligible pool point, dist = sqrt(sum_k (xe_k - p_k)^2) accumulated in k order without contraction (the host's
// torch.norm).  One warp per query; lanes stride over the eligible points; first-index tie-break.
__global__ void k_design_snap(int Q, int Ne, int d, const double* __restrict__ P, const double* __restrict__ Xe,
                              int32_t* __restrict__ out) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= Q) return;
  double p[kMaxD];
#pragma unroll
  for (int k = 0; k < kMaxD; ++k) p[k] = (k < d) ? P[(size_t)q * d + k] : 0.0;
  double best = INFINITY;
  int bi = 0x7fffffff;
  for (int e = lane; e < Ne; e += 32) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kMaxD; ++k)
      if (k < d) {
        const double t = __dsub_rn(Xe[(size_t)e * d + k], p[k]);
        acc = __dadd_rn(acc, __dmul_rn(t, t));
      }
    const double dist = __dsqrt_rn(acc);
    if (dist < best) { best = dist; bi = e; }       // ascending e per lane: strict < keeps the first
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) out[q] = bi;
}

__global__ void k_acq_values(long long count, const double* __restrict__ mean, const double* __restrict__ var,
                             double best_f, ModelArg model, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = acq_value(model.acquisition, model.beta, mean[i], var[i], best_f);
}

__global__ void k_log_ei_t(long long count, const double* __restrict__ mean, const double* __restrict__ sd,
                           const double* __restrict__ df, double best_f, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = log_ei_t(mean[i], sd[i], df[i], best_f);
}

// Opt in to > 48 KB of dynamic shared memory.  cudaFuncSetAttribute blocks while instances of the kernel are
// running, which would serialise launches on different streams, so it is issued only when the requested size
// grows (cached per kernel and device).
template <class K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return fail("campaign does not fit in shared memory (n_cap too large)");
  if (bytes <= 48 * 1024) return 0;
  struct Grant { const void* fn; int dev; size_t bytes; };
  static Grant table[64];
  static int used = 0;
  static std::mutex guard;                       // the ABI may be called from one host thread per GPU
  std::lock_guard<std::mutex> lock(guard);
  int dev = 0;
  cudaGetDevice(&dev);
  const void* fn = reinterpret_cast<const void*>(kernel);
  Grant* slot = nullptr;
  for (int i = 0; i < used; ++i)
    if (table[i].fn == fn && table[i].dev == dev) { slot = &table[i]; break; }
  if (!slot && used < 64) { slot = &table[used++]; slot->fn = fn; slot->dev = dev; slot->bytes = 0; }
  if (!slot || bytes > slot->bytes) {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail_cuda("cudaFuncSetAttribute", e);
    if (slot) slot->bytes = bytes;
  }
  return 0;
}

int check_dims(int B, int n_cap, int d, int n_limit = STP_MAX_N) {
  if (B < 0) return fail("B < 0");
  if (d < 1 || d > STP_MAX_D) return fail("d out of range [1, STP_MAX_D]");
  if (n_cap < 1 || n_cap > n_limit)
    return fail(n_limit == STP_MAX_N ? "n_cap out of range [1, STP_MAX_N] (larger training sets: the stp_exact_*_large entry points)"
                                     : (n_limit == STP_MAX_N_VAR ? "n_cap out of range [1, STP_MAX_N_VAR]"
                                                                 : "n_cap out of range [1, STP_MAX_N_LARGE]"));
  return 0;
}

int fill_bounds(BoundsArg& b, int d, const double* lower, const double* upper, const double* theta0) {
  for (int k = 0; k < kMaxP; ++k) { b.lb[k] = -HUGE_VAL; b.ub[k] = HUGE_VAL; b.theta0[k] = 0.0; }
  for (int k = 0; k < d + 3; ++k) {
    if (lower) b.lb[k] = lower[k];
    if (upper) b.ub[k] = upper[k];
    if (theta0) b.theta0[k] = theta0[k];
  }
  return 0;
}

LbfgsbOpts to_opts(const stp_lbfgsb_opts* o) {
  LbfgsbOpts r{10, 15000, 15000, 20, 2.220446049250313e-09 / 2.220446049250313e-16, 1e-5};
  if (o) { r.m = o->m; r.maxiter = o->maxiter; r.maxfun = o->maxfun; r.maxls = o->maxls; r.factr = o->factr; r.pgtol = o->pgtol; }
  if (r.m < 1) r.m = 1;
  if (r.m > kHist) r.m = kHist;
  return r;
}

int after_launch(const char* name) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(name, e);
  return 0;
}

}  // namespace

extern "C" {

int stp_version(void) { return STP_ABI_VERSION; }
const char* stp_last_error_string(void) { return g_err; }

// Shared bodies of the model-level exact entry points: work == nullptr -> the campaign lives in shared memory
// (n_cap <= STP_MAX_N), else its square is in `work` (n_cap <= STP_MAX_N_LARGE).
static int exact_check(int B, int n_cap, int d, const double* work, bool large) {
  if (check_dims(B, n_cap, d, large ? STP_MAX_N_LARGE : STP_MAX_N)) return -1;
  if (large && !work && B > 0) return fail("null workspace (stp_exact_large_workspace_bytes)");
  if (large && ((uintptr_t)work & 7)) return fail("workspace must be 8-byte aligned");
  return 0;
}

static int exact_mll_fg_impl(const char* what, bool large, int B, int n_cap, int d, const double* X, const double* y,
                             const int32_t* n, const double* theta, const double* prior, double* f, double* g,
                             int32_t* status, const stp_model_opts* model, double* work, void* stream) {
  ModelArg m;
  if (fill_model(m, model, true)) return -1;
  if (exact_check(B, n_cap, d, work, large)) return -1;
  if (!X || !y || !theta || !prior || !f || !g || !status) return fail("null argument");
  if (B == 0) return 0;
  PriorArg p; memcpy(p.v, prior, sizeof(p.v));
  if (large) {
    const size_t smem = exact_large_smem_bytes(n_cap, d, false);
    if (set_smem(k_exact_mll_fg_large, smem)) return -1;
    k_exact_mll_fg_large<<<B, threads_large(), smem, (cudaStream_t)stream>>>(n_cap, d, X, y, n, theta, p, f, g, status, m, work);
  } else {
    const int threads = threads_for(n_cap);
    const size_t smem = exact_smem_bytes(n_cap, d, threads, false);
    if (set_smem(k_exact_mll_fg, smem)) return -1;
    k_exact_mll_fg<<<B, threads, smem, (cudaStream_t)stream>>>(n_cap, d, X, y, n, theta, p, f, g, status, m);
  }
  return after_launch(what);
}

static int exact_fit_impl(const char* what, bool large, int B, int n_cap, int d, const double* X, const double* y,
                          const int32_t* n, double* theta, const double* lower, const double* upper,
                          const double* prior, const stp_lbfgsb_opts* opts, double* f_out, int32_t* nit, int32_t* nfev,
                          int32_t* status, const stp_model_opts* model, double* work, void* stream) {
  ModelArg m;
  if (fill_model(m, model, true)) return -1;
  if (exact_check(B, n_cap, d, work, large)) return -1;
  if (!X || !y || !theta || !prior || !status) return fail("null argument");
  if (B == 0) return 0;
  PriorArg p; memcpy(p.v, prior, sizeof(p.v));
  BoundsArg bnd; fill_bounds(bnd, d, lower, upper, nullptr);
  if (large) {
    const size_t smem = exact_large_smem_bytes(n_cap, d, true);
    if (set_smem(k_exact_fit<true>, smem)) return -1;
    k_exact_fit<true><<<B, threads_large(), smem, (cudaStream_t)stream>>>(n_cap, d, X, y, n, theta, bnd, p, to_opts(opts),
                                                                          f_out, nit, nfev, status, m, work);
  } else {
    const int threads = threads_for(n_cap);
    const size_t smem = exact_smem_bytes(n_cap, d, threads, true);
    if (set_smem(k_exact_fit<false>, smem)) return -1;
    k_exact_fit<false><<<B, threads, smem, (cudaStream_t)stream>>>(n_cap, d, X, y, n, theta, bnd, p, to_opts(opts), f_out,
                                                                   nit, nfev, status, m, nullptr);
  }
  return after_launch(what);
}

static int exact_acq_impl(const char* what, bool large, int B, int n_cap, int d, int N, const double* pool,
                          const int32_t* pool_id, const uint8_t* mask, const double* X, const double* y,
                          const int32_t* n, const double* theta, const double* best_f, int32_t* out_idx,
                          double* out_acq, double* mean, double* var, double* acq, int32_t* status,
                          const stp_model_opts* model, double* work, void* stream) {
  ModelArg m;
  if (fill_model(m, model, true)) return -1;
  if (exact_check(B, n_cap, d, work, large)) return -1;
  if (N < 0) return fail("N < 0");
  if (!pool || !X || !y || !theta || !status) return fail("null argument");
  if (B == 0) return 0;
  if (large) {
    const size_t smem = exact_large_smem_bytes(n_cap, d, false);
    if (set_smem(k_exact_acq_large, smem)) return -1;
    k_exact_acq_large<<<B, threads_large(), smem, (cudaStream_t)stream>>>(n_cap, d, N, pool, pool_id, mask, X, y, n, theta,
                                                                          best_f, out_idx, out_acq, mean, var, acq, status,
                                                                          m, work);
  } else {
    const int threads = threads_for(n_cap);
    const size_t smem = exact_smem_bytes(n_cap, d, threads, false);
    if (set_smem(k_exact_acq, smem)) return -1;
    k_exact_acq<<<B, threads, smem, (cudaStream_t)stream>>>(n_cap, d, N, pool, pool_id, mask, X, y, n, theta, best_f,
                                                            out_idx, out_acq, mean, var, acq, status, m);
  }
  return after_launch(what);
}

int stp_exact_mll_fg(int B, int n_cap, int d, const double* X, const double* y, const int32_t* n,
                     const double* theta, const double* prior, double* f, double* g, int32_t* status,
                     const stp_model_opts* model, void* stream) {
  return exact_mll_fg_impl("stp_exact_mll_fg", false, B, n_cap, d, X, y, n, theta, prior, f, g, status, model, nullptr, stream);
}

int stp_exact_fit(int B, int n_cap, int d, const double* X, const double* y, const int32_t* n, double* theta,
                  const double* lower, const double* upper, const double* prior, const stp_lbfgsb_opts* opts,
                  double* f_out, int32_t* nit, int32_t* nfev, int32_t* status, const stp_model_opts* model, void* stream) {
  return exact_fit_impl("stp_exact_fit", false, B, n_cap, d, X, y, n, theta, lower, upper, prior, opts, f_out, nit, nfev,
                        status, model, nullptr, stream);
}

int stp_exact_acq(int B, int n_cap, int d, int N, const double* pool, const int32_t* pool_id, const uint8_t* mask,
                  const double* X, const double* y, const int32_t* n, const double* theta, const double* best_f,
                  int32_t* out_idx, double* out_acq, double* mean, double* var, double* acq, int32_t* status,
                  const stp_model_opts* model, void* stream) {
  return exact_acq_impl("stp_exact_acq", false, B, n_cap, d, N, pool, pool_id, mask, X, y, n, theta, best_f, out_idx,
                        out_acq, mean, var, acq, status, model, nullptr, stream);
}

// ---- large capacities (row f3): the same three operations with the working square in a caller-owned workspace
size_t stp_exact_large_workspace_bytes(int B, int n_cap) {
  if (B <= 0 || n_cap < 1 || n_cap > STP_MAX_N_LARGE) return 0;
  return (size_t)B * exact_large_work_doubles(n_cap) * sizeof(double);
}

int stp_exact_mll_fg_large(int B, int n_cap, int d, const double* X, const double* y, const int32_t* n,
                           const double* theta, const double* prior, double* f, double* g, int32_t* status,
                           const stp_model_opts* model, void* work, void* stream) {
  return exact_mll_fg_impl("stp_exact_mll_fg_large", true, B, n_cap, d, X, y, n, theta, prior, f, g, status, model,
                           (double*)work, stream);
}

int stp_exact_fit_large(int B, int n_cap, int d, const double* X, const double* y, const int32_t* n, double* theta,
                        const double* lower, const double* upper, const double* prior, const stp_lbfgsb_opts* opts,
                        double* f_out, int32_t* nit, int32_t* nfev, int32_t* status, const stp_model_opts* model,
                        void* work, void* stream) {
  return exact_fit_impl("stp_exact_fit_large", true, B, n_cap, d, X, y, n, theta, lower, upper, prior, opts, f_out, nit,
                        nfev, status, model, (double*)work, stream);
}

int stp_exact_acq_large(int B, int n_cap, int d, int N, const double* pool, const int32_t* pool_id, const uint8_t* mask,
                        const double* X, const double* y, const int32_t* n, const double* theta, const double* best_f,
                        int32_t* out_idx, double* out_acq, double* mean, double* var, double* acq, int32_t* status,
                        const stp_model_opts* model, void* work, void* stream) {
  return exact_acq_impl("stp_exact_acq_large", true, B, n_cap, d, N, pool, pool_id, mask, X, y, n, theta, best_f,
                        out_idx, out_acq, mean, var, acq, status, model, (double*)work, stream);
}

static int fill_bo_args(BoArgs& a, int B, int n_cap, int d, int N, const double* pool, const double* pool_y,
                        const int32_t* pool_id, uint8_t* mask, double* X, double* y, int32_t* n, int32_t* trace,
                        const double* theta0, const double* lower, const double* upper, const double* prior,
                        const stp_lbfgsb_opts* opts, double* theta_out, double* acq_out, int32_t* nit, int32_t* nfev,
                        int32_t* status, const stp_turnover* turnover, double* work_acc, int32_t* iters_acc, int n_max,
                        const stp_model_opts* model, int n_limit = STP_MAX_N) {
  a.work = nullptr;
  if (fill_model(a.model, model, true)) return -1;
  if (check_dims(B, n_cap, d, n_limit)) return -1;
  if (N < 1) return fail("N < 1");
  if (n_max < 0 || n_max >= n_cap) return fail("n_max must be 0 (unknown) or in [1, n_cap - 1]");
  memset(&a.turn, 0, sizeof(a.turn));
  if (turnover) {
    a.turn = *turnover;
    const stp_turnover& turn = a.turn;
    if (turn.n_final > 0 && (!turn.bank || !turn.gen || turn.n_init < 1 || turn.n_init >= n_cap || turn.generations < 0 ||
                             turn.n_final > n_cap || (turn.sink && !turn.sink_count)))
      return fail("inconsistent stp_turnover");
  }
  if (!pool || !pool_y || !mask || !X || !y || !n || !trace || !theta0 || !prior || !status) return fail("null argument");
  a.n_cap = n_cap; a.d = d; a.N = N; a.s_cap = n_max > 0 ? n_max : n_cap - 1;
  a.pool = pool; a.pool_y = pool_y; a.pool_id = pool_id; a.mask = mask; a.X = X; a.y = y; a.n_arr = n; a.trace = trace;
  fill_bounds(a.bnd, d, lower, upper, theta0);
  memcpy(a.prior.v, prior, sizeof(a.prior.v));
  a.opts = to_opts(opts);
  a.theta_out = theta_out; a.acq_out = acq_out; a.nit = nit; a.nfev = nfev; a.status = status;
  a.work_acc = work_acc; a.iters_acc = iters_acc;
  return 0;
}

int stp_exact_bo_step(int B, int n_cap, int d, int N, const double* pool, const double* pool_y,
                      const int32_t* pool_id, uint8_t* mask, double* X, double* y, int32_t* n, int32_t* trace,
                      const double* theta0, const double* lower, const double* upper, const double* prior,
                      const stp_lbfgsb_opts* opts, double* theta_out, double* acq_out, int32_t* nit, int32_t* nfev,
                      int32_t* status, const int32_t* order, const stp_turnover* turnover, double* work_acc,
                      int32_t* iters_acc, int n_max, const stp_model_opts* model, void* stream) {
  BoArgs a;
  if (fill_bo_args(a, B, n_cap, d, N, pool, pool_y, pool_id, mask, X, y, n, trace, theta0, lower, upper, prior, opts,
                   theta_out, acq_out, nit, nfev, status, turnover, work_acc, iters_acc, n_max, model)) return -1;
  if (B == 0) return 0;
  const int threads = threads_for(a.s_cap);
  const size_t smem = exact_smem_bytes(a.s_cap, d, threads, true) + smem_pad();
  if (set_smem(k_exact_bo_step, smem)) return -1;
  k_exact_bo_step<<<B, threads, smem, (cudaStream_t)stream>>>(a, order);
  return after_launch("stp_exact_bo_step");
}

int stp_exact_bo_step_large(int B, int n_cap, int d, int N, const double* pool, const double* pool_y,
                            const int32_t* pool_id, uint8_t* mask, double* X, double* y, int32_t* n, int32_t* trace,
                            const double* theta0, const double* lower, const double* upper, const double* prior,
                            const stp_lbfgsb_opts* opts, double* theta_out, double* acq_out, int32_t* nit, int32_t* nfev,
                            int32_t* status, const int32_t* order, const stp_turnover* turnover, double* work_acc,
                            int32_t* iters_acc, int n_max, const stp_model_opts* model, void* work, void* stream) {
  BoArgs a;
  if (fill_bo_args(a, B, n_cap, d, N, pool, pool_y, pool_id, mask, X, y, n, trace, theta0, lower, upper, prior, opts,
                   theta_out, acq_out, nit, nfev, status, turnover, work_acc, iters_acc, n_max, model, STP_MAX_N_LARGE)) return -1;
  if (B == 0) return 0;
  if (!work || ((uintptr_t)work & 7)) return fail("null or misaligned workspace (stp_exact_large_workspace_bytes)");
  a.work = (double*)work;
  const size_t smem = exact_large_smem_bytes(a.s_cap, d, true);
  if (set_smem(k_exact_bo_step_large, smem)) return -1;
  k_exact_bo_step_large<<<B, threads_large(), smem, (cudaStream_t)stream>>>(a, order);
  return after_launch("stp_exact_bo_step_large");
}

int stp_exact_bo_run(int B, int n_cap, int d, int N, const double* pool, const double* pool_y,
                     const int32_t* pool_id, uint8_t* mask, double* X, double* y, int32_t* n, int32_t* trace,
                     const double* theta0, const double* lower, const double* upper, const double* prior,
                     const stp_lbfgsb_opts* opts, double* theta_out, double* acq_out, int32_t* nit, int32_t* nfev,
                     int32_t* status, const int32_t* order, const stp_turnover* turnover, double* work_acc,
                     int32_t* iters_acc, int n_max, int32_t* iters_left, int32_t* sched, int max_ctas,
                     const stp_model_opts* model, void* stream) {
  BoArgs a;
  if (fill_bo_args(a, B, n_cap, d, N, pool, pool_y, pool_id, mask, X, y, n, trace, theta0, lower, upper, prior, opts,
                   theta_out, acq_out, nit, nfev, status, turnover, work_acc, iters_acc, n_max, model)) return -1;
  if (!iters_left || !sched) return fail("null argument");
  if (B == 0) return 0;
  const int threads = threads_for(a.s_cap);
  const size_t smem = exact_smem_bytes(a.s_cap, d, threads, true) + smem_pad();
  if (set_smem(k_exact_bo_run, smem)) return -1;
  // one CTA per SM slot: as many as can be resident at this footprint (cached per device / footprint)
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_exact_bo_run, threads, smem);
  if (e != cudaSuccess || per_sm < 1 || sms < 1) return fail_cuda("cudaOccupancyMaxActiveBlocksPerMultiprocessor", e);
  int grid = sms * per_sm;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid > B) grid = B;
  // sched: int32[B + 1], zeroed by the caller before the FIRST launch: busy flags, then the ticket counter
  k_exact_bo_run<<<grid, threads, smem, (cudaStream_t)stream>>>(a, B, order, iters_left, sched,
                                                                reinterpret_cast<unsigned int*>(sched + B));
  return after_launch("stp_exact_bo_run");
}


static int fill_var_args(VarFitArgs& a, int d, int lik, int epochs, int fresh, double lr_ngd, double lr_adam,
                         const double* hyp0, const double* prior, const double* gh, unsigned long long seed) {
  if (lik < 0 || lik > 2) return fail("lik must be 0 (gaussian), 1 (student-t) or 2 (laplace)");
  if (!prior || !gh) return fail("null prior / gauss-hermite table");
  a.lik = lik; a.epochs = epochs; a.fresh = fresh; a.lr_ngd = lr_ngd; a.lr_adam = lr_adam; a.seed = seed;
  for (int k = 0; k < kMaxD + 4; ++k) a.hyp0[k] = (hyp0 && k < d + 4) ? hyp0[k] : 0.0;
  a.prior = VarHyperPrior{prior[0], prior[1], prior[2], prior[3], prior[4], prior[5], prior[6], prior[7]};
  for (int q = 0; q < kGH; ++q) { a.gh_x[q] = gh[q]; a.gh_w[q] = gh[kGH + q]; }
  return 0;
}

// The blocked sweep (P, and K_ZZ with the Cholesky export) needs two panels of 12 doubles per row in shared memory:
// 111 KB per CTA at n_cap = 91, d = 6 (two CTAs per SM), always carved on the device.
static int var_use_panels(int n_cap, int d) {
  (void)n_cap; (void)d;
#if defined(STP_NO_DMMA)
  return 0;
#else
  return 1;
#endif
}

// shared memory of a stp_var_acq launch
static size_t var_acq_smem_bytes(int n_cap, int d) {
  size_t smem = (var_smem_doubles(n_cap, d) + var_acq_smem_doubles(n_cap, kTC)) * sizeof(double);
  if (n_cap > kVarSmemSquareMax) {   // the tile aliases gZ and the stage region (see k_var_acq)
    const size_t with_tile = var_smem_acq_vector_doubles(n_cap, d) + var_acq_smem_doubles(n_cap, kTC);
    const size_t with_stage = var_smem_vector_doubles(n_cap, d) + contract_stage_doubles();
    smem = (with_tile > with_stage ? with_tile : with_stage) * sizeof(double);
  }
  return smem;
}

int stp_var_capacity(int n, int d) {
  if (n < 1 || d < 1 || d > STP_MAX_D) return -1;
  const size_t limit = 227 * 1024;
  for (int cap = n; cap <= STP_MAX_N_VAR; ++cap) {
    const size_t fit = var_smem_doubles(cap, d, var_use_panels(cap, d) != 0) * sizeof(double);
    const size_t fg = var_smem_doubles(cap, d, false) * sizeof(double);
    if (fit <= limit && fg <= limit && var_acq_smem_bytes(cap, d) <= limit) return cap;
  }
  return -1;
}

size_t stp_var_workspace_bytes(int B, int n_cap) {
  if (B < 0 || n_cap < 1) return 0;
  return (size_t)B * var_work_doubles(n_cap) * sizeof(double);
}

int stp_var_fit(int B, int n_cap, int d, int lik, int epochs, double lr_ngd, double lr_adam, int fresh, int t0,
                const double* X, const double* y, const int32_t* n, const double* init_noise, unsigned long long seed,
                const double* hyp0, const double* prior, const double* gh, double* Z, double* hyp, double* nat1,
                double* nat2, double* adam_m, double* adam_v, double* loss_hist, int32_t* status, double* work,
                const int32_t* order, void* stream) {
  if (check_dims(B, n_cap, d, STP_MAX_N_VAR)) return -1;
  if (epochs < 0) return fail("epochs < 0");
  if (!X || !y || !Z || !hyp || !nat1 || !nat2 || !adam_m || !adam_v || !status || !work) return fail("null argument");
  if (fresh && !hyp0) return fail("fresh model needs hyp0");
  if (B == 0) return 0;
  VarFitArgs a;
  if (fill_var_args(a, d, lik, epochs, fresh, lr_ngd, lr_adam, hyp0, prior, gh, seed)) return -1;
  const int with_panels = var_use_panels(n_cap, d);
  const size_t smem = var_smem_doubles(n_cap, d, with_panels != 0) * sizeof(double);
  const VarPtrs ptrs{Z, hyp, nat1, nat2, adam_m, adam_v};
  if (n_cap > kVarSmemSquareMax) {
    if (set_smem(k_var_fit<true>, smem)) return -1;
    k_var_fit<true><<<B, kVarBigThreads, smem, (cudaStream_t)stream>>>(n_cap, d, X, y, n, init_noise, a, t0, ptrs, loss_hist, status,
                                                            work, order, with_panels);
  } else {
    if (set_smem(k_var_fit<false>, smem)) return -1;
    k_var_fit<false><<<B, 256, smem, (cudaStream_t)stream>>>(n_cap, d, X, y, n, init_noise, a, t0, ptrs, loss_hist, status,
                                                             work, order, with_panels);
  }
  return after_launch("stp_var_fit");
}

int stp_var_elbo_fg(int B, int n_cap, int d, int lik, int standardise, const double* X, const double* y,
                    const int32_t* n, const double* Z, const double* hyp, const double* nat1, const double* nat2,
                    const double* prior, const double* gh, double* loss, double* grads, int32_t* status, double* work,
                    void* stream) {
  if (check_dims(B, n_cap, d, STP_MAX_N_VAR)) return -1;
  if (!X || !y || !Z || !hyp || !nat1 || !nat2 || !loss || !grads || !status || !work) return fail("null argument");
  if (B == 0) return 0;
  VarFitArgs a;
  if (fill_var_args(a, d, lik, 1, 0, 0.0, 0.0, nullptr, prior, gh, 0)) return -1;
  const size_t smem = var_smem_doubles(n_cap, d) * sizeof(double);
  const VarPtrs ptrs{const_cast<double*>(Z), const_cast<double*>(hyp), const_cast<double*>(nat1),
                     const_cast<double*>(nat2), nullptr, nullptr};
  if (n_cap > kVarSmemSquareMax) {
    if (set_smem(k_var_elbo_fg<true>, smem)) return -1;
    k_var_elbo_fg<true><<<B, 256, smem, (cudaStream_t)stream>>>(n_cap, d, lik, standardise, X, y, n, a, ptrs, loss, grads,
                                                                status, work);
  } else {
    if (set_smem(k_var_elbo_fg<false>, smem)) return -1;
    k_var_elbo_fg<false><<<B, 256, smem, (cudaStream_t)stream>>>(n_cap, d, lik, standardise, X, y, n, a, ptrs, loss, grads,
                                                                 status, work);
  }
  return after_launch("stp_var_elbo_fg");
}

int stp_var_acq(int B, int n_cap, int d, int N, int lik, const double* pool, const int32_t* pool_id, uint8_t* mask,
                const double* Z, const double* hyp, const double* nat1, const double* nat2, int32_t* n,
                const double* best_f, int S, const double* eps, unsigned long long seed, int32_t* out_idx,
                double* out_acq, double* mean, double* var, double* acq, int update, const double* pool_y, double* X,
                double* y, int32_t* trace, int32_t* status, double* work, const int32_t* order,
                const stp_model_opts* model, void* stream) {
  ModelArg m;
  if (fill_model(m, model, false)) return -1;
  if (check_dims(B, n_cap, d, STP_MAX_N_VAR)) return -1;
  if (N < 0 || S < 1) return fail("N < 0 or S < 1");
  if (lik < 0 || lik > 2) return fail("lik must be 0, 1 or 2");
  if (!pool || !Z || !hyp || !nat1 || !nat2 || !status || !work) return fail("null argument");
  if (update && (!pool_y || !X || !y || !trace || !mask || !n)) return fail("update needs pool_y, X, y, n, trace and mask");
  if (!best_f && !y) return fail("best_f or y required");
  if (B == 0) return 0;
  const size_t smem = var_acq_smem_bytes(n_cap, d);
  const VarPtrs ptrs{const_cast<double*>(Z), const_cast<double*>(hyp), const_cast<double*>(nat1),
                     const_cast<double*>(nat2), nullptr, nullptr};
  if (n_cap > kVarSmemSquareMax) {
    if (set_smem(k_var_acq<true>, smem)) return -1;
    k_var_acq<true><<<B, kVarBigThreads, smem, (cudaStream_t)stream>>>(n_cap, d, N, lik, pool, pool_id, mask, ptrs, n, best_f, S, eps,
                                                            seed, out_idx, out_acq, mean, var, acq, update, pool_y, X, y,
                                                            trace, status, work, order, m);
  } else {
    if (smem > 113u * 1024u) {     // one CTA per SM anyway: give it 16 warps
      if (set_smem(k_var_acq<false, true>, smem)) return -1;
      k_var_acq<false, true><<<B, kVarBigThreads, smem, (cudaStream_t)stream>>>(n_cap, d, N, lik, pool, pool_id, mask, ptrs, n, best_f, S,
                                                                                eps, seed, out_idx, out_acq, mean, var, acq, update, pool_y,
                                                                                X, y, trace, status, work, order, m);
    } else {
      if (set_smem(k_var_acq<false, false>, smem)) return -1;
      k_var_acq<false, false><<<B, 256, smem, (cudaStream_t)stream>>>(n_cap, d, N, lik, pool, pool_id, mask, ptrs, n, best_f, S, eps,
                                                                      seed, out_idx, out_acq, mean, var, acq, update, pool_y, X, y,
                                                                      trace, status, work, order, m);
    }
  }
  return after_launch("stp_var_acq");
}

int stp_philox_normal(unsigned long long seed, unsigned long long stream_id, unsigned long long counter0,
                      long long pairs, double* out, void* stream) {
  if (pairs < 0 || !out) return fail("bad argument");
  if (pairs == 0) return 0;
  long long blocks = (pairs + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_philox_normal<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(seed, stream_id, counter0, pairs, out);
  return after_launch("stp_philox_normal");
}

int stp_log_ei(long long count, const double* mean, const double* var, double best_f, double* out, void* stream) {
  if (count < 0 || !mean || !var || !out) return fail("bad argument");
  if (count == 0) return 0;
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_log_ei<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(count, mean, var, best_f, out);
  return after_launch("stp_log_ei");
}

int stp_acq_values(long long count, const double* mean, const double* var, double best_f, const stp_model_opts* model,
                   double* out, void* stream) {
  ModelArg m;
  if (fill_model(m, model, true)) return -1;
  if (count < 0 || !mean || !var || !out) return fail("bad argument");
  if (count == 0) return 0;
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_acq_values<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(count, mean, var, best_f, m, out);
  return after_launch("stp_acq_values");
}

int stp_log_ei_t(long long count, const double* mean, const double* sd, const double* df, double best_f, double* out,
                 void* stream) {
  if (count < 0 || !mean || !sd || !df || !out) return fail("bad argument");
  if (count == 0) return 0;
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_log_ei_t<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(count, mean, sd, df, best_f, out);
  return after_launch("stp_log_ei_t");
}

int stp_design_snap(int Q, int Ne, int d, const double* points, const double* eligible, int32_t* out_idx, void* stream) {
  if (Q < 0 || Ne < 1 || d < 1 || d > STP_MAX_D) return fail("bad argument");
  if (!points || !eligible || !out_idx) return fail("null argument");
  if (Q == 0) return 0;
  k_design_snap<<<(Q + 7) / 8, 256, 0, (cudaStream_t)stream>>>(Q, Ne, d, points, eligible, out_idx);
  return after_launch("stp_design_snap");
}

int stp_phase_timers(unsigned long long* out32, int reset) {
#if defined(STP_PHASE_TIMERS)
  if (out32) {
    const cudaError_t e = cudaMemcpyFromSymbol(out32, g_stp_phase, 64 * sizeof(unsigned long long));
    if (e != cudaSuccess) return fail_cuda("stp_phase_timers", e);
  }
  if (reset) {
    unsigned long long zero[64] = {0};
    const cudaError_t e = cudaMemcpyToSymbol(g_stp_phase, zero, sizeof(zero));
    if (e != cudaSuccess) return fail_cuda("stp_phase_timers", e);
  }
  return 0;
#else
  (void)out32; (void)reset;
  return fail("library built without -DSTP_PHASE_TIMERS");
#endif
}

int stp_fp64_probe(int iters, double* sink, void* stream) {
  if (iters < 1 || !sink) return fail("bad argument");
  k_fp64_probe<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(iters, sink);
  return after_launch("stp_fp64_probe");
}

int stp_fp64_mma_probe(int iters, double* sink, void* stream) {
  if (iters < 1 || !sink) return fail("bad argument");
  k_fp64_mma_probe<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(iters, sink);
  return after_launch("stp_fp64_mma_probe");
}

}  // extern "C"
