// Thread-team abstraction shared by every per-campaign routine.
//
// One BO campaign is owned by one CTA.  The numerical routines (kernel matrices, symmetric
// sweep, Cholesky, MLL/ELBO, L-BFGS-B, fused acquisition) are written once against a "team"
// with rank()/size()/sync()/sum(): on the device the team is the CTA (CtaTeam: __syncthreads
// and warp-shuffle reductions); in tests/hostemu the same source is compiled by g++ with a
// one-thread team (SoloTeam) so that the algebra can be checked on a CPU-only box.  The host
// build is test scaffolding only: the shipped library contains device code exclusively.
#pragma once

#if defined(__CUDACC__)
#define STP_HD __host__ __device__ __forceinline__
#define STP_D __device__ __forceinline__
// Out-of-line on purpose: routines run by ONE warp through a lot of branchy code (the L-BFGS-B state machine) are
// instruction-fetch bound once their inlined copies exceed the SM's instruction caches (L0 ~6 KB, L1.5 32 KB);
// one shared copy per routine keeps the optimiser's footprint small (tools/sass_footprint.py).
#define STP_HD_NOINLINE __host__ __device__ __noinline__
#else
#define STP_HD inline
#define STP_D inline
#define STP_HD_NOINLINE inline
#endif

#include <math.h>
#include <stdint.h>

namespace stp {

constexpr int kMaxWarps = 32;

// Phase timers (diagnostic builds only, -DSTP_PHASE_TIMERS; tools/gpu_phase_timers.py): thread 0 of the CTA charges the
// cycles since its previous mark to phase `id` (marks sit right after the barrier that ends a phase).
// ids: 0 load, 1 Gram build, 2 sweep (A) 8-pivot elimination, 3 sweep (B) panel products, 4 sweep (C) rank-8 update,
// 5 gradient contraction, 6 block reduction + gradient, 7 L-BFGS-B advance (what ids 11 - 15 do not cover),
// 8 factorise, 9 pool pass, 10 update, 11 BFGS matrix, 12 Cauchy point, 13 subspace step, 14 line-search start,
// 15 line-search step + BFGS pair (from the closure's end to the start of the next iteration),
// 16 closures (count), 17 sweep block steps (count), 18 campaigns (count).
#if defined(__CUDACC__) && defined(STP_PHASE_TIMERS)
__device__ unsigned long long g_stp_phase[64];
__device__ __forceinline__ void phase_mark(int id) {
  __shared__ long long last;
  if (threadIdx.x == 0) {
    const long long t = clock64();
    if (id >= 0) atomicAdd(&g_stp_phase[id], (unsigned long long)(t - last));
    last = t;
  }
}
__device__ __forceinline__ void phase_count(int id) { if (threadIdx.x == 0) atomicAdd(&g_stp_phase[id], 1ull); }
#if defined(__CUDA_ARCH__)
#define STP_PHASE(id) ::stp::phase_mark(id)
#define STP_COUNT(id) ::stp::phase_count(id)
#else
#define STP_PHASE(id) ((void)0)
#define STP_COUNT(id) ((void)0)
#endif
#else
#define STP_PHASE(id) ((void)0)
#define STP_COUNT(id) ((void)0)
#endif

// ---------------------------------------------------------------------------------------
// Lane groups: what ONE warp of a team (or the single thread of a SoloTeam) offers to the small serial-ish
// routines (the L-BFGS-B state machine): strided loops over lane() / width(), lane-uniform reductions, sync().
// ---------------------------------------------------------------------------------------
// Loops over the n <= 32 entries of a small vector are written  STP_OWNED(L, i, n) { ... }:  a warp gives entry i
// to lane i (the body runs at most once per lane -- written so that the compiler sees that and emits no unrolled
// strided loop: the optimiser's code footprint is what limits it, see STP_HD_NOINLINE), one thread walks all n.
#define STP_PRAGMA(x) _Pragma(#x)
#define STP_OWNED(L, i, n) STP_PRAGMA(unroll 1) for (int i = (L).first(), i##_end = (L).last(n); i < i##_end; ++i)
#define STP_ROLLED STP_PRAGMA(unroll 1)

struct SoloLanes {
  STP_HD int lane() const { return 0; }
  STP_HD int width() const { return 1; }
  STP_HD int first() const { return 0; }
  STP_HD int last(int n) const { return n; }
  STP_HD void sync() const {}
  STP_HD double sum(double v) const { return v; }
  STP_HD double max(double v) const { return v; }
  STP_HD int isum(int v) const { return v; }
  STP_HD int any(int p) const { return p != 0; }
};
#if defined(__CUDACC__)
// xor butterflies over the 16 low lanes (the small vectors have <= 16 entries; higher lanes contribute the neutral
// element through the first exchange): every lane ends with the bitwise-identical result.  Out of line: ~40 call sites.
static __device__ __noinline__ double warp_sum16(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static __device__ __noinline__ double warp_max16(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 16));
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
struct WarpLanes {
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int width() const { return 32; }
  __device__ __forceinline__ int first() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int last(int n) const { const int l = (threadIdx.x & 31) + 1; return n < l ? n : l; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ double sum(double v) const { return warp_sum16(v); }
  __device__ __forceinline__ double max(double v) const { return warp_max16(v); }
  __device__ __forceinline__ int isum(int v) const { return __reduce_add_sync(0xffffffffu, v); }
  __device__ __forceinline__ int any(int p) const { return __any_sync(0xffffffffu, p); }
};
#endif

// ---------------------------------------------------------------------------------------
// One-thread team: host emulation (tests/hostemu) and trivially-serial device sections.
// ---------------------------------------------------------------------------------------
struct SoloTeam {
  STP_HD int rank() const { return 0; }
  STP_HD int size() const { return 1; }
  STP_HD int first() const { return 0; }           // STP_OWNED(tm, i, n): n <= size() of the device teams
  STP_HD int last(int n) const { return n; }
  STP_HD int lane() const { return 0; }
  STP_HD int warp() const { return 0; }
  STP_HD int nwarps() const { return 1; }
  STP_HD int warp_width() const { return 1; }
  STP_HD void sync() const {}
  STP_HD void sync_warp() const {}
  STP_HD double sum(double v) const { return v; }
  STP_HD double warp_sum(double v) const { return v; }
  STP_HD double quad_sum(double v) const { return v; }
  template <int KMAX>
  STP_HD void sum_vec(double (&v)[KMAX], int k) const {}
  // argmax with torch semantics (first maximal index; NaN wins, first NaN wins)
  STP_HD void argmax(double& v, int& idx) const {}
  STP_HD int any(int p) const { return p; }
  typedef SoloLanes Lanes;
  STP_HD SoloLanes lanes() const { return SoloLanes(); }
};

// torch.argmax ordering: NaN beats everything; ties resolved towards the smaller index.
STP_HD bool acq_better(double v, int i, double w, int j) {
  const bool vn = (v != v), wn = (w != w);
  if (j < 0) return i >= 0;
  if (i < 0) return false;
  if (vn || wn) {
    if (vn && wn) return i < j;
    return vn;
  }
  if (v > w) return true;
  if (v < w) return false;
  return i < j;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------
// CTA team.  `red` points at >= 2*kMaxWarps doubles of shared scratch.
// ---------------------------------------------------------------------------------------
struct CtaTeam {
  double* red;
  __device__ __forceinline__ int rank() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ int first() const { return threadIdx.x; }
  __device__ __forceinline__ int last(int n) const { const int l = threadIdx.x + 1; return n < l ? n : l; }
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int warp() const { return threadIdx.x >> 5; }
  __device__ __forceinline__ int nwarps() const { return blockDim.x >> 5; }
  __device__ __forceinline__ int warp_width() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ void sync_warp() const { __syncwarp(); }

  __device__ __forceinline__ double warp_sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  // sum over the four lanes 4p .. 4p + 3 of a quad, returned to all four
  __device__ __forceinline__ double quad_sum(double v) const {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
  }
  // Block-wide sum, result returned to every thread.  Deterministic for a fixed block size.
  __device__ __forceinline__ double sum(double v) const {
    v = warp_sum(v);
    __syncthreads();  // protect `red` against the previous reduction's readers
    if (lane() == 0) red[warp()] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = nwarps();
    for (int w = 0; w < nw; ++w) t += red[w];
    return t;
  }
  // K simultaneous block-wide sums (K <= 16), in place, every thread gets all of them.
  // Warp stage: a halving butterfly -- at offset o the two partners split the value set, each keeps (and adds)
  // one half -- so 16 values cost 8 + 4 + 2 + 1 + 1 = 16 exchanges instead of 16 x 5; lane l ends with the warp
  // total of value l >> 1.  Block stage: lane q of every warp adds value q over the warps (fixed order) and the
  // totals are handed out by shuffle.  Deterministic for a fixed block size.
  template <int KMAX>
  __device__ __forceinline__ void sum_vec(double (&v)[KMAX], int k) const {
    static_assert(KMAX <= 16, "reduction scratch holds 16 x kMaxWarps doubles");
    const int l = lane();
    double w[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) w[q] = (q < KMAX && q < k) ? v[q < KMAX ? q : 0] : 0.0;
#pragma unroll
    for (int h = 8, o = 16; h >= 1; h >>= 1, o >>= 1) {
      const bool up = (l & o) != 0;
#pragma unroll
      for (int q = 0; q < h; ++q) {
        const double send = up ? w[q] : w[q + h];
        const double keep = up ? w[q + h] : w[q];
        w[q] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
    __syncthreads();   // protect `red` against the previous reduction's readers
    if ((l & 1) == 0) red[(l >> 1) * kMaxWarps + warp()] = w[0];
    __syncthreads();
    const int nw = nwarps();
    double t = 0.0;
    if (l < 16)
      for (int x = 0; x < nw; ++x) t += red[l * kMaxWarps + x];
#pragma unroll
    for (int q = 0; q < KMAX; ++q)
      if (q < k) v[q] = __shfl_sync(0xffffffffu, t, q);
  }
  __device__ __forceinline__ void argmax(double& v, int& idx) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double w = __shfl_xor_sync(0xffffffffu, v, o);
      const int j = __shfl_xor_sync(0xffffffffu, idx, o);
      if (acq_better(w, j, v, idx)) { v = w; idx = j; }
    }
    __syncthreads();
    int* ired = reinterpret_cast<int*>(red + kMaxWarps);
    if (lane() == 0) { red[warp()] = v; ired[warp()] = idx; }
    __syncthreads();
    const int nw = nwarps();
    v = red[0]; idx = ired[0];
    for (int w = 1; w < nw; ++w)
      if (acq_better(red[w], ired[w], v, idx)) { v = red[w]; idx = ired[w]; }
  }
  __device__ __forceinline__ int any(int p) const { return __syncthreads_or(p); }
  typedef WarpLanes Lanes;
  __device__ __forceinline__ WarpLanes lanes() const { return WarpLanes(); }
};
constexpr int kRedDoubles = 16 * kMaxWarps;  // scratch a CtaTeam needs
#endif

}  // namespace stp
