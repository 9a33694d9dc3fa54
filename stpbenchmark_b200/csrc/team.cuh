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
#else
#define STP_HD inline
#define STP_D inline
#endif

#include <math.h>
#include <stdint.h>

namespace stp {

constexpr int kMaxWarps = 32;

// ---------------------------------------------------------------------------------------
// Lane groups: what ONE warp of a team (or the single thread of a SoloTeam) offers to the small serial-ish
// routines (the L-BFGS-B state machine): strided loops over lane() / width(), lane-uniform reductions, sync().
// ---------------------------------------------------------------------------------------
struct SoloLanes {
  STP_HD int lane() const { return 0; }
  STP_HD int width() const { return 1; }
  STP_HD void sync() const {}
  STP_HD double sum(double v) const { return v; }
  STP_HD double max(double v) const { return v; }
  STP_HD int any(int p) const { return p != 0; }
};
#if defined(__CUDACC__)
struct WarpLanes {
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int width() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  // xor butterflies: every lane ends with the bitwise-identical result
  __device__ __forceinline__ double sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ __forceinline__ double max(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ int any(int p) const { return __any_sync(0xffffffffu, p); }
};
#endif

// ---------------------------------------------------------------------------------------
// One-thread team: host emulation (tests/hostemu) and trivially-serial device sections.
// ---------------------------------------------------------------------------------------
struct SoloTeam {
  STP_HD int rank() const { return 0; }
  STP_HD int size() const { return 1; }
  STP_HD int lane() const { return 0; }
  STP_HD int warp() const { return 0; }
  STP_HD int nwarps() const { return 1; }
  STP_HD int warp_width() const { return 1; }
  STP_HD void sync() const {}
  STP_HD void sync_warp() const {}
  STP_HD double sum(double v) const { return v; }
  STP_HD double warp_sum(double v) const { return v; }
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
