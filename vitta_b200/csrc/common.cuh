// Shared device/host helpers for the vitta_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vitta_b200.h"

namespace vitta {

void set_error(const char* fmt, ...);

#define VITTA_CHECK_ARG(cond, code, ...)  \
  do {                                    \
    if (!(cond)) {                        \
      ::vitta::set_error(__VA_ARGS__);    \
      return (code);                      \
    }                                     \
  } while (0)

#define VITTA_CHECK_LAUNCH()                                                        \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ::vitta::set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

constexpr int kThreads = 256;     // CTA size of the streaming kernels
constexpr int kMaxRowsPerThread = 32;

// Chunking of a channels-last (rows x C) tensor: lanes run along C (float4 each), the remaining threads
// of the CTA take different rows.  See include/vitta_b200.h (K1) for the contract.
struct ClGeom {
  int lpr;         // lanes (float4) per row handled by one CTA
  int rs;          // row slots per CTA = kThreads / lpr
  int ctiles;      // channel tiles = C / (4*lpr)
  int rpt;         // rows per thread
  int chunk_rows;  // rs * rpt
  int cpf;         // chunks per frame
  int64_t frames, frame_rows;
  int64_t n_chunks() const { return frames * cpf; }
};

static inline ClGeom cl_geom(int64_t frames, int64_t frame_rows, int C) {
  ClGeom g;
  int c4 = C / 4;
  g.lpr = c4 < 32 ? c4 : 32;
  // largest power of two <= lpr keeps kThreads % lpr == 0
  int l = 1;
  while (l * 2 <= g.lpr) l *= 2;
  g.lpr = l;
  g.rs = kThreads / g.lpr;
  g.ctiles = (c4 + g.lpr - 1) / g.lpr;
  int64_t cap = (int64_t)g.rs * kMaxRowsPerThread;
  int cpf0 = (int)((frame_rows + cap - 1) / cap);
  if (cpf0 < 1) cpf0 = 1;
  // A chunk is one CTA of the streaming kernels (K1 / K4 / K3), which run kClResident CTAs per SM: pick the chunks per
  // frame (between the register-budget minimum and twice that) whose CTA count fills whole waves of the 148 SMs best --
  // e.g. 64 channels at 56x56 x 128 frames: 7 chunks per frame = 896 CTAs = 1.51 waves, 9 chunks = 1152 = 1.95 waves.
  constexpr int64_t kClResident = 4, kSlots = 148 * kClResident;
  double best_eff = -1.0;
  for (int cand = cpf0; cand <= 2 * cpf0 + 1; ++cand) {
    int rpt = (int)((frame_rows + (int64_t)cand * g.rs - 1) / ((int64_t)cand * g.rs));
    if (rpt < 1) rpt = 1;
    if (cand > cpf0 && rpt < 4) break;   // keep at least 4 rows per thread: the per-CTA reduction epilogue is fixed cost
    const int64_t chunk = (int64_t)g.rs * rpt;
    const int64_t cpf = (frame_rows + chunk - 1) / chunk;
    const int64_t n = frames * cpf * g.ctiles;
    const double eff = (double)n / (double)(((n + kSlots - 1) / kSlots) * kSlots);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      g.rpt = rpt;
      g.chunk_rows = (int)chunk;
      g.cpf = (int)cpf;   // rounding rpt up can leave trailing chunks empty: cpf is recomputed from the chunk size
    }
  }
  g.frames = frames;
  g.frame_rows = frame_rows;
  return g;
}

// Chunking of an (O, C, I) tensor with I > 1: one warp reduces `og` consecutive o-slices of one channel.
struct OciGeom {
  int og;
  int n_entries;
};
static inline OciGeom oci_geom(int64_t O, int64_t I) {
  OciGeom g;
  int64_t og = (4096 + I - 1) / I;
  if (og < 1) og = 1;
  if (og > O) og = O;
  g.og = (int)og;
  g.n_entries = (int)((O + og - 1) / og);
  return g;
}

#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming load: read-once data, do not allocate in L1
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  // not volatile: a pure read-only load the compiler is free to hoist and batch (several in flight per thread)
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// 256-bit global accesses (sm_100): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void st8(float* p, float4 a, float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
               "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}
__device__ __forceinline__ void ld8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum_{b < n} p[b * stride] in index order (deterministic), with the loads issued sixteen at a time: the "last CTA adds the
// per-CTA partials" epilogues walk up to ~600 partials per output, and a plain `s += load` loop pays one L2 round trip per
// partial there (tens of microseconds at the tail of every launch).
__device__ __forceinline__ float ordered_sum_strided(const float* p, int n, int64_t stride) {
  float s = 0.f;
  for (int b0 = 0; b0 < n; b0 += 16) {
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = (b0 + u < n) ? __ldcg(p + (int64_t)(b0 + u) * stride) : 0.f;
#pragma unroll
    for (int u = 0; u < 16; ++u) s += v[u];
  }
  return s;
}

// Chan et al. pairwise merge of (n, mean, M2)
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float meanb, float m2b) {
  if (nb == 0.f) return;
  float nt = n + nb;
  float d = meanb - mean;
  float f = nb / nt;
  mean = fmaf(d, f, mean);
  m2 = m2 + m2b + d * d * n * f;
  n = nt;
}
#endif

}  // namespace vitta
