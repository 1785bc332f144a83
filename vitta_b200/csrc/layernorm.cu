// K9: LayerNorm over the channel axis of a (rows, C) token matrix, forward and backward, with the statistics hook
// fused in: the forward emits per-chunk per-channel (mean, M2) partials of its OUTPUT (Welford per warp, merged later by
// vitta_stats_finalize), the backward adds the closed-form hook gradient a_c + b_c (y - mu_c) while it already holds the
// row, adds the residual-branch gradient and accumulates d(gamma), d(beta).
// "merge" mode folds Video-Swin's PatchMerging gather (swin_transformer.py:293-310) into the load / scatter:
//   row (b, d, h2, w2), channel q*Cin + c  <-  x[b, d, 2*h2 + (q & 1), 2*w2 + (q >> 1), c]   (zero outside H x W).
// One warp per row, lanes along C with 128-bit accesses; a warp owns a chunk of consecutive rows (= one statistics entry).
#include "common.cuh"

namespace vitta {

constexpr int kLnThreads = 256;
constexpr int kLnWarps = kLnThreads / 32;
constexpr int kWsHeader = 64;   // floats reserved at the front of a workspace for the self-resetting tickets

struct LnGather {   // PatchMerging gather (merge != 0)
  int merge;
  int D, H, W, Cin;   // source tensor (B, D, H, W, Cin); rows index (b, d, ceil(H/2), ceil(W/2))
  int H2, W2;
};

struct LnFwdArgs {
  const float* x;
  float* y;
  const float* gamma;
  const float* beta;
  float* mean;    // [rows]
  float* rstd;    // [rows]
  float* part;    // [(entry*C + c)*2] or null
  int64_t rows;
  int C;
  int rows_per_warp;
  float eps;
  LnGather g;
  float* amax_y;   // optional: max|y| accumulated here (operand range of the fp16-split GEMM that consumes y)
};

__device__ __forceinline__ const float* ln_src_row(const float* x, int64_t r, int C, const LnGather& g, int64_t& base,
                                                   int& h0, int& w0) {
  if (!g.merge) {
    base = r * C;
    return x;
  }
  const int w2 = (int)(r % g.W2);
  int64_t t = r / g.W2;
  const int h2 = (int)(t % g.H2);
  t /= g.H2;   // = b*D + d
  h0 = 2 * h2;
  w0 = 2 * w2;
  base = ((t * g.H + h0) * g.W + w0) * (int64_t)g.Cin;
  return x;
}

template <int VPL, bool STATS>
__global__ void __launch_bounds__(kLnThreads) ln_fwd_kernel(const LnFwdArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int64_t r0 = chunk * p.rows_per_warp;
  if (r0 >= p.rows) return;
  const int64_t r1 = (r0 + p.rows_per_warp < p.rows) ? r0 + p.rows_per_warp : p.rows;
  const int C4 = p.C >> 2;
  const float invC = 1.f / (float)p.C;
  // merge mode: per-lane source offsets of each vector (row independent)
  int goff[VPL];
  int gdh[VPL], gdw[VPL];
  if (p.g.merge) {
    const int cin4 = p.g.Cin >> 2;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int i = j * 32 + lane;
      const int q = i / cin4, ci = i - q * cin4;
      gdh[j] = q & 1;
      gdw[j] = q >> 1;
      goff[j] = (gdh[j] * p.g.W + gdw[j]) * p.g.Cin + ci * 4;
    }
  }
  float4 wm[VPL], w2[VPL];   // Welford mean / M2 of y over this warp's rows
  if (STATS) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) wm[j] = w2[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float nrow = 0.f;
  float am = 0.f;
  // R rows per pass, stage by stage over all of them: loads, row sums, the two shuffle reductions, then the per-row
  // epilogue.  A row's critical path (load -> 5 shuffles -> 5 shuffles -> rsqrt -> store) is ~400 cycles; one row per pass
  // left the narrow stages (C = 96: 384 bytes per row and warp) bound by that chain at half of the copy bandwidth, whatever
  // the number of loads in flight.  Rows past the end are predicated (zero data), never branched around, so the R
  // reduction chains of a pass interleave.
  constexpr int R = VPL == 1 ? 4 : (VPL == 2 ? 2 : 1);
  for (int64_t rb = r0; rb < r1; rb += R) {
    float4 v[R][VPL];
    float s[R], q[R], mu[R], rs[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = rb + k;
      const bool live = r < r1;
      int64_t base = 0;
      int h0 = 0, w0 = 0;
      const float* src = ln_src_row(p.x, live ? r : r0, p.C, p.g, base, h0, w0);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * 32 + lane;
        v[k][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < C4 && live) {
          if (!p.g.merge) {
            v[k][j] = ld_stream4(src + base + (int64_t)i * 4);
          } else if (h0 + gdh[j] < p.g.H && w0 + gdw[j] < p.g.W) {
            v[k][j] = ld_stream4(src + base + goff[j]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      s[k] = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) s[k] += (v[k][j].x + v[k][j].y) + (v[k][j].z + v[k][j].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < R; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      mu[k] = s[k] * invC;
      q[k] = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * 32 + lane;
        if (i < C4) {
          const float a = v[k][j].x - mu[k], b = v[k][j].y - mu[k], c = v[k][j].z - mu[k], d = v[k][j].w - mu[k];
          q[k] += (a * a + b * b) + (c * c + d * d);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < R; ++k) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
    }
#pragma unroll
    for (int k = 0; k < R; ++k) rs[k] = 1.f / sqrtf(q[k] * invC + p.eps);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = rb + k;
      if (r < r1) {
        if (lane == 0) {
          p.mean[r] = mu[k];
          p.rstd[r] = rs[k];
        }
        nrow += 1.f;
        const float inv_n = 1.f / nrow;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int i = j * 32 + lane;
          if (i < C4) {
            const float4 ga = ldg4(p.gamma + i * 4), be = ldg4(p.beta + i * 4);
            float4 y;
            y.x = fmaf((v[k][j].x - mu[k]) * rs[k], ga.x, be.x);
            y.y = fmaf((v[k][j].y - mu[k]) * rs[k], ga.y, be.y);
            y.z = fmaf((v[k][j].z - mu[k]) * rs[k], ga.z, be.z);
            y.w = fmaf((v[k][j].w - mu[k]) * rs[k], ga.w, be.w);
            st4(p.y + r * p.C + (int64_t)i * 4, y);
            am = fmaxf(am, fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w))));
            if (STATS) {
              float d;
              d = y.x - wm[j].x; wm[j].x = fmaf(d, inv_n, wm[j].x); w2[j].x = fmaf(d, y.x - wm[j].x, w2[j].x);
              d = y.y - wm[j].y; wm[j].y = fmaf(d, inv_n, wm[j].y); w2[j].y = fmaf(d, y.y - wm[j].y, w2[j].y);
              d = y.z - wm[j].z; wm[j].z = fmaf(d, inv_n, wm[j].z); w2[j].z = fmaf(d, y.z - wm[j].z, w2[j].z);
              d = y.w - wm[j].w; wm[j].w = fmaf(d, inv_n, wm[j].w); w2[j].w = fmaf(d, y.w - wm[j].w, w2[j].w);
            }
          }
        }
      }
    }
  }
  if (p.amax_y) {   // one integer atomic per warp (non-negative floats order like their bit patterns)
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
    if (lane == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(p.amax_y), wmax);
  }
  if (STATS) {
    float* o = p.part + chunk * p.C * 2;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int i = j * 32 + lane;
      if (i < C4) {
        st4(o + (int64_t)i * 8, make_float4(wm[j].x, w2[j].x, wm[j].y, w2[j].y));
        st4(o + (int64_t)i * 8 + 4, make_float4(wm[j].z, w2[j].z, wm[j].w, w2[j].w));
      }
    }
  }
}

// Narrow rows (C <= 256, plain mode): LPR lanes per row, 32 / LPR rows per warp pass.  ncu on the one-warp-per-row kernel at
// C = 96 (Video-Swin-T stage 1, 308 MB per tensor): 63 % issue-slot utilisation at a third of the warp slots and 41 % of
// the DRAM bandwidth -- ~150 warp instructions per 384-byte row, so the kernel is bound by instruction issue, not by loads
// in flight.  Here one instruction stream serves 4 (C <= 128) or 2 rows, all 32 lanes hold data (24 of 32 did at C = 96),
// and the row reductions take log2(LPR) shuffle steps.  Per-column Welford statistics are kept per lane over the rows of
// its group and merged across the groups (Chan) once per chunk.
template <int LPR, int VPL, bool STATS>
__global__ void __launch_bounds__(kLnThreads) ln_fwd_narrow_kernel(const LnFwdArgs p) {
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, grp = lane / LPR;
  const int64_t chunk = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int64_t r0 = chunk * p.rows_per_warp;
  if (r0 >= p.rows) return;
  const int64_t r1 = (r0 + p.rows_per_warp < p.rows) ? r0 + p.rows_per_warp : p.rows;
  const int C4 = p.C >> 2;
  const float invC = 1.f / (float)p.C;
  float4 ga[VPL], be[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int i = j * LPR + sub;
    ga[j] = be[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < C4) {
      ga[j] = ldg4(p.gamma + i * 4);
      be[j] = ldg4(p.beta + i * 4);
    }
  }
  float4 wm[VPL], w2[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) wm[j] = w2[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float nrow = 0.f, am = 0.f;
  for (int64_t rb = r0; rb < r1; rb += G) {      // warp-uniform trip count: the shuffles below see all 32 lanes
    const int64_t r = rb + grp;
    const bool live = r < r1;
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int i = j * LPR + sub;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < C4 && live) v[j] = ld_stream4(p.x + r * p.C + (int64_t)i * 4);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s * invC;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int i = j * LPR + sub;
      if (i < C4) {
        const float a = v[j].x - mu, b = v[j].y - mu, c = v[j].z - mu, d = v[j].w - mu;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = 1.f / sqrtf(q * invC + p.eps);
    if (live) {
      if (sub == 0) {
        p.mean[r] = mu;
        p.rstd[r] = rs;
      }
      nrow += 1.f;
      const float inv_n = 1.f / nrow;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * LPR + sub;
        if (i < C4) {
          float4 y;
          y.x = fmaf((v[j].x - mu) * rs, ga[j].x, be[j].x);
          y.y = fmaf((v[j].y - mu) * rs, ga[j].y, be[j].y);
          y.z = fmaf((v[j].z - mu) * rs, ga[j].z, be[j].z);
          y.w = fmaf((v[j].w - mu) * rs, ga[j].w, be[j].w);
          st4(p.y + r * p.C + (int64_t)i * 4, y);
          am = fmaxf(am, fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w))));
          if (STATS) {
            float d;
            d = y.x - wm[j].x; wm[j].x = fmaf(d, inv_n, wm[j].x); w2[j].x = fmaf(d, y.x - wm[j].x, w2[j].x);
            d = y.y - wm[j].y; wm[j].y = fmaf(d, inv_n, wm[j].y); w2[j].y = fmaf(d, y.y - wm[j].y, w2[j].y);
            d = y.z - wm[j].z; wm[j].z = fmaf(d, inv_n, wm[j].z); w2[j].z = fmaf(d, y.z - wm[j].z, w2[j].z);
            d = y.w - wm[j].w; wm[j].w = fmaf(d, inv_n, wm[j].w); w2[j].w = fmaf(d, y.w - wm[j].w, w2[j].w);
          }
        }
      }
    }
  }
  if (p.amax_y) {
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
    if (lane == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(p.amax_y), wmax);
  }
  if (STATS) {
    // the G groups' (n, mean, M2) per column -> one: pairwise Chan merges in a fixed order
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      const float nb = __shfl_xor_sync(0xffffffffu, nrow, o);
      const float nt = nrow + nb;
      const float f = nt > 0.f ? nb / nt : 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        float* m = reinterpret_cast<float*>(&wm[j]);
        float* m2 = reinterpret_cast<float*>(&w2[j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float mb = __shfl_xor_sync(0xffffffffu, m[e], o);
          const float m2b = __shfl_xor_sync(0xffffffffu, m2[e], o);
          const float d = mb - m[e];
          m2[e] = m2[e] + m2b + d * d * nrow * f;
          m[e] = fmaf(d, f, m[e]);
        }
      }
      nrow = nt;
    }
    if (grp == 0) {
      float* o = p.part + chunk * p.C * 2;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * LPR + sub;
        if (i < C4) {
          st4(o + (int64_t)i * 8, make_float4(wm[j].x, w2[j].x, wm[j].y, w2[j].y));
          st4(o + (int64_t)i * 8 + 4, make_float4(wm[j].z, w2[j].z, wm[j].w, w2[j].w));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct LnBwdArgs {
  const float* gy;      // gradient w.r.t. the LayerNorm output (rows, C)
  const float* x;       // forward input (plain: (rows, C); merge: source tensor)
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* rstd;
  const float* gadd;    // optional gradient added to gx (residual branch), plain mode only
  const float *ca, *cb, *cm, *gs;   // hook coefficients (per channel) + gradient of the layer's r_feature; null: none
  float* gx;            // plain: (rows, C); merge: scattered into the source layout (B, D, H, W, Cin)
  float* dgamma;        // accumulated (+=)
  float* dbeta;
  float* ws;            // [grid][2][C] partials + 1 int ticket, zero-initialised once (self-resetting)
  int64_t rows;
  int C;
  LnGather g;
  float* amax_gx;   // optional: max|gx| (plain mode), the range of the GEMMs that take gx as their gradient operand
};

// R rows per warp iteration: ALL their loads (x, gy, the shortcut gradient, mean / rstd) are issued before the first
// use.  With one row at a time a warp of the narrow stages (C = 96 .. 384: one to three 16-byte loads per lane and operand)
// kept ~3 loads in flight and the kernel streamed at 1.4-1.8 TB/s (profiles/r01, r02); the rows are still visited, and
// their contributions to d(gamma) / d(beta) added, in the same order as before.
template <int VPL>
struct LnBwdRows { static constexpr int value = VPL <= 1 ? 4 : 2; };   // instantiated for VPL <= 2 (C <= 256) only

template <int VPL>
__global__ void __launch_bounds__(kLnThreads, 2) ln_bwd_rows_kernel(const LnBwdArgs p) {
  constexpr int R = LnBwdRows<VPL>::value;
  __shared__ float4 s_red[2 * 512];   // [2][C/4], C <= 2048
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int C4 = p.C >> 2;
  const float invC = 1.f / (float)p.C;
  int goff[VPL];
  int gdh[VPL], gdw[VPL];
  if (p.g.merge) {
    const int cin4 = p.g.Cin >> 2;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int i = j * 32 + lane;
      const int q = i / cin4, ci = i - q * cin4;
      gdh[j] = q & 1;
      gdw[j] = q >> 1;
      goff[j] = (gdh[j] * p.g.W + gdw[j]) * p.g.Cin + ci * 4;
    }
  }
  const float gsc = p.ca ? __ldg(p.gs) : 0.f;
  float4 dg[VPL], db[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) dg[j] = db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float am = 0.f;

  const int64_t stride = (int64_t)gridDim.x * kLnWarps;
  for (int64_t rb = (int64_t)blockIdx.x * kLnWarps + warp; rb < p.rows; rb += stride * R) {
    float4 xv[R][VPL], gyv[R][VPL], gav[R][VPL];
    float mu[R], rs[R];
    int64_t base[R];
    int h0[R], w0[R];
    // ---- phase 1: every load of the R rows
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = rb + k * stride;
      h0[k] = w0[k] = 0;
      base[k] = 0;
      mu[k] = 0.f; rs[k] = 0.f;
      if (r < p.rows) {
        const float* src = ln_src_row(p.x, r, p.C, p.g, base[k], h0[k], w0[k]);
        mu[k] = __ldg(p.mean + r);
        rs[k] = __ldg(p.rstd + r);
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int i = j * 32 + lane;
          xv[k][j] = gyv[k][j] = gav[k][j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < C4) {
            if (!p.g.merge) {
              xv[k][j] = ld_stream4(src + base[k] + (int64_t)i * 4);
              if (p.gadd) gav[k][j] = ld_stream4(p.gadd + r * p.C + (int64_t)i * 4);
            } else if (h0[k] + gdh[j] < p.g.H && w0[k] + gdw[j] < p.g.W) {
              xv[k][j] = ld_stream4(src + base[k] + goff[j]);
            }
            gyv[k][j] = ld_stream4(p.gy + r * p.C + (int64_t)i * 4);
          }
        }
      }
    }
    // ---- phase 2: the rows one after the other (same arithmetic and order as a one-row-at-a-time loop)
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = rb + k * stride;
      if (r >= p.rows) break;
      float4 xh[VPL], g[VPL];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * 32 + lane;
        xh[j] = g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < C4) {
          const float4 v = xv[k][j];
          float4 gy = gyv[k][j];
          const float4 ga = ldg4(p.gamma + i * 4);
          xh[j] = make_float4((v.x - mu[k]) * rs[k], (v.y - mu[k]) * rs[k], (v.z - mu[k]) * rs[k], (v.w - mu[k]) * rs[k]);
          if (p.ca) {
            const float4 be = ldg4(p.beta + i * 4);
            const float4 a = ldg4(p.ca + i * 4), b = ldg4(p.cb + i * 4), m = ldg4(p.cm + i * 4);
            gy.x = fmaf(gsc, fmaf(b.x, fmaf(xh[j].x, ga.x, be.x) - m.x, a.x), gy.x);
            gy.y = fmaf(gsc, fmaf(b.y, fmaf(xh[j].y, ga.y, be.y) - m.y, a.y), gy.y);
            gy.z = fmaf(gsc, fmaf(b.z, fmaf(xh[j].z, ga.z, be.z) - m.z, a.z), gy.z);
            gy.w = fmaf(gsc, fmaf(b.w, fmaf(xh[j].w, ga.w, be.w) - m.w, a.w), gy.w);
          }
          db[j].x += gy.x; db[j].y += gy.y; db[j].z += gy.z; db[j].w += gy.w;
          dg[j].x = fmaf(gy.x, xh[j].x, dg[j].x); dg[j].y = fmaf(gy.y, xh[j].y, dg[j].y);
          dg[j].z = fmaf(gy.z, xh[j].z, dg[j].z); dg[j].w = fmaf(gy.w, xh[j].w, dg[j].w);
          g[j] = make_float4(gy.x * ga.x, gy.y * ga.y, gy.z * ga.z, gy.w * ga.w);
          s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
          s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
        }
      }
      const float c1 = warp_sum(s1) * invC, c2 = warp_sum(s2) * invC;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * 32 + lane;
        if (i < C4) {
          float4 o;
          o.x = rs[k] * (g[j].x - c1 - xh[j].x * c2);
          o.y = rs[k] * (g[j].y - c1 - xh[j].y * c2);
          o.z = rs[k] * (g[j].z - c1 - xh[j].z * c2);
          o.w = rs[k] * (g[j].w - c1 - xh[j].w * c2);
          if (!p.g.merge) {
            if (p.gadd) {
              const float4 a = gav[k][j];
              o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            st4(p.gx + r * p.C + (int64_t)i * 4, o);
            am = fmaxf(am, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
          } else if (h0[k] + gdh[j] < p.g.H && w0[k] + gdw[j] < p.g.W) {
            st4(p.gx + base[k] + goff[j], o);
          }
        }
      }
    }
  }
  if (p.amax_gx) {
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
    if (lane == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(p.amax_gx), wmax);
  }
  // CTA reduction of d(gamma), d(beta): warps add in turn (fixed order), then one partial per CTA
  for (int w = 0; w < kLnWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * 32 + lane;
        if (i < C4) {
          if (w == 0) {
            s_red[i] = dg[j];
            s_red[512 + i] = db[j];
          } else {
            float4 a = s_red[i], b = s_red[512 + i];
            a.x += dg[j].x; a.y += dg[j].y; a.z += dg[j].z; a.w += dg[j].w;
            b.x += db[j].x; b.y += db[j].y; b.z += db[j].z; b.w += db[j].w;
            s_red[i] = a;
            s_red[512 + i] = b;
          }
        }
      }
    }
    __syncthreads();
  }
  float* wsp = p.ws + kWsHeader;   // ws = [ticket ints | per-CTA partials]
  float* wsb = wsp + (int64_t)blockIdx.x * 2 * p.C;
  for (int i = threadIdx.x; i < C4; i += kLnThreads) {
    st4(wsb + (int64_t)i * 4, s_red[i]);
    st4(wsb + p.C + (int64_t)i * 4, s_red[512 + i]);
  }
  // the per-CTA partials are added by ln_bwd_finish_kernel (next launch on the stream)
}

// Narrow rows (64 < C <= 192, plain mode), the backward counterpart of ln_fwd_narrow_kernel: LPR lanes per row, G = 32 / LPR
// rows per warp pass, all loads of a pass issued before the first use.  Same issue-bound picture as the forward: the
// one-warp-per-row kernel spends a full instruction stream (and two 5-step shuffle reductions) on a 384-byte row.  gamma
// stays in registers; the hook coefficients (rare on these stages) are read on use.  d(gamma) / d(beta): per lane over its
// rows, then across the warp's row groups by shuffles, then the warps of the CTA in turn -- a fixed order.
template <int LPR, int VPL>
__global__ void __launch_bounds__(kLnThreads, 2) ln_bwd_narrow_kernel(const LnBwdArgs p) {
  constexpr int G = 32 / LPR, R = 1;   // (two passes in flight spill at the 128 registers two resident CTAs allow; one
                                       //  pass is 9-12 loads of 16 bytes per lane: ~70 KB in flight per SM)
  __shared__ float4 s_red[2 * 64];   // [2][C/4], C <= 192
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  const int C4 = p.C >> 2;
  const float invC = 1.f / (float)p.C;
  const float gsc = p.ca ? __ldg(p.gs) : 0.f;
  float4 ga[VPL], dg[VPL], db[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int i = j * LPR + sub;
    dg[j] = db[j] = ga[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < C4) ga[j] = ldg4(p.gamma + i * 4);
  }
  float am = 0.f;
  const int64_t nquads = (p.rows + G - 1) / G;
  const int64_t stride = (int64_t)gridDim.x * kLnWarps;
  for (int64_t qb = (int64_t)blockIdx.x * kLnWarps + warp; qb < nquads; qb += stride * R) {   // warp-uniform trip count
    float4 xv[R][VPL], gyv[R][VPL], gav[R][VPL];
    float mu[R], rs[R];
    // ---- every load of the pass
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = (qb + k * stride) * G + grp;
      const bool live = (qb + k * stride) < nquads && r < p.rows;
      mu[k] = 0.f; rs[k] = 0.f;
      if (live) {
        mu[k] = __ldg(p.mean + r);
        rs[k] = __ldg(p.rstd + r);
      }
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * LPR + sub;
        xv[k][j] = gyv[k][j] = gav[k][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < C4 && live) {
          xv[k][j] = ld_stream4(p.x + r * p.C + (int64_t)i * 4);
          if (p.gadd) gav[k][j] = ld_stream4(p.gadd + r * p.C + (int64_t)i * 4);
          gyv[k][j] = ld_stream4(p.gy + r * p.C + (int64_t)i * 4);
        }
      }
    }
    // ---- dead rows carry zeros (gy = 0: no contribution) and store nothing
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = (qb + k * stride) * G + grp;
      const bool live = (qb + k * stride) < nquads && r < p.rows;
      float4 xh[VPL], g[VPL];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * LPR + sub;
        xh[j] = g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < C4 && live) {
          const float4 v = xv[k][j];
          float4 gy = gyv[k][j];
          xh[j] = make_float4((v.x - mu[k]) * rs[k], (v.y - mu[k]) * rs[k], (v.z - mu[k]) * rs[k], (v.w - mu[k]) * rs[k]);
          if (p.ca) {
            const float4 be = ldg4(p.beta + i * 4);
            const float4 a = ldg4(p.ca + i * 4), b = ldg4(p.cb + i * 4), m = ldg4(p.cm + i * 4);
            gy.x = fmaf(gsc, fmaf(b.x, fmaf(xh[j].x, ga[j].x, be.x) - m.x, a.x), gy.x);
            gy.y = fmaf(gsc, fmaf(b.y, fmaf(xh[j].y, ga[j].y, be.y) - m.y, a.y), gy.y);
            gy.z = fmaf(gsc, fmaf(b.z, fmaf(xh[j].z, ga[j].z, be.z) - m.z, a.z), gy.z);
            gy.w = fmaf(gsc, fmaf(b.w, fmaf(xh[j].w, ga[j].w, be.w) - m.w, a.w), gy.w);
          }
          db[j].x += gy.x; db[j].y += gy.y; db[j].z += gy.z; db[j].w += gy.w;
          dg[j].x = fmaf(gy.x, xh[j].x, dg[j].x); dg[j].y = fmaf(gy.y, xh[j].y, dg[j].y);
          dg[j].z = fmaf(gy.z, xh[j].z, dg[j].z); dg[j].w = fmaf(gy.w, xh[j].w, dg[j].w);
          g[j] = make_float4(gy.x * ga[j].x, gy.y * ga[j].y, gy.z * ga[j].z, gy.w * ga[j].w);
          s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
          s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const float c1 = s1 * invC, c2 = s2 * invC;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * LPR + sub;
        if (i < C4 && live) {
          float4 o;
          o.x = rs[k] * (g[j].x - c1 - xh[j].x * c2);
          o.y = rs[k] * (g[j].y - c1 - xh[j].y * c2);
          o.z = rs[k] * (g[j].z - c1 - xh[j].z * c2);
          o.w = rs[k] * (g[j].w - c1 - xh[j].w * c2);
          if (p.gadd) {
            const float4 a = gav[k][j];
            o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
          }
          st4(p.gx + r * p.C + (int64_t)i * 4, o);
          am = fmaxf(am, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
        }
      }
    }
  }
  if (p.amax_gx) {
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
    if (lane == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(p.amax_gx), wmax);
  }
  // the warp's row groups -> group 0 (fixed order), then the warps of the CTA in turn, then one partial per CTA
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      dg[j].x += __shfl_xor_sync(0xffffffffu, dg[j].x, o); dg[j].y += __shfl_xor_sync(0xffffffffu, dg[j].y, o);
      dg[j].z += __shfl_xor_sync(0xffffffffu, dg[j].z, o); dg[j].w += __shfl_xor_sync(0xffffffffu, dg[j].w, o);
      db[j].x += __shfl_xor_sync(0xffffffffu, db[j].x, o); db[j].y += __shfl_xor_sync(0xffffffffu, db[j].y, o);
      db[j].z += __shfl_xor_sync(0xffffffffu, db[j].z, o); db[j].w += __shfl_xor_sync(0xffffffffu, db[j].w, o);
    }
  }
  for (int w = 0; w < kLnWarps; ++w) {
    if (warp == w && grp == 0) {
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * LPR + sub;
        if (i < C4) {
          if (w == 0) {
            s_red[i] = dg[j];
            s_red[64 + i] = db[j];
          } else {
            float4 a = s_red[i], b = s_red[64 + i];
            a.x += dg[j].x; a.y += dg[j].y; a.z += dg[j].z; a.w += dg[j].w;
            b.x += db[j].x; b.y += db[j].y; b.z += db[j].z; b.w += db[j].w;
            s_red[i] = a;
            s_red[64 + i] = b;
          }
        }
      }
    }
    __syncthreads();
  }
  float* wsb = p.ws + kWsHeader + (int64_t)blockIdx.x * 2 * p.C;
  for (int i = threadIdx.x; i < C4; i += kLnThreads) {
    st4(wsb + (int64_t)i * 4, s_red[i]);
    st4(wsb + p.C + (int64_t)i * 4, s_red[64 + i]);
  }
  // the per-CTA partials are added by ln_bwd_finish_kernel (next launch on the stream)
}

// Wide rows (C > 256): a row is split over K = 1, 2, 4 or 8 warps of the CTA, 96 float4 (three per lane) each, and every
// warp iteration works on TWO rows with all their loads in flight.  The one-row-per-warp kernel above kept the whole row
// plus its d(gamma) / d(beta) accumulators in one warp's registers: 136 registers at C = 384, 218 at C = 768, 255 + spills
// at C = 1536 -- ONE resident CTA of 8 warps per SM, one row in flight per warp, 1.3-2.0 TB/s (tools/ln_probe.py).  Here
// every thread carries 3 float4 of accumulators whatever C is; the K partial row sums meet in shared memory behind a
// named barrier of the row's warps (double-buffered by iteration parity).  The reduction order of d(gamma) / d(beta) is
// fixed (rows in visiting order per warp, then the row slots of the CTA in order, then the CTAs in order).
constexpr int kLnSplitV = 3;                  // float4 per lane and warp
constexpr int kLnSplitCols = 32 * kLnSplitV;  // float4 per warp = 384 channels

template <int K>
__global__ void __launch_bounds__(kLnThreads, 2) ln_bwd_split_kernel(const LnBwdArgs p) {
  constexpr int V = kLnSplitV, R = 2, kSlots = kLnWarps / K;
  __shared__ float4 s_red[2 * 512];          // [2][C/4], C <= 2048
  __shared__ float2 s_part[2][R][kLnWarps];  // [iteration parity][row][warp]: partial (sum g, sum g * xhat)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = warp / K, part = warp % K;
  const int C4 = p.C >> 2;
  const float invC = 1.f / (float)p.C;
  int col[V];        // float4 index of this lane's j-th vector within the row
  int goff[V], gdh[V], gdw[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    col[j] = part * kLnSplitCols + j * 32 + lane;
    goff[j] = gdh[j] = gdw[j] = 0;
    if (p.g.merge && col[j] < C4) {
      const int cin4 = p.g.Cin >> 2;
      const int q = col[j] / cin4, ci = col[j] - q * cin4;
      gdh[j] = q & 1;
      gdw[j] = q >> 1;
      goff[j] = (gdh[j] * p.g.W + gdw[j]) * p.g.Cin + ci * 4;
    }
  }
  const float gsc = p.ca ? __ldg(p.gs) : 0.f;
  // gamma of this lane's columns stays in registers; the hook's per-channel constants are re-read where they are used
  // (L1-resident) -- holding all five vectors cost 60 registers and spilled under the two-CTA bound
  float4 ga[V];
#pragma unroll
  for (int j = 0; j < V; ++j) ga[j] = col[j] < C4 ? ldg4(p.gamma + col[j] * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 dg[V], db[V];
#pragma unroll
  for (int j = 0; j < V; ++j) dg[j] = db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float am = 0.f;

  const int64_t stride = (int64_t)gridDim.x * kSlots;
  uint32_t iter = 0;
  // every warp of a row slot runs the same number of iterations (rb depends on the slot only): the named barrier is safe
  for (int64_t rb = (int64_t)blockIdx.x * kSlots + slot; rb < p.rows; rb += stride * R, ++iter) {
    float4 xv[R][V], gyv[R][V];
    float mu[R], rs[R];
    int64_t base[R];
    int h0[R], w0[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = rb + k * stride;
      h0[k] = w0[k] = 0;
      base[k] = 0;
      mu[k] = rs[k] = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) xv[k][j] = gyv[k][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < p.rows) {
        const float* src = ln_src_row(p.x, r, p.C, p.g, base[k], h0[k], w0[k]);
        mu[k] = __ldg(p.mean + r);
        rs[k] = __ldg(p.rstd + r);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (col[j] < C4) {
            if (!p.g.merge) {
              xv[k][j] = ld_stream4(src + base[k] + (int64_t)col[j] * 4);
            } else if (h0[k] + gdh[j] < p.g.H && w0[k] + gdw[j] < p.g.W) {
              xv[k][j] = ld_stream4(src + base[k] + goff[j]);
            }
            gyv[k][j] = ld_stream4(p.gy + r * p.C + (int64_t)col[j] * 4);
          }
        }
      }
    }
    float4 xh[R][V], g[R][V];
    float s1[R], s2[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      s1[k] = s2[k] = 0.f;
      const bool live = rb + k * stride < p.rows;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        xh[k][j] = g[k][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && col[j] < C4) {
          const float4 v = xv[k][j];
          float4 gy = gyv[k][j];
          xh[k][j] = make_float4((v.x - mu[k]) * rs[k], (v.y - mu[k]) * rs[k], (v.z - mu[k]) * rs[k], (v.w - mu[k]) * rs[k]);
          if (p.ca) {
            const float4 be = ldg4(p.beta + col[j] * 4);
            const float4 a = ldg4(p.ca + col[j] * 4), b = ldg4(p.cb + col[j] * 4), m = ldg4(p.cm + col[j] * 4);
            gy.x = fmaf(gsc, fmaf(b.x, fmaf(xh[k][j].x, ga[j].x, be.x) - m.x, a.x), gy.x);
            gy.y = fmaf(gsc, fmaf(b.y, fmaf(xh[k][j].y, ga[j].y, be.y) - m.y, a.y), gy.y);
            gy.z = fmaf(gsc, fmaf(b.z, fmaf(xh[k][j].z, ga[j].z, be.z) - m.z, a.z), gy.z);
            gy.w = fmaf(gsc, fmaf(b.w, fmaf(xh[k][j].w, ga[j].w, be.w) - m.w, a.w), gy.w);
          }
          db[j].x += gy.x; db[j].y += gy.y; db[j].z += gy.z; db[j].w += gy.w;
          dg[j].x = fmaf(gy.x, xh[k][j].x, dg[j].x); dg[j].y = fmaf(gy.y, xh[k][j].y, dg[j].y);
          dg[j].z = fmaf(gy.z, xh[k][j].z, dg[j].z); dg[j].w = fmaf(gy.w, xh[k][j].w, dg[j].w);
          g[k][j] = make_float4(gy.x * ga[j].x, gy.y * ga[j].y, gy.z * ga[j].z, gy.w * ga[j].w);
          s1[k] += (g[k][j].x + g[k][j].y) + (g[k][j].z + g[k][j].w);
          s2[k] += (g[k][j].x * xh[k][j].x + g[k][j].y * xh[k][j].y) + (g[k][j].z * xh[k][j].z + g[k][j].w * xh[k][j].w);
        }
      }
      s1[k] = warp_sum(s1[k]);
      s2[k] = warp_sum(s2[k]);
    }
    if (K > 1) {
      const int par = iter & 1;
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < R; ++k) s_part[par][k][warp] = make_float2(s1[k], s2[k]);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(32 * K) : "memory");
#pragma unroll
      for (int k = 0; k < R; ++k) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int q = 0; q < K; ++q) {
          const float2 v = s_part[par][k][slot * K + q];
          a += v.x;
          b += v.y;
        }
        s1[k] = a;
        s2[k] = b;
      }
    }
    // the shortcut gradient is fetched only now (both rows' loads together): holding it from the first phase on cost 24
    // registers and spilled under the two-CTA bound
    float4 gav[R][V];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = rb + k * stride;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        gav[k][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.gadd && r < p.rows && col[j] < C4) gav[k][j] = ld_stream4(p.gadd + r * p.C + (int64_t)col[j] * 4);
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t r = rb + k * stride;
      if (r >= p.rows) break;
      const float c1 = s1[k] * invC, c2 = s2[k] * invC;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (col[j] < C4) {
          float4 o;
          o.x = rs[k] * (g[k][j].x - c1 - xh[k][j].x * c2) + gav[k][j].x;
          o.y = rs[k] * (g[k][j].y - c1 - xh[k][j].y * c2) + gav[k][j].y;
          o.z = rs[k] * (g[k][j].z - c1 - xh[k][j].z * c2) + gav[k][j].z;
          o.w = rs[k] * (g[k][j].w - c1 - xh[k][j].w * c2) + gav[k][j].w;
          if (!p.g.merge) {
            st4(p.gx + r * p.C + (int64_t)col[j] * 4, o);
            am = fmaxf(am, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
          } else if (h0[k] + gdh[j] < p.g.H && w0[k] + gdw[j] < p.g.W) {
            st4(p.gx + base[k] + goff[j], o);
          }
        }
      }
    }
  }
  if (p.amax_gx) {
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
    if (lane == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(p.amax_gx), wmax);
  }
  // CTA reduction of d(gamma), d(beta): the row slots add in turn (fixed order), then one partial per CTA
  for (int sl = 0; sl < kSlots; ++sl) {
    if (slot == sl) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const int i = col[j];
        if (i < C4) {
          if (sl == 0) {
            s_red[i] = dg[j];
            s_red[512 + i] = db[j];
          } else {
            float4 a = s_red[i], b = s_red[512 + i];
            a.x += dg[j].x; a.y += dg[j].y; a.z += dg[j].z; a.w += dg[j].w;
            b.x += db[j].x; b.y += db[j].y; b.z += db[j].z; b.w += db[j].w;
            s_red[i] = a;
            s_red[512 + i] = b;
          }
        }
      }
    }
    __syncthreads();
  }
  float* wsb = p.ws + kWsHeader + (int64_t)blockIdx.x * 2 * p.C;
  for (int i = threadIdx.x; i < C4; i += kLnThreads) {
    st4(wsb + (int64_t)i * 4, s_red[i]);
    st4(wsb + p.C + (int64_t)i * 4, s_red[512 + i]);
  }
  // the per-CTA partials are added by ln_bwd_finish_kernel (next launch on the stream)
}

// d(gamma)[c] += sum over the CTAs' partials, d(beta) likewise, in CTA order.  CTA = 32 columns x 8 slices of the CTA
// range (sixteen loads in flight per thread), the slices added in slice order through shared memory.  (The backward
// kernels used to elect their last CTA for this: 256 threads walking up to 592 partials of 2C columns one L2 round trip
// after the other -- 100-400 us at the tail of every launch with C >= 384, several times the streaming part.)
__global__ void __launch_bounds__(256) ln_bwd_finish_kernel(const float* __restrict__ wsp, int n_ctas, int C,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float sp[8][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int per = (n_ctas + 7) / 8;
  const int k0 = slice * per, k1 = min(n_ctas, k0 + per);
  sp[slice][lane] = (i < 2 * C && k1 > k0) ? ordered_sum_strided(wsp + (int64_t)k0 * 2 * C + i, k1 - k0, 2 * (int64_t)C) : 0.f;
  __syncthreads();
  if (slice == 0 && i < 2 * C) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += sp[q][lane];
    if (i < C) {
      if (dgamma) dgamma[i] += s;
    } else {
      if (dbeta) dbeta[i - C] += s;
    }
  }
}

static int ln_vpl(int C) {
  const int v = (C / 4 + 31) / 32;
  return v <= 1 ? 1 : v <= 2 ? 2 : v <= 4 ? 4 : v <= 8 ? 8 : 16;
}
static int ln_bwd_grid(int64_t rows) {
  int64_t g = (rows + kLnWarps - 1) / kLnWarps;
  const int64_t cap = 148 * 4;
  return (int)(g < cap ? g : cap);
}

// ------------------------------------------------------------------------------------------------
// column sums / per-frame column means
// ------------------------------------------------------------------------------------------------
// out[c] (+)= sum_r x[r, c]; two-stage, fixed order.  grid = (row chunks, channel tiles of 128)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                    float* __restrict__ ws, int64_t rows, int C, int accumulate) {
  __shared__ float4 sm[256];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
  const int c4 = blockIdx.y * 32 + lane;
  const bool active = c4 * 4 < C;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    // eight rows per batch, loads first (a one-load-per-iteration loop keeps a single 16-byte load in flight per thread)
    const int64_t rstep = (int64_t)gridDim.x * 8;
    for (int64_t r = (int64_t)blockIdx.x * 8 + slot; r < rows; r += rstep * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r + u * rstep < rows) v[u] = ld_stream4(x + (r + u * rstep) * C + (int64_t)c4 * 4);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int st = 4; st > 0; st >>= 1) {
    if (slot < st) {
      float4 a = sm[threadIdx.x], b = sm[threadIdx.x + st * 32];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      sm[threadIdx.x] = a;
    }
    __syncthreads();
  }
  float* wsp = ws + kWsHeader;
  if (slot == 0 && active) st4(wsp + (int64_t)blockIdx.x * C + (int64_t)c4 * 4, sm[threadIdx.x]);
  __threadfence();
  __syncthreads();
  int* tickets = reinterpret_cast<int*>(ws);
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets + blockIdx.y, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < 128; i += 256) {
    const int c = blockIdx.y * 128 + i;
    if (c >= C) continue;
    const float t = ordered_sum_strided(wsp + c, (int)gridDim.x, (int64_t)C);
    out[c] = accumulate ? out[c] + t : t;
  }
  if (threadIdx.x == 0) tickets[blockIdx.y] = 0;
}

// out[f, c] = mean over the `rows` rows of frame f.  grid = (frames, channel tiles of 128)
__global__ void __launch_bounds__(256) frame_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int rows,
                                                        int C) {
  __shared__ float4 sm[256];
  const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
  const int c4 = blockIdx.y * 32 + lane;
  const bool active = c4 * 4 < C;
  const float* base = x + (int64_t)blockIdx.x * rows * C;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    for (int r = slot; r < rows; r += 8) {
      const float4 v = ld_stream4(base + (int64_t)r * C + (int64_t)c4 * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int st = 4; st > 0; st >>= 1) {
    if (slot < st) {
      float4 a = sm[threadIdx.x], b = sm[threadIdx.x + st * 32];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      sm[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (slot == 0 && active) {
    const float inv = 1.f / (float)rows;
    float4 a = sm[threadIdx.x];
    st4(out + (int64_t)blockIdx.x * C + (int64_t)c4 * 4, make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv));
  }
}

// gx[f, r, c] = g[f, c] / rows
__global__ void __launch_bounds__(256) frame_mean_bwd_kernel(const float* __restrict__ g, float* __restrict__ gx,
                                                            int64_t n4, int rows, int C4) {
  const float inv = 1.f / (float)rows;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const int c4 = (int)(i % C4);
    const int64_t f = i / ((int64_t)C4 * rows);
    const float4 v = ldg4(g + (f * C4 + c4) * 4);
    st4(gx + i * 4, make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv));
  }
}

// out[r, :] = x[r, :] * scale[r / rows_per_group]   (DropPath applied to a gradient before the weight-gradient GEMM)
__global__ void __launch_bounds__(256) row_scale_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                       float* __restrict__ out, int64_t n4, int C4, int64_t rpg,
                                                       float* __restrict__ amax_out) {
  float am = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float s = __ldg(scale + (i / C4) / rpg);
    const float4 v = ld_stream4(x + i * 4);
    const float4 o = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
    st4(out + i * 4, o);
    am = fmaxf(am, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
  }
  if (amax_out) {
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
    if ((threadIdx.x & 31) == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(amax_out), wmax);
  }
}

// PatchEmbed3D gather (swin_transformer.py:432,446): video (B, 3, T, H, W) -> patches (B*D*Hp*Wp, 3*pt*ph*pw) with the
// column order of nn.Conv3d.weight.view(embed, -1): (c, i, j, k).  pw must be 4 (128-bit loads along W).
__global__ void __launch_bounds__(256) patchify3d_kernel(const float* __restrict__ v, float* __restrict__ out, int B,
                                                        int T, int H, int W, int pt, int ph) {
  const int D = T / pt, Hp = H / ph, Wp = W / 4;
  const int K4 = 3 * pt * ph;   // float4 per patch row
  const int64_t n4 = (int64_t)B * D * Hp * Wp * K4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const int k4 = (int)(i % K4);
    int64_t tok = i / K4;
    const int j = k4 % ph;
    int t2 = k4 / ph;
    const int ii = t2 % pt;
    const int c = t2 / pt;
    const int wp = (int)(tok % Wp); tok /= Wp;
    const int hp = (int)(tok % Hp); tok /= Hp;
    const int d = (int)(tok % D);
    const int64_t b = tok / D;
    const float* src = v + ((((b * 3 + c) * T + d * pt + ii) * H + hp * ph + j) * (int64_t)W + wp * 4);
    st4(out + i * 4, ld_stream4(src));
  }
}

}  // namespace vitta

using namespace vitta;

extern "C" {

int vitta_ln_chunking(int64_t rows, int C, int want_stats, VittaChunking* out) {
  VITTA_CHECK_ARG(out && rows > 0 && C > 0 && C % 4 == 0, VITTA_E_BADARG, "ln_chunking: bad shape");
  // a chunk = the rows one warp normalises = one statistics entry
  int rpw = want_stats ? 64 : 8;
  while (rpw > 4 && (rows + rpw - 1) / rpw < 148 * 16) rpw >>= 1;
  out->chunk_rows = rpw;
  out->frame_rows = rows;
  const int64_t n = (rows + rpw - 1) / rpw;
  VITTA_CHECK_ARG(n < (1ll << 31), VITTA_E_UNSUPPORTED, "ln_chunking: too many chunks");
  out->n_entries = (int32_t)n;
  out->chunks_per_frame = (int32_t)n;
  out->reserved = 0;
  return 0;
}

static int ln_check_gather(const VittaLnGather* g, int C, int64_t rows, LnGather* o) {
  o->merge = 0;
  if (!g) return 0;
  VITTA_CHECK_ARG(g->B > 0 && g->D > 0 && g->H > 0 && g->W > 0 && g->Cin > 0 && g->Cin % 4 == 0 && C == 4 * g->Cin,
                  VITTA_E_BADARG, "ln: bad PatchMerging geometry");
  o->merge = 1; o->D = g->D; o->H = g->H; o->W = g->W; o->Cin = g->Cin;
  o->H2 = (g->H + 1) / 2; o->W2 = (g->W + 1) / 2;
  VITTA_CHECK_ARG(rows == (int64_t)g->B * g->D * o->H2 * o->W2, VITTA_E_BADARG, "ln: rows do not match the merge geometry");
  return 0;
}

int vitta_ln_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean, float* rstd,
                 float* part, int64_t rows, int C, const VittaLnGather* gather, void* stream) {
  return vitta_ln_fwd_amax(x, gamma, beta, eps, y, mean, rstd, part, rows, C, gather, nullptr, stream);
}

int vitta_ln_fwd_amax(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean, float* rstd,
                      float* part, int64_t rows, int C, const VittaLnGather* gather, float* y_amax, void* stream) {
  VITTA_CHECK_ARG(x && gamma && beta && y && mean && rstd && rows > 0, VITTA_E_BADARG, "ln_fwd: null pointer");
  VITTA_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 2048, VITTA_E_UNSUPPORTED, "ln_fwd: C must be a multiple of 4, <= 2048");
  VITTA_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) && (!part || aligned16(part)),
                  VITTA_E_ALIGN, "ln_fwd: tensors must be 16-byte aligned");
  LnFwdArgs p;
  int rc = ln_check_gather(gather, C, rows, &p.g);
  if (rc) return rc;
  VittaChunking ch;
  rc = vitta_ln_chunking(rows, C, part != nullptr, &ch);
  if (rc) return rc;
  p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.mean = mean; p.rstd = rstd; p.part = part; p.amax_y = y_amax;
  p.rows = rows; p.C = C; p.rows_per_warp = ch.chunk_rows; p.eps = eps;
  const unsigned grid = (unsigned)((ch.n_entries + kLnWarps - 1) / kLnWarps);
  cudaStream_t st = (cudaStream_t)stream;
  const int vpl = ln_vpl(C);
  const int c4 = C / 4;
  if (!p.g.merge && c4 > 16 && c4 <= 64 && ch.chunk_rows % 4 == 0) {
    unsigned grid = (unsigned)((ch.n_entries + kLnWarps - 1) / kLnWarps);
    if (!part) {
      // no statistics entry to respect: 32 rows per warp (8 passes of 4 rows) amortise the per-warp prologue -- with the
      // 8-row chunks of vitta_ln_chunking a warp ran two passes and the kernel was bound by its fixed part
      int rpw = 32;
      while (rpw > 8 && (rows + rpw - 1) / rpw < (int64_t)148 * 16 * kLnWarps) rpw >>= 1;
      p.rows_per_warp = rpw;
      const int64_t n = (rows + rpw - 1) / rpw;
      grid = (unsigned)((n + kLnWarps - 1) / kLnWarps);
    }
#define VITTA_LN_NARROW(L, V)                                                            \
  if (part) ln_fwd_narrow_kernel<L, V, true><<<grid, kLnThreads, 0, st>>>(p);            \
  else ln_fwd_narrow_kernel<L, V, false><<<grid, kLnThreads, 0, st>>>(p)
    if (c4 <= 24) { VITTA_LN_NARROW(8, 3); }
    else if (c4 <= 32) { VITTA_LN_NARROW(8, 4); }
    else if (c4 <= 48) { VITTA_LN_NARROW(16, 3); }
    else { VITTA_LN_NARROW(16, 4); }
#undef VITTA_LN_NARROW
    VITTA_CHECK_LAUNCH();
    return 0;
  }
#define VITTA_LN_FWD(V)                                                        \
  if (part) ln_fwd_kernel<V, true><<<grid, kLnThreads, 0, st>>>(p);            \
  else ln_fwd_kernel<V, false><<<grid, kLnThreads, 0, st>>>(p)
  switch (vpl) {
    case 1: VITTA_LN_FWD(1); break;
    case 2: VITTA_LN_FWD(2); break;
    case 4: VITTA_LN_FWD(4); break;
    case 8: VITTA_LN_FWD(8); break;
    default: VITTA_LN_FWD(16); break;
  }
#undef VITTA_LN_FWD
  VITTA_CHECK_LAUNCH();
  return 0;
}

int64_t vitta_ln_bwd_ws_floats(int64_t rows, int C) {
  if (rows <= 0 || C <= 0) return -1;
  return (int64_t)148 * 4 * 2 * C + kWsHeader;   // independent of rows: one buffer per C serves every call
}

int vitta_ln_bwd(const float* gy, const float* x, const float* gamma, const float* beta, const float* mean,
                 const float* rstd, const float* gadd, const float* coef_a, const float* coef_b, const float* coef_mean,
                 const float* gscale, float* gx, float* dgamma, float* dbeta, float* ws, int64_t rows, int C,
                 const VittaLnGather* gather, void* stream) {
  return vitta_ln_bwd_amax(gy, x, gamma, beta, mean, rstd, gadd, coef_a, coef_b, coef_mean, gscale, gx, dgamma, dbeta, ws,
                           rows, C, gather, nullptr, stream);
}

int vitta_ln_bwd_amax(const float* gy, const float* x, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, const float* gadd, const float* coef_a, const float* coef_b,
                      const float* coef_mean, const float* gscale, float* gx, float* dgamma, float* dbeta, float* ws,
                      int64_t rows, int C, const VittaLnGather* gather, float* gx_amax, void* stream) {
  VITTA_CHECK_ARG(gy && x && gamma && beta && mean && rstd && gx && ws && rows > 0, VITTA_E_BADARG, "ln_bwd: null pointer");
  VITTA_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 2048, VITTA_E_UNSUPPORTED, "ln_bwd: C must be a multiple of 4, <= 2048");
  VITTA_CHECK_ARG((coef_a == nullptr) == (coef_b == nullptr) && (coef_a == nullptr) == (coef_mean == nullptr) &&
                      (coef_a == nullptr) == (gscale == nullptr),
                  VITTA_E_BADARG, "ln_bwd: hook coefficients must come as (a, b, mean, gscale)");
  VITTA_CHECK_ARG(aligned16(gy) && aligned16(x) && aligned16(gx) && aligned16(ws) && (!gadd || aligned16(gadd)),
                  VITTA_E_ALIGN, "ln_bwd: tensors must be 16-byte aligned");
  LnBwdArgs p;
  int rc = ln_check_gather(gather, C, rows, &p.g);
  if (rc) return rc;
  VITTA_CHECK_ARG(!(p.g.merge && gadd), VITTA_E_BADARG, "ln_bwd: gadd is not available in merge mode");
  p.gy = gy; p.x = x; p.gamma = gamma; p.beta = beta; p.mean = mean; p.rstd = rstd; p.gadd = gadd;
  p.ca = coef_a; p.cb = coef_b; p.cm = coef_mean; p.gs = gscale;
  p.gx = gx; p.dgamma = dgamma; p.dbeta = dbeta; p.ws = ws; p.rows = rows; p.C = C;
  p.amax_gx = p.g.merge ? nullptr : gx_amax;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.g.merge && ((p.g.H & 1) || (p.g.W & 1))) {
    // odd extents: every source element is still written exactly once (the padded ones do not exist)
  }
  unsigned grid = (unsigned)ln_bwd_grid(rows);
  const int vpl = ln_vpl(C);
  const int c4 = C / 4;
  if (!p.g.merge && c4 > 16 && c4 <= 48) {
    // (four float4 per lane spill at the 128 registers two resident CTAs allow: C = 128 takes 16 lanes x 2, C = 256 stays
    //  on the one-warp-per-row kernel)
    if (c4 <= 24) ln_bwd_narrow_kernel<8, 3><<<grid, kLnThreads, 0, st>>>(p);
    else if (c4 <= 32) ln_bwd_narrow_kernel<16, 2><<<grid, kLnThreads, 0, st>>>(p);
    else ln_bwd_narrow_kernel<16, 3><<<grid, kLnThreads, 0, st>>>(p);
  } else if (vpl <= 2) {
    if (vpl == 1) ln_bwd_rows_kernel<1><<<grid, kLnThreads, 0, st>>>(p);
    else ln_bwd_rows_kernel<2><<<grid, kLnThreads, 0, st>>>(p);
  } else {
    // wide rows: K warps per row, 8 / K row slots per CTA, two rows per slot and iteration
    const int need = (C / 4 + kLnSplitCols - 1) / kLnSplitCols;
    const int K = need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : 8;
    int64_t g = (rows + (kLnWarps / K) * 2 - 1) / ((kLnWarps / K) * 2);
    if (g > 148 * 2) g = 148 * 2;     // two resident CTAs per SM (launch bounds), one wave
    grid = (unsigned)g;
    switch (K) {
      case 1: ln_bwd_split_kernel<1><<<grid, kLnThreads, 0, st>>>(p); break;
      case 2: ln_bwd_split_kernel<2><<<grid, kLnThreads, 0, st>>>(p); break;
      case 4: ln_bwd_split_kernel<4><<<grid, kLnThreads, 0, st>>>(p); break;
      default: ln_bwd_split_kernel<8><<<grid, kLnThreads, 0, st>>>(p); break;
    }
  }
  VITTA_CHECK_LAUNCH();
  if (dgamma || dbeta) {
    ln_bwd_finish_kernel<<<(unsigned)((2 * C + 31) / 32), 256, 0, st>>>(p.ws + kWsHeader, (int)grid, C, dgamma, dbeta);
    VITTA_CHECK_LAUNCH();
  }
  return 0;
}

int64_t vitta_colsum_ws_floats(int64_t rows, int C) {
  if (rows <= 0 || C <= 0) return -1;
  return (int64_t)(148 * 2 + 1) * C + kWsHeader;
}

int vitta_colsum(const float* x, int64_t rows, int C, float* out, int accumulate, float* ws, void* stream) {
  VITTA_CHECK_ARG(x && out && ws && rows > 0 && C > 0 && C % 4 == 0 && C <= 128 * kWsHeader, VITTA_E_BADARG,
                  "colsum: bad arguments");
  VITTA_CHECK_ARG(aligned16(x) && aligned16(ws), VITTA_E_ALIGN, "colsum: tensors must be 16-byte aligned");
  const int ctiles = (C + 127) / 128;
  int64_t gx = (rows + 63) / 64;
  int64_t cap = (148 * 2 + ctiles - 1) / ctiles;
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)ctiles);
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, ws, rows, C, accumulate);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_frame_mean(const float* x, int64_t frames, int rows, int C, float* out, void* stream) {
  VITTA_CHECK_ARG(x && out && frames > 0 && rows > 0 && C > 0 && C % 4 == 0, VITTA_E_BADARG, "frame_mean: bad arguments");
  VITTA_CHECK_ARG(aligned16(x) && aligned16(out), VITTA_E_ALIGN, "frame_mean: tensors must be 16-byte aligned");
  dim3 grid((unsigned)frames, (unsigned)((C + 127) / 128));
  frame_mean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, rows, C);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_frame_mean_bwd(const float* g, int64_t frames, int rows, int C, float* gx, void* stream) {
  VITTA_CHECK_ARG(g && gx && frames > 0 && rows > 0 && C > 0 && C % 4 == 0, VITTA_E_BADARG, "frame_mean_bwd: bad arguments");
  const int64_t n4 = frames * rows * (C / 4);
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  frame_mean_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, gx, n4, rows, C / 4);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_row_scale(const float* x, const float* scale, int64_t rows, int64_t rows_per_group, int C, float* out,
                    void* stream) {
  return vitta_row_scale_amax(x, scale, rows, rows_per_group, C, out, nullptr, stream);
}

int vitta_row_scale_amax(const float* x, const float* scale, int64_t rows, int64_t rows_per_group, int C, float* out,
                         float* out_amax, void* stream) {
  VITTA_CHECK_ARG(x && scale && out && rows > 0 && rows_per_group > 0 && C > 0 && C % 4 == 0, VITTA_E_BADARG,
                  "row_scale: bad arguments");
  const int64_t n4 = rows * (C / 4);
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  row_scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, scale, out, n4, C / 4, rows_per_group, out_amax);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_patchify3d(const float* video, int B, int T, int H, int W, int pt, int ph, int pw, float* out, void* stream) {
  VITTA_CHECK_ARG(video && out && B > 0 && T > 0 && H > 0 && W > 0 && pt > 0 && ph > 0, VITTA_E_BADARG,
                  "patchify3d: bad arguments");
  VITTA_CHECK_ARG(pw == 4 && W % 4 == 0 && T % pt == 0 && H % ph == 0, VITTA_E_UNSUPPORTED,
                  "patchify3d: patch width must be 4 and the clip a multiple of the patch (swin_transformer.py:440-446 pads "
                  "otherwise; pad on the host)");
  VITTA_CHECK_ARG(aligned16(video) && aligned16(out), VITTA_E_ALIGN, "patchify3d: tensors must be 16-byte aligned");
  const int64_t n4 = (int64_t)B * (T / pt) * (H / ph) * (W / 4) * 3 * pt * ph;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  patchify3d_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(video, out, B, T, H, W, pt, ph);
  VITTA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
