// K4: fused BatchNorm(eval) + statistics partials + residual (optionally through a 2nd BN) + ReLU + pooled sums,
// channels-last, forward and backward.  One pass over HBM each way; 128-bit loads/stores.
#include "common.cuh"

namespace vitta {

struct BNDev {
  const float* w;
  const float* b;
  const float* rm;
  const float* rv;
  float eps;
};

struct Aff4 {  // per-thread folded BN for 4 channels: y = (x - rm)*k + b ; xhat = (x - rm)*istd
  float4 rm, k, b, istd;
};

__device__ __forceinline__ Aff4 load_aff(const BNDev& bn, int c) {
  Aff4 a;
  const float4 w = ldg4(bn.w + c), rv = ldg4(bn.rv + c);
  a.rm = ldg4(bn.rm + c);
  a.b = ldg4(bn.b + c);
  a.istd.x = 1.f / sqrtf(rv.x + bn.eps); a.istd.y = 1.f / sqrtf(rv.y + bn.eps);
  a.istd.z = 1.f / sqrtf(rv.z + bn.eps); a.istd.w = 1.f / sqrtf(rv.w + bn.eps);
  a.k.x = w.x * a.istd.x; a.k.y = w.y * a.istd.y; a.k.z = w.z * a.istd.z; a.k.w = w.w * a.istd.w;
  return a;
}
__device__ __forceinline__ float4 apply_aff(const Aff4& a, float4 x) {
  float4 y;
  y.x = fmaf(x.x - a.rm.x, a.k.x, a.b.x); y.y = fmaf(x.y - a.rm.y, a.k.y, a.b.y);
  y.z = fmaf(x.z - a.rm.z, a.k.z, a.b.z); y.w = fmaf(x.w - a.rm.w, a.k.w, a.b.w);
  return y;
}
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void acc_shift(float4& s1, float4& s2, float4 y, float4 k) {
  float d;
  d = y.x - k.x; s1.x += d; s2.x = fmaf(d, d, s2.x);
  d = y.y - k.y; s1.y += d; s2.y = fmaf(d, d, s2.y);
  d = y.z - k.z; s1.z += d; s2.z = fmaf(d, d, s2.z);
  d = y.w - k.w; s1.w += d; s2.w = fmaf(d, d, s2.w);
}

// sum `v` over the row slots of the CTA; result valid in threads with slot == 0.
__device__ __forceinline__ float4 slot_reduce(float4 v, float4* sm, int tid, int lpr, int rs) {
  __syncthreads();
  sm[tid] = v;
  __syncthreads();
  const int slot = tid / lpr;
  for (int st = rs >> 1; st > 0; st >>= 1) {
    if (slot < st) {
      float4 a = sm[tid], b = sm[tid + st * lpr];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      sm[tid] = a;
    }
    __syncthreads();
  }
  return sm[tid];
}

__device__ __forceinline__ void write_part(float* part, int64_t e, int C, int col4, float4 k, float4 a, float4 b,
                                           int nrows) {
  const float n = (float)nrows, inv = 1.f / n;
  float* o = part + (e * C + (int64_t)col4 * 4) * 2;
  float4 o0, o1;
  o0.x = k.x + a.x * inv; o0.y = fmaxf(b.x - a.x * a.x * inv, 0.f);
  o0.z = k.y + a.y * inv; o0.w = fmaxf(b.y - a.y * a.y * inv, 0.f);
  o1.x = k.z + a.z * inv; o1.y = fmaxf(b.z - a.z * a.z * inv, 0.f);
  o1.z = k.w + a.w * inv; o1.w = fmaxf(b.w - a.w * a.w * inv, 0.f);
  st4(o, o0);
  st4(o + 4, o1);
}

__device__ __forceinline__ float amax4(float m, float4 v) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}
// max over the warp, one integer atomic per warp (non-negative floats order like their bit patterns); all 32 lanes call it
__device__ __forceinline__ void amax_commit(float m, float* slot) {
  const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
  if ((threadIdx.x & 31) == 0 && w) atomicMax(reinterpret_cast<unsigned int*>(slot), w);
}

// RES: 0 none, 1 raw residual, 2 residual through its own BN
// AMAX: also accumulate max|out| into *amax_out (the operand range the fp16-split GEMM that consumes `out` needs; the
//       kernel touches every element anyway, so the extra HBM pass of vitta_amax_f32 disappears)
template <int RES, bool AMAX = false>
__global__ void __launch_bounds__(kThreads) bn_act_fwd_kernel(const float* __restrict__ x, BNDev bn,
                                                             const float* __restrict__ res, BNDev bn2, int relu,
                                                             float* __restrict__ out, float* __restrict__ part_main,
                                                             float* __restrict__ part_res, float* __restrict__ pool_part,
                                                             int C, int lpr, int rs, int chunk_rows, int cpf,
                                                             int64_t frame_rows, float* __restrict__ amax_out = nullptr) {
  __shared__ float4 sm[kThreads];
  const int tid = threadIdx.x;
  const int lane = tid % lpr;
  const int slot = tid / lpr;
  // 1-D grid, channel tile fastest: the CTAs that together cover the full rows of a chunk are scheduled side by side, so
  // a DRAM page (a row is up to 8 KB) is swept once instead of once per channel tile
  const int ctiles = (C / 4 + lpr - 1) / lpr;
  const int col4 = (int)(blockIdx.x % ctiles) * lpr + lane;
  const bool active = col4 * 4 < C;
  const int64_t e = blockIdx.x / ctiles;
  const int64_t frame = e / cpf;
  const int j = (int)(e % cpf);
  const int64_t row0 = frame * frame_rows + (int64_t)j * chunk_rows;
  const int64_t rem = frame_rows - (int64_t)j * chunk_rows;
  const int nrows = (int)(rem < chunk_rows ? rem : chunk_rows);

  float4 s1 = f4zero(), s2 = f4zero(), r1 = f4zero(), r2 = f4zero(), pool = f4zero(), ky = f4zero(), kr = f4zero();
  float am = 0.f;
  if (active) {
    const int c = col4 * 4;
    const Aff4 a = load_aff(bn, c);
    Aff4 a2;
    if (RES == 2) a2 = load_aff(bn2, c);
    const int64_t off0 = row0 * C + c;
    ky = apply_aff(a, ldg4(x + off0));
    if (RES == 2) kr = apply_aff(a2, ldg4(res + off0));
    // Rows are processed in batches whose loads are ALL issued before the first use: a plain `for (r < nrows)` loop,
    // even unrolled, keeps one 16-byte load in flight per thread (the trip count is dynamic, so every copy of the body
    // ends in an exit branch and ptxas will not hoist loads across it) -- 16 KB in flight per SM, which is what held
    // this kernel at 0.6 of the copy bandwidth in round 1 (profiles/r01_tanet_ncu_full.md).
    constexpr int kB = 4;
    auto body = [&](float4 xv, float4 rr, int64_t off) {
      float4 y = apply_aff(a, xv);
      if (part_main) acc_shift(s1, s2, y, ky);
      if (RES != 0) {
        if (RES == 2) {
          rr = apply_aff(a2, rr);
          if (part_res) acc_shift(r1, r2, rr, kr);
        }
        y.x += rr.x; y.y += rr.y; y.z += rr.z; y.w += rr.w;
      }
      if (relu) {
        y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
      }
      st4(out + off, y);
      if constexpr (AMAX) am = amax4(am, y);
      pool.x += y.x; pool.y += y.y; pool.z += y.z; pool.w += y.w;
    };
    // ... and software-pipelined: the loads of batch i+1 are issued BEFORE batch i is processed (two register sets), so
    // the memory system never idles during the arithmetic / store phase of a batch.  Rows past the chunk's end are
    // predicated off, which also replaces the remainder loop.
    const int64_t rstep = (int64_t)rs * C;
    auto load = [&](int r0, float4* xb, float4* rb) {
#pragma unroll
      for (int j = 0; j < kB; ++j) {
        if (r0 + j * rs < nrows) {
          const int64_t off = off0 + (int64_t)r0 * C + j * rstep;
          xb[j] = ld_stream4(x + off);
          if (RES != 0) rb[j] = ld_stream4(res + off);
        }
      }
    };
    auto proc = [&](int r0, const float4* xb, const float4* rb) {
#pragma unroll
      for (int j = 0; j < kB; ++j)
        if (r0 + j * rs < nrows) body(xb[j], RES != 0 ? rb[j] : f4zero(), off0 + (int64_t)r0 * C + j * rstep);
    };
    float4 xa[kB], ra[kB], xb2[kB], rb2[kB];
    const int bstep = kB * rs;
    load(slot, xa, ra);
    for (int r0 = slot; r0 < nrows; r0 += 2 * bstep) {
      load(r0 + bstep, xb2, rb2);
      proc(r0, xa, ra);
      load(r0 + 2 * bstep, xa, ra);
      proc(r0 + bstep, xb2, rb2);
    }
  }
  if constexpr (AMAX) amax_commit(am, amax_out);
  if (part_main) {
    float4 a = slot_reduce(s1, sm, tid, lpr, rs);
    float4 b = slot_reduce(s2, sm, tid, lpr, rs);
    if (slot == 0 && active) write_part(part_main, e, C, col4, ky, a, b, nrows);
  }
  if (RES == 2 && part_res) {
    float4 a = slot_reduce(r1, sm, tid, lpr, rs);
    float4 b = slot_reduce(r2, sm, tid, lpr, rs);
    if (slot == 0 && active) write_part(part_res, e, C, col4, kr, a, b, nrows);
  }
  if (pool_part) {
    float4 p = slot_reduce(pool, sm, tid, lpr, rs);
    if (slot == 0 && active) st4(pool_part + e * C + (int64_t)col4 * 4, p);
  }
}

__global__ void pool_finish_kernel(const float* __restrict__ pool_part, float* __restrict__ pool_out, int64_t frames,
                                   int cpf, int C, float inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames * C) return;
  const int64_t f = i / C;
  const int c = (int)(i % C);
  float s = 0.f;
  for (int j = 0; j < cpf; ++j) s += pool_part[(f * cpf + j) * C + c];
  pool_out[i] = s * inv;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct BwdArgs {
  const float* gout;
  const float* gpool;  // (frames, C) or null
  const float* x;
  const float* res;
  BNDev bn, bn2;
  const float *ca, *cb, *cm, *gs;     // main-branch alignment coefficients + batch mean (null: none)
  const float *ca2, *cb2, *cm2, *gs2; // residual-BN alignment coefficients
  float *gx, *gres, *gw, *gb, *gw2, *gb2;
  float* ws;
  int relu, C, lpr, rs, chunk_rows, cpf;
  int64_t frame_rows, n_chunks, ticket_off;   // tickets live at ws + ticket_off (fixed per shape, not per grid)
  float inv_frame_rows;
};

__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

// AMAX: max|gx| and max|gres| accumulate into *amax_gx / *amax_gres.  The body is shared; the two __global__ wrappers below
// keep the default kernel's signature (and with it its code) exactly as it was.
template <int RES, bool AMAX>
__device__ __forceinline__ void bn_act_bwd_body(const BwdArgs& p, float* __restrict__ amax_gx,
                                                float* __restrict__ amax_gres) {
  __shared__ float4 sm[kThreads];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int lane = tid % p.lpr;
  const int slot = tid / p.lpr;
  const int C = p.C;
  const int ctiles = (C / 4 + p.lpr - 1) / p.lpr;   // 1-D grid, channel tile fastest (see the forward kernel)
  const int by = (int)(blockIdx.x % ctiles), bx = (int)(blockIdx.x / ctiles), gdx = (int)(gridDim.x / ctiles);
  const int col4 = by * p.lpr + lane;
  const bool active = col4 * 4 < C;
  const int c = col4 * 4;
  float4 agw = f4zero(), agb = f4zero(), agw2 = f4zero(), agb2 = f4zero();
  float amx = 0.f, amr = 0.f;
  if (active) {
    const Aff4 a = load_aff(p.bn, c);
    Aff4 a2;
    if (RES == 2) a2 = load_aff(p.bn2, c);
    float4 ca = f4zero(), cb = f4zero(), ca2 = f4zero(), cb2 = f4zero(), cm = f4zero(), cm2 = f4zero();
    if (p.ca) {
      const float g = __ldg(p.gs);
      ca = ldg4(p.ca + c); cb = ldg4(p.cb + c); cm = ldg4(p.cm + c);
      ca.x *= g; ca.y *= g; ca.z *= g; ca.w *= g; cb.x *= g; cb.y *= g; cb.z *= g; cb.w *= g;
    }
    if (RES == 2 && p.ca2) {
      const float g = __ldg(p.gs2);
      ca2 = ldg4(p.ca2 + c); cb2 = ldg4(p.cb2 + c); cm2 = ldg4(p.cm2 + c);
      ca2.x *= g; ca2.y *= g; ca2.z *= g; ca2.w *= g; cb2.x *= g; cb2.y *= g; cb2.z *= g; cb2.w *= g;
    }
    for (int64_t e = bx; e < p.n_chunks; e += gdx) {
      const int64_t frame = e / p.cpf;
      const int j = (int)(e % p.cpf);
      const int64_t row0 = frame * p.frame_rows + (int64_t)j * p.chunk_rows;
      const int64_t rem = p.frame_rows - (int64_t)j * p.chunk_rows;
      const int nrows = (int)(rem < p.chunk_rows ? rem : p.chunk_rows);
      float4 gp = f4zero();
      if (p.gpool) {
        gp = ldg4(p.gpool + frame * C + c);
        gp.x *= p.inv_frame_rows; gp.y *= p.inv_frame_rows; gp.z *= p.inv_frame_rows; gp.w *= p.inv_frame_rows;
      }
      const int64_t off0 = row0 * C + c;
      // batched like the forward: all loads of kB rows are in flight before the first use
      constexpr int kB = 4;
      auto body = [&](float4 xv, float4 g, float4 rx, int64_t off) {
        const float4 y = apply_aff(a, xv);
        float4 rr = f4zero();
        if (RES != 0) rr = (RES == 2) ? apply_aff(a2, rx) : rx;
        g.x += gp.x; g.y += gp.y; g.z += gp.z; g.w += gp.w;
        if (p.relu) {
          g.x = (y.x + rr.x > 0.f) ? g.x : 0.f; g.y = (y.y + rr.y > 0.f) ? g.y : 0.f;
          g.z = (y.z + rr.z > 0.f) ? g.z : 0.f; g.w = (y.w + rr.w > 0.f) ? g.w : 0.f;
        }
        // main branch
        float4 gy = f4fma(cb, f4sub(y, cm), ca);
        gy.x += g.x; gy.y += g.y; gy.z += g.z; gy.w += g.w;
        st4(p.gx + off, make_float4(gy.x * a.k.x, gy.y * a.k.y, gy.z * a.k.z, gy.w * a.k.w));
        if constexpr (AMAX) amx = amax4(amx, make_float4(gy.x * a.k.x, gy.y * a.k.y, gy.z * a.k.z, gy.w * a.k.w));
        agb.x += gy.x; agb.y += gy.y; agb.z += gy.z; agb.w += gy.w;
        agw.x = fmaf(gy.x, (xv.x - a.rm.x) * a.istd.x, agw.x); agw.y = fmaf(gy.y, (xv.y - a.rm.y) * a.istd.y, agw.y);
        agw.z = fmaf(gy.z, (xv.z - a.rm.z) * a.istd.z, agw.z); agw.w = fmaf(gy.w, (xv.w - a.rm.w) * a.istd.w, agw.w);
        if (RES == 1) {
          st4(p.gres + off, g);
          if constexpr (AMAX) amr = amax4(amr, g);
        } else if (RES == 2) {
          float4 gr = f4fma(cb2, f4sub(rr, cm2), ca2);
          gr.x += g.x; gr.y += g.y; gr.z += g.z; gr.w += g.w;
          st4(p.gres + off, make_float4(gr.x * a2.k.x, gr.y * a2.k.y, gr.z * a2.k.z, gr.w * a2.k.w));
          if constexpr (AMAX) amr = amax4(amr, make_float4(gr.x * a2.k.x, gr.y * a2.k.y, gr.z * a2.k.z, gr.w * a2.k.w));
          agb2.x += gr.x; agb2.y += gr.y; agb2.z += gr.z; agb2.w += gr.w;
          agw2.x = fmaf(gr.x, (rx.x - a2.rm.x) * a2.istd.x, agw2.x); agw2.y = fmaf(gr.y, (rx.y - a2.rm.y) * a2.istd.y, agw2.y);
          agw2.z = fmaf(gr.z, (rx.z - a2.rm.z) * a2.istd.z, agw2.z); agw2.w = fmaf(gr.w, (rx.w - a2.rm.w) * a2.istd.w, agw2.w);
        }
      };
      const int64_t rstep = (int64_t)p.rs * C;
      int r = slot;
      for (; r + (kB - 1) * p.rs < nrows; r += kB * p.rs) {
        const int64_t off = off0 + (int64_t)r * C;
        float4 xv[kB], gv[kB], rv[kB];
#pragma unroll
        for (int j = 0; j < kB; ++j) xv[j] = ld_stream4(p.x + off + j * rstep);
#pragma unroll
        for (int j = 0; j < kB; ++j) gv[j] = ld_stream4(p.gout + off + j * rstep);
        if (RES != 0) {
#pragma unroll
          for (int j = 0; j < kB; ++j) rv[j] = ld_stream4(p.res + off + j * rstep);
        }
#pragma unroll
        for (int j = 0; j < kB; ++j) body(xv[j], gv[j], RES != 0 ? rv[j] : f4zero(), off + j * rstep);
      }
      for (; r < nrows; r += p.rs) {
        const int64_t off = off0 + (int64_t)r * C;
        const float4 xv = ld_stream4(p.x + off);
        const float4 g = ld_stream4(p.gout + off);
        float4 rx = f4zero();
        if (RES != 0) rx = ld_stream4(p.res + off);
        body(xv, g, rx, off);
      }
    }
  }
  if constexpr (AMAX) {
    amax_commit(amx, amax_gx);
    if (RES != 0) amax_commit(amr, amax_gres);
  }
  // per-CTA partial parameter gradients -> ws[bx][k][C], k in {gw, gb, gw2, gb2}
  const int nk = (RES == 2) ? 4 : 2;
  float4 red[4];
  red[0] = slot_reduce(agw, sm, tid, p.lpr, p.rs);
  red[1] = slot_reduce(agb, sm, tid, p.lpr, p.rs);
  if (RES == 2) {
    red[2] = slot_reduce(agw2, sm, tid, p.lpr, p.rs);
    red[3] = slot_reduce(agb2, sm, tid, p.lpr, p.rs);
  }
  float* wsb = p.ws + (int64_t)bx * 4 * C;
  if (slot == 0 && active) {
    for (int k = 0; k < nk; ++k) st4(wsb + (int64_t)k * C + c, red[k]);
  }
  __threadfence();
  __syncthreads();
  int* tickets = reinterpret_cast<int*>(p.ws + p.ticket_off);
  if (tid == 0) s_last = (atomicAdd(tickets + by, 1) == gdx - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA of this channel tile: sum the partials in CTA order (deterministic) and accumulate into the grads
  const int nch = p.lpr * 4;
  for (int i = tid; i < nch * nk; i += kThreads) {
    const int k = i / nch;
    const int cc = by * nch + (i % nch);
    if (cc >= C) continue;
    const float s = ordered_sum_strided(p.ws + (int64_t)k * C + cc, gdx, 4 * (int64_t)C);
    float* dst = (k == 0) ? p.gw : (k == 1) ? p.gb : (k == 2) ? p.gw2 : p.gb2;
    if (dst) dst[cc] += s;
  }
  if (tid == 0) tickets[by] = 0;
}

template <int RES, bool AMAX = false>
__global__ void __launch_bounds__(kThreads) bn_act_bwd_kernel(BwdArgs p) {
  bn_act_bwd_body<RES, false>(p, nullptr, nullptr);
}
template <int RES>
__global__ void __launch_bounds__(kThreads) bn_act_bwd_amax_kernel(BwdArgs p, float* __restrict__ amax_gx,
                                                                  float* __restrict__ amax_gres) {
  bn_act_bwd_body<RES, true>(p, amax_gx, amax_gres);
}

// upper bound of the backward grid (per channel tile): sizes the workspace, fixed per shape
static inline int bwd_grid_x(const ClGeom& g) {
  int64_t cap = (148 * 4 + g.ctiles - 1) / g.ctiles;
  if (cap < 1) cap = 1;
  int64_t n = g.n_chunks();
  return (int)(n < cap ? n : cap);
}

// The backward is a grid-stride (persistent) kernel: launch exactly one wave of what actually fits per SM for the
// variant at hand (80 / 94 / 127 registers -> 3 / 2 / 2 CTAs), otherwise the surplus CTAs form a second, mostly empty wave.
template <int RES, bool AMAX = false>
static int bwd_resident_ctas() {
  static int cached = 0;
  if (cached == 0) {
    int per_sm = 0;
    const cudaError_t e = AMAX ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_act_bwd_amax_kernel<RES>,
                                                                               kThreads, 0)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_act_bwd_kernel<RES>,
                                                                               kThreads, 0);
    if (e != cudaSuccess ||
        per_sm < 1)
      per_sm = 2;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cached = per_sm * (sms > 0 ? sms : 148);
  }
  return cached;
}

}  // namespace vitta

using namespace vitta;

static BNDev to_dev(const VittaBN& b) { return BNDev{b.weight, b.bias, b.running_mean, b.running_var, b.eps}; }

extern "C" {

static int bn_act_fwd_impl(const float* x, VittaBN bn, const float* res, const VittaBN* res_bn, int relu, float* out,
                           float* part_main, float* part_res, float* pool_part, float* pool_out, int64_t frames,
                           int64_t frame_rows, int C, float* amax_out, void* stream) {
  VITTA_CHECK_ARG(x && out && bn.weight && bn.bias && bn.running_mean && bn.running_var, VITTA_E_BADARG,
                  "bn_act_fwd: null pointer");
  VITTA_CHECK_ARG(C > 0 && C % 4 == 0 && frames > 0 && frame_rows > 0, VITTA_E_BADARG, "bn_act_fwd: bad shape");
  VITTA_CHECK_ARG(aligned16(x) && aligned16(out) && (!res || aligned16(res)), VITTA_E_ALIGN,
                  "bn_act_fwd: tensors must be 16-byte aligned");
  VITTA_CHECK_ARG(!(res_bn && !res), VITTA_E_BADARG, "bn_act_fwd: res_bn without res");
  VITTA_CHECK_ARG((pool_part == nullptr) == (pool_out == nullptr), VITTA_E_BADARG, "bn_act_fwd: pool buffers");
  ClGeom g = cl_geom(frames, frame_rows, C);
  VITTA_CHECK_ARG(g.n_chunks() * g.ctiles < (1ll << 31), VITTA_E_UNSUPPORTED, "bn_act_fwd: grid too large");
  dim3 grid((unsigned)(g.n_chunks() * g.ctiles));
  cudaStream_t st = (cudaStream_t)stream;
  BNDev b1 = to_dev(bn), b2 = res_bn ? to_dev(*res_bn) : b1;
  if (amax_out) {
    if (!res)
      bn_act_fwd_kernel<0, true><<<grid, kThreads, 0, st>>>(x, b1, nullptr, b2, relu, out, part_main, nullptr, pool_part,
                                                           C, g.lpr, g.rs, g.chunk_rows, g.cpf, g.frame_rows, amax_out);
    else if (!res_bn)
      bn_act_fwd_kernel<1, true><<<grid, kThreads, 0, st>>>(x, b1, res, b2, relu, out, part_main, nullptr, pool_part, C,
                                                           g.lpr, g.rs, g.chunk_rows, g.cpf, g.frame_rows, amax_out);
    else
      bn_act_fwd_kernel<2, true><<<grid, kThreads, 0, st>>>(x, b1, res, b2, relu, out, part_main, part_res, pool_part, C,
                                                           g.lpr, g.rs, g.chunk_rows, g.cpf, g.frame_rows, amax_out);
  } else if (!res)
    bn_act_fwd_kernel<0><<<grid, kThreads, 0, st>>>(x, b1, nullptr, b2, relu, out, part_main, nullptr, pool_part, C,
                                                   g.lpr, g.rs, g.chunk_rows, g.cpf, g.frame_rows);
  else if (!res_bn)
    bn_act_fwd_kernel<1><<<grid, kThreads, 0, st>>>(x, b1, res, b2, relu, out, part_main, nullptr, pool_part, C, g.lpr,
                                                   g.rs, g.chunk_rows, g.cpf, g.frame_rows);
  else
    bn_act_fwd_kernel<2><<<grid, kThreads, 0, st>>>(x, b1, res, b2, relu, out, part_main, part_res, pool_part, C, g.lpr,
                                                   g.rs, g.chunk_rows, g.cpf, g.frame_rows);
  VITTA_CHECK_LAUNCH();
  if (pool_part) {
    int64_t n = frames * C;
    pool_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pool_part, pool_out, frames, g.cpf, C,
                                                                   1.f / (float)frame_rows);
    VITTA_CHECK_LAUNCH();
  }
  return 0;
}

int vitta_bn_act_fwd(const float* x, VittaBN bn, const float* res, const VittaBN* res_bn, int relu, float* out,
                     float* part_main, float* part_res, float* pool_part, float* pool_out, int64_t frames,
                     int64_t frame_rows, int C, void* stream) {
  return bn_act_fwd_impl(x, bn, res, res_bn, relu, out, part_main, part_res, pool_part, pool_out, frames, frame_rows, C,
                         nullptr, stream);
}

int vitta_bn_act_fwd_amax(const float* x, VittaBN bn, const float* res, const VittaBN* res_bn, int relu, float* out,
                          float* part_main, float* part_res, float* pool_part, float* pool_out, int64_t frames,
                          int64_t frame_rows, int C, float* amax_out, void* stream) {
  VITTA_CHECK_ARG(amax_out, VITTA_E_BADARG, "bn_act_fwd_amax: amax_out is required");
  return bn_act_fwd_impl(x, bn, res, res_bn, relu, out, part_main, part_res, pool_part, pool_out, frames, frame_rows, C,
                         amax_out, stream);
}

int64_t vitta_bn_act_bwd_ws_floats(int64_t frames, int64_t frame_rows, int C) {
  if (C <= 0 || C % 4 || frames <= 0 || frame_rows <= 0) return -1;
  ClGeom g = cl_geom(frames, frame_rows, C);
  return (int64_t)bwd_grid_x(g) * 4 * C + g.ctiles + 4;
}

static int bn_act_bwd_impl(const float* gout, const float* gpool, const float* x, VittaBN bn, const float* res,
                           const VittaBN* res_bn, int relu, const float* coef_a, const float* coef_b,
                           const float* mean_main, const float* gs_main, const float* coef_a2, const float* coef_b2,
                           const float* mean_res, const float* gs_res, float* gx, float* gres, float* gw, float* gb,
                           float* gw2, float* gb2, float* ws, int64_t frames, int64_t frame_rows, int C, float* amax_gx,
                           float* amax_gres, void* stream) {
  VITTA_CHECK_ARG(gout && x && gx && ws, VITTA_E_BADARG, "bn_act_bwd: null pointer");
  VITTA_CHECK_ARG(C > 0 && C % 4 == 0 && frames > 0 && frame_rows > 0, VITTA_E_BADARG, "bn_act_bwd: bad shape");
  VITTA_CHECK_ARG(aligned16(gout) && aligned16(x) && aligned16(gx) && (!res || aligned16(res)) &&
                      (!gres || aligned16(gres)) && aligned16(ws),
                  VITTA_E_ALIGN, "bn_act_bwd: tensors must be 16-byte aligned");
  VITTA_CHECK_ARG(!(res && !gres), VITTA_E_BADARG, "bn_act_bwd: residual without gres");
  VITTA_CHECK_ARG((coef_a == nullptr) == (coef_b == nullptr) && (coef_a == nullptr) == (gs_main == nullptr) &&
                      (coef_a == nullptr) == (mean_main == nullptr),
                  VITTA_E_BADARG, "bn_act_bwd: main coefficients must come as (a, b, mean, gscale)");
  VITTA_CHECK_ARG((coef_a2 == nullptr) == (coef_b2 == nullptr) && (coef_a2 == nullptr) == (gs_res == nullptr) &&
                      (coef_a2 == nullptr) == (mean_res == nullptr),
                  VITTA_E_BADARG, "bn_act_bwd: residual coefficients must come as (a, b, mean, gscale)");
  ClGeom g = cl_geom(frames, frame_rows, C);
  BwdArgs p;
  p.gout = gout; p.gpool = gpool; p.x = x; p.res = res;
  p.bn = to_dev(bn); p.bn2 = res_bn ? to_dev(*res_bn) : p.bn;
  p.ca = coef_a; p.cb = coef_b; p.cm = mean_main; p.gs = gs_main;
  p.ca2 = coef_a2; p.cb2 = coef_b2; p.cm2 = mean_res; p.gs2 = gs_res;
  p.gx = gx; p.gres = gres; p.gw = gw; p.gb = gb; p.gw2 = gw2; p.gb2 = gb2; p.ws = ws;
  p.relu = relu; p.C = C; p.lpr = g.lpr; p.rs = g.rs; p.chunk_rows = g.chunk_rows; p.cpf = g.cpf;
  p.frame_rows = g.frame_rows; p.n_chunks = g.n_chunks(); p.inv_frame_rows = 1.f / (float)frame_rows;
  p.ticket_off = (int64_t)bwd_grid_x(g) * 4 * C;
  const bool am = amax_gx != nullptr;
  VITTA_CHECK_ARG(!am || !res || amax_gres, VITTA_E_BADARG, "bn_act_bwd_amax: amax_gres is required with a residual");
  const int resident = am ? (!res ? bwd_resident_ctas<0, true>() : !res_bn ? bwd_resident_ctas<1, true>()
                                                                            : bwd_resident_ctas<2, true>())
                          : (!res ? bwd_resident_ctas<0>() : !res_bn ? bwd_resident_ctas<1>() : bwd_resident_ctas<2>());
  int gx_ = resident / g.ctiles;
  if (gx_ < 1) gx_ = 1;
  if (gx_ > bwd_grid_x(g)) gx_ = bwd_grid_x(g);
  dim3 grid((unsigned)(gx_ * g.ctiles));
  cudaStream_t st = (cudaStream_t)stream;
  if (am) {
    if (!res)
      bn_act_bwd_amax_kernel<0><<<grid, kThreads, 0, st>>>(p, amax_gx, amax_gres);
    else if (!res_bn)
      bn_act_bwd_amax_kernel<1><<<grid, kThreads, 0, st>>>(p, amax_gx, amax_gres);
    else
      bn_act_bwd_amax_kernel<2><<<grid, kThreads, 0, st>>>(p, amax_gx, amax_gres);
  } else if (!res)
    bn_act_bwd_kernel<0><<<grid, kThreads, 0, st>>>(p);
  else if (!res_bn)
    bn_act_bwd_kernel<1><<<grid, kThreads, 0, st>>>(p);
  else
    bn_act_bwd_kernel<2><<<grid, kThreads, 0, st>>>(p);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_bn_act_bwd(const float* gout, const float* gpool, const float* x, VittaBN bn, const float* res,
                     const VittaBN* res_bn, int relu, const float* coef_a, const float* coef_b, const float* mean_main,
                     const float* gs_main, const float* coef_a2, const float* coef_b2, const float* mean_res,
                     const float* gs_res, float* gx, float* gres, float* gw, float* gb, float* gw2, float* gb2,
                     float* ws, int64_t frames, int64_t frame_rows, int C, void* stream) {
  return bn_act_bwd_impl(gout, gpool, x, bn, res, res_bn, relu, coef_a, coef_b, mean_main, gs_main, coef_a2, coef_b2,
                         mean_res, gs_res, gx, gres, gw, gb, gw2, gb2, ws, frames, frame_rows, C, nullptr, nullptr, stream);
}

int vitta_bn_act_bwd_amax(const float* gout, const float* gpool, const float* x, VittaBN bn, const float* res,
                          const VittaBN* res_bn, int relu, const float* coef_a, const float* coef_b,
                          const float* mean_main, const float* gs_main, const float* coef_a2, const float* coef_b2,
                          const float* mean_res, const float* gs_res, float* gx, float* gres, float* gw, float* gb,
                          float* gw2, float* gb2, float* ws, int64_t frames, int64_t frame_rows, int C, float* amax_gx,
                          float* amax_gres, void* stream) {
  VITTA_CHECK_ARG(amax_gx, VITTA_E_BADARG, "bn_act_bwd_amax: amax_gx is required");
  return bn_act_bwd_impl(gout, gpool, x, bn, res, res_bn, relu, coef_a, coef_b, mean_main, gs_main, coef_a2, coef_b2,
                         mean_res, gs_res, gx, gres, gw, gb, gw2, gb2, ws, frames, frame_rows, C, amax_gx, amax_gres,
                         stream);
}

}  // extern "C"
