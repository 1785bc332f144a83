// K5: TAM gated 3-tap temporal stencil in the native channels-last (N, T, HW, C) layout, forward and backward.
// The reference materialises two transposed copies, the gated tensor and a grouped conv (temporal_module.py:47-63);
// here x is read once and out written once.
#include "common.cuh"

namespace vitta {

__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

// thread = one (n, p, 4 channels) column, marching through t with a 3-frame register window
// AMAX: also accumulate max|out| into *amax_out (operand range for the fp16-split convolution that consumes `out`)
template <bool AMAX = false>
__global__ void __launch_bounds__(kThreads) tam_fwd_kernel(const float* __restrict__ x, const float* __restrict__ kern,
                                                          const float* __restrict__ act, float* __restrict__ out, int N,
                                                          int T, int64_t HW, int C4, float* __restrict__ amax_out = nullptr) {
  const int64_t per_n = HW * C4;
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= (int64_t)N * per_n) return;
  const int n = (int)(idx / per_n);
  const int64_t pc = idx % per_n;          // p*C4 + c4
  const int c = (int)(pc % C4) * 4;
  const int C = C4 * 4;
  const float4 k0 = ldg4(kern + ((int64_t)n * 3 + 0) * C + c);
  const float4 k1 = ldg4(kern + ((int64_t)n * 3 + 1) * C + c);
  const float4 k2 = ldg4(kern + ((int64_t)n * 3 + 2) * C + c);
  const float* xb = x + (int64_t)n * T * per_n * 4 + pc * 4;
  float* ob = out + (int64_t)n * T * per_n * 4 + pc * 4;
  const float* ab = act + (int64_t)n * T * C + c;
  const int64_t ts = per_n * 4;
  float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 cur = mul4(ldg4(ab), ld_stream4(xb));
  float am = 0.f;
  // four frames per batch, their loads issued together (predicated past the clip's end): a frame-by-frame loop keeps a
  // single 16-byte load in flight per thread
  constexpr int kB = 4;
  for (int t0 = 0; t0 < T; t0 += kB) {
    float4 xn[kB], an[kB];
#pragma unroll
    for (int j = 0; j < kB; ++j) {
      const int t = t0 + j + 1;
      xn[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      an[j] = xn[j];
      if (t < T) {
        xn[j] = ld_stream4(xb + (int64_t)t * ts);
        an[j] = ldg4(ab + (int64_t)t * C);
      }
    }
#pragma unroll
    for (int j = 0; j < kB; ++j) {
      const int t = t0 + j;
      if (t < T) {
        const float4 nxt = mul4(an[j], xn[j]);
        float4 o = mul4(k0, prev);
        o = fma4(k1, cur, o);
        o = fma4(k2, nxt, o);
        st4(ob + (int64_t)t * ts, o);
        if constexpr (AMAX) am = fmaxf(fmaxf(am, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
        prev = cur;
        cur = nxt;
      }
    }
  }
  if constexpr (AMAX) {   // threads past the end have returned: reduce over the lanes that are still here
    const unsigned mask = __activemask();
    const uint32_t w = __reduce_max_sync(mask, __float_as_uint(am));
    if ((int)(threadIdx.x & 31) == __ffs(mask) - 1 && w) atomicMax(reinterpret_cast<unsigned int*>(amax_out), w);
  }
}

constexpr int kTamChunkRows = 64;   // upper bound of the rows (pixels) one CTA of the backward walks

// Rows per backward CTA: at most kTamChunkRows, fewer on the small late-stage maps so that a sample still yields ~74 CTAs
// (592 = 148 SMs x 4 at the usual 8 samples) -- with 64-row chunks the 14x14 / 7x7 TAMs launched 32-64 CTAs in total.
static inline int tam_chunk_rows(int64_t HW, int C) {
  const int c4 = C / 4;
  int lpr = 1;
  while (lpr * 2 <= (c4 < 32 ? c4 : 32)) lpr *= 2;
  const int ctiles = (c4 + lpr - 1) / lpr;
  int64_t want = (74 + ctiles - 1) / ctiles;             // chunks per sample aimed at
  int64_t rows = (HW + want - 1) / want;
  if (rows < 4) rows = 4;                                // keep the per-CTA partial (3 x T x C floats) worth its rows
  if (rows > kTamChunkRows) rows = kTamChunkRows;
  return (int)rows;
}

// CTA = (n, row chunk, channel tile).  Thread = (float4 of channels, frame slot): owns frames t = slot, slot+rs, ...
//   gx[t]    = act[t] * (k0*g[t+1] + k1*g[t] + k2*g[t-1])
//   D[t][k]  = sum_p g[t-k+1][p] * x[t][p]                 (partial over the chunk's rows)
__global__ void __launch_bounds__(kThreads, 2) tam_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ x,
                                                          const float* __restrict__ kern, const float* __restrict__ act,
                                                          float* __restrict__ gx, float* __restrict__ dpart, int N, int T,
                                                          int64_t HW, int C, int lpr, int rs, int nchunks, int chunk_rows) {
  const int tid = threadIdx.x;
  const int lane = tid % lpr;
  const int slot = tid / lpr;
  const int col4 = blockIdx.y * lpr + lane;
  if (col4 * 4 >= C) return;
  const int c = col4 * 4;
  const int n = blockIdx.x / nchunks;
  const int ch = blockIdx.x % nchunks;
  const int64_t p0 = (int64_t)ch * chunk_rows;
  const int64_t left = HW - p0;
  const int np = (int)(left < chunk_rows ? left : chunk_rows);
  const float4 k0 = ldg4(kern + ((int64_t)n * 3 + 0) * C + c);
  const float4 k1 = ldg4(kern + ((int64_t)n * 3 + 1) * C + c);
  const float4 k2 = ldg4(kern + ((int64_t)n * 3 + 2) * C + c);
  const int64_t ts = HW * C;
  const int64_t nb = (int64_t)n * T * ts;
  for (int t = slot; t < T; t += rs) {
    const float4 a = ldg4(act + ((int64_t)n * T + t) * C + c);
    float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0, d2 = d0;
    const int64_t base = nb + (int64_t)t * ts + p0 * C + c;
    const bool hp = t + 1 < T, hm = t > 0;
    auto body = [&](float4 g1, float4 gp, float4 gm, float4 xv, int64_t off) {
      float4 s = mul4(k0, gp);
      s = fma4(k1, g1, s);
      s = fma4(k2, gm, s);
      st4(gx + off, mul4(a, s));
      d0 = fma4(gp, xv, d0);
      d1 = fma4(g1, xv, d1);
      d2 = fma4(gm, xv, d2);
    };
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    // Two rows per batch.  The loads that go to HBM (this frame's g and x) are issued one batch AHEAD of their use, so
    // they stay in flight across the arithmetic and the stores of the current batch; the neighbour frames' g rows are
    // being fetched by the adjacent frame slots of this CTA at the same moment (L1 / L2 hits) and are loaded in place.
    int p = 0;
    float4 g1a = z, g1b = z, xa = z, xb = z;
    if (np >= 2) {
      g1a = ldg4(gout + base); g1b = ldg4(gout + base + C);
      xa = ld_stream4(x + base); xb = ld_stream4(x + base + C);
    }
    for (; p + 1 < np; p += 2) {
      const int64_t o0 = base + (int64_t)p * C, o1 = o0 + C;
      const float4 gpa = hp ? ldg4(gout + o0 + ts) : z, gpb = hp ? ldg4(gout + o1 + ts) : z;
      const float4 gma = hm ? ldg4(gout + o0 - ts) : z, gmb = hm ? ldg4(gout + o1 - ts) : z;
      float4 ng1a = z, ng1b = z, nxa = z, nxb = z;
      if (p + 3 < np) {
        const int64_t n0 = o0 + 2 * (int64_t)C, n1 = n0 + C;
        ng1a = ldg4(gout + n0); ng1b = ldg4(gout + n1);
        nxa = ld_stream4(x + n0); nxb = ld_stream4(x + n1);
      }
      body(g1a, gpa, gma, xa, o0);
      body(g1b, gpb, gmb, xb, o1);
      g1a = ng1a; g1b = ng1b; xa = nxa; xb = nxb;
    }
    if (p < np) {
      const int64_t off = base + (int64_t)p * C;
      body(ldg4(gout + off), hp ? ldg4(gout + off + ts) : z, hm ? ldg4(gout + off - ts) : z, ld_stream4(x + off), off);
    }
    float* dp = dpart + ((((int64_t)n * nchunks + ch) * T + t) * 3) * C + c;
    st4(dp, d0);
    st4(dp + C, d1);
    st4(dp + 2 * (int64_t)C, d2);
  }
}


// After tam_bwd_kernel: D[n, t, k, c] = sum over the row chunks of dpart (fixed order), then
//   gkern[n, k, c] = sum_t act[n, t, c] * D[n, t, k, c]        gact[n, t, c] = sum_k kern[n, k, c] * D[n, t, k, c]
// (three eager reductions / products per TAM before).  CTA = (channel tile of <= 32 quads, frame, video) x 8 chunk
// slices: a slice adds every 8th chunk, five chunks (15 x 16 bytes) in flight per thread; the slices are added in slice
// order through shared memory; slice 0 writes gact and parks act * D in the frame's chunk-0 slot of dpart (already
// consumed, and read by no other CTA).  The last CTA of a (video, channel tile) to arrive -- one ticket per pair, left
// at zero again -- adds those over the frames in frame order.  (Round 2's first version walked all chunks of all frames
// in ONE CTA per (video, tile): 8-32 CTAs, up to 19 dependent L2 round trips, 32 us per TAM.)
constexpr int kTamFinMaxT = 16;
constexpr int kTamFinSlices = 8;

__global__ void __launch_bounds__(32 * kTamFinSlices) tam_bwd_finish_kernel(float* __restrict__ dpart,
                                                                           const float* __restrict__ kern,
                                                                           const float* __restrict__ act,
                                                                           float* __restrict__ gkern,
                                                                           float* __restrict__ gact, int* tickets, int N,
                                                                           int T, int nch, int C4) {
  __shared__ float4 sd[kTamFinSlices][3][32];
  __shared__ int s_last;
  const int lane = threadIdx.x, s = threadIdx.y;
  const int t = blockIdx.y, n = blockIdx.z, c4 = blockIdx.x * 32 + lane;
  const int C = C4 * 4, c = c4 * 4;
  const bool on = c4 < C4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 d[3] = {z, z, z};
  if (on) {
    constexpr int kB = 5;
    for (int ch0 = s; ch0 < nch; ch0 += kB * kTamFinSlices) {
      float4 v[kB][3];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int ch = ch0 + u * kTamFinSlices;
        const bool ok = ch < nch;
        const float* q = dpart + ((((int64_t)n * nch + (ok ? ch : 0)) * T + t) * 3) * C + c;
#pragma unroll
        for (int k = 0; k < 3; ++k) v[u][k] = ok ? __ldcg(reinterpret_cast<const float4*>(q + (int64_t)k * C)) : z;
      }
#pragma unroll
      for (int u = 0; u < kB; ++u)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          d[k].x += v[u][k].x; d[k].y += v[u][k].y; d[k].z += v[u][k].z; d[k].w += v[u][k].w;
        }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) sd[s][k][lane] = d[k];
  __syncthreads();
  if (on && s < 3) {      // slice-threads 0..2 each own one tap k = s: total over the slices, in slice order
    float4 g = z;
#pragma unroll
    for (int q = 0; q < kTamFinSlices; ++q) {
      const float4 v = sd[q][s][lane];
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
    }
    sd[0][s][lane] = g;       // only this thread reads sd[*][s][lane] in the loop above
    const float4 a = ldg4(act + ((int64_t)n * T + t) * C + c);
    st4(dpart + ((((int64_t)n * nch) * T + t) * 3 + s) * C + c, mul4(a, g));
  }
  __syncthreads();
  if (on && s == 0) {
    const float4 k0 = ldg4(kern + ((int64_t)n * 3 + 0) * C + c), k1 = ldg4(kern + ((int64_t)n * 3 + 1) * C + c),
                 k2 = ldg4(kern + ((int64_t)n * 3 + 2) * C + c);
    float4 ga = mul4(k0, sd[0][0][lane]);
    ga = fma4(k1, sd[0][1][lane], ga);
    ga = fma4(k2, sd[0][2][lane], ga);
    st4(gact + ((int64_t)n * T + t) * C + c, ga);
  }
  __threadfence();
  __syncthreads();
  int* ticket = tickets + n * gridDim.x + blockIdx.x;
  if (lane == 0 && s == 0) s_last = (atomicAdd(ticket, 1) == T - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (on && s < 3) {
    float4 v[kTamFinMaxT];
#pragma unroll
    for (int tt = 0; tt < kTamFinMaxT; ++tt)
      v[tt] = tt < T ? __ldcg(reinterpret_cast<const float4*>(dpart + ((((int64_t)n * nch) * T + tt) * 3 + s) * C + c)) : z;
    float4 g = z;
#pragma unroll
    for (int tt = 0; tt < kTamFinMaxT; ++tt) { g.x += v[tt].x; g.y += v[tt].y; g.z += v[tt].z; g.w += v[tt].w; }
    st4(gkern + ((int64_t)n * 3 + s) * C + c, g);
  }
  if (lane == 0 && s == 0) *ticket = 0;
}

}  // namespace vitta

using namespace vitta;

extern "C" {

static int tam_fwd_impl(const float* x, const float* kern, const float* act, float* out, int N, int T, int64_t HW, int C,
                        float* amax_out, void* stream) {
  VITTA_CHECK_ARG(x && kern && act && out, VITTA_E_BADARG, "tam_fwd: null pointer");
  VITTA_CHECK_ARG(N > 0 && T > 0 && HW > 0 && C > 0 && C % 4 == 0, VITTA_E_BADARG, "tam_fwd: bad shape (C %% 4 != 0?)");
  VITTA_CHECK_ARG(aligned16(x) && aligned16(kern) && aligned16(act) && aligned16(out), VITTA_E_ALIGN,
                  "tam_fwd: tensors must be 16-byte aligned");
  const int64_t total = (int64_t)N * HW * (C / 4);
  const int64_t blocks = (total + kThreads - 1) / kThreads;
  VITTA_CHECK_ARG(blocks < (1ll << 31), VITTA_E_UNSUPPORTED, "tam_fwd: grid too large");
  if (amax_out)
    tam_fwd_kernel<true><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(x, kern, act, out, N, T, HW, C / 4,
                                                                                 amax_out);
  else
    tam_fwd_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(x, kern, act, out, N, T, HW, C / 4);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_tam_fwd(const float* x, const float* kern, const float* act, float* out, int N, int T, int64_t HW, int C,
                  void* stream) {
  return tam_fwd_impl(x, kern, act, out, N, T, HW, C, nullptr, stream);
}

int vitta_tam_fwd_amax(const float* x, const float* kern, const float* act, float* out, int N, int T, int64_t HW, int C,
                       float* amax_out, void* stream) {
  VITTA_CHECK_ARG(amax_out, VITTA_E_BADARG, "tam_fwd_amax: amax_out is required");
  return tam_fwd_impl(x, kern, act, out, N, T, HW, C, amax_out, stream);
}

int vitta_tam_num_chunks(int64_t HW, int C) {
  if (HW <= 0 || C < 4) return 0;
  const int rows = tam_chunk_rows(HW, C);
  return (int)((HW + rows - 1) / rows);
}

int vitta_tam_bwd(const float* gout, const float* x, const float* kern, const float* act, float* gx, float* dpart,
                  int N, int T, int64_t HW, int C, void* stream) {
  VITTA_CHECK_ARG(gout && x && kern && act && gx && dpart, VITTA_E_BADARG, "tam_bwd: null pointer");
  VITTA_CHECK_ARG(N > 0 && T > 0 && HW > 0 && C > 0 && C % 4 == 0, VITTA_E_BADARG, "tam_bwd: bad shape");
  VITTA_CHECK_ARG(aligned16(gout) && aligned16(x) && aligned16(gx) && aligned16(dpart) && aligned16(kern) &&
                      aligned16(act),
                  VITTA_E_ALIGN, "tam_bwd: tensors must be 16-byte aligned");
  ClGeom g = cl_geom(1, 1, C);  // only lpr / rs / ctiles are used
  const int nchunks = vitta_tam_num_chunks(HW, C);
  dim3 grid((unsigned)(N * nchunks), (unsigned)g.ctiles);
  tam_bwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(gout, x, kern, act, gx, dpart, N, T, HW, C, g.lpr, g.rs,
                                                             nchunks, tam_chunk_rows(HW, C));
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_tam_bwd_finish(float* dpart, const float* kern, const float* act, float* gkern, float* gact, int* tickets, int N,
                         int T, int nch, int C, void* stream) {
  VITTA_CHECK_ARG(dpart && kern && act && gkern && gact && tickets, VITTA_E_BADARG, "tam_bwd_finish: null pointer");
  VITTA_CHECK_ARG(N > 0 && T > 0 && nch > 0 && C > 0 && C % 4 == 0, VITTA_E_BADARG, "tam_bwd_finish: bad shape");
  VITTA_CHECK_ARG(aligned16(dpart) && aligned16(kern) && aligned16(act) && aligned16(gkern) && aligned16(gact), VITTA_E_ALIGN,
                  "tam_bwd_finish: tensors must be 16-byte aligned");
  VITTA_CHECK_ARG(T >= 3 && T <= kTamFinMaxT, VITTA_E_UNSUPPORTED, "tam_bwd_finish: 3 <= T <= 16");
  const int C4 = C / 4;
  const int ctiles = (C4 + 31) / 32;
  VITTA_CHECK_ARG((int64_t)N * ctiles <= vitta_tam_bwd_finish_tickets(), VITTA_E_UNSUPPORTED,
                  "tam_bwd_finish: more (video, channel tile) pairs than tickets");
  VITTA_CHECK_ARG(N <= 65535, VITTA_E_UNSUPPORTED, "tam_bwd_finish: grid too large");
  dim3 block(32, kTamFinSlices), grid((unsigned)ctiles, (unsigned)T, (unsigned)N);
  tam_bwd_finish_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(dpart, kern, act, gkern, gact, tickets, N, T, nch, C4);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_tam_bwd_finish_tickets(void) { return 4096; }

}  // extern "C"
