// K1/K2/K3: per-channel spatio-temporal statistics, merge + EMA + alignment loss, standalone backward.
// HBM-bound: the feature is read exactly once with 128-bit loads (channels-last) / coalesced rows (NCHW).
#include <stdarg.h>

#include "common.cuh"

namespace vitta {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// K1, channels-last: CTA = (chunk e, channel tile).  Thread = one float4 of channels x every rs-th row.
// Sums are taken about a per-chunk shift K (first row of the chunk) so that M2 = S2 - S1^2/n is
// well conditioned in fp32; chunks are merged later with Chan's formula.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) stats_cl_kernel(const float* __restrict__ x, float* __restrict__ part, int C,
                                                           int lpr, int rs, int chunk_rows, int cpf,
                                                           int64_t frame_rows) {
  __shared__ float4 sm1[kThreads];
  __shared__ float4 sm2[kThreads];
  const int tid = threadIdx.x;
  const int lane = tid % lpr;
  const int slot = tid / lpr;
  const int col4 = blockIdx.y * lpr + lane;
  const bool active = col4 * 4 < C;
  const int64_t e = blockIdx.x;
  const int64_t frame = e / cpf;
  const int j = (int)(e % cpf);
  const int64_t row0 = frame * frame_rows + (int64_t)j * chunk_rows;
  int64_t rem = frame_rows - (int64_t)j * chunk_rows;
  const int nrows = (int)(rem < chunk_rows ? rem : chunk_rows);

  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, k = s1;
  if (active) {
    const float* base = x + row0 * C + (int64_t)col4 * 4;
    k = ldg4(base);
    auto acc = [&](float4 v) {
      float d;
      d = v.x - k.x; s1.x += d; s2.x = fmaf(d, d, s2.x);
      d = v.y - k.y; s1.y += d; s2.y = fmaf(d, d, s2.y);
      d = v.z - k.z; s1.z += d; s2.z = fmaf(d, d, s2.z);
      d = v.w - k.w; s1.w += d; s2.w = fmaf(d, d, s2.w);
    };
    // batches of 4 rows, software-pipelined over two register sets: the loads of batch i+1 are in flight while batch i is
    // accumulated (a dynamic-trip-count loop keeps one load in flight per thread, see bn_act.cu); rows past the chunk's
    // end are predicated off
    constexpr int kB = 4;
    const int64_t rstep = (int64_t)rs * C;
    auto load = [&](int r0, float4* v) {
#pragma unroll
      for (int j = 0; j < kB; ++j)
        if (r0 + j * rs < nrows) v[j] = ld_stream4(base + (int64_t)r0 * C + j * rstep);
    };
    auto proc = [&](int r0, const float4* v) {
#pragma unroll
      for (int j = 0; j < kB; ++j)
        if (r0 + j * rs < nrows) acc(v[j]);
    };
    float4 va[kB], vb[kB];
    const int bstep = kB * rs;
    load(slot, va);
    for (int r0 = slot; r0 < nrows; r0 += 2 * bstep) {
      load(r0 + bstep, vb);
      proc(r0, va);
      load(r0 + 2 * bstep, va);
      proc(r0 + bstep, vb);
    }
  }
  sm1[tid] = s1;
  sm2[tid] = s2;
  __syncthreads();
  for (int st = rs >> 1; st > 0; st >>= 1) {
    if (slot < st) {
      float4 a = sm1[tid], b = sm1[tid + st * lpr];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      sm1[tid] = a;
      a = sm2[tid]; b = sm2[tid + st * lpr];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      sm2[tid] = a;
    }
    __syncthreads();
  }
  if (slot == 0 && active) {
    float4 a = sm1[tid], b = sm2[tid];
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
}

// K1, (O, C, I) with I > 1 (NCHW / NCTHW): one warp per (entry, channel); lanes run along the contiguous I.
__global__ void __launch_bounds__(kThreads) stats_oci_kernel(const float* __restrict__ x, float* __restrict__ part,
                                                            int64_t O, int C, int64_t I, int og, int n_entries) {
  const int64_t wg = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wg >= (int64_t)n_entries * C) return;
  const int c = (int)(wg % C);
  const int64_t e = wg / C;
  const int64_t o0 = e * og;
  const int64_t left = O - o0;
  const int no = (int)(left < og ? left : og);
  const float k = __ldg(x + (o0 * C + c) * I);
  float s1 = 0.f, s2 = 0.f;
  for (int o = 0; o < no; ++o) {
    const float* row = x + ((o0 + o) * C + c) * I;
#pragma unroll 4
    for (int64_t i = lane; i < I; i += 32) {
      float d = __ldg(row + i) - k;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    const float n = (float)no * (float)I;
    part[(e * C + c) * 2 + 0] = k + s1 / n;
    part[(e * C + c) * 2 + 1] = fmaxf(s2 - s1 * s1 / n, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// K2: grid = (layers, channel groups of 32).  Thread = (channel, entry slice): the layer's entries are merged
// with Chan's formula by 8 slices in parallel (coalesced float2 loads across the 32 channels), then tree-merged in
// shared memory; 32 threads apply the meter, the loss term and the backward coefficients.  Per-CTA loss partials
// are summed in a fixed order by the last CTA to finish (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int kFinThreads = 256;
constexpr int kFinCh = 32;
constexpr int kFinSlices = kFinThreads / kFinCh;

__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(kFinThreads) stats_finalize_kernel(
    const VittaLayerDesc* __restrict__ descs, int n_layers, const float* __restrict__ part,
    const int32_t* __restrict__ counts, const float* __restrict__ src_mean, const float* __restrict__ src_var,
    float* __restrict__ ema_mean, float* __restrict__ ema_var, float* __restrict__ batch_mean,
    float* __restrict__ batch_var, float* __restrict__ coef_a, float* __restrict__ coef_b, float* __restrict__ loss,
    int merge_only, float* __restrict__ merged, int32_t* __restrict__ merged_counts) {
  __shared__ float s_n[kFinThreads], s_mean[kFinThreads], s_m2[kFinThreads];
  __shared__ int s_last;
  const VittaLayerDesc d = descs[blockIdx.x];
  const int C = d.C;
  const int cs = threadIdx.x % kFinCh, es = threadIdx.x / kFinCh;
  const int c = blockIdx.y * kFinCh + cs;
  const int max_groups = gridDim.y;
  float* loss_part = loss + n_layers + 2;
  float l = 0.f;
  if (blockIdx.y * kFinCh < C) {
    float n = 0.f, mean = 0.f, m2 = 0.f;
    if (c < C) {
      const float* p = part + d.part_off + (int64_t)c * 2;
#pragma unroll 4
      for (int e = es; e < d.n_entries; e += kFinSlices) {
        float cnt;
        if (counts) {
          cnt = (float)counts[d.cnt_off + (int64_t)e * d.cnt_stride];
        } else {
          int64_t rem = d.frame_rows - (int64_t)(e % d.chunks_per_frame) * d.chunk_rows;
          cnt = (float)(rem < d.chunk_rows ? rem : (int64_t)d.chunk_rows);
        }
        const float2 v = __ldg(reinterpret_cast<const float2*>(p + (int64_t)e * d.entry_stride));
        chan_merge(n, mean, m2, cnt, v.x, v.y);
      }
    }
    s_n[threadIdx.x] = n; s_mean[threadIdx.x] = mean; s_m2[threadIdx.x] = m2;
    __syncthreads();
    for (int st = kFinSlices >> 1; st > 0; st >>= 1) {
      if (es < st) {
        const int o = threadIdx.x + st * kFinCh;
        chan_merge(n, mean, m2, s_n[o], s_mean[o], s_m2[o]);
        s_n[threadIdx.x] = n; s_mean[threadIdx.x] = mean; s_m2[threadIdx.x] = m2;
      }
      __syncthreads();
    }
    if (es == 0 && c < C) {
      const int64_t ci = d.ch_off + c;
      if (merge_only) {
        merged[ci * 2 + 0] = mean;
        merged[ci * 2 + 1] = m2;
        if (c == 0) merged_counts[blockIdx.x] = (int32_t)n;
      } else {
        const float var = m2 / n;
        batch_mean[ci] = mean;
        batch_var[ci] = var;
        if (d.has_source) {
          const float em = fmaf(d.w_new, mean, d.w_old * ema_mean[ci]);
          const float ev = fmaf(d.w_new, var, d.w_old * ema_var[ci]);
          ema_mean[ci] = em;
          ema_var[ci] = ev;
          const float sm = src_mean[ci], sv = src_var[ci];
          const float dm = em - sm, dv = ev - sv;
          float gm, gv;
          if (d.reg_type == VITTA_REG_L1) {
            const float invC = 1.f / (float)C;
            l = (fabsf(dv) + fabsf(dm)) * invC;
            gm = sgnf(dm) * invC;
            gv = sgnf(dv) * invC;
          } else if (d.reg_type == VITTA_REG_MSE) {
            const float invC = 1.f / (float)C;
            l = (dv * dv + dm * dm) * invC;
            gm = 2.f * dm * invC;
            gv = 2.f * dv * invC;
          } else {  // KLD: true = source, pred = ema (norm_stats_utils.py:8-16); summed over channels
            const float q = sv + dm * dm;
            l = 0.5f * logf(ev / sv) + q / (2.f * ev) - 0.5f;
            gm = dm / ev;
            gv = 0.5f / ev - q / (2.f * ev * ev);
          }
          // dLoss/dy = coef_a + coef_b * (y - batch_mean): the centred form avoids the cancellation of a' + b*y
          coef_b[ci] = d.w_new * gv * 2.f / n;
          coef_a[ci] = d.w_new * gm / n;
        }
      }
    }
  }
  if (merge_only) return;
  // loss partial of this CTA: fixed shuffle tree over its 32 channels (warp 0 holds them: es == 0)
  if (threadIdx.x < 32) {
    l = warp_sum(l);
    if (threadIdx.x == 0) {
      loss_part[(int64_t)blockIdx.x * max_groups + blockIdx.y] = l;
      __threadfence();
      int* ticket = reinterpret_cast<int*>(loss + n_layers + 1);
      s_last = (atomicAdd(ticket, 1) == (int)(gridDim.x * gridDim.y) - 1);
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int ly = threadIdx.x; ly < n_layers; ly += kFinThreads) {
    const int groups = (descs[ly].C + kFinCh - 1) / kFinCh;
    float t = 0.f;
    for (int g = 0; g < groups; ++g) t += __ldcg(loss_part + (int64_t)ly * max_groups + g);
    loss[ly] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int ly = 0; ly < n_layers; ++ly) t += loss[ly];
    loss[n_layers] = t;
    *reinterpret_cast<int*>(loss + n_layers + 1) = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// K3: gy = gscale * (a[c] + b[c]*(y - mean[c])), y = x or yscale[c]*x + yshift[c]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) inject_cl_kernel(const float* __restrict__ x, const float* __restrict__ ys,
                                                            const float* __restrict__ yt, const float* __restrict__ ca,
                                                            const float* __restrict__ cb, const float* __restrict__ cm,
                                                            const float* __restrict__ gscale, float* __restrict__ gy,
                                                            int64_t n4, int C4) {
  const float g = __ldg(gscale);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(i % C4) * 4;
    float4 v = ld_stream4(x + i * 4);
    if (ys) {
      const float4 s = ldg4(ys + c), t = ldg4(yt + c);
      v.x = fmaf(v.x, s.x, t.x); v.y = fmaf(v.y, s.y, t.y); v.z = fmaf(v.z, s.z, t.z); v.w = fmaf(v.w, s.w, t.w);
    }
    const float4 a = ldg4(ca + c), b = ldg4(cb + c), m = ldg4(cm + c);
    float4 o;
    o.x = g * fmaf(b.x, v.x - m.x, a.x); o.y = g * fmaf(b.y, v.y - m.y, a.y);
    o.z = g * fmaf(b.z, v.z - m.z, a.z); o.w = g * fmaf(b.w, v.w - m.w, a.w);
    st4(gy + i * 4, o);
  }
}

__global__ void __launch_bounds__(kThreads) inject_oci_kernel(const float* __restrict__ x, const float* __restrict__ ys,
                                                             const float* __restrict__ yt, const float* __restrict__ ca,
                                                             const float* __restrict__ cb, const float* __restrict__ cm,
                                                             const float* __restrict__ gscale, float* __restrict__ gy,
                                                             int64_t n, int C, int64_t I) {
  const float g = __ldg(gscale);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const int c = (int)((i / I) % C);
    float v = __ldg(x + i);
    if (ys) v = fmaf(v, __ldg(ys + c), __ldg(yt + c));
    gy[i] = g * fmaf(__ldg(cb + c), v - __ldg(cm + c), __ldg(ca + c));
  }
}

}  // namespace vitta

using namespace vitta;

extern "C" {

int vitta_version(void) { return 100; }
const char* vitta_last_error(void) { return g_err; }

int vitta_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

int vitta_stats_chunking(int64_t O, int C, int64_t I, int64_t frames, VittaChunking* out) {
  VITTA_CHECK_ARG(out && O > 0 && C > 0 && I > 0, VITTA_E_BADARG, "stats_chunking: bad shape O=%lld C=%d I=%lld",
                  (long long)O, C, (long long)I);
  if (I == 1 && C % 4 == 0) {
    if (frames < 1) frames = 1;
    VITTA_CHECK_ARG(O % frames == 0, VITTA_E_BADARG, "stats_chunking: frames=%lld does not divide rows=%lld",
                    (long long)frames, (long long)O);
    ClGeom g = cl_geom(frames, O / frames, C);
    out->chunk_rows = g.chunk_rows;
    out->chunks_per_frame = g.cpf;
    out->frame_rows = g.frame_rows;
    VITTA_CHECK_ARG(g.n_chunks() < (1ll << 31), VITTA_E_UNSUPPORTED, "too many chunks");
    out->n_entries = (int32_t)g.n_chunks();
  } else {
    OciGeom g = oci_geom(O, I);
    VITTA_CHECK_ARG((int64_t)g.og * I < (1ll << 31), VITTA_E_UNSUPPORTED, "inner extent too large");
    out->chunk_rows = (int32_t)(g.og * I);
    out->chunks_per_frame = g.n_entries;
    out->frame_rows = O * I;
    out->n_entries = g.n_entries;
  }
  out->reserved = 0;
  return 0;
}

int vitta_stats_partial(const float* x, int64_t O, int C, int64_t I, int64_t frames, float* part, void* stream) {
  VITTA_CHECK_ARG(x && part, VITTA_E_BADARG, "stats_partial: null pointer");
  VittaChunking ch;
  int rc = vitta_stats_chunking(O, C, I, frames, &ch);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (I == 1 && C % 4 == 0) {
    VITTA_CHECK_ARG(aligned16(x) && aligned16(part), VITTA_E_ALIGN, "stats_partial: pointers must be 16-byte aligned");
    ClGeom g = cl_geom(frames < 1 ? 1 : frames, O / (frames < 1 ? 1 : frames), C);
    dim3 grid((unsigned)g.n_chunks(), (unsigned)g.ctiles);
    stats_cl_kernel<<<grid, kThreads, 0, st>>>(x, part, C, g.lpr, g.rs, g.chunk_rows, g.cpf, g.frame_rows);
  } else {
    OciGeom g = oci_geom(O, I);
    int64_t warps = (int64_t)g.n_entries * C;
    int64_t blocks = (warps + kThreads / 32 - 1) / (kThreads / 32);
    VITTA_CHECK_ARG(blocks < (1ll << 31), VITTA_E_UNSUPPORTED, "grid too large");
    stats_oci_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(x, part, O, C, I, g.og, g.n_entries);
  }
  VITTA_CHECK_LAUNCH();
  return 0;
}

int64_t vitta_stats_finalize_loss_floats(int n_layers, int max_channels) {
  if (n_layers <= 0 || max_channels <= 0) return -1;
  return (int64_t)n_layers + 2 + (int64_t)n_layers * ((max_channels + kFinCh - 1) / kFinCh);
}

int vitta_stats_finalize(const VittaLayerDesc* descs, int n_layers, const float* part, const int32_t* counts,
                         const float* src_mean, const float* src_var, float* ema_mean, float* ema_var,
                         float* batch_mean, float* batch_var, float* coef_a, float* coef_b, float* loss,
                         int merge_only, float* merged, int32_t* merged_counts, int max_channels, void* stream) {
  VITTA_CHECK_ARG(descs && part && n_layers > 0 && max_channels > 0, VITTA_E_BADARG, "stats_finalize: bad arguments");
  if (merge_only) {
    VITTA_CHECK_ARG(merged && merged_counts, VITTA_E_BADARG, "stats_finalize: merge_only needs merged buffers");
  } else {
    VITTA_CHECK_ARG(batch_mean && batch_var && loss, VITTA_E_BADARG, "stats_finalize: null output");
  }
  dim3 grid((unsigned)n_layers, (unsigned)((max_channels + kFinCh - 1) / kFinCh));
  stats_finalize_kernel<<<grid, kFinThreads, 0, (cudaStream_t)stream>>>(
      descs, n_layers, part, counts, src_mean, src_var, ema_mean, ema_var, batch_mean, batch_var, coef_a, coef_b, loss,
      merge_only, merged, merged_counts);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_stats_inject(const float* x, const float* yscale, const float* yshift, const float* coef_a,
                       const float* coef_b, const float* mean, const float* gscale, float* gy, int64_t O, int C,
                       int64_t I, void* stream) {
  VITTA_CHECK_ARG(x && coef_a && coef_b && mean && gscale && gy, VITTA_E_BADARG, "stats_inject: null pointer");
  VITTA_CHECK_ARG((yscale == nullptr) == (yshift == nullptr), VITTA_E_BADARG, "stats_inject: scale/shift mismatch");
  const int64_t n = O * C * I;
  const int sms = vitta_sm_count() > 0 ? vitta_sm_count() : 148;
  if (I == 1 && C % 4 == 0 && aligned16(x) && aligned16(gy)) {
    int64_t n4 = n / 4;
    int64_t blocks = (n4 + kThreads - 1) / kThreads;
    if (blocks > sms * 16) blocks = sms * 16;
    inject_cl_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(x, yscale, yshift, coef_a, coef_b, mean,
                                                                             gscale, gy, n4, C / 4);
  } else {
    int64_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > sms * 16) blocks = sms * 16;
    inject_oci_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(x, yscale, yshift, coef_a, coef_b,
                                                                              mean, gscale, gy, n, C, I);
  }
  VITTA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
