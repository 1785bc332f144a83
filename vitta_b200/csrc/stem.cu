// ResNet-50 stem of TANet (reference models/tanet_models/tanet.py:129: torchvision conv1 / bn1 / relu / maxpool):
//   * vitta_stem_pack          image (F, 3, H, W) -> zero-padded 4-channel channels-last (F, H+6, W+6, 4), the operand
//                              layout of vitta_stem_conv_tf32x3 (gemm_tf32.cu)
//   * vitta_stem_pack_weight   conv1 weight (64, 3, 7, 7) -> tf32 hi/lo [64][7 kh][8 kw][4 c] (zeros in the padding)
//   * vitta_bn_relu_pool_fwd   eval-mode BN + ReLU + MaxPool(3, 2, 1) in one pass: reads the conv output once
//                              (411 MB at the bench shape), writes the pooled map (103 MB) and a one-byte argmax code;
//                              the un-pooled activation is never materialised
//   * vitta_bn_relu_pool_bwd   its backward: routes the pooled gradient through the recorded argmax (PyTorch's rule:
//                              first maximum in scan order), the ReLU mask and the BN affine map, and reduces the BN
//                              parameter gradients deterministically (per-CTA partials, fixed-order final sum)
// All HBM-bound; channels-last, 128-bit accesses along C = 64.
#include "common.cuh"

namespace vitta {

__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ x, float4* __restrict__ xp, int F,
                                                       int H, int W) {
  const int Hp = H + 6, Wp = W + 6;
  const int64_t total = (int64_t)F * Hp * Wp;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int wp = (int)(i % Wp);
    const int64_t q = i / Wp;
    const int hp = (int)(q % Hp);
    const int64_t f = q / Hp;
    const int h = hp - 3, w = wp - 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const int64_t plane = (int64_t)H * W;
      const float* b = x + f * 3 * plane + (int64_t)h * W + w;
      v.x = __ldg(b);
      v.y = __ldg(b + plane);
      v.z = __ldg(b + 2 * plane);
    }
    xp[i] = v;
  }
}

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void stem_pack_weight_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // dst [co][kh][kw8][c4]
  if (i >= 64 * 224) return;
  const int c = i & 3, kw = (i >> 2) & 7, kh = (i >> 5) % 7, co = i / 224;
  float v = 0.f;
  if (c < 3 && kw < 7) v = __ldg(w + ((co * 3 + c) * 7 + kh) * 7 + kw);
  const float h = tf32_round(v);
  hi[i] = h;
  lo[i] = v - h;
}

struct StemBN {
  const float *w, *b, *rm, *rv;
  float eps;
};

// thread = (pooled pixel, 4 channels).  Window rows 2ho-1 .. 2ho+1, cols 2wo-1 .. 2wo+1, scanned h-major like PyTorch;
// the first maximum wins (strict >).  code = 3*dh + dw of the winner.
// I = uint32_t when the flat index fits (the index decomposition is three divisions per thread and iteration, and 64-bit
// ones cost more issue slots than the nine loads they address).
// amax_out (optional): max|out| accumulates there -- the operand range the fp16-split convolutions of layer1.0 need.
template <typename I>
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_kernel(const float* __restrict__ x, StemBN bn,
                                                              float* __restrict__ out, uint8_t* __restrict__ code,
                                                              int F, int H, int W, int C4,
                                                              float* __restrict__ amax_out) {
  float am = 0.f;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const I total = (I)F * Ho * Wo * C4;
  const int C = C4 * 4;
  for (I i = (I)blockIdx.x * 256 + threadIdx.x; i < total; i += (I)gridDim.x * 256) {
    const int c = (int)(i % (I)C4) * 4;
    I q = i / (I)C4;
    const int wo = (int)(q % (I)Wo); q /= (I)Wo;
    const int ho = (int)(q % (I)Ho);
    const int64_t f = (int64_t)(q / (I)Ho);
    const float4 w = ldg4(bn.w + c), b = ldg4(bn.b + c), rm = ldg4(bn.rm + c), rv = ldg4(bn.rv + c);
    float4 k;
    k.x = w.x * (1.f / sqrtf(rv.x + bn.eps)); k.y = w.y * (1.f / sqrtf(rv.y + bn.eps));
    k.z = w.z * (1.f / sqrtf(rv.z + bn.eps)); k.w = w.w * (1.f / sqrtf(rv.w + bn.eps));
    float4 v[9];
    bool ok[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int h = 2 * ho - 1 + t / 3, ww = 2 * wo - 1 + t % 3;
      ok[t] = h >= 0 && h < H && ww >= 0 && ww < W;
      v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[t]) v[t] = ldg4(x + ((f * H + h) * W + ww) * C + c);
    }
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uint32_t cx = 0, cy = 0, cz = 0, cw = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (!ok[t]) continue;
      // The ReLU is applied to the maximum, not to the nine candidates: when the largest y is positive the first maximum
      // of y and of relu(y) coincide; when it is not, the output is 0 either way and the recorded position is irrelevant
      // -- the backward masks that pixel's gradient with y > 0.
      const float yx = fmaf(v[t].x - rm.x, k.x, b.x), yy = fmaf(v[t].y - rm.y, k.y, b.y);
      const float yz = fmaf(v[t].z - rm.z, k.z, b.z), yw = fmaf(v[t].w - rm.w, k.w, b.w);
      if (yx > m.x) { m.x = yx; cx = t; }
      if (yy > m.y) { m.y = yy; cy = t; }
      if (yz > m.z) { m.z = yz; cz = t; }
      if (yw > m.w) { m.w = yw; cw = t; }
    }
    m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f); m.z = fmaxf(m.z, 0.f); m.w = fmaxf(m.w, 0.f);
    st4(out + (int64_t)i * 4, m);
    reinterpret_cast<uint32_t*>(code)[i] = cx | (cy << 8) | (cz << 16) | (cw << 24);
    am = fmaxf(fmaxf(am, fmaxf(m.x, m.y)), fmaxf(m.z, m.w));      // out >= 0
  }
  if (amax_out) {   // non-negative floats order like their bit patterns; the loop above has re-converged the warp
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
    if ((threadIdx.x & 31) == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(amax_out), wmax);
  }
}

// CTA = 256 threads = 16 column slots x 16 channel quads (C = 64).  A thread handles a 2 x 2 block of input pixels
// (rows 2i, 2i+1; columns 2j, 2j+1): the block lies inside the four windows (i..i+1) x (j..j+1) and no other, so four
// pooled-gradient / code loads serve four pixels (a pixel on its own needs 1, 2, 2 or 4 of them: nine per block), and
// every window position is a compile-time constant.  A CTA owns a contiguous range of row pairs (the pooled row i+1 it
// gathers from is in its L1 again for the next pair); all twelve loads of a block are issued before the first use.
// Per pixel: sum the pooled gradients of the windows whose recorded winner it is, apply the ReLU mask and the BN map.
__global__ void __launch_bounds__(256) bn_relu_pool_bwd_kernel(const float* __restrict__ gpool,
                                                              const uint8_t* __restrict__ code,
                                                              const float* __restrict__ x, StemBN bn,
                                                              float* __restrict__ gx, float* __restrict__ ws,
                                                              float* __restrict__ gw, float* __restrict__ gb, int F,
                                                              int H, int W, int C4, int ticket_off) {
  __shared__ float4 sm[256];
  __shared__ int s_last;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int C = C4 * 4;
  const int lane = threadIdx.x % C4, slot = threadIdx.x / C4, slots = 256 / C4;
  const int c = lane * 4;
  const float4 w = ldg4(bn.w + c), b = ldg4(bn.b + c), rm = ldg4(bn.rm + c), rv = ldg4(bn.rv + c);
  float4 k, istd;
  istd.x = 1.f / sqrtf(rv.x + bn.eps); istd.y = 1.f / sqrtf(rv.y + bn.eps);
  istd.z = 1.f / sqrtf(rv.z + bn.eps); istd.w = 1.f / sqrtf(rv.w + bn.eps);
  k.x = w.x * istd.x; k.y = w.y * istd.y; k.z = w.z * istd.z; k.w = w.w * istd.w;
  float4 agw = make_float4(0.f, 0.f, 0.f, 0.f), agb = agw;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t npairs = (int64_t)F * Ho;
  const int64_t ppc = (npairs + gridDim.x - 1) / gridDim.x;
  const int64_t pr0 = (int64_t)blockIdx.x * ppc, pr1 = pr0 + ppc < npairs ? pr0 + ppc : npairs;
  for (int64_t pr = pr0; pr < pr1; ++pr) {
    const int i = (int)(pr % Ho);
    const int64_t f = pr / Ho;
    const bool row1 = 2 * i + 1 < H, win1 = i + 1 < Ho;        // second pixel row / second window row exist
    const float* xr = x + ((f * H + 2 * i) * W) * C + c;
    float* gr = gx + ((f * H + 2 * i) * W) * C + c;
    const int64_t o0 = ((f * Ho + i) * Wo) * C4 + lane;         // quad index of window (i, 0); + Wo*C4 for row i+1
    for (int j = slot; j < Wo; j += slots) {
      const bool col1 = 2 * j + 1 < W, wcol1 = j + 1 < Wo;
      // pixel (dy, dx) -> xv[2*dy + dx]; window (a, bb) -> gp / cd[2*a + bb]
      float4 xv[4], gp[4];
      uint32_t cd[4];
      const int64_t xo = (int64_t)(2 * j) * C;
      xv[0] = ld_stream4(xr + xo);
      xv[1] = col1 ? ld_stream4(xr + xo + C) : z4;
      xv[2] = row1 ? ld_stream4(xr + xo + (int64_t)W * C) : z4;
      xv[3] = (row1 && col1) ? ld_stream4(xr + xo + (int64_t)W * C + C) : z4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool ok = ((q >> 1) == 0 || win1) && ((q & 1) == 0 || wcol1);
        const int64_t o = o0 + (int64_t)(q >> 1) * Wo * C4 + (int64_t)(j + (q & 1)) * C4;
        cd[q] = 0xffffffffu;                     // matches no window position
        gp[q] = z4;
        if (ok) {
          cd[q] = __ldg(reinterpret_cast<const uint32_t*>(code) + o);
          gp[q] = ldg4(gpool + o * 4);
        }
      }
      // position of pixel (dy, dx) inside window (a, bb): 3 * (dy - 2a + 1) + (dx - 2bb + 1); outside it when negative
      float4 g[4] = {z4, z4, z4, z4};
#pragma unroll
      for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ry = (px >> 1) - 2 * (q >> 1) + 1, rx = (px & 1) - 2 * (q & 1) + 1;
          if (ry < 0 || rx < 0) continue;
          const uint32_t mine = (uint32_t)(3 * ry + rx);
          if ((cd[q] & 0xffu) == mine) g[px].x += gp[q].x;
          if (((cd[q] >> 8) & 0xffu) == mine) g[px].y += gp[q].y;
          if (((cd[q] >> 16) & 0xffu) == mine) g[px].z += gp[q].z;
          if ((cd[q] >> 24) == mine) g[px].w += gp[q].w;
        }
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        if (((px >> 1) && !row1) || ((px & 1) && !col1)) continue;
        const float4 v = xv[px];
        float4 gg = g[px];
        // ReLU mask on y = BN(x)
        const float yx = fmaf(v.x - rm.x, k.x, b.x), yy = fmaf(v.y - rm.y, k.y, b.y);
        const float yz = fmaf(v.z - rm.z, k.z, b.z), yw = fmaf(v.w - rm.w, k.w, b.w);
        gg.x = yx > 0.f ? gg.x : 0.f; gg.y = yy > 0.f ? gg.y : 0.f; gg.z = yz > 0.f ? gg.z : 0.f; gg.w = yw > 0.f ? gg.w : 0.f;
        st4(gr + xo + (int64_t)(px >> 1) * W * C + (px & 1) * C, make_float4(gg.x * k.x, gg.y * k.y, gg.z * k.z, gg.w * k.w));
        agb.x += gg.x; agb.y += gg.y; agb.z += gg.z; agb.w += gg.w;
        agw.x = fmaf(gg.x, (v.x - rm.x) * istd.x, agw.x); agw.y = fmaf(gg.y, (v.y - rm.y) * istd.y, agw.y);
        agw.z = fmaf(gg.z, (v.z - rm.z) * istd.z, agw.z); agw.w = fmaf(gg.w, (v.w - rm.w) * istd.w, agw.w);
      }
    }
  }
  // per-CTA partials -> ws[cta][2][C]; the last CTA sums them in CTA order (deterministic)
  float4 red[2] = {agw, agb};
  for (int r = 0; r < 2; ++r) {
    __syncthreads();
    sm[threadIdx.x] = red[r];
    __syncthreads();
    for (int st = slots >> 1; st > 0; st >>= 1) {
      if (slot < st) {
        float4 a = sm[threadIdx.x], bq = sm[threadIdx.x + st * C4];
        a.x += bq.x; a.y += bq.y; a.z += bq.z; a.w += bq.w;
        sm[threadIdx.x] = a;
      }
      __syncthreads();
    }
    if (slot == 0) st4(ws + ((int64_t)blockIdx.x * 2 + r) * C + c, sm[threadIdx.x]);
  }
  __threadfence();
  __syncthreads();
  int* ticket = reinterpret_cast<int*>(ws + ticket_off);   // fixed slot at the end of the workspace, whatever the grid
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    const int r = i / C, cc = i % C;
    const float s = ordered_sum_strided(ws + (int64_t)r * C + cc, (int)gridDim.x, 2 * (int64_t)C);
    (r == 0 ? gw : gb)[cc] = s;
  }
  if (threadIdx.x == 0) *ticket = 0;
}


// ------------------------------------------------------------------------------------------------
// Stem weight gradient  dW[co][c][kh][kw] = sum over (f, ho, wo) of dY[f, ho, wo, co] * XP[f, 2ho + kh, 2wo + kw, c]
// (cuDNN's NHWC wgrad engine before: 1.49 ms at the bench shape).  64 x 147 outputs reduced over 1.6 M pixels: a
// register-blocked fp32 kernel -- exact fp32 products, no operand split needed.  CTA = 64 consecutive output pixels per
// tile: the dY rows (64 x 64) and the compacted 7 x 7 x 3 input windows (64 x 147, padded to 156) are staged in shared
// memory.  Thread (pixel half, co-group of 8, k-group of 10) keeps an 8 x 10 accumulator block over its 32 pixels of every
// tile as 40 packed pairs and updates it with FFMA2 (fma.rn.f32x2): the three-register FFMA issues every other cycle per
// scheduler on this part, so a scalar-FFMA kernel tops out at half the fp32 peak (the first version: 28 TFLOP/s, 1.08 ms);
// two 128-bit loads of dY + five 64-bit loads of the window feed 40 FFMA2 (the loads are broadcast / conflict-free: dY is stored as
// [pixel][half of the co-group][co-group][4]).  Per-CTA partials go to a workspace and are summed in CTA order
// (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int kSwP = 64;            // pixels per tile
constexpr int kSwKG = 10;           // window values per k-group (five packed pairs)
constexpr int kSwGroups = 15;       // 147 window values in 15 groups of 10
constexpr int kSwK = 152;           // row length of the window buffer / the partials: 150 used + 2 (16-byte rows)
constexpr int kSwThreads = 256;     // 2 pixel halves x 8 co-groups x 15 k-groups = 240 compute threads; all 256 stage
constexpr int kSwCompute = 8 * kSwGroups;     // compute threads per pixel half

__global__ void __launch_bounds__(kSwThreads, 2) stem_wgrad_kernel(const float4* __restrict__ xp,
                                                                  const float* __restrict__ gy, float* __restrict__ ws,
                                                                  int F, int H, int W) {
  extern __shared__ __align__(16) float stem_smem[];
  float* sA = stem_smem;                                                 // [64 pixels][2][8][4]
  float (*sB)[kSwK] = reinterpret_cast<float (*)[kSwK]>(stem_smem + kSwP * 64);
  const int Hp = H + 6, Wp = W + 6, Ho = H / 2, Wo = W / 2;
  const int64_t npix = (int64_t)F * Ho * Wo;
  const int64_t ntiles = (npix + kSwP - 1) / kSwP;
  const int tid = threadIdx.x;
  const bool compute = tid < 2 * kSwCompute;
  const int ph = tid / kSwCompute, t = tid - ph * kSwCompute;
  const int cg = t & 7, kg = t >> 3;               // co-group (8 channels), k-group (10 window values)
  float2 acc[8][kSwKG / 2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < kSwKG / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);
  // zero the padding columns once (147..151)
  for (int i = tid; i < kSwP * (kSwK - 147); i += kSwThreads) sB[i / (kSwK - 147)][147 + i % (kSwK - 147)] = 0.f;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * kSwP;
    __syncthreads();                               // the previous tile has been consumed
    // dY rows: 64 pixels x 16 float4; float4 q of a row (channels 4q .. 4q+3) goes to [half = q & 1][co-group = q >> 1]
    for (int i = tid; i < kSwP * 16; i += kSwThreads) {
      const int p = i >> 4, q = i & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p0 + p < npix) v = ld_stream4(gy + (p0 + p) * 64 + q * 4);
      *reinterpret_cast<float4*>(sA + p * 64 + (q & 1) * 32 + (q >> 1) * 4) = v;
    }
    // input windows: (pixel, filter row) pairs, seven 4-channel pixels each, compacted to 3 channels
    for (int i = tid; i < kSwP * 7; i += kSwThreads) {
      const int p = i / 7, kh = i - p * 7;
      const int64_t pix = p0 + p;
      float* dst = &sB[p][kh * 21];
      if (pix < npix) {
        const int wo = (int)(pix % Wo);
        const int64_t tt = pix / Wo;
        const int ho = (int)(tt % Ho);
        const int64_t f = tt / Ho;
        const float4* src = xp + (f * Hp + 2 * ho + kh) * Wp + 2 * wo;
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) {
          const float4 v = __ldg(src + kw);
          dst[kw * 3 + 0] = v.x; dst[kw * 3 + 1] = v.y; dst[kw * 3 + 2] = v.z;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 21; ++j) dst[j] = 0.f;
      }
    }
    __syncthreads();
    if (compute) {
      const float* pa = sA + ph * 32 * 64 + cg * 4;
      const float* pb = &sB[ph * 32][kg * kSwKG];
#pragma unroll 2
      for (int p = 0; p < kSwP / 2; ++p) {
        const float4 a0 = *reinterpret_cast<const float4*>(pa + p * 64);
        const float4 a1 = *reinterpret_cast<const float4*>(pa + p * 64 + 32);
        float2 bp[kSwKG / 2];
#pragma unroll
        for (int j = 0; j < kSwKG / 2; ++j) bp[j] = *reinterpret_cast<const float2*>(pb + p * kSwK + 2 * j);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 a2 = make_float2(av[i], av[i]);
#pragma unroll
          for (int j = 0; j < kSwKG / 2; ++j) acc[i][j] = __ffma2_rn(a2, bp[j], acc[i][j]);
        }
      }
    }
  }
  // the two pixel halves of a (co-group, k-group) are added: half 1 parks its block in shared memory (the window buffer),
  // half 0 adds it to its own in a fixed order
  __syncthreads();
  float* park = stem_smem + kSwP * 64;
  static_assert(kSwCompute * 8 * kSwKG <= kSwP * kSwK, "parked accumulators fit the window buffer");
  if (compute && ph == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < kSwKG / 2; ++j)
        *reinterpret_cast<float2*>(park + (i * (kSwKG / 2) + j) * 2 * kSwCompute + t * 2) = acc[i][j];
  }
  __syncthreads();
  if (compute && ph == 0) {
    float* o = ws + (int64_t)blockIdx.x * 64 * kSwK;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < kSwKG / 2; ++j) {
        const float2 other = *reinterpret_cast<const float2*>(park + (i * (kSwKG / 2) + j) * 2 * kSwCompute + t * 2);
        *reinterpret_cast<float2*>(o + (cg * 8 + i) * kSwK + kg * kSwKG + 2 * j) =
            make_float2(acc[i][j].x + other.x, acc[i][j].y + other.y);
      }
  }
}

// dW[co][c][kh][kw] = sum over the CTAs (in order) of ws[cta][co][kh*21 + kw*3 + c]
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int nctas) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // dst index [co][c][kh][kw]
  if (i >= 64 * 147) return;
  const int kw = i % 7, kh = (i / 7) % 7, c = (i / 49) % 3, co = i / 147;
  const int k = kh * 21 + kw * 3 + c;
  float s = 0.f;
  for (int b0 = 0; b0 < nctas; b0 += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (b0 + u < nctas) ? __ldg(ws + ((int64_t)(b0 + u) * 64 + co) * kSwK + k) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  dw[i] = s;
}

static int stem_grid() {
  int sms = vitta_sm_count();
  if (sms <= 0) sms = 148;
  return sms * 4;
}

}  // namespace vitta

using namespace vitta;

extern "C" {

int vitta_stem_pack(const float* x, float* xp, int F, int H, int W, void* stream) {
  VITTA_CHECK_ARG(x && xp && F > 0 && H > 0 && W > 0, VITTA_E_BADARG, "stem_pack: bad arguments");
  VITTA_CHECK_ARG(aligned16(xp), VITTA_E_ALIGN, "stem_pack: output must be 16-byte aligned");
  const int64_t total = (int64_t)F * (H + 6) * (W + 6);
  int64_t blocks = (total + 255) / 256;
  if (blocks > stem_grid() * 8) blocks = stem_grid() * 8;
  stem_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<float4*>(xp), F, H, W);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_stem_pack_weight(const float* w, float* hi, float* lo, void* stream) {
  VITTA_CHECK_ARG(w && hi && lo, VITTA_E_BADARG, "stem_pack_weight: null pointer");
  stem_pack_weight_kernel<<<(64 * 224 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, hi, lo);
  VITTA_CHECK_LAUNCH();
  return 0;
}

static int bn_relu_pool_fwd_impl(const float* x, VittaBN bn, float* out, uint8_t* code, int F, int H, int W, int C,
                                 float* amax_out, void* stream) {
  VITTA_CHECK_ARG(x && out && code && bn.weight && bn.bias && bn.running_mean && bn.running_var, VITTA_E_BADARG,
                  "bn_relu_pool_fwd: null pointer");
  VITTA_CHECK_ARG(F > 0 && H > 1 && W > 1 && C > 0 && C % 4 == 0, VITTA_E_BADARG, "bn_relu_pool_fwd: bad shape");
  VITTA_CHECK_ARG(aligned16(x) && aligned16(out) && (reinterpret_cast<uintptr_t>(code) & 3u) == 0, VITTA_E_ALIGN,
                  "bn_relu_pool_fwd: alignment");
  const int64_t total = (int64_t)F * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > stem_grid() * 4) blocks = stem_grid() * 4;
  StemBN b{bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps};
  if (total + (int64_t)blocks * 256 < (int64_t)1 << 32)
    bn_relu_pool_fwd_kernel<uint32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, b, out, code, F, H, W, C / 4,
                                                                                        amax_out);
  else
    bn_relu_pool_fwd_kernel<int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, b, out, code, F, H, W, C / 4,
                                                                                       amax_out);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_bn_relu_pool_fwd(const float* x, VittaBN bn, float* out, uint8_t* code, int F, int H, int W, int C,
                           void* stream) {
  return bn_relu_pool_fwd_impl(x, bn, out, code, F, H, W, C, nullptr, stream);
}

// ... and max|out| accumulated into *amax_out (a device scalar the caller zeroed)
int vitta_bn_relu_pool_fwd_amax(const float* x, VittaBN bn, float* out, uint8_t* code, int F, int H, int W, int C,
                                float* amax_out, void* stream) {
  VITTA_CHECK_ARG(amax_out, VITTA_E_BADARG, "bn_relu_pool_fwd_amax: amax_out is required");
  return bn_relu_pool_fwd_impl(x, bn, out, code, F, H, W, C, amax_out, stream);
}

int64_t vitta_bn_relu_pool_bwd_ws_floats(int C) {
  if (C <= 0 || C % 4) return -1;
  return (int64_t)stem_grid() * 2 * C + 4;
}

int vitta_bn_relu_pool_bwd(const float* gpool, const uint8_t* code, const float* x, VittaBN bn, float* gx, float* gw,
                           float* gb, float* ws, int F, int H, int W, int C, void* stream) {
  VITTA_CHECK_ARG(gpool && code && x && gx && gw && gb && ws, VITTA_E_BADARG, "bn_relu_pool_bwd: null pointer");
  VITTA_CHECK_ARG(F > 0 && H > 1 && W > 1 && C >= 4 && C % 4 == 0 && 256 % (C / 4) == 0 && C <= 1024, VITTA_E_UNSUPPORTED,
                  "bn_relu_pool_bwd: C / 4 must divide 256");
  VITTA_CHECK_ARG(aligned16(gpool) && aligned16(x) && aligned16(gx) && aligned16(ws), VITTA_E_ALIGN,
                  "bn_relu_pool_bwd: alignment");
  StemBN b{bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps};
  // Row pairs in contiguous ranges; exactly the CTAs that are resident at once (the register count decides: a grid of
  // 4 per SM with 3 resident ran a second, one-third-full wave), never more than the workspace has partial slots for.
  static int per_sm = 0;
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bn_relu_pool_bwd_kernel, 256, 0) != cudaSuccess || n < 1) n = 1;
    per_sm = n;
  }
  int64_t blocks = (int64_t)F * ((H + 1) / 2);
  const int64_t resident = (int64_t)per_sm * (stem_grid() / 4);
  if (blocks > resident) blocks = resident;
  if (blocks > stem_grid()) blocks = stem_grid();
  bn_relu_pool_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(gpool, code, x, b, gx, ws, gw, gb, F, H, W,
                                                                            C / 4, stem_grid() * 2 * C);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int64_t vitta_stem_wgrad_ws_floats(void) { return (int64_t)(stem_grid() / 2) * 64 * kSwK; }

int vitta_stem_wgrad(const float* XP, const float* dY, float* dW, float* ws, int F, int H, int W, void* stream) {
  VITTA_CHECK_ARG(XP && dY && dW && ws && F > 0 && H >= 8 && W >= 8 && H % 2 == 0 && W % 2 == 0, VITTA_E_BADARG,
                  "stem_wgrad: bad arguments");
  VITTA_CHECK_ARG(aligned16(XP) && aligned16(dY) && aligned16(ws), VITTA_E_ALIGN, "stem_wgrad: tensors must be 16-byte aligned");
  const int64_t npix = (int64_t)F * (H / 2) * (W / 2);
  const int64_t ntiles = (npix + kSwP - 1) / kSwP;
  int grid = stem_grid() / 2;                      // two CTAs per SM
  if (ntiles < grid) grid = (int)ntiles;
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int kSmem = kSwP * (64 + kSwK) * (int)sizeof(float);   // 56 KB
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) {
      set_error("stem_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  stem_wgrad_kernel<<<grid, kSwThreads, kSmem, st>>>(reinterpret_cast<const float4*>(XP), dY, ws, F, H, W);
  VITTA_CHECK_LAUNCH();
  stem_wgrad_reduce_kernel<<<(64 * 147 + 255) / 256, 256, 0, st>>>(ws, dW, grid);
  VITTA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
