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
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_kernel(const float* __restrict__ x, StemBN bn,
                                                              float* __restrict__ out, uint8_t* __restrict__ code,
                                                              int F, int H, int W, int C4) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int64_t total = (int64_t)F * Ho * Wo * C4;
  const int C = C4 * 4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C4) * 4;
    int64_t q = i / C4;
    const int wo = (int)(q % Wo); q /= Wo;
    const int ho = (int)(q % Ho);
    const int64_t f = q / Ho;
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
      const float yx = fmaxf(fmaf(v[t].x - rm.x, k.x, b.x), 0.f), yy = fmaxf(fmaf(v[t].y - rm.y, k.y, b.y), 0.f);
      const float yz = fmaxf(fmaf(v[t].z - rm.z, k.z, b.z), 0.f), yw = fmaxf(fmaf(v[t].w - rm.w, k.w, b.w), 0.f);
      if (yx > m.x) { m.x = yx; cx = t; }
      if (yy > m.y) { m.y = yy; cy = t; }
      if (yz > m.z) { m.z = yz; cz = t; }
      if (yw > m.w) { m.w = yw; cw = t; }
    }
    st4(out + i * 4, m);
    reinterpret_cast<uint32_t*>(code)[i] = cx | (cy << 8) | (cz << 16) | (cw << 24);
  }
}

// CTA = 256 threads = 16 pixel slots x 16 channel quads (C = 64); grid-stride over blocks of pixels.  Per input pixel:
// gather the pooled gradients of the <= 4 windows whose recorded winner it is, apply the ReLU mask and the BN map.
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
  const int64_t npix = (int64_t)F * H * W;
  for (int64_t p = (int64_t)blockIdx.x * slots + slot; p < npix; p += (int64_t)gridDim.x * slots) {
    const int ww = (int)(p % W);
    const int64_t q = p / W;
    const int h = (int)(q % H);
    const int64_t f = q / H;
    const float4 xv = ld_stream4(x + p * C + c);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    // windows containing (h, ww): ho in {floor(h/2), and (h+1)/2 when h is odd}, likewise for the columns.  All (up to
    // four) code / gradient loads are issued before the first use.
    const int ho0 = h >> 1, nho = (h & 1) ? 2 : 1;
    const int wo0 = ww >> 1, nwo = (ww & 1) ? 2 : 1;
    uint32_t cd[4], mine[4];
    float4 gp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = i >> 1, bb = i & 1;
      const int ho = ho0 + a, wo = wo0 + bb;
      const bool ok = a < nho && bb < nwo && ho < Ho && wo < Wo;
      mine[i] = (uint32_t)(3 * (h - (2 * ho - 1)) + (ww - (2 * wo - 1)));
      cd[i] = 0xffffffffu;                       // matches no window position
      gp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const int64_t o = ((f * Ho + ho) * Wo + wo) * C4 + lane;
        cd[i] = __ldg(reinterpret_cast<const uint32_t*>(code) + o);
        gp[i] = ldg4(gpool + o * 4);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if ((cd[i] & 0xffu) == mine[i]) g.x += gp[i].x;
      if (((cd[i] >> 8) & 0xffu) == mine[i]) g.y += gp[i].y;
      if (((cd[i] >> 16) & 0xffu) == mine[i]) g.z += gp[i].z;
      if ((cd[i] >> 24) == mine[i]) g.w += gp[i].w;
    }
    // ReLU mask on y = BN(x)
    const float yx = fmaf(xv.x - rm.x, k.x, b.x), yy = fmaf(xv.y - rm.y, k.y, b.y);
    const float yz = fmaf(xv.z - rm.z, k.z, b.z), yw = fmaf(xv.w - rm.w, k.w, b.w);
    g.x = yx > 0.f ? g.x : 0.f; g.y = yy > 0.f ? g.y : 0.f; g.z = yz > 0.f ? g.z : 0.f; g.w = yw > 0.f ? g.w : 0.f;
    st4(gx + p * C + c, make_float4(g.x * k.x, g.y * k.y, g.z * k.z, g.w * k.w));
    agb.x += g.x; agb.y += g.y; agb.z += g.z; agb.w += g.w;
    agw.x = fmaf(g.x, (xv.x - rm.x) * istd.x, agw.x); agw.y = fmaf(g.y, (xv.y - rm.y) * istd.y, agw.y);
    agw.z = fmaf(g.z, (xv.z - rm.z) * istd.z, agw.z); agw.w = fmaf(g.w, (xv.w - rm.w) * istd.w, agw.w);
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
    float s = 0.f;
    for (unsigned bI = 0; bI < gridDim.x; ++bI) s += __ldcg(ws + ((int64_t)bI * 2 + r) * C + cc);
    (r == 0 ? gw : gb)[cc] = s;
  }
  if (threadIdx.x == 0) *ticket = 0;
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

int vitta_bn_relu_pool_fwd(const float* x, VittaBN bn, float* out, uint8_t* code, int F, int H, int W, int C,
                           void* stream) {
  VITTA_CHECK_ARG(x && out && code && bn.weight && bn.bias && bn.running_mean && bn.running_var, VITTA_E_BADARG,
                  "bn_relu_pool_fwd: null pointer");
  VITTA_CHECK_ARG(F > 0 && H > 1 && W > 1 && C > 0 && C % 4 == 0, VITTA_E_BADARG, "bn_relu_pool_fwd: bad shape");
  VITTA_CHECK_ARG(aligned16(x) && aligned16(out) && (reinterpret_cast<uintptr_t>(code) & 3u) == 0, VITTA_E_ALIGN,
                  "bn_relu_pool_fwd: alignment");
  const int64_t total = (int64_t)F * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > stem_grid() * 4) blocks = stem_grid() * 4;
  StemBN b{bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps};
  bn_relu_pool_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, b, out, code, F, H, W, C / 4);
  VITTA_CHECK_LAUNCH();
  return 0;
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
  const int slots = 256 / (C / 4);
  const int64_t npix = (int64_t)F * H * W;
  int64_t blocks = (npix + slots - 1) / slots;
  if (blocks > stem_grid()) blocks = stem_grid();
  bn_relu_pool_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(gpool, code, x, b, gx, ws, gw, gb, F, H, W,
                                                                            C / 4, stem_grid() * 2 * C);
  VITTA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
