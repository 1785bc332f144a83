// View gathering + normalisation on the GPU: the step just before the hot path.
//   replaces, for frames that are already decoded and at the target scale: container.get_batch(frame_indices) ->
//   per-view crop -> Stack -> ToTorchFormatTensor (/255) -> GroupNormalize (models/tanet_models/video_dataset.py:318-345,
//   transforms.py:627-690) and the Swin pipeline's Normalize + FormatShape('NCTHW').
// Input: decoded frames (F, H, W, 3) uint8 on the device and the frame indices of all views (V*T, from
// vitta_b200.corpus.views.sample_tta_view_indices).  Output (fp32), crop window (crop_y, crop_x, out_h, out_w):
//   layout 0 (TANet loader):  (V*T*3, out_h, out_w)   planes ordered [view][frame][rgb]
//   layout 1 (Swin loader):   (V, 3, T, out_h, out_w)
// One thread per output pixel: 3 contiguous bytes in, one float per colour plane out (coalesced both ways).
#include "common.cuh"

namespace vitta {

__global__ void __launch_bounds__(256) gather_normalize_kernel(const uint8_t* __restrict__ frames, int F, int H, int W,
                                                              const int32_t* __restrict__ idx, int n_idx, int crop_y,
                                                              int crop_x, int out_h, int out_w, float3 scale, float3 shift,
                                                              int layout, int T, float* __restrict__ out) {
  const int64_t plane = (int64_t)out_h * out_w;
  const int64_t total = (int64_t)n_idx * plane;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int k = (int)(i / plane);          // which (view, frame)
    const int64_t px = i - (int64_t)k * plane;
    const int y = (int)(px / out_w), x = (int)(px - (int64_t)y * out_w);
    int f = __ldg(idx + k);
    f = f < 0 ? 0 : (f >= F ? F - 1 : f);    // np.minimum(frame_indices, num_frames - 1) (video_dataset.py:328)
    const uint8_t* src = frames + (((int64_t)f * H + crop_y + y) * W + crop_x + x) * 3;
    const float r = (float)src[0] * scale.x + shift.x;
    const float g = (float)src[1] * scale.y + shift.y;
    const float b = (float)src[2] * scale.z + shift.z;
    if (layout == 0) {
      float* o = out + (int64_t)k * 3 * plane + px;
      o[0] = r;
      o[plane] = g;
      o[2 * plane] = b;
    } else {
      const int v = k / T, t = k - v * T;
      float* o = out + (((int64_t)v * 3) * T + t) * plane + px;
      o[0] = r;
      o[(int64_t)T * plane] = g;
      o[2 * (int64_t)T * plane] = b;
    }
  }
}

// ---- per-view crop + bilinear resize (SubgroupWise_MultiScaleCrop_TANet, transforms.py:277-384) ----
// Bit-exact restatement of what the reference's PIL pipeline computes for 8-bit frames: img.crop(box).resize((S, S),
// Image.BILINEAR) is Pillow's two-pass fixed-point resampler (Resample.c: 22-bit coefficients, horizontal pass rounded to
// uint8 before the vertical pass).  The coefficient tables come from vitta_resample_coeffs_u8 (host, double precision like
// Pillow's precompute_coeffs) with the crop offset already added to the window starts, one table set per view.
// One thread per output pixel: for each of its <= ks source rows it forms the horizontally resampled uint8 value, then
// combines the rows with the vertical coefficients.  Consecutive threads read consecutive source pixels.
constexpr int kResampleBits = 32 - 8 - 2;

__device__ __forceinline__ int clip8_fixed(int v) {
  v >>= kResampleBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) gather_crop_resize_kernel(
    const uint8_t* __restrict__ frames, int F, int H, int W, const int32_t* __restrict__ idx, int n_idx,
    const int32_t* __restrict__ hb, const int32_t* __restrict__ hk, const int32_t* __restrict__ vb,
    const int32_t* __restrict__ vk, int ks, int out_h, int out_w, float3 mean, float3 stdv, int layout, int T,
    float* __restrict__ out) {
  const int64_t plane = (int64_t)out_h * out_w;
  const int64_t total = (int64_t)n_idx * plane;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int k = (int)(i / plane);          // which (view, frame)
    const int64_t px = i - (int64_t)k * plane;
    const int y = (int)(px / out_w), x = (int)(px - (int64_t)y * out_w);
    const int v = k / T;
    int f = __ldg(idx + k);
    f = f < 0 ? 0 : (f >= F ? F - 1 : f);    // np.minimum(frame_indices, num_frames - 1) (video_dataset.py:328)
    const int32_t* hbv = hb + ((int64_t)v * out_w + x) * 2;
    const int32_t* vbv = vb + ((int64_t)v * out_h + y) * 2;
    const int32_t* hkv = hk + ((int64_t)v * out_w + x) * ks;
    const int32_t* vkv = vk + ((int64_t)v * out_h + y) * ks;
    const int x0 = __ldg(hbv), nx = min(__ldg(hbv + 1), ks);
    const int y0 = __ldg(vbv), ny = min(__ldg(vbv + 1), ks);
    const uint8_t* fr = frames + (int64_t)f * H * W * 3;
    int ar = 1 << (kResampleBits - 1), ag = ar, ab = ar;
    for (int yy = 0; yy < ny; ++yy) {
      const int sy = min(max(y0 + yy, 0), H - 1);          // tables built by the library never leave the frame;
      const uint8_t* row = fr + (int64_t)sy * W * 3;        // the clamp only guards against foreign tables
      int hr = 1 << (kResampleBits - 1), hg = hr, hbl = hr;
      for (int xx = 0; xx < nx; ++xx) {
        const int sx = min(max(x0 + xx, 0), W - 1);
        const int c = __ldg(hkv + xx);
        hr += (int)row[sx * 3] * c;
        hg += (int)row[sx * 3 + 1] * c;
        hbl += (int)row[sx * 3 + 2] * c;
      }
      const int c = __ldg(vkv + yy);
      ar += clip8_fixed(hr) * c;
      ag += clip8_fixed(hg) * c;
      ab += clip8_fixed(hbl) * c;
    }
    // ToTorchFormatTensor (img.float().div(255)) then GroupNormalize (t.sub_(m).div_(s)): three correctly rounded float32
    // operations, spelled with intrinsics so that nothing is contracted or turned into a reciprocal multiply -- the result
    // is bit-identical to the reference loader's tensor (tests/golden/loader.npz)
    const float r = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8_fixed(ar), 255.f), mean.x), stdv.x);
    const float g = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8_fixed(ag), 255.f), mean.y), stdv.y);
    const float b = __fdiv_rn(__fsub_rn(__fdiv_rn((float)clip8_fixed(ab), 255.f), mean.z), stdv.z);
    if (layout == 0) {
      float* o = out + (int64_t)k * 3 * plane + px;
      o[0] = r;
      o[plane] = g;
      o[2 * plane] = b;
    } else {
      const int t = k - v * T;
      float* o = out + (((int64_t)v * 3) * T + t) * plane + px;
      o[0] = r;
      o[(int64_t)T * plane] = g;
      o[2 * (int64_t)T * plane] = b;
    }
  }
}

// ---- OpenCV's 8-bit INTER_LINEAR (the Video-Swin loader's mmcv.imresize, transforms_backup.py:794-798) ----
// Bit-exact restatement of resize.cpp's fixed-point path (oracle/cv2_resample.py has the derivation): 11-bit weights, the
// horizontal pass keeps 32-bit sums, rows clipped (not re-weighted) at the top / bottom, and the vertical combine
// ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  The source is the region (x0, y0, cw, ch) of frame
// idx[k] (or frame k); tables from vitta_cv_linear_tables.  OUT_F32: normalise like mmcv.imnormalize_ ((x - mean) * 1/std
// on the 0..255 scale, float32 mean / std as mmcv's Normalize stores them) and write the Swin loader layout (V, 3, T, h, w) or the TANet one; else write uint8 (n, h, w, 3).
template <bool OUT_F32>
__global__ void __launch_bounds__(256) cv_resize_kernel(const uint8_t* __restrict__ src, int F, int H, int W,
                                                       const int32_t* __restrict__ idx, int n, int x0, int y0, int cw, int ch,
                                                       const int32_t* __restrict__ xofs, const int32_t* __restrict__ xw,
                                                       const int32_t* __restrict__ yofs, const int32_t* __restrict__ yw,
                                                       int out_h, int out_w, float3 mean, double3 stdinv, int layout, int T,
                                                       void* __restrict__ out_) {
  const int64_t plane = (int64_t)out_h * out_w;
  const int64_t total = (int64_t)n * plane;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int k = (int)(i / plane);
    const int64_t px = i - (int64_t)k * plane;
    const int y = (int)(px / out_w), x = (int)(px - (int64_t)y * out_w);
    int f = idx ? __ldg(idx + k) : k;
    f = f < 0 ? 0 : (f >= F ? F - 1 : f);
    const int sy = __ldg(yofs + y), sx = __ldg(xofs + x);
    const int r0 = y0 + min(max(sy, 0), ch - 1), r1 = y0 + min(max(sy + 1, 0), ch - 1);
    const int c0 = x0 + min(max(sx, 0), cw - 1), c1 = x0 + min(max(sx + 1, 0), cw - 1);
    const int a0 = __ldg(xw + 2 * x), a1 = __ldg(xw + 2 * x + 1), b0 = __ldg(yw + 2 * y), b1 = __ldg(yw + 2 * y + 1);
    const uint8_t* fr = src + (int64_t)f * H * W * 3;
    const uint8_t* p00 = fr + ((int64_t)r0 * W + c0) * 3;
    const uint8_t* p01 = fr + ((int64_t)r0 * W + c1) * 3;
    const uint8_t* p10 = fr + ((int64_t)r1 * W + c0) * 3;
    const uint8_t* p11 = fr + ((int64_t)r1 * W + c1) * 3;
    int v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int s0 = (int)p00[c] * a0 + (int)p01[c] * a1;
      const int s1 = (int)p10[c] * a0 + (int)p11[c] * a1;
      const int o = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
      v[c] = o < 0 ? 0 : (o > 255 ? 255 : o);
    }
    if constexpr (OUT_F32) {
      float* out = reinterpret_cast<float*>(out_);
      // mmcv.imnormalize_ = cv2.subtract then cv2.multiply with float64 scalars on a float32 image: the difference is
      // rounded to float32, the product is formed in double and rounded once (checked against OpenCV for all 256 values)
      const float r = (float)((double)((float)v[0] - mean.x) * stdinv.x), g = (float)((double)((float)v[1] - mean.y) * stdinv.y),
                  b = (float)((double)((float)v[2] - mean.z) * stdinv.z);
      if (layout == 0) {
        float* o = out + (int64_t)k * 3 * plane + px;
        o[0] = r; o[plane] = g; o[2 * plane] = b;
      } else {
        const int vw = k / T, t = k - vw * T;
        float* o = out + (((int64_t)vw * 3) * T + t) * plane + px;
        o[0] = r; o[(int64_t)T * plane] = g; o[2 * (int64_t)T * plane] = b;
      }
    } else {
      uint8_t* o = reinterpret_cast<uint8_t*>(out_) + i * 3;
      o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
    }
  }
}

}  // namespace vitta

using namespace vitta;

extern "C" int vitta_gather_normalize_u8(const uint8_t* frames, int F, int H, int W, const int32_t* idx, int n_idx,
                                         int crop_y, int crop_x, int out_h, int out_w, const float* mean3_host,
                                         const float* std3_host, int layout, int T, float* out, void* stream) {
  VITTA_CHECK_ARG(frames && idx && out && mean3_host && std3_host, VITTA_E_BADARG, "gather_normalize: null pointer");
  VITTA_CHECK_ARG(F > 0 && H > 0 && W > 0 && n_idx > 0 && out_h > 0 && out_w > 0 && T > 0 && n_idx % T == 0,
                  VITTA_E_BADARG, "gather_normalize: bad shape");
  VITTA_CHECK_ARG(crop_y >= 0 && crop_x >= 0 && crop_y + out_h <= H && crop_x + out_w <= W, VITTA_E_BADARG,
                  "gather_normalize: crop window outside the frame");
  VITTA_CHECK_ARG(layout == 0 || layout == 1, VITTA_E_BADARG, "gather_normalize: layout must be 0 (TANet) or 1 (Swin)");
  // x/255 normalised with mean/std given on the [0, 1] scale (utils/opts.py:4-5): (x/255 - m)/s = x * 1/(255 s) - m/s
  const float3 scale = make_float3(1.f / (255.f * std3_host[0]), 1.f / (255.f * std3_host[1]), 1.f / (255.f * std3_host[2]));
  const float3 shift = make_float3(-mean3_host[0] / std3_host[0], -mean3_host[1] / std3_host[1], -mean3_host[2] / std3_host[2]);
  const int64_t total = (int64_t)n_idx * out_h * out_w;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_normalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(frames, F, H, W, idx, n_idx, crop_y, crop_x,
                                                                             out_h, out_w, scale, shift, layout, T, out);
  VITTA_CHECK_LAUNCH();
  return 0;
}


// ---- Pillow's bilinear coefficient tables (host only: no CUDA call) ----
extern "C" int vitta_resample_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return -1;
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return (int)ceil(filterscale) * 2 + 1;
}

extern "C" int vitta_resample_coeffs_u8(int in_size, int out_size, int in_offset, int slots, int32_t* bounds_host,
                                        int32_t* kk_host) {
  const int ksize = vitta_resample_ksize(in_size, out_size);
  VITTA_CHECK_ARG(ksize > 0 && bounds_host && kk_host && in_offset >= 0, VITTA_E_BADARG, "resample_coeffs: bad arguments");
  VITTA_CHECK_ARG(slots >= ksize, VITTA_E_BADARG, "resample_coeffs: fewer coefficient slots than vitta_resample_ksize");
  // Resample.c precompute_coeffs (bilinear: support 1.0) + normalize_coeffs_8bpc, in double precision like Pillow
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const double ss = 1.0 / filterscale;
  double w[64];
  VITTA_CHECK_ARG(ksize <= 64, VITTA_E_UNSUPPORTED, "resample_coeffs: reduction factor above 31");
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      w[x] = a < 1.0 ? 1.0 - a : 0.0;
      ww += w[x];
    }
    int32_t* k = kk_host + (int64_t)xx * slots;
    for (int x = 0; x < slots; ++x) {
      double c = 0.0;
      if (x < xmax) c = (ww != 0.0) ? w[x] / ww : w[x];
      k[x] = c < 0 ? (int32_t)(-0.5 + c * (1 << kResampleBits)) : (int32_t)(0.5 + c * (1 << kResampleBits));
    }
    bounds_host[xx * 2] = xmin + in_offset;     // window start in frame coordinates (crop offset folded in)
    bounds_host[xx * 2 + 1] = xmax;
  }
  return 0;
}

extern "C" int vitta_gather_crop_resize_normalize_u8(const uint8_t* frames, int F, int H, int W, const int32_t* idx,
                                                     int n_idx, const int32_t* boxes_host, int n_views,
                                                     const int32_t* hbounds, const int32_t* hk, const int32_t* vbounds,
                                                     const int32_t* vk, int slots, int out_h, int out_w,
                                                     const float* mean3_host, const float* std3_host, int layout, int T,
                                                     float* out, void* stream) {
  VITTA_CHECK_ARG(frames && idx && out && mean3_host && std3_host && boxes_host && hbounds && hk && vbounds && vk,
                  VITTA_E_BADARG, "gather_crop_resize: null pointer");
  VITTA_CHECK_ARG(F > 0 && H > 0 && W > 0 && n_idx > 0 && out_h > 0 && out_w > 0 && T > 0 && n_views > 0 &&
                      n_idx == n_views * T && slots >= 3,
                  VITTA_E_BADARG, "gather_crop_resize: bad shape (n_idx must be n_views * T)");
  VITTA_CHECK_ARG(layout == 0 || layout == 1, VITTA_E_BADARG, "gather_crop_resize: layout must be 0 (TANet) or 1 (Swin)");
  for (int v = 0; v < n_views; ++v) {   // source region of each view: (crop_w, crop_h, offset_w, offset_h), as _sample_crop_size returns it
    const int32_t* b = boxes_host + v * 4;
    VITTA_CHECK_ARG(b[0] > 0 && b[1] > 0 && b[2] >= 0 && b[3] >= 0 && b[2] + b[0] <= W && b[3] + b[1] <= H, VITTA_E_BADARG,
                    "gather_crop_resize: crop box outside the frame");
  }
  const float3 mean = make_float3(mean3_host[0], mean3_host[1], mean3_host[2]);
  const float3 stdv = make_float3(std3_host[0], std3_host[1], std3_host[2]);
  const int64_t total = (int64_t)n_idx * out_h * out_w;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_crop_resize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      frames, F, H, W, idx, n_idx, hbounds, hk, vbounds, vk, slots, out_h, out_w, mean, stdv, layout, T, out);
  VITTA_CHECK_LAUNCH();
  return 0;
}


// ---- OpenCV INTER_LINEAR tap tables (host only: no CUDA call) ----
extern "C" int vitta_cv_linear_tables(int src, int dst, int horizontal, int32_t* ofs_host, int32_t* w_host) {
  VITTA_CHECK_ARG(src > 0 && dst > 0 && ofs_host && w_host, VITTA_E_BADARG, "cv_linear_tables: bad arguments");
  const double scale = 1.0 / ((double)dst / src);          // resize(): inv_scale = dst / src, scale = 1 / inv_scale
  for (int d = 0; d < dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (horizontal) {                                      // the vertical pass clips the ROW INDICES instead
      if (s < 0) { f = 0.f; s = 0; }
      if (s >= src - 1) { f = 0.f; s = src - 1; }
    }
    ofs_host[d] = s;
    w_host[2 * d] = (int32_t)lrintf((1.f - f) * 2048.f);   // saturate_cast<short>(float): round half to even
    w_host[2 * d + 1] = (int32_t)lrintf(f * 2048.f);
  }
  return 0;
}

static int cv_resize_impl(const uint8_t* src, int F, int H, int W, const int32_t* idx, int n, int x0, int y0, int cw, int ch,
                          const int32_t* xofs, const int32_t* xw, const int32_t* yofs, const int32_t* yw, int out_h,
                          int out_w, const float* mean3_host, const float* std3_host, int layout, int T, void* out,
                          bool f32, void* stream) {
  VITTA_CHECK_ARG(src && out && xofs && xw && yofs && yw, VITTA_E_BADARG, "cv_resize: null pointer");
  VITTA_CHECK_ARG(F > 0 && H > 0 && W > 0 && n > 0 && out_h > 0 && out_w > 0, VITTA_E_BADARG, "cv_resize: bad shape");
  VITTA_CHECK_ARG(cw > 0 && ch > 0 && x0 >= 0 && y0 >= 0 && x0 + cw <= W && y0 + ch <= H, VITTA_E_BADARG,
                  "cv_resize: source region outside the frame");
  VITTA_CHECK_ARG(idx || n <= F, VITTA_E_BADARG, "cv_resize: without an index vector n must not exceed the frame count");
  float3 mean = make_float3(0.f, 0.f, 0.f);
  double3 stdinv = make_double3(1.0, 1.0, 1.0);
  if (f32) {
    VITTA_CHECK_ARG(mean3_host && std3_host && T > 0 && n % T == 0 && (layout == 0 || layout == 1), VITTA_E_BADARG,
                    "cv_resize_normalize: needs mean / std, layout 0 or 1 and n a multiple of T");
    mean = make_float3(mean3_host[0], mean3_host[1], mean3_host[2]);
    stdinv = make_double3(1.0 / (double)std3_host[0], 1.0 / (double)std3_host[1], 1.0 / (double)std3_host[2]);
  }
  const int64_t total = (int64_t)n * out_h * out_w;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (f32)
    cv_resize_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, F, H, W, idx, n, x0, y0, cw, ch, xofs, xw,
                                                                              yofs, yw, out_h, out_w, mean, stdinv, layout,
                                                                              T, out);
  else
    cv_resize_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, F, H, W, idx, n, x0, y0, cw, ch, xofs, xw,
                                                                               yofs, yw, out_h, out_w, mean, stdinv, 0, 1,
                                                                               out);
  VITTA_CHECK_LAUNCH();
  return 0;
}

extern "C" int vitta_cv_resize_u8(const uint8_t* src, int F, int H, int W, const int32_t* idx, int n, int x0, int y0, int cw,
                                  int ch, const int32_t* xofs, const int32_t* xw, const int32_t* yofs, const int32_t* yw,
                                  int out_h, int out_w, uint8_t* out, void* stream) {
  return cv_resize_impl(src, F, H, W, idx, n, x0, y0, cw, ch, xofs, xw, yofs, yw, out_h, out_w, nullptr, nullptr, 0, 1, out,
                        false, stream);
}

extern "C" int vitta_cv_resize_normalize_u8(const uint8_t* src, int F, int H, int W, const int32_t* idx, int n, int x0,
                                            int y0, int cw, int ch, const int32_t* xofs, const int32_t* xw,
                                            const int32_t* yofs, const int32_t* yw, int out_h, int out_w,
                                            const float* mean3_host, const float* std3_host, int layout, int T, float* out,
                                            void* stream) {
  return cv_resize_impl(src, F, H, W, idx, n, x0, y0, cw, ch, xofs, xw, yofs, yw, out_h, out_w, mean3_host, std3_host, layout,
                        T, out, true, stream);
}
