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
