// K10 prediction-consistency loss (+gradient) and K11 multi-tensor SGD.
#include "common.cuh"

namespace vitta {

constexpr int kConsisThreads = 512;

__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];  // same order in every thread: deterministic
  return t;
}
__device__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmaxf(t, red[w]);
  return t;
}

// One CTA, videos processed one after the other (B*V*K is at most a few 10^4 elements).
//   p_v = softmax(z_v); m = mean_v p_v; L += (1/V) sum_v sum_k |p_v - m|
//   dL/dp_u = (1/V) (s_u - mean_v s_v), s = sign(p - m);  dL/dz_u = p_u * (q_u - <q_u, p_u>)
__global__ void __launch_bounds__(kConsisThreads) consis_kernel(const float* __restrict__ preds, int B, int V, int K,
                                                               float* __restrict__ loss, float* __restrict__ grad) {
  extern __shared__ float sp[];  // V*K probabilities
  __shared__ float red[kConsisThreads / 32];
  float total = 0.f;
  const float invV = 1.f / (float)V;
  for (int b = 0; b < B; ++b) {
    const float* z = preds + (int64_t)b * V * K;
    for (int v = 0; v < V; ++v) {
      float mx = -INFINITY;
      for (int k = threadIdx.x; k < K; k += blockDim.x) mx = fmaxf(mx, z[v * K + k]);
      mx = block_max(mx, red);
      float s = 0.f;
      for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float e = expf(z[v * K + k] - mx);
        sp[v * K + k] = e;
        s += e;
      }
      s = block_sum(s, red);
      const float inv = 1.f / s;
      for (int k = threadIdx.x; k < K; k += blockDim.x) sp[v * K + k] *= inv;
    }
    __syncthreads();
    float l = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float m = 0.f;
      for (int v = 0; v < V; ++v) m += sp[v * K + k];
      m *= invV;
      for (int v = 0; v < V; ++v) l += fabsf(sp[v * K + k] - m);
    }
    total += block_sum(l, red) * invV;
    if (grad) {
      for (int u = 0; u < V; ++u) {
        float dot = 0.f;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
          float m = 0.f;
          for (int v = 0; v < V; ++v) m += sp[v * K + k];
          m *= invV;
          float ss = 0.f;
          for (int v = 0; v < V; ++v) {
            const float d = sp[v * K + k] - m;
            ss += (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
          }
          const float du = sp[u * K + k] - m;
          const float su = (du > 0.f) ? 1.f : ((du < 0.f) ? -1.f : 0.f);
          const float q = invV * (su - invV * ss);
          grad[((int64_t)b * V + u) * K + k] = q;  // stash q, finished below
          dot = fmaf(q, sp[u * K + k], dot);
        }
        dot = block_sum(dot, red);
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
          const int64_t gi = ((int64_t)b * V + u) * K + k;
          grad[gi] = sp[u * K + k] * (grad[gi] - dot);
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = total;
}

// ------------------------------------------------------------------------------------------------
// K11: every CTA owns one 4096-element block of one tensor (block table prefix in the tensor entries).
// ------------------------------------------------------------------------------------------------
constexpr int kSgdBlock = 4096;

struct SgdEntry {  // mirrors VittaSgdTensor
  float* p;
  const float* g;
  float* buf;
  int64_t n;
};

__global__ void __launch_bounds__(kThreads) sgd_kernel(const SgdEntry* __restrict__ tensors,
                                                      const int32_t* __restrict__ block_start, int n_tensors, float lr,
                                                      float momentum, float wd, int first, float gscale) {
  // binary search: last tensor whose block_start <= blockIdx.x
  int lo = 0, hi = n_tensors - 1;
  const int bid = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (block_start[mid] <= bid) lo = mid; else hi = mid - 1;
  }
  const SgdEntry t = tensors[lo];
  const int64_t e0 = (int64_t)(bid - block_start[lo]) * kSgdBlock;
  const int64_t rem = t.n - e0;
  const int cnt = (int)(rem < kSgdBlock ? rem : kSgdBlock);
  float* p = t.p + e0;
  const float* g = t.g + e0;
  float* m = t.buf + e0;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m)) & 15u) == 0;
  if (vec) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += kThreads) {
      float4 pv = *reinterpret_cast<float4*>(p + i * 4);
      const float4 gv = ld_stream4(g + i * 4);
      float4 d, b;
      d.x = fmaf(wd, pv.x, gv.x * gscale); d.y = fmaf(wd, pv.y, gv.y * gscale);
      d.z = fmaf(wd, pv.z, gv.z * gscale); d.w = fmaf(wd, pv.w, gv.w * gscale);
      if (first) {
        b = d;
      } else {
        b = *reinterpret_cast<float4*>(m + i * 4);
        b.x = fmaf(momentum, b.x, d.x); b.y = fmaf(momentum, b.y, d.y);
        b.z = fmaf(momentum, b.z, d.z); b.w = fmaf(momentum, b.w, d.w);
      }
      *reinterpret_cast<float4*>(m + i * 4) = b;
      pv.x = fmaf(-lr, b.x, pv.x); pv.y = fmaf(-lr, b.y, pv.y); pv.z = fmaf(-lr, b.z, pv.z); pv.w = fmaf(-lr, b.w, pv.w);
      *reinterpret_cast<float4*>(p + i * 4) = pv;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += kThreads) {
      const float d = fmaf(wd, p[i], g[i] * gscale);
      const float b = first ? d : fmaf(momentum, m[i], d);
      m[i] = b;
      p[i] = fmaf(-lr, b, p[i]);
    }
  } else {
    for (int i = threadIdx.x; i < cnt; i += kThreads) {
      const float d = fmaf(wd, p[i], g[i] * gscale);
      const float b = first ? d : fmaf(momentum, m[i], d);
      m[i] = b;
      p[i] = fmaf(-lr, b, p[i]);
    }
  }
}

}  // namespace vitta

using namespace vitta;

extern "C" {

int vitta_pred_consis(const float* preds, int B, int V, int K, float* loss, float* grad, void* stream) {
  VITTA_CHECK_ARG(preds && loss, VITTA_E_BADARG, "pred_consis: null pointer");
  VITTA_CHECK_ARG(B > 0 && V > 0 && K > 0, VITTA_E_BADARG, "pred_consis: bad shape");
  const size_t smem = (size_t)V * K * sizeof(float);
  VITTA_CHECK_ARG(smem <= 200 * 1024, VITTA_E_UNSUPPORTED, "pred_consis: V*K too large for shared memory");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(consis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("pred_consis: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  consis_kernel<<<1, kConsisThreads, smem, (cudaStream_t)stream>>>(preds, B, V, K, loss, grad);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_sgd_block_elems(void) { return kSgdBlock; }

int vitta_sgd_step(const VittaSgdTensor* tensors, const int32_t* block_start, int n_tensors, int total_blocks,
                   float lr, float momentum, float weight_decay, int first_step, float grad_scale, void* stream) {
  VITTA_CHECK_ARG(tensors && block_start && n_tensors > 0 && total_blocks > 0, VITTA_E_BADARG, "sgd_step: bad arguments");
  static_assert(sizeof(SgdEntry) == sizeof(VittaSgdTensor), "layout");
  sgd_kernel<<<(unsigned)total_blocks, kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const SgdEntry*>(tensors), block_start, n_tensors, lr, momentum, weight_decay, first_step,
      grad_scale);
  VITTA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
