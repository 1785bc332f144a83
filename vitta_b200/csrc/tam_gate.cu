// K5b: the two small networks of the Temporal Adaptive Module that produce the operands of the stencil kernel (tam.cu)
// from the spatially pooled activation p (N, T, C)  (reference models/tanet_models/temporal_module.py:27-41, 49-55):
//   G (global branch): per (video, channel)   softmax( W2 . relu(BN1d( W1 . p[n, :, c] )) )            -> kern (N, 3, C)
//   L (local branch):  per (video, frame)     sigmoid( Wb . relu(BN1d( conv1d_k3(Wa, p)[n, t, :] )) )   -> act  (N, T, C)
// with the BatchNorm1d layers in eval mode.  In eager PyTorch this is ~20 tiny kernels forward and ~25 backward per TAM
// (cutlass simt sgemm, ATen batch-norm / elementwise / reduce kernels: 16 TAMs -> ~700 launches per step); here it is
// 3 launches forward and 8 backward (tam_bwd_finish of tam.cu included).  Everything is L2-resident (p is N*T*C <= 64 K floats, the largest weight 786 KB),
// so the kernels are latency-bound; a generic 32x32-tile fp32 GEMM with stage-specific operand loaders and epilogues does
// the L branch, one thread per (video, channel) does the G branch.  All reductions run in a fixed order (deterministic).
#include "common.cuh"

namespace vitta {

constexpr int kGateMaxT = 16;

struct GateBN {
  const float *w, *b, *rm, *rv;
  float eps;
};

__device__ __forceinline__ float gate_bn_k(const GateBN& bn, int i) { return __ldg(bn.w + i) * (1.f / sqrtf(__ldg(bn.rv + i) + bn.eps)); }
__device__ __forceinline__ float gate_bn_apply(const GateBN& bn, int i, float x) {
  return fmaf(x - __ldg(bn.rm + i), gate_bn_k(bn, i), __ldg(bn.b + i));
}

struct GateArgs {
  const float* p;        // (N, T, C)
  const float *W1, *W2;  // (2T, T), (3, 2T)
  const float *Wa, *Wb;  // (C/4, C, 3), (C, C/4)
  GateBN bn1, bn2;
  float *kern, *act;     // (N, 3, C), (N, T, C)
  float* pre;            // (N*T, C/4): L hidden layer before its BatchNorm
  // backward
  const float *gkern, *gact;
  float *gz, *gpre, *ghm;     // (N*T, C), (N*T, C/4), (N*T, C/4)
  float *gp;                  // (N, T, C): gradient of p (G part written first, L part added)
  float *gW1, *gW2, *gbn1w, *gbn1b, *gWa, *gWb, *gbn2w, *gbn2b;
  float* ws;                  // G-branch per-CTA partials + ticket
  int N, T, C;
};

// ------------------------------------------------------------------------------------------------
// G branch
// ------------------------------------------------------------------------------------------------
constexpr int kGThreads = 128;

// recompute the forward of one (n, c) column; returns softmax s[3]; fills v[T], hb[2T] (BN output, pre-ReLU)
__device__ __forceinline__ void g_forward(const GateArgs& a, const float* sW1, const float* sW2, int n, int c, float* v,
                                          float* hb, float* s) {
  const int T = a.T, H = 2 * a.T;
#pragma unroll
  for (int t = 0; t < kGateMaxT; ++t) v[t] = (t < T) ? __ldg(a.p + ((int64_t)n * T + t) * a.C + c) : 0.f;
  float z[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 2 * kGateMaxT; ++j) {
    hb[j] = 0.f;
    if (j < H) {
      float h = 0.f;
#pragma unroll
      for (int t = 0; t < kGateMaxT; ++t)
        if (t < T) h = fmaf(sW1[j * T + t], v[t], h);
      h = gate_bn_apply(a.bn1, j, h);
      hb[j] = h;
      const float r = fmaxf(h, 0.f);
      z[0] = fmaf(sW2[j], r, z[0]);
      z[1] = fmaf(sW2[H + j], r, z[1]);
      z[2] = fmaf(sW2[2 * H + j], r, z[2]);
    }
  }
  const float m = fmaxf(z[0], fmaxf(z[1], z[2]));
  const float e0 = expf(z[0] - m), e1 = expf(z[1] - m), e2 = expf(z[2] - m);
  const float inv = 1.f / (e0 + e1 + e2);
  s[0] = e0 * inv; s[1] = e1 * inv; s[2] = e2 * inv;
}

__global__ void __launch_bounds__(kGThreads) tam_g_fwd_kernel(GateArgs a) {
  __shared__ float sW1[2 * kGateMaxT * kGateMaxT], sW2[3 * 2 * kGateMaxT];
  const int T = a.T, H = 2 * T;
  for (int i = threadIdx.x; i < H * T; i += kGThreads) sW1[i] = __ldg(a.W1 + i);
  for (int i = threadIdx.x; i < 3 * H; i += kGThreads) sW2[i] = __ldg(a.W2 + i);
  __syncthreads();
  const int idx = blockIdx.x * kGThreads + threadIdx.x;
  if (idx >= a.N * a.C) return;
  const int n = idx / a.C, c = idx % a.C;
  float v[kGateMaxT], hb[2 * kGateMaxT], s[3];
  g_forward(a, sW1, sW2, n, c, v, hb, s);
  a.kern[((int64_t)n * 3 + 0) * a.C + c] = s[0];
  a.kern[((int64_t)n * 3 + 1) * a.C + c] = s[1];
  a.kern[((int64_t)n * 3 + 2) * a.C + c] = s[2];
}

// per-thread backward + CTA-level reduction of the parameter gradients through shared memory, per-CTA partials to the
// workspace, the last CTA adds them in CTA order.  Partial layout per CTA: gW1 [2T*T] | gW2 [3*2T] | gbn_w [2T] | gbn_b [2T]
__global__ void __launch_bounds__(kGThreads) tam_g_bwd_kernel(GateArgs a, int n_part) {
  extern __shared__ float sm[];
  __shared__ int s_last;
  const int T = a.T, H = 2 * T;
  float* sW1 = sm;                         // H*T
  float* sW2 = sW1 + H * T;                // 3*H
  float* sv = sW2 + 3 * H;                 // [kGThreads][T]
  float* sg = sv + kGThreads * T;          // [kGThreads][H]   gh_pre
  float* sr = sg + kGThreads * H;          // [kGThreads][H]   relu(hb)
  float* sb = sr + kGThreads * H;          // [kGThreads][H]   ghb (gradient at the BN output)
  float* sx = sb + kGThreads * H;          // [kGThreads][H]   xhat
  float* sz = sx + kGThreads * H;          // [kGThreads][3]   gz
  for (int i = threadIdx.x; i < H * T; i += kGThreads) sW1[i] = __ldg(a.W1 + i);
  for (int i = threadIdx.x; i < 3 * H; i += kGThreads) sW2[i] = __ldg(a.W2 + i);
  __syncthreads();
  const int idx = blockIdx.x * kGThreads + threadIdx.x;
  const bool live = idx < a.N * a.C;
  const int tid = threadIdx.x;
  {
    float v[kGateMaxT], hb[2 * kGateMaxT], s[3];
    float gz[3] = {0.f, 0.f, 0.f};
    int n = 0, c = 0;
    if (live) {
      n = idx / a.C; c = idx % a.C;
      g_forward(a, sW1, sW2, n, c, v, hb, s);
      const float g0 = __ldg(a.gkern + ((int64_t)n * 3 + 0) * a.C + c), g1 = __ldg(a.gkern + ((int64_t)n * 3 + 1) * a.C + c),
                  g2 = __ldg(a.gkern + ((int64_t)n * 3 + 2) * a.C + c);
      const float dot = g0 * s[0] + g1 * s[1] + g2 * s[2];
      gz[0] = s[0] * (g0 - dot); gz[1] = s[1] * (g1 - dot); gz[2] = s[2] * (g2 - dot);
    } else {
#pragma unroll
      for (int t = 0; t < kGateMaxT; ++t) v[t] = 0.f;
#pragma unroll
      for (int j = 0; j < 2 * kGateMaxT; ++j) hb[j] = 0.f;
    }
    float gv[kGateMaxT];
#pragma unroll
    for (int t = 0; t < kGateMaxT; ++t) gv[t] = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * kGateMaxT; ++j) {
      if (j < H) {
        const float ghr = sW2[j] * gz[0] + sW2[H + j] * gz[1] + sW2[2 * H + j] * gz[2];
        const float ghb = (live && hb[j] > 0.f) ? ghr : 0.f;
        const float k1 = gate_bn_k(a.bn1, j);
        const float ghp = ghb * k1;
        // xhat from the recomputed pre-BN value (dividing (hb - beta) by gamma would fail for gamma == 0)
        float hp = 0.f;
#pragma unroll
        for (int t = 0; t < kGateMaxT; ++t)
          if (t < T) hp = fmaf(sW1[j * T + t], v[t], hp);
        const float xh = (hp - __ldg(a.bn1.rm + j)) * (1.f / sqrtf(__ldg(a.bn1.rv + j) + a.bn1.eps));
        sg[tid * H + j] = ghp;
        sr[tid * H + j] = live ? fmaxf(hb[j], 0.f) : 0.f;
        sb[tid * H + j] = ghb;
        sx[tid * H + j] = xh;
#pragma unroll
        for (int t = 0; t < kGateMaxT; ++t)
          if (t < T) gv[t] = fmaf(sW1[j * T + t], ghp, gv[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < kGateMaxT; ++t)
      if (t < T) {
        sv[tid * T + t] = v[t];
        if (live) a.gp[((int64_t)n * T + t) * a.C + c] = gv[t];
      }
    sz[tid * 3 + 0] = gz[0]; sz[tid * 3 + 1] = gz[1]; sz[tid * 3 + 2] = gz[2];
  }
  __syncthreads();
  // CTA partial of every parameter gradient: output o summed over the CTA's threads in thread order
  float* part = a.ws + (int64_t)blockIdx.x * n_part;
  for (int o = tid; o < n_part; o += kGThreads) {
    float acc = 0.f;
    if (o < H * T) {
      const int j = o / T, t = o % T;
      for (int q = 0; q < kGThreads; ++q) acc = fmaf(sg[q * H + j], sv[q * T + t], acc);
    } else if (o < H * T + 3 * H) {
      const int k = (o - H * T) / H, j = (o - H * T) % H;
      for (int q = 0; q < kGThreads; ++q) acc = fmaf(sz[q * 3 + k], sr[q * H + j], acc);
    } else if (o < H * T + 4 * H) {
      const int j = o - H * T - 3 * H;
      for (int q = 0; q < kGThreads; ++q) acc = fmaf(sb[q * H + j], sx[q * H + j], acc);
    } else {
      const int j = o - H * T - 4 * H;
      for (int q = 0; q < kGThreads; ++q) acc += sb[q * H + j];
    }
    part[o] = acc;
  }
  __threadfence();
  __syncthreads();
  int* ticket = reinterpret_cast<int*>(a.ws + (int64_t)gridDim.x * n_part);
  if (tid == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int o = tid; o < n_part; o += kGThreads) {
    // CTA order is kept (deterministic), but the loads go out eight at a time: a plain `acc += load` loop pays one L2
    // round trip per CTA and output (32 CTAs x 6 outputs per thread = ~60 us of pure latency on a 512-channel TAM)
    float acc = 0.f;
    for (unsigned b0 = 0; b0 < gridDim.x; b0 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (b0 + u < gridDim.x) ? __ldcg(a.ws + (int64_t)(b0 + u) * n_part + o) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    if (o < H * T) a.gW1[o] = acc;
    else if (o < H * T + 3 * H) a.gW2[o - H * T] = acc;
    else if (o < H * T + 4 * H) a.gbn1w[o - H * T - 3 * H] = acc;
    else a.gbn1b[o - H * T - 4 * H] = acc;
  }
  if (tid == 0) *ticket = 0;
}

// ------------------------------------------------------------------------------------------------
// L branch: generic 32x32-tile GEMM  D[m, n] = sum_k A(m, k) * B(n, k)  with stage-specific loaders / epilogues
// ------------------------------------------------------------------------------------------------
enum GateStage { kL2Fwd = 1, kGradWb, kGradHid, kGradWa, kGradP };   // (the hidden layer has its own kernel, tam_l1_fwd_kernel)

template <int STAGE>
__device__ __forceinline__ float gate_A(const GateArgs& a, int m, int k) {
  const int T = a.T, C = a.C, Hc = a.C / 4;
  if (STAGE == kL2Fwd) {            // hid[m, k] = relu(bn2(pre[m, k]))
    return fmaxf(gate_bn_apply(a.bn2, k, __ldg(a.pre + (int64_t)m * Hc + k)), 0.f);
  } else if (STAGE == kGradWb) {    // m = c, k = r  ->  gz[r, c]
    return a.gz[(int64_t)k * C + m];
  } else if (STAGE == kGradHid) {   // m = r, k = c  ->  gz[r, c]
    return a.gz[(int64_t)m * C + k];
  } else if (STAGE == kGradWa) {    // m = o, k = r  ->  gpre[r, o]
    return a.gpre[(int64_t)k * Hc + m];
  } else {                          // kGradP: m = (n, t'), k = j*Hc + o  ->  gpre[n, t' - j + 1, o]
    const int j = k / Hc, o = k - j * Hc;
    const int t = m % T - j + 1;
    return (t >= 0 && t < T) ? a.gpre[(int64_t)(m - m % T + t) * Hc + o] : 0.f;
  }
}

template <int STAGE>
__device__ __forceinline__ float gate_B(const GateArgs& a, int n, int k) {
  const int T = a.T, C = a.C, Hc = a.C / 4;
  if (STAGE == kL2Fwd) {            // n = c, k = o  ->  Wb[c][o]
    return __ldg(a.Wb + (int64_t)n * Hc + k);
  } else if (STAGE == kGradWb) {    // n = o, k = r  ->  hid[r, o]
    return fmaxf(gate_bn_apply(a.bn2, n, __ldg(a.pre + (int64_t)k * Hc + n)), 0.f);
  } else if (STAGE == kGradHid) {   // n = o, k = c  ->  Wb[c][o]
    return __ldg(a.Wb + (int64_t)k * Hc + n);
  } else if (STAGE == kGradWa) {    // n = c*3 + j, k = r = (n', t)  ->  p[n', t + j - 1, c]
    const int c = n / 3, j = n - 3 * c;
    const int t = k % T + j - 1;
    return (t >= 0 && t < T) ? __ldg(a.p + ((int64_t)(k - k % T + t)) * C + c) : 0.f;
  } else {                          // kGradP: n = c, k = j*Hc + o  ->  Wa[o][c][j]
    const int j = k / Hc, o = k - j * Hc;
    return __ldg(a.Wa + ((int64_t)o * C + n) * 3 + j);
  }
}

template <int STAGE>
__device__ __forceinline__ void gate_store(const GateArgs& a, int m, int n, float acc) {
  const int C = a.C, Hc = a.C / 4;
  if (STAGE == kL2Fwd) {
    a.act[(int64_t)m * C + n] = 1.f / (1.f + expf(-acc));
  } else if (STAGE == kGradWb) {
    a.gWb[(int64_t)m * Hc + n] = acc;
  } else if (STAGE == kGradHid) {   // through ReLU and the eval-mode BatchNorm
    const float hb = gate_bn_apply(a.bn2, n, __ldg(a.pre + (int64_t)m * Hc + n));
    const float g = hb > 0.f ? acc : 0.f;
    a.ghm[(int64_t)m * Hc + n] = g;
    a.gpre[(int64_t)m * Hc + n] = g * gate_bn_k(a.bn2, n);
  } else if (STAGE == kGradWa) {
    a.gWa[(int64_t)m * 3 * C + n] = acc;
  } else {
    a.gp[(int64_t)m * C + n] += acc;      // the G branch wrote its part first (same stream)
  }
}

// 32 x 32 output tile per CTA (16 x 16 threads, 2 x 2 outputs each), K walked in steps of kGateBK = 128: every thread has
// 32 independent operand loads in flight per step -- the operands are L2-resident, so the kernel is bound by L2 latency
// times the number of K steps (a 32-wide step made the K = 3C stage of a 512-channel TAM take ~50 us).
constexpr int kGateBK = 128;

template <int STAGE>
__global__ void __launch_bounds__(256) tam_gate_gemm_kernel(GateArgs a, int M, int Nn, int K) {
  __shared__ float As[kGateBK][33], Bs[kGateBK][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < K; k0 += kGateBK) {
    float av[kGateBK / 8], bv[kGateBK / 8];
#pragma unroll
    for (int i = 0; i < kGateBK / 8; ++i) {
      const int idx = threadIdx.x + i * 256;
      const int kk = idx % kGateBK, r = idx / kGateBK;
      const int k = k0 + kk;
      av[i] = (m0 + r < M && k < K) ? gate_A<STAGE>(a, m0 + r, k) : 0.f;
      bv[i] = (n0 + r < Nn && k < K) ? gate_B<STAGE>(a, n0 + r, k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kGateBK / 8; ++i) {
      const int idx = threadIdx.x + i * 256;
      As[idx % kGateBK][idx / kGateBK] = av[i];
      Bs[idx % kGateBK][idx / kGateBK] = bv[i];
    }
    __syncthreads();
#pragma unroll 16
    for (int kk = 0; kk < kGateBK; ++kk) {
      const float a0 = As[kk][ty], a1 = As[kk][ty + 16], b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m < M && n < Nn) gate_store<STAGE>(a, m, n, acc[i][j]);
    }
}

// gz = gact * act * (1 - act)
__global__ void __launch_bounds__(256) tam_gate_gz_kernel(GateArgs a, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float s = __ldg(a.act + i);
    a.gz[i] = __ldg(a.gact + i) * s * (1.f - s);
  }
}

// BatchNorm1d (eval) parameter gradients of the L branch.  CTA = 32 hidden channels x 8 row slices; a slice walks its
// rows eight at a time (loads first), the slices are added in slice order through shared memory (deterministic).
__global__ void __launch_bounds__(256) tam_gate_bn2_kernel(GateArgs a, int R) {
  __shared__ float sgw[8][33], sgb[8][33];
  const int Hc = a.C / 4;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + lane;
  float gw = 0.f, gb = 0.f;
  if (o < Hc) {
    const float rm = __ldg(a.bn2.rm + o), istd = 1.f / sqrtf(__ldg(a.bn2.rv + o) + a.bn2.eps);
    const int per = (R + 7) / 8;
    const int r0 = slice * per, r1 = min(R, r0 + per);
    for (int rb = r0; rb < r1; rb += 8) {
      float g[8], x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ok = rb + u < r1;
        g[u] = ok ? a.ghm[(int64_t)(rb + u) * Hc + o] : 0.f;
        x[u] = ok ? __ldg(a.pre + (int64_t)(rb + u) * Hc + o) : rm;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        gb += g[u];
        gw = fmaf(g[u], (x[u] - rm) * istd, gw);
      }
    }
  }
  sgw[slice][lane] = gw;
  sgb[slice][lane] = gb;
  __syncthreads();
  if (slice == 0 && o < Hc) {
    float tw = 0.f, tb = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { tw += sgw[q][lane]; tb += sgb[q][lane]; }
    a.gbn2w[o] = tw;
    a.gbn2b[o] = tb;
  }
}

// L branch, hidden layer (the K = 3C stage): CTA = (slice of hidden channels, video).  The video's pooled rows
// p[n] (T x C, plus a zero row on either side for the temporal padding) are staged in shared memory once; a warp owns
// one hidden channel at a time, its lanes stride over the input channels of one temporal tap (conflict-free smem rows,
// weights read once from L2) and keep T accumulators, reduced across the lanes at the end.
constexpr int kL1Slices = 8;

__global__ void __launch_bounds__(256) tam_l1_fwd_kernel(GateArgs a) {
  extern __shared__ float sp[];   // [(T + 2)][C]
  const int T = a.T, C = a.C, Hc = a.C / 4;
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (T + 2) * C; i += 256) {
    const int t = i / C - 1;
    sp[i] = (t >= 0 && t < T) ? __ldg(a.p + ((int64_t)n * T + t) * C + (i % C)) : 0.f;
  }
  __syncthreads();
  const int per = (Hc + kL1Slices - 1) / kL1Slices;
  const int o0 = blockIdx.x * per, o1 = min(Hc, o0 + per);
  for (int o = o0 + warp; o < o1; o += 8) {
    float acc[kGateMaxT];
#pragma unroll
    for (int t = 0; t < kGateMaxT; ++t) acc[t] = 0.f;
    const float* wrow = a.Wa + (int64_t)o * 3 * C;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      for (int c = lane; c < C; c += 32) {
        const float w = __ldg(wrow + c * 3 + j);
#pragma unroll
        for (int t = 0; t < kGateMaxT; ++t)
          if (t < T) acc[t] = fmaf(w, sp[(t + j) * C + c], acc[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < kGateMaxT; ++t) {
      if (t < T) {
        const float v = warp_sum(acc[t]);
        if (lane == 0) a.pre[((int64_t)n * T + t) * Hc + o] = v;
      }
    }
  }
}

template <int STAGE>
static void launch_stage(const GateArgs& a, int M, int Nn, int K, cudaStream_t st) {
  dim3 grid((unsigned)((Nn + 31) / 32), (unsigned)((M + 31) / 32));
  tam_gate_gemm_kernel<STAGE><<<grid, 256, 0, st>>>(a, M, Nn, K);
}

static int g_part_floats(int T) { return 2 * T * T + 3 * 2 * T + 2 * 2 * T; }
static size_t g_bwd_smem(int T) {
  const int H = 2 * T;
  return sizeof(float) * (size_t)(H * T + 3 * H + kGThreads * T + 4 * kGThreads * H + kGThreads * 3);
}

}  // namespace vitta

using namespace vitta;

static GateBN to_gate_bn(const VittaBN& b) { return GateBN{b.weight, b.bias, b.running_mean, b.running_var, b.eps}; }

static int check_gate_shape(int N, int T, int C) {
  VITTA_CHECK_ARG(N > 0 && T > 0 && T <= kGateMaxT && C >= 4 && C % 4 == 0, VITTA_E_UNSUPPORTED,
                  "tam_gate: needs T <= %d and C %% 4 == 0 (got T=%d C=%d)", kGateMaxT, T, C);
  return 0;
}

extern "C" {

int vitta_tam_gate_fwd(const float* p, const float* W1, VittaBN bn1, const float* W2, const float* Wa, VittaBN bn2,
                       const float* Wb, float* kern, float* act, float* pre, int N, int T, int C, void* stream) {
  VITTA_CHECK_ARG(p && W1 && W2 && Wa && Wb && kern && act && pre, VITTA_E_BADARG, "tam_gate_fwd: null pointer");
  int rc = check_gate_shape(N, T, C);
  if (rc) return rc;
  GateArgs a{};
  a.p = p; a.W1 = W1; a.W2 = W2; a.Wa = Wa; a.Wb = Wb; a.bn1 = to_gate_bn(bn1); a.bn2 = to_gate_bn(bn2);
  a.kern = kern; a.act = act; a.pre = pre; a.N = N; a.T = T; a.C = C;
  cudaStream_t st = (cudaStream_t)stream;
  tam_g_fwd_kernel<<<(unsigned)((N * C + kGThreads - 1) / kGThreads), kGThreads, 0, st>>>(a);
  VITTA_CHECK_LAUNCH();
  {
    const size_t smem = sizeof(float) * (size_t)(T + 2) * C;
    VITTA_CHECK_ARG(smem <= 200 * 1024, VITTA_E_UNSUPPORTED, "tam_gate_fwd: (T + 2) * C floats exceed shared memory");
    static size_t attr_smem = 48 * 1024;
    if (smem > attr_smem) {
      cudaError_t e = cudaFuncSetAttribute(tam_l1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_error("tam_gate_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
      }
      attr_smem = smem;
    }
    tam_l1_fwd_kernel<<<dim3(kL1Slices, (unsigned)N), 256, smem, st>>>(a);
  }
  VITTA_CHECK_LAUNCH();
  launch_stage<kL2Fwd>(a, N * T, C, C / 4, st);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int64_t vitta_tam_gate_bwd_ws_floats(int N, int T, int C) {
  if (N <= 0 || T <= 0 || T > kGateMaxT || C <= 0) return -1;
  const int64_t ctas = ((int64_t)N * C + kGThreads - 1) / kGThreads;
  return ctas * g_part_floats(T) + 4;
}

int vitta_tam_gate_bwd(const float* p, const float* W1, VittaBN bn1, const float* W2, const float* Wa, VittaBN bn2,
                       const float* Wb, const float* act, const float* pre, const float* gkern, const float* gact,
                       float* gp, float* gW1, float* gbn1w, float* gbn1b, float* gW2, float* gWa, float* gbn2w,
                       float* gbn2b, float* gWb, float* gz, float* gpre, float* ghm, float* ws, int N, int T, int C,
                       void* stream) {
  VITTA_CHECK_ARG(p && W1 && W2 && Wa && Wb && act && pre && gkern && gact && gp && gW1 && gbn1w && gbn1b && gW2 && gWa &&
                      gbn2w && gbn2b && gWb && gz && gpre && ghm && ws,
                  VITTA_E_BADARG, "tam_gate_bwd: null pointer");
  int rc = check_gate_shape(N, T, C);
  if (rc) return rc;
  GateArgs a{};
  a.p = p; a.W1 = W1; a.W2 = W2; a.Wa = Wa; a.Wb = Wb; a.bn1 = to_gate_bn(bn1); a.bn2 = to_gate_bn(bn2);
  a.act = const_cast<float*>(act); a.pre = const_cast<float*>(pre); a.gkern = gkern; a.gact = gact;
  a.gp = gp; a.gW1 = gW1; a.gbn1w = gbn1w; a.gbn1b = gbn1b; a.gW2 = gW2; a.gWa = gWa; a.gbn2w = gbn2w; a.gbn2b = gbn2b;
  a.gWb = gWb; a.gz = gz; a.gpre = gpre; a.ghm = ghm; a.ws = ws; a.N = N; a.T = T; a.C = C;
  cudaStream_t st = (cudaStream_t)stream;
  const int R = N * T, Hc = C / 4;
  // G branch: writes gp (its part), gW1, gW2, gbn1w, gbn1b
  static bool attr_done = false;
  const size_t smem = g_bwd_smem(kGateMaxT);
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tam_g_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("tam_gate_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  const unsigned g_ctas = (unsigned)((N * C + kGThreads - 1) / kGThreads);
  tam_g_bwd_kernel<<<g_ctas, kGThreads, g_bwd_smem(T), st>>>(a, g_part_floats(T));
  VITTA_CHECK_LAUNCH();
  // L branch
  const int64_t nz = (int64_t)R * C;
  tam_gate_gz_kernel<<<(unsigned)((nz + 255) / 256 < 592 ? (nz + 255) / 256 : 592), 256, 0, st>>>(a, nz);
  VITTA_CHECK_LAUNCH();
  launch_stage<kGradWb>(a, C, Hc, R, st);
  VITTA_CHECK_LAUNCH();
  launch_stage<kGradHid>(a, R, Hc, C, st);
  VITTA_CHECK_LAUNCH();
  tam_gate_bn2_kernel<<<(unsigned)((Hc + 31) / 32), 256, 0, st>>>(a, R);
  VITTA_CHECK_LAUNCH();
  launch_stage<kGradWa>(a, Hc, 3 * C, R, st);
  VITTA_CHECK_LAUNCH();
  launch_stage<kGradP>(a, R, C, 3 * Hc, st);
  VITTA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
