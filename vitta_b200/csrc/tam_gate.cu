// K5b: the two small networks of the Temporal Adaptive Module that produce the operands of the stencil kernel (tam.cu)
// from the spatially pooled activation p (N, T, C)  (reference models/tanet_models/temporal_module.py:27-41, 49-55):
//   G (global branch): per (video, channel)   softmax( W2 . relu(BN1d( W1 . p[n, :, c] )) )            -> kern (N, 3, C)
//   L (local branch):  per (video, frame)     sigmoid( Wb . relu(BN1d( conv1d_k3(Wa, p)[n, t, :] )) )   -> act  (N, T, C)
// with the BatchNorm1d layers in eval mode.  In eager PyTorch this is ~20 tiny kernels forward and ~25 backward per TAM
// (cutlass simt sgemm, ATen batch-norm / elementwise / reduce kernels: 16 TAMs -> ~700 launches per step).
//
// Everything is L2-resident (p is N*T*C <= 64 K floats, the largest weight 786 KB), so the cost is the NUMBER of dependent
// launches times their latency, not bytes or flops.  The work is therefore grouped by data dependence only -- every launch
// runs all the stages whose inputs are ready as different CTA "roles" of one grid:
//   forward   launch 1: L hidden layer (K = 3C)  |  G branch                              launch 2: L output layer
//   backward  launch 1: G branch (+ its parameter gradients)  |  dWb  |  gradient at the hidden layer
//             launch 2: dWa  |  gradient of p (L part, added to the G part)  |  BatchNorm1d(L) parameter gradients
// (3 + 7 launches before: the ncu launch list of round 2 had 62 us forward and 134 us backward per TAM, with 8-64 CTAs
// per launch.)  Inside the stages: the eval-mode BatchNorm constants are folded once per thread / CTA (an IEEE sqrt and
// divide per element and hidden unit dominated the per-column G code), every operand tile is loaded along its contiguous
// dimension, and the next K step's loads are in flight while the current one is multiplied.  All reductions run in a
// fixed order (deterministic).
#include "common.cuh"

namespace vitta {

constexpr int kGateMaxT = 16;
constexpr int kGateThreads = 256;

struct GateBN {
  const float *w, *b, *rm, *rv;
  float eps;
};

// folded eval-mode BatchNorm of unit i:  y = (x - rm) * k + b,  k = w / sqrt(rv + eps)
struct BNc {
  float rm, k, b, istd;
};
__device__ __forceinline__ BNc gate_bn_load(const GateBN& bn, int i) {
  BNc c;
  c.istd = 1.f / sqrtf(__ldg(bn.rv + i) + bn.eps);
  c.rm = __ldg(bn.rm + i);
  c.k = __ldg(bn.w + i) * c.istd;
  c.b = __ldg(bn.b + i);
  return c;
}
__device__ __forceinline__ float gate_bn_apply(const BNc& c, float x) { return fmaf(x - c.rm, c.k, c.b); }

struct GateArgs {
  const float* p;        // (N, T, C)
  const float *W1, *W2;  // (2T, T), (3, 2T)
  const float *Wa, *Wb;  // (C/4, C, 3), (C, C/4)
  GateBN bn1, bn2;
  float *kern, *act;     // (N, 3, C), (N, T, C)
  float* pre;            // (N*T, C/4): L hidden layer before its BatchNorm
  // backward
  const float *gkern, *gact;
  float *gpre, *ghm;          // (N*T, C/4), (N*T, C/4)
  float *gp;                  // (N, T, C): gradient of p (G part written first, L part added)
  float *gW1, *gW2, *gbn1w, *gbn1b, *gWa, *gWb, *gbn2w, *gbn2b;
  float* ws;                  // G-branch per-CTA partials + ticket
  int N, T, C;
};

// ------------------------------------------------------------------------------------------------
// G branch
// ------------------------------------------------------------------------------------------------
// shared memory of the G roles (floats): W1 [2T*T] | W2 [3*2T] | BN1 rm, k, b, istd [4][2T]
__device__ __forceinline__ void g_stage_params(const GateArgs& a, float* sW1, float* sW2, float* sbn) {
  const int T = a.T, H = 2 * T;
  for (int i = threadIdx.x; i < H * T; i += kGateThreads) sW1[i] = __ldg(a.W1 + i);
  for (int i = threadIdx.x; i < 3 * H; i += kGateThreads) sW2[i] = __ldg(a.W2 + i);
  if ((int)threadIdx.x < H) {
    const BNc c = gate_bn_load(a.bn1, threadIdx.x);
    sbn[threadIdx.x] = c.rm;
    sbn[H + threadIdx.x] = c.k;
    sbn[2 * H + threadIdx.x] = c.b;
    sbn[3 * H + threadIdx.x] = c.istd;
  }
}

// The per-column code keeps the loop over the 2T hidden units ROLLED (only the T-long dot products are unrolled): a CTA
// runs it once, so fully unrolled it was ~270 KB of straight-line SASS whose instruction fetch, not its arithmetic, set
// the 50 us floor of the first version of this file.
// No `t < T` predicates on the unrolled T loops (here, in the gv update and in the L hidden layer): v is zero-filled past T
// and w[t >= T] reads the next rows of the parameter block (finite), so the extra terms are exact zeros -- whereas the
// compiler turns every predicate into a uniform BRANCH around its LDS + FFMA pair, which exposes the full shared-memory
// latency 256 times per weight batch (the L hidden layer ran 12 us per batch that way).
__device__ __forceinline__ float g_dot(const float* w, const float* v) {
  float h = 0.f;
#pragma unroll
  for (int t = 0; t < kGateMaxT; ++t) h = fmaf(w[t], v[t], h);
  return h;
}

// forward of one (n, c) column: softmax s[3] of W2 . relu(BN(W1 . v))
__device__ __forceinline__ void g_forward(int T, const float* sW1, const float* sW2, const float* sbn, const float* v,
                                          float* s) {
  const int H = 2 * T;
  float z0 = 0.f, z1 = 0.f, z2 = 0.f;
#pragma unroll 1
  for (int j = 0; j < H; ++j) {
    const float h = g_dot(sW1 + j * T, v);
    const float r = fmaxf(fmaf(h - sbn[j], sbn[H + j], sbn[2 * H + j]), 0.f);
    z0 = fmaf(sW2[j], r, z0);
    z1 = fmaf(sW2[H + j], r, z1);
    z2 = fmaf(sW2[2 * H + j], r, z2);
  }
  const float m = fmaxf(z0, fmaxf(z1, z2));
  const float e0 = expf(z0 - m), e1 = expf(z1 - m), e2 = expf(z2 - m);
  const float inv = 1.f / (e0 + e1 + e2);
  s[0] = e0 * inv; s[1] = e1 * inv; s[2] = e2 * inv;
}

__device__ __forceinline__ void g_load_column(const GateArgs& a, int n, int c, bool live, float* v) {
#pragma unroll
  for (int t = 0; t < kGateMaxT; ++t) v[t] = (live && t < a.T) ? __ldg(a.p + ((int64_t)n * a.T + t) * a.C + c) : 0.f;
}

// forward role: CTA = 256 columns
__device__ __forceinline__ void g_fwd_role(const GateArgs& a, int cta, float* sm) {
  const int T = a.T, H = 2 * T;
  float *sW1 = sm, *sW2 = sW1 + H * T, *sbn = sW2 + 3 * H;
  const int idx = cta * kGateThreads + threadIdx.x;
  const bool live = idx < a.N * a.C;
  const int n = live ? idx / a.C : 0, c = live ? idx % a.C : 0;
  float v[kGateMaxT], s[3];
  g_load_column(a, n, c, live, v);          // in flight while the parameters are staged
  g_stage_params(a, sW1, sW2, sbn);
  __syncthreads();
  if (!live) return;
  g_forward(T, sW1, sW2, sbn, v, s);
  a.kern[((int64_t)n * 3 + 0) * a.C + c] = s[0];
  a.kern[((int64_t)n * 3 + 1) * a.C + c] = s[1];
  a.kern[((int64_t)n * 3 + 2) * a.C + c] = s[2];
}

// backward role: CTA = kGCols columns (warps 0-3, one column per thread), all 8 warps reduce.
// Partial layout per CTA: gW1 [2T*T] | gW2 [3*2T] | gbn_w [2T] | gbn_b [2T]; the last CTA adds the partials in CTA order.
// The column threads leave their contributions to the parameter gradients in shared memory (ghp and the column for the
// outer-product sum gW1, five more values per hidden unit for the small ones); the sums over the columns then run in
// column order, one output per thread (no warp shuffles: inside a role branch the compiler brackets every shuffle with
// WARPSYNC / ENDCOLLECTIVE, 800 of them in a row in the first version).
constexpr int kGCols = 128;
__host__ __device__ constexpr int g_part_floats(int T) { return 2 * T * T + 3 * 2 * T + 2 * 2 * T; }
__host__ __device__ constexpr int g_small_floats(int T) { return 3 * 2 * T + 2 * 2 * T; }
__host__ __device__ constexpr int g_bwd_smem_floats(int T) {
  return 2 * T * T + 3 * 2 * T + 4 * 2 * T + kGCols * (2 * T + 1) + kGCols * (T + 1) + kGCols * (g_small_floats(T) + 1);
}

__device__ __forceinline__ void g_bwd_role(const GateArgs& a, int cta, int n_ctas, float* sm) {
  __shared__ int s_last;
  const int T = a.T, H = 2 * T;
  const int n_part = g_part_floats(T), n_small = g_small_floats(T), qs = n_small + 1;
  float* sW1 = sm;                          // H*T
  float* sW2 = sW1 + H * T;                 // 3*H
  float* sbn = sW2 + 3 * H;                 // 4*H
  float* sg = sbn + 4 * H;                  // [kGCols][H + 1]   gradient at the hidden layer before its BatchNorm
  float* sv = sg + kGCols * (H + 1);        // [kGCols][T + 1]   the column
  float* sq = sv + kGCols * (T + 1);        // [kGCols][5H + 1]  per-column terms of gW2 | gbn_w | gbn_b
  const int tid = threadIdx.x;
  const int idx = cta * kGCols + tid;
  const bool col = tid < kGCols, live = col && idx < a.N * a.C;
  const int n = live ? idx / a.C : 0, c = live ? idx % a.C : 0;
  float v[kGateMaxT];
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  g_load_column(a, n, c, live, v);          // in flight while the parameters are staged
  if (live) {
    g0 = __ldg(a.gkern + ((int64_t)n * 3 + 0) * a.C + c);
    g1 = __ldg(a.gkern + ((int64_t)n * 3 + 1) * a.C + c);
    g2 = __ldg(a.gkern + ((int64_t)n * 3 + 2) * a.C + c);
  }
  g_stage_params(a, sW1, sW2, sbn);
  __syncthreads();
  if (col) {
    float s[3], gz0 = 0.f, gz1 = 0.f, gz2 = 0.f;
    g_forward(T, sW1, sW2, sbn, v, s);
    if (live) {
      const float dot = g0 * s[0] + g1 * s[1] + g2 * s[2];
      gz0 = s[0] * (g0 - dot); gz1 = s[1] * (g1 - dot); gz2 = s[2] * (g2 - dot);
    }
    float gv[kGateMaxT];
#pragma unroll
    for (int t = 0; t < kGateMaxT; ++t) gv[t] = 0.f;
    float* q = sq + tid * qs;
#pragma unroll 1
    for (int j = 0; j < H; ++j) {
      const float* w = sW1 + j * T;
      const float h = g_dot(w, v);
      const float hb = fmaf(h - sbn[j], sbn[H + j], sbn[2 * H + j]);
      const float r = fmaxf(hb, 0.f);
      const float ghr = sW2[j] * gz0 + sW2[H + j] * gz1 + sW2[2 * H + j] * gz2;
      const float ghb = hb > 0.f ? ghr : 0.f;       // gz == 0 on dead columns
      const float ghp = ghb * sbn[H + j];
      const float xh = (h - sbn[j]) * sbn[3 * H + j];
      sg[tid * (H + 1) + j] = ghp;
      q[j] = gz0 * r; q[H + j] = gz1 * r; q[2 * H + j] = gz2 * r;
      q[3 * H + j] = ghb * xh; q[4 * H + j] = ghb;
#pragma unroll
      for (int t = 0; t < kGateMaxT; ++t) gv[t] = fmaf(w[t], ghp, gv[t]);     // gv[t >= T] is never stored
    }
#pragma unroll
    for (int t = 0; t < kGateMaxT; ++t)
      if (t < T) {
        sv[tid * (T + 1) + t] = v[t];
        if (live) a.gp[((int64_t)n * T + t) * a.C + c] = gv[t];
      }
  }
  __syncthreads();
  // CTA partial of every parameter gradient, columns added in column order
  float* part = a.ws + (int64_t)cta * n_part;
  for (int o = tid; o < H * T; o += kGateThreads) {      // gW1[j][t] = sum_q ghp[q][j] * v[q][t]
    const int j = o / T, t = o - j * T;
    float acc = 0.f;
#pragma unroll 8
    for (int qq = 0; qq < kGCols; ++qq) acc = fmaf(sg[qq * (H + 1) + j], sv[qq * (T + 1) + t], acc);
    part[o] = acc;
  }
  if (tid < n_small) {
    float acc = 0.f;
#pragma unroll 8
    for (int qq = 0; qq < kGCols; ++qq) acc += sq[qq * qs + tid];
    part[H * T + tid] = acc;
  }
  __threadfence();
  __syncthreads();
  int* ticket = reinterpret_cast<int*>(a.ws + (int64_t)n_ctas * n_part);
  if (tid == 0) s_last = (atomicAdd(ticket, 1) == n_ctas - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // a thread owns outputs tid, tid + 256, tid + 512 (n_part <= 672): the partials of all three are loaded together, eight
  // CTAs at a time, and added in CTA order
  static_assert(g_part_floats(kGateMaxT) <= 3 * kGateThreads, "three outputs per thread cover the partial vector");
  float acc[3] = {0.f, 0.f, 0.f};
  for (int b0 = 0; b0 < n_ctas; b0 += 8) {
    float pv[3][8];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int o = tid + i * kGateThreads;
        pv[i][u] = (o < n_part && b0 + u < n_ctas) ? __ldcg(a.ws + (int64_t)(b0 + u) * n_part + o) : 0.f;
      }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[i] += pv[i][u];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int o = tid + i * kGateThreads;
    if (o >= n_part) continue;
    if (o < H * T) a.gW1[o] = acc[i];
    else if (o < H * T + 3 * H) a.gW2[o - H * T] = acc[i];
    else if (o < H * T + 4 * H) a.gbn1w[o - H * T - 3 * H] = acc[i];
    else a.gbn1b[o - H * T - 4 * H] = acc[i];
  }
  if (tid == 0) *ticket = 0;
}

// ------------------------------------------------------------------------------------------------
// L branch, hidden layer (the K = 3C stage): CTA = (slice of hidden channels, video).  The video's pooled rows
// p[n] (T x C, plus a zero row on either side for the temporal padding) are staged in shared memory once; a warp owns
// TWO hidden channels at a time (one shared-memory read feeds both), its lanes walk the contiguous weight rows
// Wa[o][c][j] with 128-bit loads (eight in flight per lane) and keep 2 x T accumulators, reduced across the lanes at the
// end.  (Before: one 4-byte weight load per lane and iteration with nothing else in flight -- 35 us at C = 512.)
// ------------------------------------------------------------------------------------------------
constexpr int kL1Slices = 8;

__device__ __forceinline__ float f4c(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

__device__ __forceinline__ void l1_fwd_role(const GateArgs& a, int slice, int n, float* sp) {
  const int T = a.T, C = a.C, Hc = a.C / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* scr = sp + (kGateMaxT + 2) * C + warp * (32 * 33);     // per-warp reduction tile, behind the staged rows
  {
    const int C4 = C / 4;
    float4* sp4 = reinterpret_cast<float4*>(sp);
    for (int i = threadIdx.x; i < (T + 2) * C4; i += kGateThreads) {
      const int t = i / C4 - 1;
      sp4[i] = (t >= 0 && t < T) ? ldg4(a.p + ((int64_t)n * T + t) * C + (i % C4) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  const int per = (Hc + kL1Slices - 1) / kL1Slices;
  const int o0 = slice * per, o1 = min(Hc, o0 + per);
  const int n4 = 3 * C / 4;     // float4s per weight row
  for (int o = o0 + 2 * warp; o < o1; o += 2 * (kGateThreads / 32)) {
    const bool two = o + 1 < o1;
    const float4* wr0 = reinterpret_cast<const float4*>(a.Wa + (int64_t)o * 3 * C);
    const float4* wr1 = wr0 + (two ? n4 : 0);
    float acc0[kGateMaxT], acc1[kGateMaxT];
#pragma unroll
    for (int t = 0; t < kGateMaxT; ++t) acc0[t] = acc1[t] = 0.f;
    for (int b = 0; b < n4; b += 128) {
      float4 wa[4], wb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = b + lane + 32 * u;
        const bool ok = q < n4;
        wa[u] = ok ? __ldg(wr0 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        wb[u] = (ok && two) ? __ldg(wr1 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e0 = 4 * (b + lane + 32 * u);
        if (e0 < 3 * C) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = e0 + k;
            const int c = e / 3, j = e - 3 * c;
            const float* s = sp + j * C + c;
            const float x0 = f4c(wa[u], k), x1 = f4c(wb[u], k);
#pragma unroll
            for (int t = 0; t < kGateMaxT; ++t) {     // rows past T + 1 are allocated but never written: acc[t >= T] is dropped
              const float pv = s[t * C];
              acc0[t] = fmaf(x0, pv, acc0[t]);
              acc1[t] = fmaf(x1, pv, acc1[t]);
            }
          }
        }
      }
    }
    // sums over the lanes through the warp's scratch tile (lane m adds column m: 2 x T outputs), no shuffles
    __syncwarp();
#pragma unroll
    for (int t = 0; t < kGateMaxT; ++t) {
      scr[lane * 33 + t] = acc0[t];
      scr[lane * 33 + kGateMaxT + t] = acc1[t];
    }
    __syncwarp();
    {
      float tot = 0.f;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) tot += scr[l * 33 + lane];
      const int t = lane & (kGateMaxT - 1), second = lane >> 4;
      if (t < T && (!second || two)) a.pre[((int64_t)n * T + t) * Hc + o + second] = tot;
    }
  }
}

// forward launch 1: blocks [0, kL1Slices * N) = L hidden layer, the rest = G branch
__global__ void __launch_bounds__(kGateThreads) tam_gate_fwd1_kernel(GateArgs a) {
  extern __shared__ __align__(16) float dsm[];
  const int nl1 = kL1Slices * a.N;
  if ((int)blockIdx.x < nl1) l1_fwd_role(a, blockIdx.x % kL1Slices, blockIdx.x / kL1Slices, dsm);
  else g_fwd_role(a, blockIdx.x - nl1, dsm);
}

// ------------------------------------------------------------------------------------------------
// L branch: 32x32-tile GEMM  D[m, n] = sum_k A(m, k) * B(n, k)  with stage-specific loaders / epilogues
// ------------------------------------------------------------------------------------------------
enum GateStage { kL2Fwd = 1, kGradWb, kGradHid, kGradWa, kGradP };   // (the hidden layer has its own role, l1_fwd_role)

// which index of an operand element is contiguous in memory: the K index (false) or the tile row (true).  The loader
// threads walk that one, so a warp reads whole 128-byte lines either way.
template <int STAGE> struct GateLayout;
template <> struct GateLayout<kL2Fwd>   { static constexpr bool a_row = false, b_row = false; };
template <> struct GateLayout<kGradWb>  { static constexpr bool a_row = true,  b_row = true;  };
template <> struct GateLayout<kGradHid> { static constexpr bool a_row = false, b_row = true;  };
template <> struct GateLayout<kGradWa>  { static constexpr bool a_row = true,  b_row = true;  };
template <> struct GateLayout<kGradP>   { static constexpr bool a_row = false, b_row = true;  };

// gradient at the output layer before the sigmoid: gz = gact * act * (1 - act) (formed on the fly: no gz pass / buffer)
__device__ __forceinline__ float gate_gz(const GateArgs& a, int64_t i) {
  const float s = __ldg(a.act + i);
  return __ldg(a.gact + i) * s * (1.f - s);
}

template <int STAGE>
__device__ __forceinline__ float gate_A(const GateArgs& a, int m, int k) {
  const int T = a.T, C = a.C, Hc = a.C / 4;
  if (STAGE == kL2Fwd) {            // pre[m, k]  (BatchNorm + ReLU applied by the loader: see gate_gemm_tile)
    return __ldg(a.pre + (int64_t)m * Hc + k);
  } else if (STAGE == kGradWb) {    // m = c, k = r  ->  gz[r, c]
    return gate_gz(a, (int64_t)k * C + m);
  } else if (STAGE == kGradHid) {   // m = r, k = c  ->  gz[r, c]
    return gate_gz(a, (int64_t)m * C + k);
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
  } else if (STAGE == kGradWb) {    // n = o, k = r  ->  pre[r, o]  (BatchNorm + ReLU applied by the loader)
    return __ldg(a.pre + (int64_t)k * Hc + n);
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
  } else if (STAGE == kGradWa) {
    a.gWa[(int64_t)m * 3 * C + n] = acc;
  } else if (STAGE == kGradP) {
    a.gp[(int64_t)m * C + n] += acc;      // the G branch wrote its part first (previous launch, same stream)
  }
}

// 32 x 32 output tile (16 x 16 threads, 2 x 2 outputs each), K walked in steps of kGateBK = 128: every thread has 32
// independent operand loads in flight per step, and the loads of step i+1 are issued before step i is multiplied.
constexpr int kGateBK = 128;
constexpr int kGateLd = kGateBK / 8;      // elements per thread, operand and step
constexpr int kGemmSmemFloats = 2 * kGateBK * 33;

template <bool ROW>
__device__ __forceinline__ void gate_map(int i, int& kk, int& r) {
  const int idx = threadIdx.x + i * kGateThreads;
  if (ROW) { r = idx & 31; kk = idx >> 5; }             // thread's row fixed (tid & 31), kk = (tid >> 5) + 8 i
  else { kk = idx % kGateBK; r = idx / kGateBK; }       // thread's kk fixed (tid & 127), r = (tid >> 7) + 2 i
}

template <int STAGE>
__device__ __forceinline__ void gate_gemm_tile(const GateArgs& a, int m0, int n0, int M, int Nn, int K, float* sm) {
  using L = GateLayout<STAGE>;
  float (*As)[33] = reinterpret_cast<float (*)[33]>(sm);
  float (*Bs)[33] = reinterpret_cast<float (*)[33]>(sm + kGateBK * 33);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int Hc = a.C / 4;
  // folded BatchNorm(L) constants this thread needs: kL2Fwd -- of hidden unit k0 + (tid & 127) (reloaded per K step);
  // kGradWb -- of hidden unit n0 + (tid & 31) (B operand rows); kGradHid -- of the two output columns of the epilogue
  BNc bnc = {0.f, 0.f, 0.f, 0.f}, bnc2 = {0.f, 0.f, 0.f, 0.f};
  if (STAGE == kGradWb) bnc = gate_bn_load(a.bn2, min(n0 + (int)(threadIdx.x & 31), Hc - 1));
  if (STAGE == kGradHid) {
    bnc = gate_bn_load(a.bn2, min(n0 + tx, Hc - 1));
    bnc2 = gate_bn_load(a.bn2, min(n0 + tx + 16, Hc - 1));
  }
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float av[kGateLd], bv[kGateLd];
  auto load = [&](int k0) {
    if (STAGE == kL2Fwd) bnc = gate_bn_load(a.bn2, min(k0 + (int)(threadIdx.x & 127), Hc - 1));
#pragma unroll
    for (int i = 0; i < kGateLd; ++i) {
      int kk, r;
      gate_map<L::a_row>(i, kk, r);
      av[i] = (m0 + r < M && k0 + kk < K) ? gate_A<STAGE>(a, m0 + r, k0 + kk) : 0.f;
      gate_map<L::b_row>(i, kk, r);
      bv[i] = (n0 + r < Nn && k0 + kk < K) ? gate_B<STAGE>(a, n0 + r, k0 + kk) : 0.f;
    }
  };
  load(0);
  for (int k0 = 0; k0 < K; k0 += kGateBK) {
#pragma unroll
    for (int i = 0; i < kGateLd; ++i) {
      int kk, r;
      gate_map<L::a_row>(i, kk, r);
      float x = av[i];
      if (STAGE == kL2Fwd) x = (m0 + r < M && k0 + kk < K) ? fmaxf(gate_bn_apply(bnc, x), 0.f) : 0.f;
      As[kk][r] = x;
      gate_map<L::b_row>(i, kk, r);
      float y = bv[i];
      if (STAGE == kGradWb) y = (n0 + r < Nn && k0 + kk < K) ? fmaxf(gate_bn_apply(bnc, y), 0.f) : 0.f;
      Bs[kk][r] = y;
    }
    __syncthreads();
    if (k0 + kGateBK < K) load(k0 + kGateBK);
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
      if (m < M && n < Nn) {
        if (STAGE == kGradHid) {   // through ReLU and the eval-mode BatchNorm
          const BNc& c = j ? bnc2 : bnc;
          const float hb = gate_bn_apply(c, __ldg(a.pre + (int64_t)m * Hc + n));
          const float g = hb > 0.f ? acc[i][j] : 0.f;
          a.ghm[(int64_t)m * Hc + n] = g;
          a.gpre[(int64_t)m * Hc + n] = g * c.k;
        } else {
          gate_store<STAGE>(a, m, n, acc[i][j]);
        }
      }
    }
}

__host__ __device__ inline int gate_tiles(int M, int Nn) { return ((M + 31) / 32) * ((Nn + 31) / 32); }

// tile index -> (m0, n0), n fastest
template <int STAGE>
__device__ __forceinline__ void gate_gemm_role(const GateArgs& a, int tile, int M, int Nn, int K, float* sm) {
  const int tn = (Nn + 31) / 32;
  gate_gemm_tile<STAGE>(a, (tile / tn) * 32, (tile % tn) * 32, M, Nn, K, sm);
}

// forward launch 2: act = sigmoid(relu(BN2(pre)) . Wb^T)
__global__ void __launch_bounds__(kGateThreads) tam_gate_fwd2_kernel(GateArgs a) {
  __shared__ __align__(16) float sm[kGemmSmemFloats];
  gate_gemm_role<kL2Fwd>(a, blockIdx.x, a.N * a.T, a.C, a.C / 4, sm);
}

// BatchNorm1d (eval) parameter gradients of the L branch.  CTA = 32 hidden channels x 8 row slices; a slice walks its
// rows eight at a time (loads first), the slices are added in slice order through shared memory (deterministic).
__device__ __forceinline__ void gate_bn2_role(const GateArgs& a, int cta, float* sm) {
  float (*sgw)[33] = reinterpret_cast<float (*)[33]>(sm);
  float (*sgb)[33] = reinterpret_cast<float (*)[33]>(sm + 8 * 33);
  const int Hc = a.C / 4, R = a.N * a.T;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int o = cta * 32 + lane;
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

// backward launch 1: [gradient at the hidden layer (K = C: the longest role first)] [G branch] [dWb]
__global__ void __launch_bounds__(kGateThreads) tam_gate_bwd1_kernel(GateArgs a, int n_hid, int n_g) {
  extern __shared__ __align__(16) float sm[];     // max(GEMM tile buffers, G backward role): see bwd1_smem_bytes
  const int R = a.N * a.T, Hc = a.C / 4;
  int b = blockIdx.x;
  if (b < n_hid) { gate_gemm_role<kGradHid>(a, b, R, Hc, a.C, sm); return; }
  b -= n_hid;
  if (b < n_g) { g_bwd_role(a, b, n_g, sm); return; }
  b -= n_g;
  gate_gemm_role<kGradWb>(a, b, a.C, Hc, R, sm);
}

// backward launch 2: [gradient of p, L part (K = 3C/4)] [dWa] [BatchNorm1d(L) parameter gradients]
__global__ void __launch_bounds__(kGateThreads) tam_gate_bwd2_kernel(GateArgs a, int n_gp, int n_wa) {
  __shared__ __align__(16) float sm[kGemmSmemFloats];
  const int R = a.N * a.T, Hc = a.C / 4;
  int b = blockIdx.x;
  if (b < n_gp) { gate_gemm_role<kGradP>(a, b, R, a.C, 3 * Hc, sm); return; }
  b -= n_gp;
  if (b < n_wa) { gate_gemm_role<kGradWa>(a, b, Hc, 3 * a.C, R, sm); return; }
  b -= n_wa;
  gate_bn2_role(a, b, sm);
}

}  // namespace vitta

using namespace vitta;

static GateBN to_gate_bn(const VittaBN& b) { return GateBN{b.weight, b.bias, b.running_mean, b.running_var, b.eps}; }

static int check_gate_shape(int N, int T, int C) {
  // T >= 2: the unpredicated T loops read up to 16 - T floats past a W1 row, which must stay inside W1 | W2 | BN1
  VITTA_CHECK_ARG(N > 0 && T >= 2 && T <= kGateMaxT && C >= 4 && C % 4 == 0, VITTA_E_UNSUPPORTED,
                  "tam_gate: needs 2 <= T <= %d and C %% 4 == 0 (got T=%d C=%d)", kGateMaxT, T, C);
  return 0;
}

static int g_bwd_ctas(int N, int C) { return (N * C + kGCols - 1) / kGCols; }
static size_t bwd1_smem_bytes(int T) {
  const int g = g_bwd_smem_floats(T);
  return sizeof(float) * (size_t)(g > kGemmSmemFloats ? g : kGemmSmemFloats);
}

extern "C" {

int vitta_tam_gate_fwd(const float* p, const float* W1, VittaBN bn1, const float* W2, const float* Wa, VittaBN bn2,
                       const float* Wb, float* kern, float* act, float* pre, int N, int T, int C, void* stream) {
  VITTA_CHECK_ARG(p && W1 && W2 && Wa && Wb && kern && act && pre, VITTA_E_BADARG, "tam_gate_fwd: null pointer");
  int rc = check_gate_shape(N, T, C);
  if (rc) return rc;
  VITTA_CHECK_ARG(aligned16(p) && aligned16(Wa), VITTA_E_ALIGN, "tam_gate_fwd: p and Wa must be 16-byte aligned");
  GateArgs a{};
  a.p = p; a.W1 = W1; a.W2 = W2; a.Wa = Wa; a.Wb = Wb; a.bn1 = to_gate_bn(bn1); a.bn2 = to_gate_bn(bn2);
  a.kern = kern; a.act = act; a.pre = pre; a.N = N; a.T = T; a.C = C;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int H = 2 * T;
    size_t smem = sizeof(float) * ((size_t)(kGateMaxT + 2) * C + (kGateThreads / 32) * 32 * 33);
    const size_t g_smem = sizeof(float) * (size_t)(H * T + 3 * H + 4 * H);
    if (smem < g_smem) smem = g_smem;
    VITTA_CHECK_ARG(smem <= 200 * 1024, VITTA_E_UNSUPPORTED, "tam_gate_fwd: 18 * C floats exceed shared memory");
    static size_t attr_smem = 48 * 1024;
    if (smem > attr_smem) {
      cudaError_t e = cudaFuncSetAttribute(tam_gate_fwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_error("tam_gate_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
      }
      attr_smem = smem;
    }
    const unsigned n_g = (unsigned)((N * C + kGateThreads - 1) / kGateThreads);
    tam_gate_fwd1_kernel<<<(unsigned)(kL1Slices * N) + n_g, kGateThreads, smem, st>>>(a);
  }
  VITTA_CHECK_LAUNCH();
  tam_gate_fwd2_kernel<<<(unsigned)gate_tiles(N * T, C), kGateThreads, 0, st>>>(a);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int64_t vitta_tam_gate_bwd_ws_floats(int N, int T, int C) {
  if (N <= 0 || T < 2 || T > kGateMaxT || C <= 0) return -1;
  return (int64_t)g_bwd_ctas(N, C) * g_part_floats(T) + 4;
}

int vitta_tam_gate_bwd(const float* p, const float* W1, VittaBN bn1, const float* W2, const float* Wa, VittaBN bn2,
                       const float* Wb, const float* act, const float* pre, const float* gkern, const float* gact,
                       float* gp, float* gW1, float* gbn1w, float* gbn1b, float* gW2, float* gWa, float* gbn2w,
                       float* gbn2b, float* gWb, float* gpre, float* ghm, float* ws, int N, int T, int C,
                       void* stream) {
  VITTA_CHECK_ARG(p && W1 && W2 && Wa && Wb && act && pre && gkern && gact && gp && gW1 && gbn1w && gbn1b && gW2 && gWa &&
                      gbn2w && gbn2b && gWb && gpre && ghm && ws,
                  VITTA_E_BADARG, "tam_gate_bwd: null pointer");
  int rc = check_gate_shape(N, T, C);
  if (rc) return rc;
  GateArgs a{};
  a.p = p; a.W1 = W1; a.W2 = W2; a.Wa = Wa; a.Wb = Wb; a.bn1 = to_gate_bn(bn1); a.bn2 = to_gate_bn(bn2);
  a.act = const_cast<float*>(act); a.pre = const_cast<float*>(pre); a.gkern = gkern; a.gact = gact;
  a.gp = gp; a.gW1 = gW1; a.gbn1w = gbn1w; a.gbn1b = gbn1b; a.gW2 = gW2; a.gWa = gWa; a.gbn2w = gbn2w; a.gbn2b = gbn2b;
  a.gWb = gWb; a.gpre = gpre; a.ghm = ghm; a.ws = ws; a.N = N; a.T = T; a.C = C;
  cudaStream_t st = (cudaStream_t)stream;
  const int R = N * T, Hc = C / 4;
  // launch 1 -- G branch: gp (its part), gW1, gW2, gbn1w, gbn1b;  L branch: gWb, ghm, gpre
  const int n_hid = gate_tiles(R, Hc), n_g = g_bwd_ctas(N, C), n_wb = gate_tiles(C, Hc);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tam_gate_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bwd1_smem_bytes(kGateMaxT));
    if (e != cudaSuccess) {
      set_error("tam_gate_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  tam_gate_bwd1_kernel<<<(unsigned)(n_hid + n_g + n_wb), kGateThreads, bwd1_smem_bytes(T), st>>>(a, n_hid, n_g);
  VITTA_CHECK_LAUNCH();
  // launch 2 -- L branch: gp += its part, gWa, gbn2w, gbn2b
  const int n_gp = gate_tiles(R, C), n_wa = gate_tiles(Hc, 3 * C), n_bn = (Hc + 31) / 32;
  tam_gate_bwd2_kernel<<<(unsigned)(n_gp + n_wa + n_bn), kGateThreads, 0, st>>>(a, n_gp, n_wa);
  VITTA_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
