// K7: Video-Swin 3-D (shifted-)window multi-head self-attention, head_dim 32, fp32-grade (3xTF32) on tcgen05.
//
//   replaces WindowAttention3D.forward's core (swin_transformer.py:145-166) AND the data movement around it in
//   SwinTransformerBlock3D.forward_part1 (:229-248): torch.roll, window_partition, the qkv permute, the relative-position
//   bias gather, the (0 / -100) shift mask tensor, softmax, attn @ v, window_reverse and the roll back.  All of those are
//   index maps here: a work item = (batch, window, head); its N = Wd*Wh*Ww tokens are gathered straight from the
//   (B, D, H, W, 3, heads, 32) qkv tensor with the cyclic shift folded into the addresses, and the output rows are
//   scattered back to their un-shifted token positions.
//
// Forward kernel (480 threads, 1 CTA / SM, persistent over a contiguous item range, head-major so the bias table of a
// head is loaded once; details at wmsa3d_fwd_kernel):
//   warps 8-11 loaders: gather K (whole window), per 128-row tile Q (pre-scaled), per 32-key chunk V; Q and K are split
//              into tf32 hi / lo in the K-major SWIZZLE_128B layout, V into fp16 hi / lo (16-bit MN-major SWIZZLE_128B)
//   warp  12   issues tcgen05.mma.kind::tf32:  S[128 x N] = Q K^T  (3 MMAs per k-step: lo*hi, hi*lo, hi*hi) into TMEM
//              columns [0, 400)
//   warps 0-7  two softmax groups (thread = query row = TMEM lane) taking the 32-column chunks alternately: pass 1 adds
//              bias[rel(i, j)] and the shift mask to S in place and finds the row maximum; pass 2 exponentiates,
//              accumulates the row sum and writes P (fp16 hi / lo pairs) in place over the S columns
//   warps 13, 14  one PV issuer per group: O_g[128 x 32] += P_chunk V_chunk on kind::f16 (columns [400, 432), [432, 464));
//              the epilogue adds the two accumulators and writes O / rowsum with the log-sum-exp
//   The whole score row lives in TMEM, so there is no online-softmax rescaling and the N x N matrix never touches HBM.
// Backward: wmsa3d_bwd3_kernel<MODE> (tcgen05, kind::f16 operand pairs; two launches); wmsa3d_bwd_kernel is an exact fp32
// FFMA2 kernel (one CTA per item with Q, K, V, dO resident in shared memory) kept as an on-device cross-check.
#include <cuda_fp16.h>

#include "tc05.cuh"

namespace vitta {

constexpr int kAtMaxKeys = 400;     // 392 padded to a multiple of 16 (UMMA N granularity)
constexpr int kAtMaxRel = 2560;     // (2*8-1)*(2*7-1)*(2*7-1) = 2535 table rows

struct WmsaGeom {
  int B, D, H, W, heads;
  int ws0, ws1, ws2;   // window, clamped to the volume (get_window_size, swin_transformer.py:71-84)
  int ss0, ss1, ss2;   // cyclic shift, clamped the same way (0 where the window covers the dimension)
  int fw0, fw1, fw2;   // configured window: geometry of relative_position_index[:N, :N] (:113-125, :148-149)
  int nw0, nw1, nw2;   // windows per dimension
  int N;               // tokens per window
  int NP;              // N rounded up to 16
  int nrel;            // rows of the bias table
};

// per-token bookkeeping of one window: source token offset and packed (relative-index base | region id << 16)
__device__ __forceinline__ void window_token(const WmsaGeom& g, int b, int wd, int wh, int ww, int i, int& tok, int& info) {
  if (i >= g.N) {
    tok = -1;
    info = (int)0x80000000;
    return;
  }
  const int hw = g.ws1 * g.ws2;
  const int a = i / hw, rem = i - a * hw;
  const int bb = rem / g.ws2, cc = rem - bb * g.ws2;
  const int p0 = wd * g.ws0 + a, p1 = wh * g.ws1 + bb, p2 = ww * g.ws2 + cc;   // position in the rolled volume
  // region ids of compute_mask (swin_transformer.py:316-329)
  const int r0 = g.ss0 ? (p0 < g.D - g.ws0 ? 0 : (p0 < g.D - g.ss0 ? 1 : 2)) : 0;
  const int r1 = g.ss1 ? (p1 < g.H - g.ws1 ? 0 : (p1 < g.H - g.ss1 ? 1 : 2)) : 0;
  const int r2 = g.ss2 ? (p2 < g.W - g.ws2 ? 0 : (p2 < g.W - g.ss2 ? 1 : 2)) : 0;
  int s0 = p0 + g.ss0, s1 = p1 + g.ss1, s2 = p2 + g.ss2;                       // torch.roll(x, -shift): source index
  if (s0 >= g.D) s0 -= g.D;
  if (s1 >= g.H) s1 -= g.H;
  if (s2 >= g.W) s2 -= g.W;
  tok = ((b * g.D + s0) * g.H + s1) * g.W + s2;
  const int fhw = g.fw1 * g.fw2;
  const int fd = i / fhw, fr = i - fd * fhw;
  const int fh = fr / g.fw2, fw = fr - fh * g.fw2;
  const int bj = (fd * (2 * g.fw1 - 1) + fh) * (2 * g.fw2 - 1) + fw;
  info = bj | ((r0 * 9 + r1 * 3 + r2) << 16);
}
__device__ __forceinline__ int rel_row_base(const WmsaGeom& g) {
  return ((g.fw0 - 1) * (2 * g.fw1 - 1) + (g.fw1 - 1)) * (2 * g.fw2 - 1) + (g.fw2 - 1);
}

__device__ __forceinline__ uint32_t sw128_off(int row, int q) {   // K-major SWIZZLE_128B, 128-B rows, q = float4 index
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((q ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint32_t mn32_off(int row, int q) {    // MN-major SWIZZLE_128B_BASE32B, 128-B rows
  return (uint32_t)(row * 128 + (((((q >> 1) ^ (row & 3)) << 1) | (q & 1)) << 4));
}
__device__ __forceinline__ void split4(float4 v, float4& h, float4& l) {
  h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
  l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
}

// ------------------------------------------------------------------------------------------------
// shared memory plan of the forward kernel (bytes from the 1024-aligned base)
// ------------------------------------------------------------------------------------------------
constexpr int kAtColPad = 416;                           // per-token arrays padded to a multiple of the 32-column chunk
constexpr int kFwdVStages = 8;
constexpr int kOffKhi = 0;
constexpr int kOffKlo = kOffKhi + kAtMaxKeys * 128;      //  51200
constexpr int kOffQhi = kOffKlo + kAtMaxKeys * 128;      // 102400
constexpr int kOffQlo = kOffQhi + 128 * 128;             // 118784
constexpr int kOffV = kOffQlo + 128 * 128;               // 135168: 8 stages x (hi atom 4 KB, lo atom 4 KB)
constexpr int kOffTab = kOffV + kFwdVStages * 8192;      // bias table of the current head (* log2 e)
constexpr int kOffInfo = kOffTab + kAtMaxRel * 4;
constexpr int kOffTok = kOffInfo + 2 * kAtColPad * 4;    // (info and tok: two buffers each, by item parity)
constexpr int kOffBar = kOffTok + 2 * kAtMaxKeys * 4;
constexpr int kOffXch = kOffBar + 512;                   // row maximum / row sum exchange between the two softmax groups
constexpr int kAtSmemBytes = kOffXch + 2 * 2 * 128 * 4 + 1024;   // ~218 KB
// TMEM columns of the forward kernel: S [0, 400); the P chunk c (fp16 hi | lo, two keys per column) overwrites S columns
// [32c, 32c + 32) once the softmax has consumed them; one O accumulator per softmax group / PV issuer at 400 and 432
constexpr uint32_t kFT_O = 400;

struct WmsaFwdParams {
  const float* qkv;     // (B, D, H, W, 3, heads, 32)
  const float* table;   // (nrel, heads)
  const float* qkv_amax;   // device scalar >= max|qkv|: scale of V's fp16 operand split
  float* out;           // (B, D, H, W, heads*32)
  float* lse;           // ((b*nW + w)*heads + head)*N + i
  float scale;
  int items, items_per_cta;
  WmsaGeom g;
  float* amax_out;      // optional: max|out| (range of the fp16-split proj GEMM and its weight gradient)
  unsigned long long* trace;   // TRACE instantiation only: 15 warps x trace_cap records (0 = unused)
  int trace_cap;
};

constexpr int kPRing = 8;   // P_READY barriers per group: a group runs at most 7 chunks (one tile) ahead of its issuer
enum { B_KV_READY = 0, B_KV_FREE, B_TAB_FREE0, B_TAB_FREE1, B_Q_READY, B_Q_FREE, B_S_FULL, B_S_FREE, B_O_FULL, B_O_FREE,
       B_V_READY0, B_V_FREE0 = B_V_READY0 + kFwdVStages, B_P_READY0 = B_V_FREE0 + kFwdVStages,
       B_COUNT = B_P_READY0 + 2 * kPRing };
static_assert(B_COUNT * 8 + 8 <= 512, "barrier block");

// D[tmem] (+)= A[tmem, packed fp16 pairs along K] * B[smem], kind::f16, issued by the lane with pe != 0
__device__ __forceinline__ void umma_f16_ts_p(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum,
                                              uint32_t pe) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum), "r"(pe)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// warps 0-3 softmax group A, 4-7 softmax group B (TMEM lane quadrant = warp & 3), 8-11 loaders, 12 S issuer (+ TMEM
// owner), 13 PV issuer of group A, 14 PV issuer of group B.
//
// What bounds this kernel is the ISSUE of the P V MMAs, not the tensor pipe and not the softmax arithmetic (the kernel's
// own time stamps: tools/wmsa_trace.py, profiles/r02_wmsa_fwd_timeline.md).  With tf32 operands a 32-key chunk is 4 K steps
// x 3 split products = 12 MMAs of 128 x 32 x 8, each ~100 cycles of issue in one thread (~16 cycles of tensor work), plus
// ~750 cycles of hand-over per chunk: 27 K of the 40 K cycles of a 128-row tile, during which eight softmax warps wait.
// Hence:
//   * P and V enter the P V product as fp16 hi / lo pairs (kind::f16, K = 16 per MMA: 6 MMAs per chunk).  P = 2^(t - m + 10)
//     lies in [0, 1024]; V is scaled by the power of two that puts max|qkv| below 2^14 (the split
//     the convolution GEMMs use; hi + lo carry 22 mantissa bits).  Output and log-sum-exp undo both scales exactly.
//   * the 32-column chunks of a tile alternate between two softmax groups (parity of a global chunk counter), in both
//     passes, and EACH GROUP HAS ITS OWN PV ISSUER AND ITS OWN O ACCUMULATOR: two independent issue streams; the epilogue
//     adds the two accumulators.
//   * a P chunk (16 packed hi + 16 packed lo columns) overwrites the S columns it was computed from, so there is no P
//     buffer to hand back: a group writes all its chunks back to back, announcing each on a ring of mbarriers.
//   pass 1 adds bias + mask in place (log2 domain) and takes the partial row maximum of the group's chunks; the two
//   partial maxima (and later the partial sums) of a row meet in shared memory behind a 64-thread named barrier.
// Because P lives in the S columns, S is released for the next tile's Q K^T by the PV issuers (commit after their last
// chunk), not by the softmax warps.  The epilogue is split too: each group normalises and stores 16 of the 32 channels.
constexpr int kFwdThreads = 480;

#define WMSA_TR(id)                                                                                              \
  do {                                                                                                           \
    if (TRACE && blockIdx.x == 0 && lane == 0 && tr_n < (uint32_t)p.trace_cap) {                                 \
      p.trace[(size_t)warp * p.trace_cap + tr_n] = ((unsigned long long)clock64() << 16) | ((unsigned long long)warp << 8) | (id); \
      ++tr_n;                                                                                                    \
    }                                                                                                            \
  } while (0)

template <bool TRACE>
__global__ void __launch_bounds__(kFwdThreads, 1) wmsa3d_fwd_kernel(const WmsaFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the extern array (no integer round trip) keeps the shared address space: LDS/STS, not LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
  float* tab = reinterpret_cast<float*>(smem + kOffTab);
  int* info_all = reinterpret_cast<int*>(smem + kOffInfo);   // [2][kAtColPad]: per-token arrays of the item, double-buffered
  int* tok_all = reinterpret_cast<int*>(smem + kOffTok);     // [2][kAtMaxKeys]  by item parity
  float* xch_max = reinterpret_cast<float*>(smem + kOffXch);     // [2][128]
  float* xch_sum = xch_max + 2 * 128;                            // [2][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t tr_n = 0;
  (void)tr_n;
  const WmsaGeom& g = p.g;
  const int C = g.heads * 32;
  const int nwin = g.nw0 * g.nw1 * g.nw2;
  const int nwin_total = g.B * nwin;
  const int n_tiles = (g.N + 127) >> 7;
  const int n_chunks = (g.N + 31) >> 5;
  const int item0 = blockIdx.x * p.items_per_cta;
  const int item1 = min(p.items, item0 + p.items_per_cta);

  if (threadIdx.x == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      int cnt = 1;   // tcgen05.commit barriers
      if (i == B_KV_READY || i == B_Q_READY || (i >= B_V_READY0 && i < B_V_FREE0) || i >= B_P_READY0)
        cnt = 4;     // one elected arrive per warp of a 4-warp role
      if (i == B_TAB_FREE0 || i == B_TAB_FREE0 + 1 || i == B_O_FREE) cnt = 8;   // both softmax groups
      if (i == B_S_FREE || i == B_O_FULL) cnt = 2;     // both PV issuers
      mbar_init(&bar[i], cnt);
    }
    fence_barrier_init();
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // broadcast from lane 0: the value is the same in every lane, but only a shuffle tells the compiler so -- without it
  // everything derived from it travels through R2UR.BROADCAST + ELECT in front of every lane-predicated tcgen05.mma
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp >= 8 && warp < 12) {
    // =========================== loaders ===========================
    const int lt = threadIdx.x - 256;
    const int rslot = lt >> 3, q4 = lt & 7;
    int cur_head = -1;
    uint32_t it = 0, tile_ctr = 0, chunk_ctr = 0;
    float sv, sv_inv_unused;
    f16_split_scale(__ldg(p.qkv_amax), sv, sv_inv_unused);
    // Set-up of one item: token maps, bias table on a change of head, the K gather.  tok[] / info[] are double-buffered by
    // item parity and the set-up of item i + 1 runs INSIDE the last tile of item i, between its first eight V chunks (the
    // ring is full then: the loaders would only wait) and the rest: it needs just the S MMAs of that tile to have retired
    // (KV_FREE), while the softmax passes and PV products of the tile are still running.  Done at the item boundary it cost
    // +9.5 K cycles per item (the K gather alone is four dependent batches of global loads): profiles/r02_wmsa_fwd_timeline_*.
    // The bias table is single: a change of head (once per ~100 items) waits for the previous item's TAB_FREE too.
    auto prepare_item = [&](int item, uint32_t it) {
      const int head = item / nwin_total;
      const int wg = item - head * nwin_total;
      const int b = wg / nwin;
      int w = wg - b * nwin;
      const int ww = w % g.nw2; w /= g.nw2;
      const int wh = w % g.nw1;
      const int wd = w / g.nw1;
      int* info = info_all + (it & 1) * kAtColPad;
      int* tok = tok_all + (it & 1) * kAtMaxKeys;
      mbar_wait(&bar[B_KV_FREE], (it & 1) ^ 1);
      mbar_wait(&bar[B_TAB_FREE0 + (it & 1)], ((it >> 1) & 1) ^ 1);
      if (head != cur_head && it > 0) mbar_wait(&bar[B_TAB_FREE0 + ((it - 1) & 1)], ((it - 1) >> 1) & 1);
      WMSA_TR(40);
      for (int i = lt; i < kAtColPad; i += 128) {
        int t = -1, f = 31 << 16;   // padding columns: region id 31 = always masked (and their V rows are zero)
        if (i < g.N) window_token(g, b, wd, wh, ww, i, t, f);
        if (i < kAtMaxKeys) tok[i] = t;
        info[i] = f;
      }
      if (head != cur_head) {
        for (int i = lt; i < g.nrel; i += 128)
          tab[i] = __ldg(p.table + (int64_t)i * g.heads + head) * 1.4426950408889634f;   // bias * log2(e)
        cur_head = head;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // tok[] complete (loader warps only)
      const float* qkv_h = p.qkv + head * 32 + q4 * 4;
      // Every loop below issues a BATCH of independent global loads into registers before touching shared memory, so a
      // thread pays the HBM/L2 latency once per batch instead of once per row.
      // ---- K: all keys of the window (8 rows per thread and batch)
      for (int r0 = rslot; r0 < g.NP; r0 += 128) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + u * 16;
          const int t = (r < g.NP) ? tok[r] : -1;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t >= 0) v[u] = ldg4(qkv_h + ((int64_t)t * 3 + 1) * C);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + u * 16;
          if (r < g.NP) {
            float4 h, l;
            split4(v[u], h, l);
            const uint32_t o = sw128_off(r, q4);
            *reinterpret_cast<float4*>(smem + kOffKhi + o) = h;
            *reinterpret_cast<float4*>(smem + kOffKlo + o) = l;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[B_KV_READY]);
      WMSA_TR(41);
    };
    if (item0 < item1) prepare_item(item0, 0);
    for (int item = item0; item < item1; ++item, ++it) {
      const int head = item / nwin_total;
      const int* tok = tok_all + (it & 1) * kAtMaxKeys;
      const float* qkv_h = p.qkv + head * 32 + q4 * 4;
      // V chunks: groups of 4 chunks (8 float4 per thread), the next group in flight while the current one is stored
      auto v_issue = [&](float4 (&v)[8], int grp) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = grp * 128 + (u >> 1) * 32 + rslot + (u & 1) * 16;
          const int t = (grp * 4 + (u >> 1) < n_chunks && j < g.N) ? tok[j] : -1;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t >= 0) v[u] = ldg4(qkv_h + ((int64_t)t * 3 + 2) * C);
        }
      };
      // a V chunk in shared memory: two 16-bit MN-major SWIZZLE_128B atoms (hi, lo) of 32 keys x 128 bytes -- a row is one
      // key holding its 32 channels as fp16 in the first four 16-byte chunks (chunk index XOR key % 8), the layout the
      // weight-gradient kernel feeds its X operand with (wgrad_tf32.cu); the MMAs read N = 32 of the 64 columns
      auto v_drain = [&](float4 (&v)[8], int grp) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (grp * 4 + cc >= n_chunks) break;
          const int st = chunk_ctr & (kFwdVStages - 1);
          mbar_wait(&bar[B_V_FREE0 + st], ((chunk_ctr / kFwdVStages) & 1) ^ 1);
          uint8_t* vb = smem + kOffV + st * 8192;
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int r = rslot + rr * 16;
            const float4 x = v[cc * 2 + rr];
            const float x0 = x.x * sv, x1 = x.y * sv, x2 = x.z * sv, x3 = x.w * sv;
            const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
            const uint32_t o = (uint32_t)r * 128u + ((((uint32_t)q4 >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)q4 & 1u) * 8u;
            *reinterpret_cast<uint2*>(vb + o) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            *reinterpret_cast<uint2*>(vb + 4096 + o) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar[B_V_READY0 + st]);
          ++chunk_ctr;
        }
      };
      const int n_groups = (n_chunks + 3) >> 2;
      for (int tile = 0; tile < n_tiles; ++tile, ++tile_ctr) {
        // Q tile: loads go to registers BEFORE the wait for the previous tile's S MMAs to release the buffer
        float4 qv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = tile * 128 + rslot + u * 16;
          const int t = (i < g.N) ? tok[i] : -1;
          qv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t >= 0) qv[u] = ldg4(qkv_h + (int64_t)t * 3 * C);
        }
        float4 va[8], vb8[8];
        v_issue(va, 0);
        mbar_wait(&bar[B_Q_FREE], (tile_ctr & 1) ^ 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = rslot + u * 16;
          float4 v = qv[u];
          v.x *= p.scale; v.y *= p.scale; v.z *= p.scale; v.w *= p.scale;   // q = q * scale (:146)
          float4 h, l;
          split4(v, h, l);
          const uint32_t o = sw128_off(r, q4);
          *reinterpret_cast<float4*>(smem + kOffQhi + o) = h;
          *reinterpret_cast<float4*>(smem + kOffQlo + o) = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[B_Q_READY]);
        WMSA_TR(42);
        for (int grp = 0; grp < n_groups; grp += 2) {
          if (grp + 1 < n_groups) v_issue(vb8, grp + 1);
          v_drain(va, grp);
          if (grp + 2 < n_groups) v_issue(va, grp + 2);
          if (grp + 1 < n_groups) v_drain(vb8, grp + 1);
          // the next item's set-up, once, after the first (up to) eight V chunks of this item's last tile
          if (grp == 0 && tile == n_tiles - 1 && item + 1 < item1) prepare_item(item + 1, it + 1);
        }
      }
    }
  } else if (warp == 12) {
    // =========================== S issuer: S[128 x NP] = Q K^T ===========================
    // every lane runs the warp-uniform control flow, only the instruction is predicated on lane 0 (uniform operands)
    const uint32_t pe = (lane == 0) ? 1u : 0u;
    const int part0 = g.NP < 256 ? g.NP : 256;
    const int part1 = g.NP - part0;
    const uint32_t idesc_s0 = umma_idesc_tf32(128, part0);
    const uint32_t idesc_s1 = part1 ? umma_idesc_tf32(128, part1) : 0;
    const uint32_t sbase = smem_u32(smem);
    const uint64_t q_hi = umma_desc_sw128(sbase + kOffQhi), q_lo = umma_desc_sw128(sbase + kOffQlo);
    uint32_t it = 0, tile_ctr = 0;
    for (int item = item0; item < item1; ++item, ++it) {
      for (int tile = 0; tile < n_tiles; ++tile, ++tile_ctr) {
        if (tile == 0) mbar_wait(&bar[B_KV_READY], it & 1);
        mbar_wait(&bar[B_Q_READY], tile_ctr & 1);
        mbar_wait(&bar[B_S_FREE], (tile_ctr & 1) ^ 1);   // the previous tile's PV MMAs (which read P from these columns) retired
        tc_fence_after();
        WMSA_TR(30);
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          if (part == 1 && part1 == 0) break;
          const uint32_t koff = part ? 256 * 128 : 0;
          const uint64_t k_hi = umma_desc_sw128(sbase + kOffKhi + koff), k_lo = umma_desc_sw128(sbase + kOffKlo + koff);
          const uint32_t d = tmem_base + (part ? 256u : 0u);
          const uint32_t idesc = part ? idesc_s1 : idesc_s0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            umma_tf32_p(d, q_lo + adv, k_hi + adv, idesc, k != 0, pe);
            umma_tf32_p(d, q_hi + adv, k_lo + adv, idesc, 1, pe);
            umma_tf32_p(d, q_hi + adv, k_hi + adv, idesc, 1, pe);
          }
        }
        umma_commit_p(&bar[B_Q_FREE], pe);
        if (tile == n_tiles - 1) umma_commit_p(&bar[B_KV_FREE], pe);
        umma_commit_p(&bar[B_S_FULL], pe);
        WMSA_TR(31);
      }
    }
  } else if (warp == 13 || warp == 14) {
    // =========================== PV issuers: O_grp[128 x 32] += P_chunk V_chunk over the group's chunks ===============
    const int grp = __shfl_sync(0xffffffffu, warp - 13, 0);   // warp-uniform by construction; the shuffle makes it provable
    const uint32_t pe = (lane == 0) ? 1u : 0u;
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 32) | (1u << 16);   // A in TMEM (packed fp16 pairs), B (= V) MN-major
    const uint32_t sbase = smem_u32(smem);
    const uint32_t d = tmem_base + kFT_O + (uint32_t)grp * 32u;
    uint32_t tile_ctr = 0, chunk_ctr = 0;
    for (int item = item0; item < item1; ++item) {
      for (int tile = 0; tile < n_tiles; ++tile, ++tile_ctr) {
        uint32_t first = 1;
        for (int c = 0; c < n_chunks; ++c, ++chunk_ctr) {
          const uint32_t ccu = __shfl_sync(0xffffffffu, chunk_ctr, 0);   // uniform registers for the descriptors / addresses
          if ((ccu & 1u) != (uint32_t)grp) continue;
          const int st = ccu & (kFwdVStages - 1);
          const uint32_t own = ccu >> 1;                                  // index of this chunk among the group's chunks
          const uint32_t cu = __shfl_sync(0xffffffffu, (uint32_t)c, 0);
          mbar_wait(&bar[B_P_READY0 + grp * kPRing + (own & (kPRing - 1))], (own / kPRing) & 1);
          WMSA_TR(20);
          mbar_wait(&bar[B_V_READY0 + st], (chunk_ctr / kFwdVStages) & 1);
          if (first) mbar_wait(&bar[B_O_FREE], (tile_ctr & 1) ^ 1);
          tc_fence_after();
          WMSA_TR(21);
          const uint32_t vb = sbase + kOffV + st * 8192;
          const uint64_t v_hi = umma_desc_mn_sw128_f16(vb, 4096), v_lo = umma_desc_mn_sw128_f16(vb + 4096, 4096);
          const int left = g.N - c * 32;
          const int ksteps = left >= 32 ? 2 : (left + 15) >> 4;
          const int cols = min(32, g.NP - c * 32);
          const uint32_t p_hi = tmem_base + cu * 32u, p_lo = p_hi + (uint32_t)(cols >> 1);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (k < ksteps) {
              const uint64_t advb = (uint64_t)(k * (2048 >> 4));   // 16 keys = two 8-row atoms of 1024 B
              const uint32_t ka = (uint32_t)(k * 8);               // 16 fp16 of P = 8 TMEM columns
              umma_f16_ts_p(d, p_lo + ka, v_hi + advb, idesc_o, (first && k == 0) ? 0u : 1u, pe);
              umma_f16_ts_p(d, p_hi + ka, v_lo + advb, idesc_o, 1, pe);
              umma_f16_ts_p(d, p_hi + ka, v_hi + advb, idesc_o, 1, pe);
            }
          }
          first = 0;
          umma_commit_p(&bar[B_V_FREE0 + st], pe);
          WMSA_TR(22);
        }
        // (a group without a chunk in this tile -- single-chunk windows -- commits at once and its accumulator is not read)
        umma_commit_p(&bar[B_O_FULL], pe);
        umma_commit_p(&bar[B_S_FREE], pe);
      }
    }
  } else {
    // =========================== softmax / epilogue (thread = query row = TMEM lane, group = chunk parity) ===========
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int rel0 = rel_row_base(g);
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kMask2 = -100.f * kLog2e;
    constexpr uint32_t kNegInf = 0xff800000u;
    uint32_t it = 0, tile_ctr = 0, chunk_ctr = 0;
    float out_amax = 0.f;
    float sv_unused, sv_inv;
    f16_split_scale(__ldg(p.qkv_amax), sv_unused, sv_inv);
    for (int item = item0; item < item1; ++item, ++it) {
      const int head = item / nwin_total;
      const int wg = item - head * nwin_total;
      mbar_wait(&bar[B_KV_READY], it & 1);   // tab / info / tok of this item are in place
      const int* info = info_all + (it & 1) * kAtColPad;
      const int* tok = tok_all + (it & 1) * kAtMaxKeys;
      for (int tile = 0; tile < n_tiles; ++tile, ++tile_ctr) {
        const int i = tile * 128 + row;
        const bool valid = i < g.N;
        // a warp whose 32 rows all lie past the window (last tile: 392 = 3 x 128 + 8) only keeps the handshakes going
        const bool warp_live = tile * 128 + quad * 32 < g.N;
        const int fi = info[valid ? i : 0];
        const int a_i = (fi & 0xffff) + rel0;
        const int r_i = fi & 0x1f0000;
        const int my_tok = tok[valid ? i : 0];   // read before TAB_FREE is released (the loaders reuse tok[] / info[])
        mbar_wait(&bar[B_S_FULL], tile_ctr & 1);
        tc_fence_after();
        WMSA_TR(1);
        // ---- pass 1 (log2 domain): t = s*log2e + bias2 + mask2, in place; partial row maximum over this group's chunks.
        //      Padding columns carry region id 31 and are therefore always masked; a 16-column last chunk is completed
        //      with -inf in registers (no predicates inside the unrolled loops: they become branches around every
        //      shared-memory load).
        float m2 = -INFINITY;
        if (warp_live) {
          for (int c = 0; c < n_chunks; ++c) {
            if (((chunk_ctr + c) & 1) != (uint32_t)grp) continue;
            const int cols = min(32, g.NP - c * 32);
            uint32_t r[32];
            tmem_ld16(t_lane + (uint32_t)(c * 32), r);
            if (cols > 16) {
              tmem_ld16(t_lane + (uint32_t)(c * 32 + 16), r + 16);
            } else {
#pragma unroll
              for (int jj = 16; jj < 32; ++jj) r[jj] = kNegInf;
            }
            tmem_ld_wait();
            const int* ic = info + c * 32;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const int fj = ic[jj];
              float t = fmaf(__uint_as_float(r[jj]), kLog2e, tab[a_i - (fj & 0xffff)]);
              t += ((fj & 0x1f0000) != r_i) ? kMask2 : 0.f;
              m2 = fmaxf(m2, t);
              r[jj] = __float_as_uint(t);
            }
            tmem_st16(t_lane + (uint32_t)(c * 32), r);
            if (cols > 16) tmem_st16(t_lane + (uint32_t)(c * 32 + 16), r + 16);
          }
          tmem_st_wait();
        }
        WMSA_TR(2);
        // the two partial maxima of a row meet (warps quad and quad + 4)
        xch_max[grp * 128 + row] = m2;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
        m2 = fmaxf(m2, xch_max[(grp ^ 1) * 128 + row]);
        WMSA_TR(3);
        if (tile == n_tiles - 1) {   // last use of tab / info by this warp for this item
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar[B_TAB_FREE0 + (it & 1)]);
        }
        // ---- pass 2: p = 2^(t - m2 + 10), partial row sum; P chunks (fp16 hi / lo pairs) go back to tensor memory, over
        //      the columns they came from, as the A operand of P V
        float l = 0.f;
        const float m2s = m2 - 10.f;
        for (int c = 0; c < n_chunks; ++c, ++chunk_ctr) {
          if ((chunk_ctr & 1) != (uint32_t)grp) continue;
          const uint32_t own = chunk_ctr >> 1;
          if (warp_live) {
            const int cols = min(32, g.NP - c * 32);
            uint32_t r[32], hi[16], lo[16];
            tmem_ld16(t_lane + (uint32_t)(c * 32), r);
            if (cols > 16) {
              tmem_ld16(t_lane + (uint32_t)(c * 32 + 16), r + 16);
            } else {
#pragma unroll
              for (int jj = 16; jj < 32; ++jj) r[jj] = kNegInf;
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float p0, p1;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(__uint_as_float(r[2 * j]) - m2s));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(__uint_as_float(r[2 * j + 1]) - m2s));
              l += p0;
              l += p1;
              const __half2 h = __floats2half2_rn(p0, p1);          // key 2j in the low half (lower K index)
              const float2 hf = __half22float2(h);
              const __half2 lw = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
              hi[j] = *reinterpret_cast<const uint32_t*>(&h);
              lo[j] = *reinterpret_cast<const uint32_t*>(&lw);
            }
            if (cols > 16) {
              tmem_st16(t_lane + (uint32_t)(c * 32), hi);
              tmem_st16(t_lane + (uint32_t)(c * 32 + 16), lo);
            } else {   // 16 keys: 8 + 8 columns
              tmem_st8(t_lane + (uint32_t)(c * 32), hi);
              tmem_st8(t_lane + (uint32_t)(c * 32 + 8), lo);
            }
            tmem_st_wait();
            tc_fence_before();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar[B_P_READY0 + grp * kPRing + (own & (kPRing - 1))]);
          WMSA_TR(12);
        }
        // the two partial sums of a row meet
        xch_sum[grp * 128 + row] = l;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
        l += xch_sum[(grp ^ 1) * 128 + row];
        WMSA_TR(5);
        // ---- (O_A + O_B) / l -> global: this group's 16 of the 32 head channels
        mbar_wait(&bar[B_O_FULL], tile_ctr & 1);
        tc_fence_after();
        WMSA_TR(6);
        uint32_t oa[16], ob[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) oa[q] = ob[q] = 0u;
        if (warp_live) {
          // single-chunk windows: only the accumulator of the chunk's parity was written in this tile
          const uint32_t first_par = (chunk_ctr - (uint32_t)n_chunks) & 1u;
          if (n_chunks > 1 || first_par == 0u) tmem_ld16(t_lane + kFT_O + (uint32_t)(grp * 16), oa);
          if (n_chunks > 1 || first_par == 1u) tmem_ld16(t_lane + kFT_O + 32u + (uint32_t)(grp * 16), ob);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[B_O_FREE]);
        if (valid) {
          const float inv = sv_inv / l;    // l carries the factor 2^10 of P, the accumulators that factor and V's scale
          float* dst = p.out + (int64_t)my_tok * C + head * 32 + grp * 16;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 ov;
            ov.x = (__uint_as_float(oa[q * 4]) + __uint_as_float(ob[q * 4])) * inv;
            ov.y = (__uint_as_float(oa[q * 4 + 1]) + __uint_as_float(ob[q * 4 + 1])) * inv;
            ov.z = (__uint_as_float(oa[q * 4 + 2]) + __uint_as_float(ob[q * 4 + 2])) * inv;
            ov.w = (__uint_as_float(oa[q * 4 + 3]) + __uint_as_float(ob[q * 4 + 3])) * inv;
            st4(dst + q * 4, ov);
            out_amax = fmaxf(out_amax, fmaxf(fmaxf(fabsf(ov.x), fabsf(ov.y)), fmaxf(fabsf(ov.z), fabsf(ov.w))));
          }
          if (grp == 0) p.lse[((int64_t)wg * g.heads + head) * g.N + i] = (m2s + log2f(l)) * 0.6931471805599453f;
        }
        WMSA_TR(7);
      }
    }
    if (p.amax_out) {   // one integer atomic per softmax warp (non-negative floats order like their bit patterns)
      const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(out_amax));
      if (lane == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(p.amax_out), wmax);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// backward (v0): exact fp32 on the FFMA2 pipe.  One CTA (384 threads) per item; Qs = scale*Q, K, V, dO of the window live
// in shared memory (broadcast reads), the bias table and its gradient too.
//   pass A (thread = query i):  dQ_i = scale * sum_j dS_ij K_j,   dTable[rel(i,j)] += dS_ij
//   pass B (thread = key j):    dK_j = sum_i dS_ij Qs_i,          dV_j = sum_i P_ij dO_i
//   with P_ij = exp(S_ij - lse_i), dS_ij = P_ij (dO_i . V_j - dO_i . O_i)
// ------------------------------------------------------------------------------------------------
constexpr int kAtBwdThreads = 384;   // 12 warps = 3 per SM sub-partition -> 168 registers / thread

struct WmsaBwdParams {
  const float* qkv;
  const float* table;
  const float* out;     // forward output (B, D, H, W, C)
  const float* dout;    // gradient of it
  const float* lse;
  float* dqkv;          // (B, D, H, W, 3, heads, 32), every element written exactly once
  float* dtable;        // (nrel, heads), accumulated with red.global.add
  float scale;
  int items, items_per_cta;
  WmsaGeom g;
};

constexpr int kBwRows = 392;
constexpr int kBwOffQ = 0;
constexpr int kBwOffK = kBwOffQ + kBwRows * 128;
constexpr int kBwOffV = kBwOffK + kBwRows * 128;
constexpr int kBwOffDO = kBwOffV + kBwRows * 128;
constexpr int kBwOffTab = kBwOffDO + kBwRows * 128;        // 200704
constexpr int kBwOffDTab = kBwOffTab + kAtMaxRel * 4;
constexpr int kBwOffLse = kBwOffDTab + kAtMaxRel * 4;
constexpr int kBwOffDsum = kBwOffLse + kAtMaxKeys * 4;
constexpr int kBwOffInfo = kBwOffDsum + kAtMaxKeys * 4;
constexpr int kBwOffTok = kBwOffInfo + kAtMaxKeys * 4;
constexpr int kBwSmemBytes = kBwOffTok + kAtMaxKeys * 4 + 16;   // 227600

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__global__ void __launch_bounds__(kAtBwdThreads, 1) wmsa3d_bwd_kernel(const WmsaBwdParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* sQ = reinterpret_cast<float*>(smem + kBwOffQ);
  float* sK = reinterpret_cast<float*>(smem + kBwOffK);
  float* sV = reinterpret_cast<float*>(smem + kBwOffV);
  float* sDO = reinterpret_cast<float*>(smem + kBwOffDO);
  float* tab = reinterpret_cast<float*>(smem + kBwOffTab);
  float* dtab = reinterpret_cast<float*>(smem + kBwOffDTab);
  float* sLse = reinterpret_cast<float*>(smem + kBwOffLse);
  float* sDs = reinterpret_cast<float*>(smem + kBwOffDsum);
  int* info = reinterpret_cast<int*>(smem + kBwOffInfo);
  int* tok = reinterpret_cast<int*>(smem + kBwOffTok);
  const WmsaGeom& g = p.g;
  const int tid = threadIdx.x;
  const int C = g.heads * 32;
  const int nwin = g.nw0 * g.nw1 * g.nw2;
  const int nwin_total = g.B * nwin;
  const int rel0 = rel_row_base(g);
  const int item0 = blockIdx.x * p.items_per_cta;
  const int item1 = min(p.items, item0 + p.items_per_cta);
  int cur_head = -1;

  auto flush_dtab = [&](int head) {
    for (int i = tid; i < g.nrel; i += kAtBwdThreads) {
      const float v = dtab[i];
      if (v != 0.f) atomicAdd(p.dtable + (int64_t)i * g.heads + head, v);
    }
  };

  for (int item = item0; item < item1; ++item) {
    const int head = item / nwin_total;
    const int wg = item - head * nwin_total;
    const int b = wg / nwin;
    int w = wg - b * nwin;
    const int ww = w % g.nw2; w /= g.nw2;
    const int wh = w % g.nw1;
    const int wd = w / g.nw1;
    __syncthreads();   // previous item fully consumed
    if (head != cur_head) {
      if (cur_head >= 0) flush_dtab(cur_head);
      __syncthreads();
      for (int i = tid; i < g.nrel; i += kAtBwdThreads) {
        tab[i] = __ldg(p.table + (int64_t)i * g.heads + head);
        dtab[i] = 0.f;
      }
      cur_head = head;
    }
    for (int i = tid; i < g.N; i += kAtBwdThreads) {
      int t, f;
      window_token(g, b, wd, wh, ww, i, t, f);
      tok[i] = t;
      info[i] = f;
      sLse[i] = __ldg(p.lse + ((int64_t)wg * g.heads + head) * g.N + i);
    }
    __syncthreads();
    // cooperative gather: 8 lanes per row
    {
      const int rslot = tid >> 3, q4 = tid & 7;   // 52 row slots
      for (int r0 = 0; r0 < g.N; r0 += kAtBwdThreads / 8) {   // uniform trip count: the shuffles below need full warps
        const int r = (r0 + rslot < g.N) ? r0 + rslot : g.N - 1;
        const bool own = r0 + rslot < g.N;
        const int64_t t = tok[r];
        const float* base = p.qkv + t * 3 * C + head * 32 + q4 * 4;
        float4 q = ldg4(base);
        q.x *= p.scale; q.y *= p.scale; q.z *= p.scale; q.w *= p.scale;
        const float4 dO = ldg4(p.dout + t * C + head * 32 + q4 * 4);
        const float4 O = ldg4(p.out + t * C + head * 32 + q4 * 4);
        if (own) {
          *reinterpret_cast<float4*>(sQ + r * 32 + q4 * 4) = q;
          *reinterpret_cast<float4*>(sK + r * 32 + q4 * 4) = ldg4(base + C);
          *reinterpret_cast<float4*>(sV + r * 32 + q4 * 4) = ldg4(base + 2 * C);
          *reinterpret_cast<float4*>(sDO + r * 32 + q4 * 4) = dO;
        }
        float d = (dO.x * O.x + dO.y * O.y) + (dO.z * O.z + dO.w * O.w);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if (q4 == 0 && own) sDs[r] = d;
      }
    }
    __syncthreads();
    for (int base = 0; base < g.N; base += kAtBwdThreads) {   // 392 rows = one full sweep + one 8-row sweep
    const bool act = base + tid < g.N;
    const int me = act ? base + tid : 0;
    const int f_me = info[me];
    const int b_me = f_me & 0xffff;
    const int r_me = (f_me >> 16) & 0x1f;
    // ---------------- pass A: thread = query row ----------------
    {
      float2 q[16], d_o[16], dq[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        q[k] = *reinterpret_cast<const float2*>(sQ + me * 32 + k * 2);
        d_o[k] = *reinterpret_cast<const float2*>(sDO + me * 32 + k * 2);
        dq[k] = make_float2(0.f, 0.f);
      }
      const float lse_i = sLse[me], dsum_i = sDs[me];
      const int a_i = b_me + rel0;
      for (int j = 0; j < g.N; ++j) {
        const float2* kj = reinterpret_cast<const float2*>(sK + j * 32);
        const float2* vj = reinterpret_cast<const float2*>(sV + j * 32);
        float2 kr[16];
        float2 s2 = make_float2(0.f, 0.f), p2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          kr[k] = kj[k];
          s2 = ffma2(q[k], kr[k], s2);
          p2 = ffma2(d_o[k], vj[k], p2);
        }
        const int fj = info[j];
        const int idx = a_i - (fj & 0xffff);
        float s = (s2.x + s2.y) + tab[idx];
        s += (((fj >> 16) & 0x1f) != r_me) ? -100.f : 0.f;
        const float pij = expf(s - lse_i);
        const float ds = pij * ((p2.x + p2.y) - dsum_i);
        const float2 ds2 = make_float2(ds, ds);
#pragma unroll
        for (int k = 0; k < 16; ++k) dq[k] = ffma2(ds2, kr[k], dq[k]);
        if (act) atomicAdd(&dtab[idx], ds);
      }
      if (act) {
        float* dst = p.dqkv + (int64_t)tok[me] * 3 * C + head * 32;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          st4(dst + k * 4, make_float4(dq[2 * k].x * p.scale, dq[2 * k].y * p.scale, dq[2 * k + 1].x * p.scale,
                                       dq[2 * k + 1].y * p.scale));
      }
    }
    // ---------------- pass B: thread = key row ----------------
    {
      float2 kk[16], vv[16], dk[16], dv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        kk[k] = *reinterpret_cast<const float2*>(sK + me * 32 + k * 2);
        vv[k] = *reinterpret_cast<const float2*>(sV + me * 32 + k * 2);
        dk[k] = dv[k] = make_float2(0.f, 0.f);
      }
      for (int i = 0; i < g.N; ++i) {
        const float2* qi = reinterpret_cast<const float2*>(sQ + i * 32);
        const float2* doi = reinterpret_cast<const float2*>(sDO + i * 32);
        float2 s2 = make_float2(0.f, 0.f), p2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          s2 = ffma2(qi[k], kk[k], s2);
          p2 = ffma2(doi[k], vv[k], p2);
        }
        const int fi = info[i];
        float s = (s2.x + s2.y) + tab[(fi & 0xffff) + rel0 - b_me];
        s += (((fi >> 16) & 0x1f) != r_me) ? -100.f : 0.f;
        const float pij = expf(s - sLse[i]);
        const float ds = pij * ((p2.x + p2.y) - sDs[i]);
        const float2 ds2 = make_float2(ds, ds), pp2 = make_float2(pij, pij);
#pragma unroll
        for (int k = 0; k < 16; ++k) {   // rows re-read from shared memory (broadcast) instead of held in registers
          dk[k] = ffma2(ds2, qi[k], dk[k]);
          dv[k] = ffma2(pp2, doi[k], dv[k]);
        }
      }
      if (act) {
        float* dst = p.dqkv + ((int64_t)tok[me] * 3 + 1) * C + head * 32;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          st4(dst + k * 4, make_float4(dk[2 * k].x, dk[2 * k].y, dk[2 * k + 1].x, dk[2 * k + 1].y));
          st4(dst + C + k * 4, make_float4(dv[2 * k].x, dv[2 * k].y, dv[2 * k + 1].x, dv[2 * k + 1].y));
        }
      }
    }
    }   // row sweeps
  }
  __syncthreads();
  if (cur_head >= 0) flush_dtab(cur_head);
}


// dsum[token, head] = sum_d dout[token, head*32 + d] * out[token, head*32 + d]; one 8-lane group per (token, head)
__global__ void __launch_bounds__(256) wmsa3d_dsum_kernel(const float* __restrict__ out, const float* __restrict__ dout,
                                                         float* __restrict__ dsum, int64_t n_pairs) {
  const int q4 = threadIdx.x & 7;
  for (int64_t pr0 = (int64_t)blockIdx.x * 32; pr0 < n_pairs; pr0 += (int64_t)gridDim.x * 32) {
    const int64_t pr = pr0 + (threadIdx.x >> 3);
    float d = 0.f;
    if (pr < n_pairs) {
      const float4 a = ldg4(out + pr * 32 + q4 * 4), b = ldg4(dout + pr * 32 + q4 * 4);
      d = (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (q4 == 0 && pr < n_pairs) dsum[pr] = d;
  }
}

// ------------------------------------------------------------------------------------------------
// backward: tcgen05, two launches of one skeleton (template MODE), EVERY product on kind::f16.
//   MODE 0 "query-outer": thread = query row i.  Row operands (resident per 128-row tile in tensor memory):  Qs_t, dO_t.
//           per 32-key chunk c:  S = Qs_t K_c^T,  dP = dO_t V_c^T  (TMEM, four rotating buffers)
//           dS = P o (dP - D_i), P = exp(S + bias + mask - lse_i);  dQ_t += dS K_c ;  dTable[rel(i,j)] += dS
//   MODE 1 "key-outer":   thread = key row j.    Row operands: K_t, V_t.
//           per 32-query chunk c:  S^T = K_t Qs_c^T,  dP^T = V_t dO_c^T
//           P^T, dS^T with the per-column lse_i, D_i;  dV_t += P^T dO_c ;  dK_t += dS^T Qs_c
// D_i = dO_i . O_i comes from wmsa3d_dsum_kernel.  (v1 of this kernel ran the same skeleton on 3xTF32.)
//
// What bounds these kernels is the number of dependent tcgen05.mma instructions and of hand-overs per 32-column chunk, not
// tensor work (128 x 32 x 16 per MMA) or the row threads' arithmetic -- the forward kernel's time stamps show ~130 cycles
// per dependent MMA and several hundred per mbarrier hand-over (profiles/r02_wmsa_fwd_timeline_*.txt).  v1 issued 12 tf32
// MMAs per chunk and stream (4 K steps x 3 split products) and passed dS / P through ONE tensor-memory buffer
// (store -> accumulate MMAs -> "buffer free" -> next store).  Here:
//   * all operands are fp16 hi / lo pairs under per-tensor power-of-two scales (q, k, v: 2^14 / max|qkv|, dO: 2^14 /
//     max|dO|; P * 2^10; dS * (those scales) * 2^-20 -- every factor is undone exactly in the epilogues): K = 16 per MMA,
//     6 MMAs per chunk and stream;
//   * dS (and P in the key-outer launch) are packed IN PLACE over the score columns they were computed from (each row
//     thread owns 16 of a chunk's 32 columns: hi pairs in the first 8, lo pairs in the next 8), so there is no operand
//     buffer to hand back: four score buffers rotate, released by the accumulate issuers' commits;
//   * a column chunk is stored ONCE per tensor (fp16 hi + lo): 128-byte rows of 32 channels with the 16-byte chunk index
//     XORed with row % 8 is at the same time the K-major SWIZZLE_128B operand of the score MMAs (row = N index) and the
//     MN-major SWIZZLE_128B operand of the accumulate MMAs (row = K index); v1 wrote every chunk twice.
// ------------------------------------------------------------------------------------------------
struct WmsaBwd3Params {
  const float* qkv;
  const float* table;
  const float* dout;
  const float* lse;
  const float* dsum;    // (tokens, heads)
  const float* qkv_amax;
  const float* dout_amax;
  float* dqkv;
  float* dtable;
  float scale;
  int items, items_per_cta;
  WmsaGeom g;
  float* amax_out;      // optional: max|dqkv| over both launches (range of the qkv data / weight gradient GEMMs)
  unsigned long long* trace;   // TRACE instantiation only: 16 warps x trace_cap records per launch (0 = unused)
  int trace_cap;
};

constexpr int kB3Stages = 6;                        // column-chunk stages: C1 hi | lo, C2 hi | lo (4 KB each)
constexpr int kB3StageBytes = 16384;
constexpr int kB3Bufs = 4;                          // score buffers in tensor memory
constexpr int kB3OffC = 0;
constexpr int kB3OffDTab = kB3OffC + kB3Stages * kB3StageBytes;   // MODE 0: 8 warp-private table gradients
constexpr int kB3OffTab = kB3OffDTab + 8 * kAtMaxRel * 4;         // bias table of the head (* log2 e)
constexpr int kB3OffLse = kB3OffTab + kAtMaxRel * 4;              // float2 (lse * log2e, dsum * scales) per token
constexpr int kB3OffInfo = kB3OffLse + 2 * kAtColPad * 8;         // (lse / info / tok: two buffers each, by item parity)
constexpr int kB3OffTok = kB3OffInfo + 2 * kAtColPad * 4;
constexpr int kB3OffBar = kB3OffTok + 2 * kAtColPad * 4;
constexpr int kB3SmemBytes = kB3OffBar + 512 + 1024;              // 205312 <= 232448
// TMEM columns: row tiles (packed fp16 pairs, 16 columns per 32 channels), score buffers b at 64 + 64 b (S | dP, 32 + 32),
// accumulators
constexpr uint32_t kT3R1hi = 0, kT3R1lo = 16, kT3R2hi = 32, kT3R2lo = 48, kT3SC = 64, kT3ACC1 = 320, kT3ACC2 = 352;

enum { D_ITEM_READY0 = 0, D_ITEM_READY1, D_ITEM_FREE0, D_ITEM_FREE1, D_ROWS_READY, D_ROWS_FREE, D_ACC_FULL, D_ACC_FREE, D_COL_READY0,
       D_COL_FREE0 = D_COL_READY0 + kB3Stages, D_SC_FULL0 = D_COL_FREE0 + kB3Stages, D_E_READY0 = D_SC_FULL0 + kB3Bufs,
       D_SC_FREE0 = D_E_READY0 + kB3Bufs, D_COUNT = D_SC_FREE0 + kB3Bufs };
static_assert(D_COUNT * 8 + 8 <= 512, "barrier block");

// fp16 hi / lo pair words of (x0, x1): element 0 in the low half (lower K index)
__device__ __forceinline__ void f16_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

constexpr int kB3Threads = 512;   // warps 0-7 row threads (TMEM lane quadrant = w & 3, column half = w >> 2), 8-11 loaders,
                                  // 12-15 MMA issuers (S, dP, ACC1, ACC2)

template <int MODE, bool TRACE>
__global__ void __launch_bounds__(kB3Threads, 1) wmsa3d_bwd3_kernel(const WmsaBwd3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kB3OffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + D_COUNT);
  float* tab = reinterpret_cast<float*>(smem + kB3OffTab);
  float* dtab = reinterpret_cast<float*>(smem + kB3OffDTab);
  float2* sLD_all = reinterpret_cast<float2*>(smem + kB3OffLse);   // [2][kAtColPad]: per-token arrays of the item,
  int* info_all = reinterpret_cast<int*>(smem + kB3OffInfo);       // double-buffered by item parity (the loaders start
  int* tok_all = reinterpret_cast<int*>(smem + kB3OffTok);         // on the next item while the row threads finish this one)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t tr_n = 0;     // WMSA_TR: time stamps of the TRACE instantiation (vitta_wmsa3d_bwd_trace)
  (void)tr_n;
  const WmsaGeom& g = p.g;
  const int C = g.heads * 32;
  const int nwin = g.nw0 * g.nw1 * g.nw2;
  const int nwin_total = g.B * nwin;
  const int n_tiles = (g.N + 127) >> 7;
  const int n_chunks = (g.N + 31) >> 5;
  const int item0 = blockIdx.x * p.items_per_cta;
  const int item1 = min(p.items, item0 + p.items_per_cta);
  constexpr int kAccIssuers = MODE == 1 ? 2 : 1;
  // operand scales (powers of two): q, k, v -> * sq;  dO -> * sd
  float sq, inv_sq, sd, inv_sd;
  f16_split_scale(__ldg(p.qkv_amax), sq, inv_sq);
  f16_split_scale(__ldg(p.dout_amax), sd, inv_sd);

  if (threadIdx.x == 0) {
    for (int i = 0; i < D_COUNT; ++i) {
      int cnt = kAccIssuers;   // COL_FREE, SC_FREE, ACC_FULL: one tcgen05.commit per accumulate issuer
      if ((i >= D_SC_FULL0 && i < D_E_READY0) || i == D_ROWS_FREE) cnt = 2;   // one commit per score issuer
      if (i == D_ITEM_READY0 || i == D_ITEM_READY1 || i == D_ROWS_READY || (i >= D_COL_READY0 && i < D_COL_FREE0))
        cnt = 4;     // one elected arrive per loader warp
      if (i == D_ITEM_FREE0 || i == D_ITEM_FREE0 + 1 || (i >= D_E_READY0 && i < D_SC_FREE0) || i == D_ACC_FREE)
        cnt = 8;     // one elected arrive per row warp
      mbar_init(&bar[i], cnt);
    }
    fence_barrier_init();
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // (uniform value made provably uniform)

  if (warp >= 8 && warp < 12) {
    // =========================== loaders ===========================
    const int lt = threadIdx.x - 256;
    const int rslot = lt >> 3, q4 = lt & 7;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp - 8) * 32) << 16);   // this warp's TMEM lane quadrant
    int cur_head = -1;
    uint32_t it = 0, tile_ctr = 0, chunk_ctr = 0;
    // scale of the operand (which 0: R1 / C1, 1: R2 / C2) -- MODE 0 rows: Qs, dO; columns: K, V.  MODE 1 rows: K, V;
    // columns: Qs, dO
    const float s_r1 = (MODE == 0) ? p.scale * sq : sq, s_r2 = (MODE == 0) ? sd : sq;
    const float s_c1 = (MODE == 0) ? sq : p.scale * sq, s_c2 = (MODE == 0) ? sq : sd;
    auto flush_dtab = [&](int head) {
      for (int i = lt; i < g.nrel; i += 128) {
        float v = 0.f;
#pragma unroll
        for (int cpy = 0; cpy < 8; ++cpy) v += dtab[cpy * kAtMaxRel + i];
        if (v != 0.f) atomicAdd(p.dtable + (int64_t)i * g.heads + head, v);
      }
    };
    for (int item = item0; item < item1; ++item, ++it) {
      const int head = item / nwin_total;
      const int wg = item - head * nwin_total;
      const int b = wg / nwin;
      int w = wg - b * nwin;
      const int ww = w % g.nw2; w /= g.nw2;
      const int wh = w % g.nw1;
      const int wd = w / g.nw1;
      float2* sLD = sLD_all + (it & 1) * kAtColPad;
      int* info = info_all + (it & 1) * kAtColPad;
      int* tok = tok_all + (it & 1) * kAtColPad;
      mbar_wait(&bar[D_ITEM_FREE0 + (it & 1)], ((it >> 1) & 1) ^ 1);      // this buffer's previous item (it - 2) is done
      WMSA_TR(40);
      if (head != cur_head) {
        // the bias table and the table gradients are single: a change of head waits for the previous item as well
        if (it > 0) mbar_wait(&bar[D_ITEM_FREE0 + ((it - 1) & 1)], ((it - 1) >> 1) & 1);
        if (MODE == 0 && cur_head >= 0) flush_dtab(cur_head);
        for (int i = lt; i < g.nrel; i += 128) {
          tab[i] = __ldg(p.table + (int64_t)i * g.heads + head) * 1.4426950408889634f;   // bias * log2(e)
          if (MODE == 0) {
#pragma unroll
            for (int cpy = 0; cpy < 8; ++cpy) dtab[cpy * kAtMaxRel + i] = 0.f;
          }
        }
        cur_head = head;
      }
      for (int i = lt; i < kAtColPad; i += 128) {
        int t = -1, f = 31 << 16;                 // padding columns: region id 31 = always masked
        float2 q = make_float2(INFINITY, 0.f);    // ... and lse = +inf: p = 0 exactly when they are queries
        if (i < g.N) {
          window_token(g, b, wd, wh, ww, i, t, f);
          q.x = __ldg(p.lse + ((int64_t)wg * g.heads + head) * g.N + i) * 1.4426950408889634f;
          q.y = (__ldg(p.dsum + (int64_t)t * g.heads + head) * sd) * sq;      // D_i in the units of the dP accumulator
        }
        tok[i] = t;
        info[i] = f;
        sLD[i] = q;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[D_ITEM_READY0 + (it & 1)]);   // (one barrier per buffer: the loaders run an item ahead)
      const float* qkv_h = p.qkv + head * 32;
      const float* do_h = p.dout + head * 32;
      // row operand (which 0: R1, 1: R2) / column operand of token t; q = float4 index inside the 32-float head slice
      auto load_row = [&](int which, int t, int q) -> float4 {
        if (t < 0) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 0) return which == 0 ? ldg4(qkv_h + (int64_t)t * 3 * C + q * 4) : ldg4(do_h + (int64_t)t * C + q * 4);
        return ldg4(qkv_h + ((int64_t)t * 3 + 1 + which) * C + q * 4);
      };
      auto load_col = [&](int which, int t) -> float4 {
        if (t < 0) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 0) return ldg4(qkv_h + ((int64_t)t * 3 + 1 + which) * C + q4 * 4);
        return which == 0 ? ldg4(qkv_h + (int64_t)t * 3 * C + q4 * 4) : ldg4(do_h + (int64_t)t * C + q4 * 4);
      };
      // column chunks: 2 chunks per group (2 tensors x 2 rows x 2 chunks = 8 float4 per thread)
      auto c_issue = [&](float4 (&v)[8], int grp) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int cc = u >> 2, which = (u >> 1) & 1, rr = u & 1;
          const int j = (grp * 2 + cc) * 32 + rslot + rr * 16;
          const int t = (grp * 2 + cc < n_chunks) ? tok[j] : -1;
          v[u] = load_col(which, t);
        }
      };
      auto c_drain = [&](float4 (&v)[8], int grp) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (grp * 2 + cc >= n_chunks) break;
          const int st = chunk_ctr % kB3Stages;
          mbar_wait(&bar[D_COL_FREE0 + st], ((chunk_ctr / kB3Stages) & 1) ^ 1);
          uint8_t* cb = smem + kB3OffC + st * kB3StageBytes;
#pragma unroll
          for (int which = 0; which < 2; ++which) {
            const float sc = which ? s_c2 : s_c1;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int r = rslot + rr * 16;
              const float4 x = v[cc * 4 + which * 2 + rr];
              uint32_t h0, l0, h1, l1;
              f16_pair(x.x * sc, x.y * sc, h0, l0);
              f16_pair(x.z * sc, x.w * sc, h1, l1);
              const uint32_t o = (uint32_t)r * 128u + ((((uint32_t)q4 >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)q4 & 1u) * 8u;
              uint8_t* base = cb + which * 8192;
              *reinterpret_cast<uint2*>(base + o) = make_uint2(h0, h1);
              *reinterpret_cast<uint2*>(base + 4096 + o) = make_uint2(l0, l1);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar[D_COL_READY0 + st]);
          ++chunk_ctr;
        }
      };
      const int n_groups = (n_chunks + 1) >> 1;
      for (int tile = 0; tile < n_tiles; ++tile, ++tile_ctr) {
        // row tile: thread = row (TMEM lane); both rows' global loads are in flight before the wait
        const int i = tile * 128 + (warp - 8) * 32 + lane;
        const int trow = (i < g.N) ? tok[i] : -1;
        float4 ra[8], rb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          ra[u] = load_row(0, trow, u);
          rb[u] = load_row(1, trow, u);
        }
        mbar_wait(&bar[D_ROWS_FREE], (tile_ctr & 1) ^ 1);
        tc_fence_after();
        {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            f16_pair(ra[u].x * s_r1, ra[u].y * s_r1, hi[2 * u], lo[2 * u]);
            f16_pair(ra[u].z * s_r1, ra[u].w * s_r1, hi[2 * u + 1], lo[2 * u + 1]);
          }
          tmem_st16(t_lane + kT3R1hi, hi);
          tmem_st16(t_lane + kT3R1lo, lo);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            f16_pair(rb[u].x * s_r2, rb[u].y * s_r2, hi[2 * u], lo[2 * u]);
            f16_pair(rb[u].z * s_r2, rb[u].w * s_r2, hi[2 * u + 1], lo[2 * u + 1]);
          }
          tmem_st16(t_lane + kT3R2hi, hi);
          tmem_st16(t_lane + kT3R2lo, lo);
        }
        float4 va[8], vb8[8];
        c_issue(va, 0);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[D_ROWS_READY]);
        WMSA_TR(42);
        for (int grp = 0; grp < n_groups; grp += 2) {
          if (grp + 1 < n_groups) c_issue(vb8, grp + 1);
          c_drain(va, grp);
          if (grp + 2 < n_groups) c_issue(va, grp + 2);
          if (grp + 1 < n_groups) c_drain(vb8, grp + 1);
        }
      }
    }
    if (MODE == 0 && cur_head >= 0) {
      mbar_wait(&bar[D_ITEM_FREE0 + ((it - 1) & 1)], ((it - 1) >> 1) & 1);   // the row threads finished the last item
      flush_dtab(cur_head);
    }
  } else if (warp >= 12) {
    // =========================== MMA issuers ===========================
    // warp 12: S / S^T, warp 13: dP / dP^T, warp 14: ACC1 (dQ | dK), warp 15: ACC2 (dV, MODE 1 only).  They touch disjoint
    // accumulators; ordering against the other roles goes through the mbarriers.
    const int which = __shfl_sync(0xffffffffu, warp & 1, 0);   // warp-uniform by construction; the shuffle makes it provable
    const bool is_score = warp < 14;
    const uint32_t pe = (lane == 0) ? 1u : 0u;
    const uint32_t sbase = smem_u32(smem);
    uint32_t tile_ctr = 0, chunk_ctr = 0;
    if (is_score) {
      constexpr uint32_t idesc_sc = umma_idesc_f16(128, 32);      // A in TMEM (packed pairs along the channels), B K-major
      const uint32_t a_hi = tmem_base + (which ? kT3R2hi : kT3R1hi), a_lo = tmem_base + (which ? kT3R2lo : kT3R1lo);
      for (int item = item0; item < item1; ++item) {
        for (int tile = 0; tile < n_tiles; ++tile, ++tile_ctr) {
          mbar_wait(&bar[D_ROWS_READY], tile_ctr & 1);
          for (int c = 0; c < n_chunks; ++c, ++chunk_ctr) {
            const uint32_t ccu = __shfl_sync(0xffffffffu, chunk_ctr, 0);   // uniform registers for descriptors / addresses
            const int st = ccu % kB3Stages;
            const int sb = ccu & (kB3Bufs - 1);
            mbar_wait(&bar[D_COL_READY0 + st], (chunk_ctr / kB3Stages) & 1);
            WMSA_TR(20);
            mbar_wait(&bar[D_SC_FREE0 + sb], ((chunk_ctr / kB3Bufs) & 1) ^ 1);
            tc_fence_after();
            WMSA_TR(21);
            const uint32_t cb = sbase + kB3OffC + st * kB3StageBytes + which * 8192;
            const uint64_t c_hi = umma_desc_sw128(cb), c_lo = umma_desc_sw128(cb + 4096);
            const uint32_t d = tmem_base + kT3SC + (uint32_t)(sb * 64 + which * 32);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);      // 16 channels = 32 bytes along the row
              const uint32_t ka = (uint32_t)(k * 8);
              umma_f16_ts_p(d, a_lo + ka, c_hi + adv, idesc_sc, k != 0, pe);
              umma_f16_ts_p(d, a_hi + ka, c_lo + adv, idesc_sc, 1, pe);
              umma_f16_ts_p(d, a_hi + ka, c_hi + adv, idesc_sc, 1, pe);
            }
            umma_commit_p(&bar[D_SC_FULL0 + sb], pe);
            if (c == n_chunks - 1) umma_commit_p(&bar[D_ROWS_FREE], pe);   // the row tile is reusable when these retire
            WMSA_TR(22);
          }
        }
      }
    } else if (which == 0 || MODE == 1) {
      constexpr uint32_t idesc_ac = umma_idesc_f16(128, 32) | (1u << 16);   // B MN-major
      const uint32_t acc = tmem_base + (which ? kT3ACC2 : kT3ACC1);
      for (int item = item0; item < item1; ++item) {
        for (int tile = 0; tile < n_tiles; ++tile, ++tile_ctr) {
          for (int c = 0; c < n_chunks; ++c, ++chunk_ctr) {
            const uint32_t ccu = __shfl_sync(0xffffffffu, chunk_ctr, 0);
            const int st = ccu % kB3Stages;
            const int sb = ccu & (kB3Bufs - 1);
            mbar_wait(&bar[D_E_READY0 + sb], (chunk_ctr / kB3Bufs) & 1);
            if (c == 0) mbar_wait(&bar[D_ACC_FREE], (tile_ctr & 1) ^ 1);
            tc_fence_after();
            WMSA_TR(30);
            const uint32_t cb = sbase + kB3OffC + st * kB3StageBytes + which * 8192;
            const uint64_t m_hi = umma_desc_mn_sw128_f16(cb, 4096), m_lo = umma_desc_mn_sw128_f16(cb + 4096, 4096);
            // E operand of this stream: dS over the S columns (which 0), P over the dP columns (which 1); per 16-key K step
            // the hi pairs sit in the first 8 of the step's 16 columns, the lo pairs in the next 8
            const uint32_t e = tmem_base + kT3SC + (uint32_t)(sb * 64 + which * 32);
            const int left = g.N - c * 32;
            const int ksteps = left >= 32 ? 2 : (left + 15) >> 4;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (k < ksteps) {
                const uint64_t advb = (uint64_t)(k * (2048 >> 4));   // 16 keys = two 8-row atoms of 1024 B
                const uint32_t ka = (uint32_t)(k * 16);
                umma_f16_ts_p(acc, e + ka + 8u, m_hi + advb, idesc_ac, (c | k) != 0, pe);
                umma_f16_ts_p(acc, e + ka, m_lo + advb, idesc_ac, 1, pe);
                umma_f16_ts_p(acc, e + ka, m_hi + advb, idesc_ac, 1, pe);
              }
            }
            umma_commit_p(&bar[D_SC_FREE0 + sb], pe);
            umma_commit_p(&bar[D_COL_FREE0 + st], pe);
            if (c == n_chunks - 1) umma_commit_p(&bar[D_ACC_FULL], pe);
            WMSA_TR(31);
          }
        }
      }
    }
  } else {
    // =========================== row threads ===========================
    // thread = (row of the tile = TMEM lane, column half): warps w and w + 4 share a lane quadrant and split the 32 columns
    // of every chunk, which doubles the warps available to hide the latency of the per-element chain.
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int rel0 = rel_row_base(g);
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kMask2 = -100.f * kLog2e;
    const float* __restrict__ tab2 = tab;                        // bias table, pre-multiplied by log2(e) by the loaders
    float* __restrict__ mytab = dtab + warp * kAtMaxRel;         // this warp's private table gradient (no atomics)
    const float c_s = (inv_sq * inv_sq) * kLog2e;                // score accumulator -> log2 domain
    constexpr float kDsDown = 9.5367431640625e-07f;              // 2^-20: dS in the units of its fp16 operand
    constexpr float kDsUp = 1048576.f;
    uint32_t it = 0, tile_ctr = 0, chunk_ctr = 0;
    float out_amax = 0.f;
    for (int item = item0; item < item1; ++item, ++it) {
      const int head = item / nwin_total;
      mbar_wait(&bar[D_ITEM_READY0 + (it & 1)], (it >> 1) & 1);
      const float2* sLD = sLD_all + (it & 1) * kAtColPad;
      const int* info = info_all + (it & 1) * kAtColPad;
      const int* tok = tok_all + (it & 1) * kAtColPad;
      for (int tile = 0; tile < n_tiles; ++tile, ++tile_ctr) {
        const int i = tile * 128 + row;
        const bool valid = i < g.N;
        const int ii = valid ? i : 0;
        const int f_row = info[ii];
        const int b_row = f_row & 0xffff;
        const int r_row = f_row & 0x1f0000;
        const int my_tok = tok[ii];
        const float2 ld_row = sLD[ii];
        const int kidx = (MODE == 0) ? (b_row + rel0) : (rel0 - b_row);
        for (int c = 0; c < n_chunks; ++c, ++chunk_ctr) {
          const int sb = chunk_ctr & (kB3Bufs - 1);
          mbar_wait(&bar[D_SC_FULL0 + sb], (chunk_ctr / kB3Bufs) & 1);
          tc_fence_after();
          WMSA_TR(1);
          const uint32_t sc = t_lane + kT3SC + (uint32_t)(sb * 64 + half * 16);
          uint32_t s[16], d[16];
          tmem_ld16(sc, s);
          tmem_ld16(sc + 32u, d);
          tmem_ld_wait();
          // per element: p = 2^(s*log2e + bias2 + mask2 - lse2), ds = p * (dp - dsum).  The per-token arrays are padded to
          // a multiple of 32 columns: padding columns carry region id 31 (always masked: p flushes to 0) and, as
          // queries (MODE 1), lse = +inf (p = 0 exactly), so no bounds selects are needed here.
          const int col0 = c * 32 + half * 16;
          const int* ic = info + col0;
          float pv[16], dsv[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int fc = ic[jj];
            const int idx = (MODE == 0) ? (kidx - (fc & 0xffff)) : (kidx + (fc & 0xffff));
            float t = fmaf(__uint_as_float(s[jj]), c_s, tab2[idx]);
            t += ((fc & 0x1f0000) != r_row) ? kMask2 : 0.f;
            float lse2, dsum;
            if (MODE == 0) {
              lse2 = ld_row.x;
              dsum = ld_row.y;
            } else {
              const float2 q = sLD[col0 + jj];
              lse2 = q.x;
              dsum = q.y;
            }
            float pij;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pij) : "f"(t - lse2));
            pv[jj] = pij;
            dsv[jj] = pij * (__uint_as_float(d[jj]) - dsum) * kDsDown;     // dS * sd * sq * 2^-20
          }
          if (MODE == 0) {
            // dTable[rel(i, j)] += dS_ij.  Lanes of a warp are distinct rows => distinct entries for one column, so a
            // predicated straight-line LDS / FFMA / STS on the warp-private copy is race free; program order keeps the
            // successive columns coherent -- (row, column) pairs of DIFFERENT columns do share entries (same relative
            // position), which is why the 16 updates cannot be batched (all loads first loses updates: tried) and why this
            // chain costs ~1000 cycles per chunk.  Shared-memory atomicAdd (a CAS loop for fp32) was slower still, and so
            // was a fixed-point variant on the native 32-bit integer atomics (two digits per element, drained into a 64-bit
            // sum by the loaders once per item: exact and order independent, but 32 ATOMS per thread and chunk cost more
            // than the chain -- 1.77 ms against 1.60 ms for the stage-1 geometry of tools/one_wmsa.py).
            const float k_dt = (kDsUp * inv_sd) * inv_sq;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int idx = kidx - (ic[jj] & 0xffff);
              const bool ok = valid && (col0 + jj < g.N);
              const float o = mytab[ok ? idx : 0];
              if (ok) mytab[idx] = fmaf(dsv[jj], k_dt, o);
            }
          }
          WMSA_TR(2);
          // fp16 hi / lo pairs, straight back into tensor memory over this thread's 16 score columns: the A operand of the
          // accumulate MMAs (key 2u in the low half of column u)
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) f16_pair(dsv[2 * u], dsv[2 * u + 1], hi[u], lo[u]);
          tmem_st8(sc, hi);
          tmem_st8(sc + 8u, lo);
          if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) f16_pair(pv[2 * u] * 1024.f, pv[2 * u + 1] * 1024.f, hi[u], lo[u]);
            tmem_st8(sc + 32u, hi);
            tmem_st8(sc + 40u, lo);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar[D_E_READY0 + sb]);
          WMSA_TR(3);
        }
        // ---- accumulators -> global (this thread's 16 of the 32 head channels)
        mbar_wait(&bar[D_ACC_FULL], tile_ctr & 1);
        tc_fence_after();
        WMSA_TR(6);
        uint32_t a1[16], a2[16];
        tmem_ld16(t_lane + kT3ACC1 + (uint32_t)(half * 16), a1);
        if (MODE == 1) tmem_ld16(t_lane + kT3ACC2 + (uint32_t)(half * 16), a2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[D_ACC_FREE]);
        if (valid) {
          // undo the operand scales (exact: powers of two, applied one after the other)
          const float k1 = ((MODE == 0 ? p.scale : 1.f) * kDsUp * inv_sd) * inv_sq;      // (* inv_sq once more below)
          const float k2 = 0.0009765625f * inv_sd;                                       // dV: P * 2^10, dO * sd
          float* dst = p.dqkv + ((int64_t)my_tok * 3 + (MODE == 0 ? 0 : 1)) * C + head * 32 + half * 16;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o1;
            o1.x = (__uint_as_float(a1[q * 4]) * k1) * inv_sq; o1.y = (__uint_as_float(a1[q * 4 + 1]) * k1) * inv_sq;
            o1.z = (__uint_as_float(a1[q * 4 + 2]) * k1) * inv_sq; o1.w = (__uint_as_float(a1[q * 4 + 3]) * k1) * inv_sq;
            st4(dst + q * 4, o1);
            out_amax = fmaxf(out_amax, fmaxf(fmaxf(fabsf(o1.x), fabsf(o1.y)), fmaxf(fabsf(o1.z), fabsf(o1.w))));
            if (MODE == 1) {
              float4 o2;
              o2.x = __uint_as_float(a2[q * 4]) * k2; o2.y = __uint_as_float(a2[q * 4 + 1]) * k2;
              o2.z = __uint_as_float(a2[q * 4 + 2]) * k2; o2.w = __uint_as_float(a2[q * 4 + 3]) * k2;
              st4(dst + C + q * 4, o2);
              out_amax = fmaxf(out_amax, fmaxf(fmaxf(fabsf(o2.x), fabsf(o2.y)), fmaxf(fabsf(o2.z), fabsf(o2.w))));
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[D_ITEM_FREE0 + (it & 1)]);   // dtab / info / tok / lse of this item no longer needed
    }
    if (p.amax_out) {
      const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(out_amax));
      if (lane == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(p.amax_out), wmax);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

static unsigned long long* g_bwd_trace = nullptr;   // see vitta_wmsa3d_bwd_set_trace
static int g_bwd_trace_cap = 0;

static int wmsa_geom(int B, int D, int H, int W, int heads, const int* window, const int* shift, WmsaGeom* g) {
  VITTA_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0 && heads > 0 && window && shift, VITTA_E_BADARG, "wmsa3d: bad shape");
  const int dims[3] = {D, H, W};
  int ws[3], ss[3];
  for (int i = 0; i < 3; ++i) {
    VITTA_CHECK_ARG(window[i] > 0 && shift[i] >= 0 && shift[i] < window[i], VITTA_E_BADARG, "wmsa3d: bad window / shift");
    ws[i] = window[i];
    ss[i] = shift[i];
    if (dims[i] <= window[i]) {   // get_window_size (swin_transformer.py:71-84)
      ws[i] = dims[i];
      ss[i] = 0;
    }
    VITTA_CHECK_ARG(dims[i] % ws[i] == 0, VITTA_E_UNSUPPORTED,
                    "wmsa3d: the token volume must be a multiple of the window -- pad it first, as the reference does "
                    "(swin_transformer.py:222-227; vitta_b200.ops_swin.SwinAttentionFn does so on the host side)");
  }
  g->B = B; g->D = D; g->H = H; g->W = W; g->heads = heads;
  g->ws0 = ws[0]; g->ws1 = ws[1]; g->ws2 = ws[2];
  g->ss0 = ss[0]; g->ss1 = ss[1]; g->ss2 = ss[2];
  g->fw0 = window[0]; g->fw1 = window[1]; g->fw2 = window[2];
  g->nw0 = D / ws[0]; g->nw1 = H / ws[1]; g->nw2 = W / ws[2];
  g->N = ws[0] * ws[1] * ws[2];
  g->NP = (g->N + 15) / 16 * 16;
  g->nrel = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1);
  VITTA_CHECK_ARG(g->N <= 392 && g->NP <= kAtMaxKeys && g->nrel <= kAtMaxRel, VITTA_E_UNSUPPORTED,
                  "wmsa3d: windows of at most 392 tokens / 2560 relative positions");
  VITTA_CHECK_ARG((int64_t)B * D * H * W < (1ll << 31) / 4, VITTA_E_UNSUPPORTED, "wmsa3d: too many tokens");
  return 0;
}

}  // namespace vitta

using namespace vitta;

extern "C" {


static int wmsa3d_fwd_impl(const float* qkv, const float* qkv_amax, const float* bias_table, float* out, float* lse, int B,
                           int D, int H, int W, int heads, int head_dim, const int* window, const int* shift, float scale,
                           float* out_amax, unsigned long long* trace, int trace_cap, void* stream) {
  VITTA_CHECK_ARG(qkv && qkv_amax && bias_table && out && lse, VITTA_E_BADARG, "wmsa3d_fwd: null pointer");
  VITTA_CHECK_ARG(head_dim == 32, VITTA_E_UNSUPPORTED, "wmsa3d: head_dim must be 32 (every Video-Swin configuration)");
  VITTA_CHECK_ARG(aligned16(qkv) && aligned16(out), VITTA_E_ALIGN, "wmsa3d_fwd: tensors must be 16-byte aligned");
  WmsaFwdParams p;
  int rc = wmsa_geom(B, D, H, W, heads, window, shift, &p.g);
  if (rc) return rc;
  p.qkv = qkv; p.qkv_amax = qkv_amax; p.table = bias_table; p.out = out; p.lse = lse; p.scale = scale; p.amax_out = out_amax;
  p.trace = trace; p.trace_cap = trace_cap;
  p.items = B * p.g.nw0 * p.g.nw1 * p.g.nw2 * heads;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wmsa3d_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wmsa3d_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes);
    if (e != cudaSuccess) {
      set_error("wmsa3d_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  const int sms = cached_sm_count();
  int grid = p.items < sms ? p.items : sms;
  p.items_per_cta = (p.items + grid - 1) / grid;
  grid = (p.items + p.items_per_cta - 1) / p.items_per_cta;
  if (trace)
    wmsa3d_fwd_kernel<true><<<grid, kFwdThreads, kAtSmemBytes, (cudaStream_t)stream>>>(p);
  else
    wmsa3d_fwd_kernel<false><<<grid, kFwdThreads, kAtSmemBytes, (cudaStream_t)stream>>>(p);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_wmsa3d_fwd(const float* qkv, const float* qkv_amax, const float* bias_table, float* out, float* lse, int B, int D,
                     int H, int W, int heads, int head_dim, const int* window, const int* shift, float scale, void* stream) {
  return wmsa3d_fwd_impl(qkv, qkv_amax, bias_table, out, lse, B, D, H, W, heads, head_dim, window, shift, scale, nullptr,
                         nullptr, 0, stream);
}

int vitta_wmsa3d_fwd_amax(const float* qkv, const float* qkv_amax, const float* bias_table, float* out, float* lse, int B,
                          int D, int H, int W, int heads, int head_dim, const int* window, const int* shift, float scale,
                          float* out_amax, void* stream) {
  return wmsa3d_fwd_impl(qkv, qkv_amax, bias_table, out, lse, B, D, H, W, heads, head_dim, window, shift, scale, out_amax,
                         nullptr, 0, stream);
}

int vitta_wmsa3d_fwd_trace(const float* qkv, const float* qkv_amax, const float* bias_table, float* out, float* lse, int B,
                           int D, int H, int W, int heads, int head_dim, const int* window, const int* shift, float scale,
                           unsigned long long* trace, int trace_cap, void* stream) {
  VITTA_CHECK_ARG(trace && trace_cap > 1, VITTA_E_BADARG, "wmsa3d_fwd_trace: trace buffer required");
  return wmsa3d_fwd_impl(qkv, qkv_amax, bias_table, out, lse, B, D, H, W, heads, head_dim, window, shift, scale, nullptr,
                         trace, trace_cap, stream);
}

static int wmsa3d_bwd_v0(const WmsaGeom& g, const float* qkv, const float* bias_table, const float* out,
                         const float* dout, const float* lse, float* dqkv, float* dbias_table, float scale,
                         cudaStream_t st) {
  WmsaBwdParams p;
  p.g = g;
  p.qkv = qkv; p.table = bias_table; p.out = out; p.dout = dout; p.lse = lse; p.dqkv = dqkv; p.dtable = dbias_table;
  p.scale = scale;
  p.items = g.B * g.nw0 * g.nw1 * g.nw2 * g.heads;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wmsa3d_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwSmemBytes);
    if (e != cudaSuccess) {
      set_error("wmsa3d_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  const int sms = cached_sm_count();
  int grid = p.items < sms ? p.items : sms;
  p.items_per_cta = (p.items + grid - 1) / grid;
  grid = (p.items + p.items_per_cta - 1) / p.items_per_cta;
  wmsa3d_bwd_kernel<<<grid, kAtBwdThreads, kBwSmemBytes, st>>>(p);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int64_t vitta_wmsa3d_bwd_ws_floats(int B, int D, int H, int W, int heads) {
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || heads <= 0) return -1;
  return (int64_t)B * D * H * W * heads;
}

int vitta_wmsa3d_bwd(const float* qkv, const float* qkv_amax, const float* bias_table, const float* out, const float* dout,
                     const float* dout_amax, const float* lse, float* dqkv, float* dbias_table, float* ws, int B, int D, int H,
                     int W, int heads, int head_dim, const int* window, const int* shift, float scale, int impl,
                     void* stream) {
  return vitta_wmsa3d_bwd_amax(qkv, qkv_amax, bias_table, out, dout, dout_amax, lse, dqkv, dbias_table, ws, B, D, H, W, heads,
                               head_dim, window, shift, scale, impl, nullptr, stream);
}

int vitta_wmsa3d_bwd_amax(const float* qkv, const float* qkv_amax, const float* bias_table, const float* out,
                          const float* dout, const float* dout_amax, const float* lse, float* dqkv, float* dbias_table,
                          float* ws, int B, int D, int H, int W, int heads, int head_dim, const int* window, const int* shift,
                          float scale, int impl, float* dqkv_amax, void* stream) {
  VITTA_CHECK_ARG(!(impl == 1 && dqkv_amax), VITTA_E_UNSUPPORTED, "wmsa3d_bwd: the fp32 cross-check kernel emits no range");
  VITTA_CHECK_ARG(qkv && bias_table && out && dout && lse && dqkv && dbias_table, VITTA_E_BADARG, "wmsa3d_bwd: null pointer");
  VITTA_CHECK_ARG(head_dim == 32, VITTA_E_UNSUPPORTED, "wmsa3d: head_dim must be 32 (every Video-Swin configuration)");
  VITTA_CHECK_ARG(aligned16(qkv) && aligned16(out) && aligned16(dout) && aligned16(dqkv), VITTA_E_ALIGN,
                  "wmsa3d_bwd: tensors must be 16-byte aligned");
  WmsaGeom g;
  int rc = wmsa_geom(B, D, H, W, heads, window, shift, &g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 1) return wmsa3d_bwd_v0(g, qkv, bias_table, out, dout, lse, dqkv, dbias_table, scale, st);
  VITTA_CHECK_ARG(ws, VITTA_E_BADARG, "wmsa3d_bwd: workspace of vitta_wmsa3d_bwd_ws_floats() floats required");
  VITTA_CHECK_ARG(qkv_amax && dout_amax, VITTA_E_BADARG, "wmsa3d_bwd: the operand ranges (qkv_amax, dout_amax) are required");
  const int64_t n_pairs = (int64_t)B * D * H * W * heads;
  {
    int64_t blocks = (n_pairs + 31) / 32;
    if (blocks > 148 * 16) blocks = 148 * 16;
    wmsa3d_dsum_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, dout, ws, n_pairs);
    VITTA_CHECK_LAUNCH();
  }
  WmsaBwd3Params p;
  p.g = g;
  p.qkv = qkv; p.table = bias_table; p.dout = dout; p.lse = lse; p.dsum = ws; p.dqkv = dqkv; p.dtable = dbias_table;
  p.qkv_amax = qkv_amax; p.dout_amax = dout_amax;
  p.scale = scale; p.amax_out = dqkv_amax;
  p.trace = g_bwd_trace; p.trace_cap = g_bwd_trace_cap;
  p.items = B * g.nw0 * g.nw1 * g.nw2 * heads;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wmsa3d_bwd3_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB3SmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wmsa3d_bwd3_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB3SmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wmsa3d_bwd3_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB3SmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wmsa3d_bwd3_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB3SmemBytes);
    if (e != cudaSuccess) {
      set_error("wmsa3d_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  const int sms = cached_sm_count();
  int grid = p.items < sms ? p.items : sms;
  p.items_per_cta = (p.items + grid - 1) / grid;
  grid = (p.items + p.items_per_cta - 1) / p.items_per_cta;
  if (p.trace) {   // profiling aid (vitta_wmsa3d_bwd_set_trace): launch 0 writes the first 16 x cap records, launch 1 the next
    wmsa3d_bwd3_kernel<0, true><<<grid, kB3Threads, kB3SmemBytes, st>>>(p);
    VITTA_CHECK_LAUNCH();
    p.trace += (size_t)16 * p.trace_cap;
    wmsa3d_bwd3_kernel<1, true><<<grid, kB3Threads, kB3SmemBytes, st>>>(p);
    VITTA_CHECK_LAUNCH();
    return 0;
  }
  wmsa3d_bwd3_kernel<0, false><<<grid, kB3Threads, kB3SmemBytes, st>>>(p);
  VITTA_CHECK_LAUNCH();
  wmsa3d_bwd3_kernel<1, false><<<grid, kB3Threads, kB3SmemBytes, st>>>(p);
  VITTA_CHECK_LAUNCH();
  return 0;
}

// Profiling aid: the next vitta_wmsa3d_bwd calls run the traced instantiation (time stamps of CTA 0, as in
// vitta_wmsa3d_fwd_trace) into trace = 2 launches x 16 warps x trace_cap zeroed records; null switches it off again.
int vitta_wmsa3d_bwd_set_trace(unsigned long long* trace, int trace_cap) {
  g_bwd_trace = trace;
  g_bwd_trace_cap = trace ? trace_cap : 0;
  return 0;
}

}  // extern "C"
