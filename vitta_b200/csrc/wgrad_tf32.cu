// K6 (weight gradient): dW[co][tap][ci] = sum over output pixels p of dY[p, co] * X[p shifted by tap, ci]
// as a split-K GEMM on the tcgen05 tensor cores with the 3xTF32 split (see gemm_tf32.cu for the numerics).
//
// The reduction dimension is the pixel index, so BOTH operands are "MN-major" in UMMA terms: a smem row is one pixel
// (K index) holding 32 consecutive channels (128 B), exactly what a TMA box {32 channels, BW, BH, BF pixels} of the
// channels-last activation delivers (swizzle mode 128B_ATOM_32B, the one layout tf32 MN-major operands accept).  A stage = 32 pixels; the A tile (128 output channels) is 4 such boxes, the B
// tile (BN input channels) BN/32 boxes, the filter tap only shifts the X box (OOB zero fill = padding).  Both
// operands are activations, so both are split in-kernel (hi in place, lo beside it) by four warps.
// Work item = (Cout tile, tap, Cin tile, K split); every item writes its fp32 partial [128 x BN] to a workspace and
// wgrad_reduce_kernel sums the K splits in a fixed order (deterministic) straight into the NCHW weight gradient.
//
// F16 = true (vitta_conv2d_wgrad_f16x3, tensor-memory form only): the same pipeline on kind::f16 with the fp16 operand split
// of gemm_tf32.cu (x*s = hi + lo, s a per-tensor power of two from a device amax scalar).  dY^T goes to tensor memory as
// PACKED fp16 pairs (column j of a stage = pixels 2j, 2j+1: 16 hi + 16 lo columns per 32-pixel stage); the X boxes land
// with the plain 128B swizzle and are converted IN PLACE into 16-bit MN-major SWIZZLE_128B atoms of 64 channels x 32
// pixels: box pair (2g, 2g+1) -> hi atom in box area g, lo atom in box area BN/64 + g (hi atoms adjacent, then lo atoms,
// so [x_hi ; x_lo] is again one descriptor).  K = 16 pixels per MMA: half the MMAs and half the tensor-pipe work.
#include <cuda_fp16.h>

#include "tc05.cuh"

namespace vitta {

constexpr int kWgThreads = 448;   // warp 0 TMA, 1 MMA, 2-9 operand split (8 warps: both operands are split in-kernel), 10-13 epilogue
constexpr int kWgSplitThreads = 256;
constexpr int kWgRows = 32;                       // pixels per stage
constexpr int kWgBoxBytes = kWgRows * 128;        // one {32 ch x 32 px} box = 4 KB
constexpr int kWgBM = 128;

struct WgradParams {
  float* ws;               // [splits][Cout][taps*Cin]
  int Cout, Cin;
  int taps_h, taps_w, stride, pad;
  int BW, BH, BF;          // pixel box (product = 32)
  int boxes_w, boxes_h, boxes_f;
  int m_tiles, n_ctiles;   // Cout tiles, Cin tiles (per tap)
  int splits;
  int tap_group;           // filter taps covered by one item (1, or 3 in the BN = 192 mode)
  int box_base, box_rem;   // total_boxes = splits * box_base + box_rem: split sp covers box_base (+1 if sp < box_rem) boxes
  const float* a_amax;     // F16 kernels: device scalars >= max|dY|, >= max|X|
  const float* b_amax;
  float* bias_ws;          // optional (F16 kernels): [splits][Cout] partial sums of dY over the pixels = bias gradient
};

// TS = true: the dY tile (A operand, M = output channel) is transposed into TENSOR MEMORY by four of the split warps
// (thread = channel = TMEM lane, one conflict-free LDS.32 per pixel, hi | lo stored with tcgen05.st) and the MMAs read
// it from there; shared memory keeps the raw dY landing zone and the X tile (hi in place + lo).  This takes the
// 12 A-operand reads and the A hi/lo write-back per stage off the shared-memory port, which is what bounds the SS form
// (224 KB of smem traffic per 768 MMA cycles at BN = 128), and the freed space buys a fourth stage.
// BN = 192 ("tap group" mode, TS only, Cin = 64): the B tile is the X boxes of THREE filter taps side by side
// (3 x 64 channels), so one dY tile -- loaded, transposed and split once -- feeds three taps, and the 24 narrow MMAs per
// pixel stage of three separate items become 12 MMAs of N = 192.  The accumulator is single-buffered there (192 columns +
// 3 x 64 columns of A stages); an item runs for ~100 stages, so the unoverlapped epilogue is noise.
// F16 = true: the X tile is converted in place (hi atoms + lo atoms fill exactly the raw boxes), so a stage is the raw dY
// landing zone + BN/32 X boxes -- which makes an N tile of 256 channels affordable (48 KB stages, 4 of them; accumulator
// single-buffered like the tap-group mode).  At fp16 speed the 128x128 tile is bound by the L2 feed of its two fp32
// operands (32 KB per 384 MMA cycles); 128x256 halves the dY bytes per flop.
template <int BN, bool TS, bool F16 = false>
struct WgSmem {
  static_assert(BN != 192 || TS, "the tap-group mode exists for the tensor-memory form only");
  static_assert(BN != 256 || F16, "the 256-channel N tile exists for the fp16 split only");
  static constexpr int kGroup = (BN == 192) ? 3 : 1;
  static constexpr int kStages = TS ? ((BN == 192 && !F16) ? 3 : 4) : 3;
  static constexpr int kABytes = 4 * kWgBoxBytes;            // 16 KB raw/hi (SS: + same for lo)
  static constexpr int kBBytes = (BN / 32) * kWgBoxBytes;
  static constexpr int kBOff = TS ? kABytes : 2 * kABytes;
  static constexpr int kStageBytes = kBOff + (F16 ? 1 : 2) * kBBytes;
  static constexpr int kTotal = kStages * kStageBytes + 1024 + 1024;
  // kCat (see GemmSmem in gemm_tf32.cu): a_hi x [b_hi ; b_lo] as one MMA of N = 2*BN into two column sets
  static constexpr bool kCat = (BN >= 192) ? false : (TS ? (BN == 64) : true);
  static constexpr int kChains = kCat ? 2 : 1;
  static constexpr int kAccCols = kChains * BN;
  static constexpr int kAccBufs = (BN >= 192) ? 1 : 2;
  static constexpr int kATmem = kAccBufs * kAccCols;
  static constexpr int kTmemNeed = TS ? (kAccBufs * kAccCols + kStages * 64) : kAccBufs * kAccCols;
  static constexpr int kTmemCols = (kTmemNeed <= 128) ? 128 : (kTmemNeed <= 256) ? 256 : 512;
  static_assert(kTmemNeed <= 512, "tensor memory budget");
};

template <int BN, bool TS, bool F16 = false>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                    const WgradParams p) {
  static_assert(!F16 || (TS && BN % 64 == 0), "the fp16 split exists for the tensor-memory form, N tiles of 64 channels");
  using S = WgSmem<BN, TS, F16>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the extern array (no integer round trip) keeps the shared address space: LDS/STS, not LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bars_mem = smem + kStages * S::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bars_mem);
  uint64_t* split_bar = full_bar + kStages;
  uint64_t* empty_bar = split_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.taps_h * p.taps_w;
  const int n_tiles = (taps / p.tap_group) * p.n_ctiles;
  const int total_items = p.m_tiles * n_tiles * p.splits;
  constexpr uint32_t stage_tx = (uint32_t)(S::kABytes + S::kBBytes);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], kWgSplitThreads / 32);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(S::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (m tile, split, tap, cin tile), (tap, cin tile) fastest: CTAs running together work on the SAME pixel
  // range for different taps / channel tiles, so dY and the (overlapping) shifted X boxes are shared through L2 and
  // every pixel range streams from HBM once.
  auto decode = [&](int item, int& mt, int& tap, int& ct, int& sp, int& b0, int& b1) {
    const int nt = item % n_tiles;
    const int r = item / n_tiles;
    sp = r % p.splits;
    mt = r / p.splits;
    tap = nt / p.n_ctiles;
    ct = nt - tap * p.n_ctiles;
    tap *= p.tap_group;   // first tap of the item
    // 32-bit arithmetic only: a 64-bit division is a subroutine call, behind which the compiler no longer treats the
    // loop state as warp-uniform (descriptors then travel through per-MMA R2UR moves in the issuing thread)
    b0 = sp * p.box_base + (sp < p.box_rem ? sp : p.box_rem);
    b1 = b0 + p.box_base + (sp < p.box_rem ? 1 : 0);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int mt, tap, ct, sp, b0, b1;
        decode(item, mt, tap, ct, sp, b0, b1);
        for (int b = b0; b < b1; ++b) {
          const int wb = b % p.boxes_w;
          const int r = b / p.boxes_w;
          const int hb = r % p.boxes_h;
          const int fb = r / p.boxes_h;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * S::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            tma_load_4d(&tmDY, &full_bar[stage], st + g * kWgBoxBytes, mt * kWgBM + g * 32, wb * p.BW, hb * p.BH,
                        fb * p.BF);
          uint8_t* sb = st + S::kBOff;
#pragma unroll
          for (int g = 0; g < BN / 32; ++g) {
            constexpr int kBoxesPerTap = (BN / 32) / S::kGroup;
            const int tapg = tap + g / kBoxesPerTap;
            const int th = tapg / p.taps_w, tw = tapg - th * p.taps_w;
            tma_load_4d(&tmX, &full_bar[stage], sb + g * kWgBoxBytes, ct * (BN / S::kGroup) + (g % kBoxesPerTap) * 32,
                        wb * p.BW * p.stride - p.pad + tw, hb * p.BH * p.stride - p.pad + th, fb * p.BF);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_tf32_mn(kWgBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int mt, tap, ct, sp, b0, b1;
      decode(item, mt, tap, ct, sp, b0, b1);
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * S::kAccCols);
      for (int b = b0; b < b1; ++b) {
        mbar_wait(&full_bar[stage], phase);
        mbar_wait(&split_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = smem_u32(smem + stage * S::kStageBytes);
          const uint64_t b_hi = umma_desc_mn_sw128(st + S::kBOff, kWgBoxBytes);
          const uint64_t b_lo = umma_desc_mn_sw128(st + S::kBOff + S::kBBytes, kWgBoxBytes);
          const uint32_t first = (uint32_t)(b != b0);
          if constexpr (F16) {
            constexpr uint32_t id16 = umma_idesc_f16(kWgBM, BN) | (1u << 16);   // A in TMEM, B MN-major
            constexpr uint32_t id16_cat = umma_idesc_f16(kWgBM, S::kCat ? 2 * BN : BN) | (1u << 16);
            const uint64_t x_hi = umma_desc_mn_sw128_f16(st + S::kBOff, kWgBoxBytes);
            const uint64_t x_lo = umma_desc_mn_sw128_f16(st + S::kBOff + (BN / 64) * kWgBoxBytes, kWgBoxBytes);
            const uint32_t a_hi = tmem_base + (uint32_t)(S::kATmem + stage * 64);
            const uint32_t a_lo = a_hi + 32u;
#pragma unroll
            for (int k = 0; k < kWgRows / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * (2048 >> 4));   // 16 pixels = two 8-row atoms of 1024 B
              const uint32_t ka = (uint32_t)(k * 8);              // 16 fp16 of A = 8 TMEM columns
              if constexpr (S::kCat) {
                umma_f16_ts(d_tmem, a_hi + ka, x_hi + adv, id16_cat, first | (uint32_t)(k != 0));
                umma_f16_ts(d_tmem, a_lo + ka, x_hi + adv, id16, 1);
              } else {
                umma_f16_ts(d_tmem, a_lo + ka, x_hi + adv, id16, first | (uint32_t)(k != 0));
                umma_f16_ts(d_tmem, a_hi + ka, x_lo + adv, id16, 1);
                umma_f16_ts(d_tmem, a_hi + ka, x_hi + adv, id16, 1);
              }
            }
          } else if constexpr (TS) {
            constexpr uint32_t idesc_ts = umma_idesc_tf32(kWgBM, BN) | (1u << 16);   // A in TMEM, B MN-major
            constexpr uint32_t idesc_ts_cat = umma_idesc_tf32(kWgBM, S::kCat ? 2 * BN : BN) | (1u << 16);
            const uint32_t a_hi = tmem_base + (uint32_t)(S::kATmem + stage * 64);
            const uint32_t a_lo = a_hi + 32u;
#pragma unroll
            for (int k = 0; k < kWgRows / 8; ++k) {
              const uint64_t adv = (uint64_t)(k * (1024 >> 4));
              const uint32_t ka = (uint32_t)(k * 8);
              if constexpr (S::kCat) {
                umma_tf32_ts(d_tmem, a_hi + ka, b_hi + adv, idesc_ts_cat, first | (uint32_t)(k != 0));
                umma_tf32_ts(d_tmem, a_lo + ka, b_hi + adv, idesc_ts, 1);
              } else {
                umma_tf32_ts(d_tmem, a_lo + ka, b_hi + adv, idesc_ts, first | (uint32_t)(k != 0));
                umma_tf32_ts(d_tmem, a_hi + ka, b_lo + adv, idesc_ts, 1);
                umma_tf32_ts(d_tmem, a_hi + ka, b_hi + adv, idesc_ts, 1);
              }
            }
          } else {
            constexpr uint32_t idesc_cat = umma_idesc_tf32_mn(kWgBM, 2 * BN);   // B = [x_hi ; x_lo]: N = 2*BN
            const uint64_t a_hi = umma_desc_mn_sw128(st, kWgBoxBytes);
            const uint64_t a_lo = umma_desc_mn_sw128(st + S::kABytes, kWgBoxBytes);
#pragma unroll
            for (int k = 0; k < kWgRows / 8; ++k) {
              const uint64_t adv = (uint64_t)(k * (1024 >> 4));   // 8 pixels = 8 rows of 128 B (two 4-row swizzle atoms)
              umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc_cat, first | (uint32_t)(k != 0));
              umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, 1);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (b == b1 - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (b1 > b0) {
        if (++acc == S::kAccBufs) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 2 + kWgSplitThreads / 32) {
    const int t = threadIdx.x - 64;
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int mt, tap, ct, sp, b0, b1;
      decode(item, mt, tap, ct, sp, b0, b1);
      // bias gradient for free: the threads that transpose the dY tile (one output channel each) see every dY element of
      // the item's pixel range; the items of the first tap / first Cin tile add them up (pixel order, then a fixed-order
      // sum over the K splits in wgrad_reduce_kernel) -- no separate column-sum pass over dY
      float bsum = 0.f;
      const bool want_bias = F16 && p.bias_ws != nullptr && tap == 0 && ct == 0;
      for (int b = b0; b < b1; ++b) {
        mbar_wait(&full_bar[stage], phase);
        uint8_t* st = smem + stage * S::kStageBytes;
        if constexpr (F16) {
          if (warp < 6) {
            // warps 2-5: dY tile -> tensor memory, transposed and packed.  thread = output channel m (TMEM lane); column j
            // of the stage holds pixels (2j, 2j+1) as an fp16 pair (lower K index in the low half), hi at +0, lo at +32.
            const int q = warp & 3;
            const uint8_t* box = st + q * kWgBoxBytes + (lane & 7) * 4;
            const int c8 = lane >> 3;
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(S::kATmem + stage * 64);
            float sa, inv_unused;
            f16_split_scale(__ldg(p.a_amax), sa, inv_unused);
            tc_fence_after();
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int r0 = 2 * u, r1 = 2 * u + 1;
              const float x0 = *reinterpret_cast<const float*>(box + r0 * 128 + ((c8 ^ (r0 & 3)) << 5));
              const float x1 = *reinterpret_cast<const float*>(box + r1 * 128 + ((c8 ^ (r1 & 3)) << 5));
              bsum += x0;
              bsum += x1;
              const float v0 = x0 * sa, v1 = x1 * sa;
              const __half2 h = __floats2half2_rn(v0, v1);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
              hi[u] = *reinterpret_cast<const uint32_t*>(&h);
              lo[u] = *reinterpret_cast<const uint32_t*>(&l);
            }
            tmem_st16(t_lane, hi);
            tmem_st16(t_lane + 32u, lo);
            tmem_st_wait();
            tc_fence_before();
          } else {
            // warps 6-9: X boxes (raw fp32, plain 128B swizzle) -> fp16 MN-major atoms, in place.  A warp owns 8 pixel
            // rows of EVERY box: lane = (pair slot, half, row): lanes 0-15 work on box pair 2i, lanes 16-31 on pair 2i+1;
            // within 16 lanes, lane l and l+8 share a pixel row and own the two boxes of the pair.  All 128-byte lines of
            // the warp's rows are read first, then __syncwarp, then every lane writes its four 16-byte chunks of the hi
            // atom (box area g) and of the lo atom (box area BN/64 + g) at chunk ^ (row % 8).  Because a row of any box
            // is only ever touched by one warp, no cross-warp barrier is needed (a named barrier here made the compiler
            // drop warp-uniformity for the whole kernel: descriptors then travel through R2UR moves in the MMA warp).
            constexpr int kPairs = BN / 64;
            constexpr int kIters = (kPairs + 1) / 2;
            const int row = (warp - 6) * 8 + (lane & 7);
            const int half = (lane >> 3) & 1;
            const uint32_t sw = (uint32_t)(row & 7);
            float sx, inv_unused;
            f16_split_scale(__ldg(p.b_amax), sx, inv_unused);
            uint8_t* xb = st + S::kBOff;
            float4 v[kIters][8];
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
              const int g = 2 * i + (lane >> 4);
              if (g < kPairs) {
                const uint8_t* src = xb + (2 * g + half) * kWgBoxBytes + row * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) v[i][c] = *reinterpret_cast<const float4*>(src + (((uint32_t)c ^ sw) << 4));
              }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
              const int g = 2 * i + (lane >> 4);
              if (g < kPairs) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float x[8] = {v[i][2 * j].x, v[i][2 * j].y, v[i][2 * j].z, v[i][2 * j].w,
                                      v[i][2 * j + 1].x, v[i][2 * j + 1].y, v[i][2 * j + 1].z, v[i][2 * j + 1].w};
                  uint32_t hw[4], lw[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float x0 = x[2 * e] * sx, x1 = x[2 * e + 1] * sx;
                    const __half2 h = __floats2half2_rn(x0, x1);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                    hw[e] = *reinterpret_cast<const uint32_t*>(&h);
                    lw[e] = *reinterpret_cast<const uint32_t*>(&l);
                  }
                  const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)(half * 4 + j)) ^ sw) << 4);
                  *reinterpret_cast<uint4*>(xb + g * kWgBoxBytes + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                  *reinterpret_cast<uint4*>(xb + (kPairs + g) * kWgBoxBytes + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
              }
            }
          }
        } else if constexpr (TS) {
          if (warp < 6) {
            // warps 2-5: dY tile -> tensor memory, transposed.  thread = output channel m (TMEM lane): box m / 32,
            // channel c = m % 32; pixel r of that box sits at r*128 + (((c >> 3) ^ (r & 3)) << 5) + (c & 7)*4
            // (SWIZZLE_128B_ATOM_32B); the 32 lanes of a warp read one 128-byte line per pixel: no bank conflicts.
            const int q = warp & 3;
            const uint8_t* box = st + q * kWgBoxBytes + (lane & 7) * 4;
            const int c8 = lane >> 3;
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(S::kATmem + stage * 64);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                const int r = half * 16 + u;
                const float v = *reinterpret_cast<const float*>(box + r * 128 + ((c8 ^ (r & 3)) << 5));
                const float h = tf32_rna_fast(v);
                hi[u] = __float_as_uint(h);
                lo[u] = __float_as_uint(v - h);
              }
              tmem_st16(t_lane + (uint32_t)(half * 16), hi);
              tmem_st16(t_lane + (uint32_t)(32 + half * 16), lo);
            }
            tmem_st_wait();
            tc_fence_before();
          } else {
            // warps 6-9: X tile, elementwise hi (in place) / lo
            const int tb = threadIdx.x - 6 * 32;
            float4* a = reinterpret_cast<float4*>(st + S::kBOff);
            float4* lo = reinterpret_cast<float4*>(st + S::kBOff + S::kBBytes);
#pragma unroll
            for (int j = 0; j < (S::kBBytes / 16) / 128; ++j) {
              const int i = j * 128 + tb;
              const float4 v = a[i];
              float4 h, l;
              h.x = tf32_rna_fast(v.x); h.y = tf32_rna_fast(v.y); h.z = tf32_rna_fast(v.z); h.w = tf32_rna_fast(v.w);
              l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
              a[i] = h;
              lo[i] = l;
            }
          }
        } else {
          // both tiles' loads are issued before the first use: one exposed shared-memory round trip per stage
          constexpr int kA4 = (S::kABytes / 16) / kWgSplitThreads, kB4 = (S::kBBytes / 16) / kWgSplitThreads;
          float4* a = reinterpret_cast<float4*>(st);
          float4* alo = reinterpret_cast<float4*>(st + S::kABytes);
          float4* x = reinterpret_cast<float4*>(st + S::kBOff);
          float4* xlo = reinterpret_cast<float4*>(st + S::kBOff + S::kBBytes);
          float4 va[kA4], vx[kB4];
#pragma unroll
          for (int j = 0; j < kA4; ++j) va[j] = a[j * kWgSplitThreads + t];
#pragma unroll
          for (int j = 0; j < kB4; ++j) vx[j] = x[j * kWgSplitThreads + t];
#pragma unroll
          for (int j = 0; j < kA4; ++j) {
            const float4 v = va[j];
            float4 h, l;
            h.x = tf32_rna_fast(v.x); h.y = tf32_rna_fast(v.y); h.z = tf32_rna_fast(v.z); h.w = tf32_rna_fast(v.w);
            l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
            a[j * kWgSplitThreads + t] = h;
            alo[j * kWgSplitThreads + t] = l;
          }
#pragma unroll
          for (int j = 0; j < kB4; ++j) {
            const float4 v = vx[j];
            float4 h, l;
            h.x = tf32_rna_fast(v.x); h.y = tf32_rna_fast(v.y); h.z = tf32_rna_fast(v.z); h.w = tf32_rna_fast(v.w);
            l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
            x[j * kWgSplitThreads + t] = h;
            xlo[j * kWgSplitThreads + t] = l;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (want_bias && warp < 6) {
        const int ch = mt * kWgBM + (warp & 3) * 32 + lane;
        if (ch < p.Cout) p.bias_ws[(int64_t)sp * p.Cout + ch] = bsum;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int64_t ktot = (int64_t)taps * p.Cin;
    float inv_a = 1.f, inv_b = 1.f;   // F16: exact power-of-two rescale of the partial sums
    if constexpr (F16) {
      float s_unused;
      f16_split_scale(__ldg(p.a_amax), s_unused, inv_a);
      f16_split_scale(__ldg(p.b_amax), s_unused, inv_b);
    }
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int mt, tap, ct, sp, b0, b1;
      decode(item, mt, tap, ct, sp, b0, b1);
      const int co = mt * kWgBM + row;
      float* dst = p.ws + ((int64_t)sp * p.Cout + co) * ktot + (int64_t)tap * p.Cin + ct * (BN / S::kGroup);
      // valid columns of this item: the rest of Cin, or (tap group) Cin == BN / 3 exactly and all three taps
      const int ncols = (S::kGroup > 1) ? BN : (p.Cin - ct * BN);
      if (b1 <= b0) {   // empty K range (more splits than boxes): the partial is zero
        if (co < p.Cout)
          for (int j = 0; j < BN; ++j)
            if (j < ncols) dst[j] = 0.f;
        continue;
      }
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * S::kAccCols);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_addr + (uint32_t)c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 1; ch < S::kChains; ++ch) {   // fixed summation order over the accumulator chains
          uint32_t r2[32];
          tmem_ld32(t_addr + (uint32_t)(ch * BN + c0), r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        }
        if constexpr (F16) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * inv_a * inv_b);
        }
        if (co < p.Cout) {
          if (c0 + 32 <= ncols && (p.Cin & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              st4(dst + c0 + j, make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                            __uint_as_float(r[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < ncols) dst[c0 + j] = __uint_as_float(r[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == S::kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::kTmemCols));
  }
}

// dW (NCHW: [co][ci][tap]) (+)= sum_s ws[s][co][tap][ci], splits summed in a fixed order.
// SLICED = false (few splits): thread = output, all its partials in flight at once (sixteen at a time).
// SLICED = true (the small early layers have up to ~300 K splits for a few thousand outputs): CTA = 32 outputs x 8 slices;
// a slice adds a contiguous range of splits, the slices are added in slice order through shared memory -- a thread per
// output walking 296 partials was 19 dependent L2 round trips on 16 CTAs.
__device__ __forceinline__ void wgrad_store(float* __restrict__ dw, int64_t i, int Cin, int taps, int accumulate, float s) {
  const int ci = (int)(i % Cin);
  const int64_t r = i / Cin;
  const int tap = (int)(r % taps);
  const int64_t co = r / taps;
  float* d = dw + (co * Cin + ci) * taps + tap;
  *d = accumulate ? *d + s : s;
}

template <bool SLICED>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cout,
                                                          int Cin, int taps, int splits, int accumulate,
                                                          const float* __restrict__ bias_ws, float* __restrict__ dbias) {
  const int64_t n = (int64_t)Cout * taps * Cin;
  if (dbias) {   // bias gradient: the K-split partials of sum_pixels dY, in split order
    for (int c = blockIdx.x * 256 + threadIdx.x; c < Cout; c += gridDim.x * 256)
      dbias[c] = ordered_sum_strided(bias_ws + c, splits, Cout);
  }
  if (!SLICED) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
      wgrad_store(dw, i, Cin, taps, accumulate, ordered_sum_strided(ws + i, splits, n));
  } else {
    __shared__ float sp[8][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int per = (splits + 7) / 8;
    const int k0 = slice * per, k1 = min(splits, k0 + per);
    for (int64_t i0 = (int64_t)blockIdx.x * 32; i0 < n; i0 += (int64_t)gridDim.x * 32) {
      const int64_t i = i0 + lane;
      sp[slice][lane] = (i < n && k1 > k0) ? ordered_sum_strided(ws + (int64_t)k0 * n + i, k1 - k0, n) : 0.f;
      __syncthreads();
      if (slice == 0 && i < n) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += sp[q][lane];
        wgrad_store(dw, i, Cin, taps, accumulate, s);
      }
      __syncthreads();
    }
  }
}

static void wgrad_box(int Wo, int Ho, int& BW, int& BH, int& BF) {
  BW = 1;
  for (int c = 8; c >= 1; c >>= 1)
    if (Wo % c == 0) { BW = c; break; }
  BH = 1;
  for (int c = kWgRows / BW; c >= 1; c >>= 1)
    if (Ho % c == 0) { BH = c; break; }
  BF = kWgRows / (BW * BH);
}

struct WgPlan {
  WgradParams p;
  int bn;
};

// pointwise convolution: the pixels form one dense range, so a stage is simply 32 consecutive pixels (one 2-D TMA tile)
// instead of a {BW, BH, BF} box that has to divide the image geometry
static void flatten_pointwise(int& F, int& H, int& W, int KH, int KW, int stride, int pad) {
  if (KH == 1 && KW == 1 && stride == 1 && pad == 0 && (int64_t)F * H * W < (1ll << 31)) {
    W = F * H * W;
    H = 1;
    F = 1;
  }
}

static int wgrad_plan(int F, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, WgPlan* out,
                      bool f16 = false) {
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  if (Ho <= 0 || Wo <= 0) return VITTA_E_BADARG;
  WgradParams& p = out->p;
  p = WgradParams{};
  p.Cout = Cout; p.Cin = Cin; p.taps_h = KH; p.taps_w = KW; p.stride = stride; p.pad = pad;
  if (KH == 1 && KW == 1 && stride == 1 && pad == 0 && H == 1 && F == 1) {
    p.BW = kWgRows; p.BH = 1; p.BF = 1;
  } else {
    wgrad_box(Wo, Ho, p.BW, p.BH, p.BF);
  }
  p.boxes_w = (Wo + p.BW - 1) / p.BW; p.boxes_h = (Ho + p.BH - 1) / p.BH; p.boxes_f = (F + p.BF - 1) / p.BF;
  out->bn = (Cin <= 64) ? 64 : 128;
  if (f16 && Cin % 256 == 0) out->bn = 256;   // fp16 split: wider N tile (see WgSmem)
  p.tap_group = 1;
  if (Cin == 64 && (KH * KW) % 3 == 0 && g_gemm_operand_form != 1) {   // three taps per item (tensor-memory form only)
    out->bn = 192;
    p.tap_group = 3;
  }
  p.m_tiles = (Cout + kWgBM - 1) / kWgBM;
  p.n_ctiles = (p.tap_group > 1) ? 1 : (Cin + out->bn - 1) / out->bn;
  const int64_t base_items = (int64_t)p.m_tiles * (KH * KW / p.tap_group) * p.n_ctiles;
  const int64_t boxes = (int64_t)p.boxes_w * p.boxes_h * p.boxes_f;
  // K splits: the persistent grid walks base_items * splits items round-robin, so the launch lasts
  //   ceil(items / SMs) * (stages per item + fixed cost per item)
  // -- minimised over the split count.  (Round 2 aimed at "about two items per SM", ceil(2 * SMs / base_items): 297, 304,
  // 306, 320 or 360 items on 148 SMs for most of ResNet-50's layers, i.e. a THIRD pass in which 1-64 CTAs work and the
  // others wait: +25 ... +50 % on those launches.)  The fixed cost stands for the pipeline fill and the un-overlapped
  // part of the 128 x BN partial-tile store; fewer items also mean fewer partial bytes for wgrad_reduce_kernel.
  const int64_t max_by_k = boxes / 8 > 0 ? boxes / 8 : 1;   // at least 8 stages of work per item
  const int64_t sms = cached_sm_count();
  const int64_t kItemCost = 6;
  int64_t splits = 1, best = -1;
  for (int64_t s = 1; s <= max_by_k && s <= 1024; ++s) {
    const int64_t waves = (base_items * s + sms - 1) / sms;
    const int64_t cost = waves * ((boxes + s - 1) / s + kItemCost);
    if (best < 0 || cost < best) {
      best = cost;
      splits = s;
    }
    if (waves > 4) break;   // more passes only add per-item cost
  }
  p.splits = (int)splits;
  p.box_base = (int)(boxes / splits);
  p.box_rem = (int)(boxes % splits);
  return 0;
}

template <int BN, bool TS, bool F16 = false>
static int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, const WgradParams& p, cudaStream_t st) {
  using S = WgSmem<BN, TS, F16>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tf32x3_kernel<BN, TS, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::kTotal);
    if (e != cudaSuccess) {
      set_error("wgrad_tf32x3: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  const int64_t items = (int64_t)p.m_tiles * (p.taps_h * p.taps_w / p.tap_group) * p.n_ctiles * p.splits;
  const int grid = (int)(items < cached_sm_count() ? items : cached_sm_count());
  wgrad_tf32x3_kernel<BN, TS, F16><<<grid, kWgThreads, S::kTotal, st>>>(tdy, tx, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("wgrad_tf32x3 launch: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace vitta

using namespace vitta;

extern "C" {

int64_t vitta_conv2d_wgrad_ws_floats(int F, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
  WgPlan pl;
  if (F <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 || stride < 1) return -1;
  flatten_pointwise(F, H, W, KH, KW, stride, pad);
  if (wgrad_plan(F, H, W, Cin, Cout, KH, KW, stride, pad, &pl)) return -1;
  int splits = pl.p.splits;
  // the fp16 kernel may pick a wider N tile and therefore more K splits: size the workspace for either plan
  if (wgrad_plan(F, H, W, Cin, Cout, KH, KW, stride, pad, &pl, true)) return -1;
  if (pl.p.splits > splits) splits = pl.p.splits;
  return (int64_t)splits * Cout * KH * KW * Cin + (int64_t)splits * Cout;   // weight partials + bias partials
}

// Host-only: the split-K plan of a weight gradient (no device work; callable without a GPU, the SM count then defaults to
// 148).  out = { N tile, base items (Cout tiles x taps x Cin tiles), K splits, 32-pixel stages in total }.
int vitta_conv2d_wgrad_plan(int F, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int f16,
                            int* out) {
  if (!out || F <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 || stride < 1) return VITTA_E_BADARG;
  flatten_pointwise(F, H, W, KH, KW, stride, pad);
  WgPlan pl;
  if (wgrad_plan(F, H, W, Cin, Cout, KH, KW, stride, pad, &pl, f16 != 0)) return VITTA_E_BADARG;
  const WgradParams& p = pl.p;
  out[0] = pl.bn;
  out[1] = p.m_tiles * (KH * KW / p.tap_group) * p.n_ctiles;
  out[2] = p.splits;
  out[3] = p.boxes_w * p.boxes_h * p.boxes_f;
  return 0;
}

static int wgrad_impl(const float* X, const float* dY, int F, int H, int W, int Cin, int Cout, int KH, int KW,
                      int stride, int pad, float* dW, int accumulate, float* ws, void* stream, const float* x_amax,
                      const float* dy_amax, float* dbias = nullptr) {
  const bool f16 = x_amax != nullptr;
  VITTA_CHECK_ARG(!f16 || dy_amax, VITTA_E_BADARG, "conv2d_wgrad_f16x3: both amax scalars are required");
  VITTA_CHECK_ARG(X && dY && dW && ws && F > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, VITTA_E_BADARG,
                  "conv2d_wgrad: bad arguments");
  VITTA_CHECK_ARG(KH > 0 && KW > 0 && stride >= 1 && stride <= 8 && pad >= 0, VITTA_E_BADARG, "conv2d_wgrad: bad filter");
  VITTA_CHECK_ARG(Cin % 4 == 0 && Cout % 4 == 0 && aligned16(X) && aligned16(dY) && aligned16(ws), VITTA_E_ALIGN,
                  "conv2d_wgrad: channel counts must be multiples of 4 and tensors 16-byte aligned");
  flatten_pointwise(F, H, W, KH, KW, stride, pad);
  WgPlan pl;
  int rc = wgrad_plan(F, H, W, Cin, Cout, KH, KW, stride, pad, &pl, f16);
  VITTA_CHECK_ARG(rc == 0, VITTA_E_BADARG, "conv2d_wgrad: empty output");
  WgradParams& p = pl.p;
  p.ws = ws;
  p.a_amax = dy_amax; p.b_amax = x_amax;
  VITTA_CHECK_ARG(!dbias || f16, VITTA_E_UNSUPPORTED, "conv2d_wgrad: the fused bias gradient exists for the fp16 split only");
  // the bias partials live behind the weight partials of the workspace (vitta_conv2d_wgrad_ws_floats sizes both)
  p.bias_ws = dbias ? ws + (int64_t)p.splits * Cout * KH * KW * Cin : nullptr;
  if (f16 && pl.bn == 192 && p.tap_group != 3) return VITTA_E_BADARG;   // (cannot happen: 192 is the tap-group mode)
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  VITTA_CHECK_ARG(p.BW * stride <= 256 && p.BH * stride <= 256, VITTA_E_UNSUPPORTED, "conv2d_wgrad: box too large");
  CUtensorMap tdy, tx;
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)F};
    const uint64_t str[3] = {(uint64_t)Cout * 4, (uint64_t)Cout * 4 * Wo, (uint64_t)Cout * 4 * Wo * Ho};
    const uint32_t box[4] = {32, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BF};
    const uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tensor_map_f32(&tdy, dY, 4, dims, str, box, es, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)F};
    const uint64_t str[3] = {(uint64_t)Cin * 4, (uint64_t)Cin * 4 * W, (uint64_t)Cin * 4 * W * H};
    const uint32_t box[4] = {32, (uint32_t)(p.BW * stride), (uint32_t)(p.BH * stride), (uint32_t)p.BF};
    const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    rc = make_tensor_map_f32(&tx, X, 4, dims, str, box, es, !f16);   // fp16 split: plain 128B swizzle (converted in place)
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (f16)
    rc = (pl.bn == 256) ? launch_wgrad<256, true, true>(tdy, tx, p, st)
         : (pl.bn == 192) ? launch_wgrad<192, true, true>(tdy, tx, p, st)
         : (pl.bn == 64) ? launch_wgrad<64, true, true>(tdy, tx, p, st)
                         : launch_wgrad<128, true, true>(tdy, tx, p, st);
  else if (g_gemm_operand_form == 1)   // automatic = dY through tensor memory (12-15 % faster: profiles/r01_conv_shapes.md)
    rc = (pl.bn == 64) ? launch_wgrad<64, false>(tdy, tx, p, st) : launch_wgrad<128, false>(tdy, tx, p, st);
  else
    rc = (pl.bn == 192) ? launch_wgrad<192, true>(tdy, tx, p, st)
                        : (pl.bn == 64) ? launch_wgrad<64, true>(tdy, tx, p, st) : launch_wgrad<128, true>(tdy, tx, p, st);
  if (rc) return rc;
  const int64_t n = (int64_t)Cout * KH * KW * Cin;
  if (p.splits >= 32 && n <= 148 * 8 * 256) {
    int64_t blocks = (n + 31) / 32;
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(ws, dW, Cout, Cin, KH * KW, p.splits, accumulate, p.bias_ws,
                                                                dbias);
  } else {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(ws, dW, Cout, Cin, KH * KW, p.splits, accumulate, p.bias_ws,
                                                                 dbias);
  }
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_conv2d_wgrad_tf32x3(const float* X, const float* dY, int F, int H, int W, int Cin, int Cout, int KH, int KW,
                              int stride, int pad, float* dW, int accumulate, float* ws, void* stream) {
  return wgrad_impl(X, dY, F, H, W, Cin, Cout, KH, KW, stride, pad, dW, accumulate, ws, stream, nullptr, nullptr);
}

int vitta_conv2d_wgrad_f16x3(const float* X, const float* x_amax, const float* dY, const float* dy_amax, int F, int H,
                             int W, int Cin, int Cout, int KH, int KW, int stride, int pad, float* dW, int accumulate,
                             float* ws, void* stream) {
  VITTA_CHECK_ARG(x_amax && dy_amax, VITTA_E_BADARG, "conv2d_wgrad_f16x3: amax scalars are required");
  return wgrad_impl(X, dY, F, H, W, Cin, Cout, KH, KW, stride, pad, dW, accumulate, ws, stream, x_amax, dy_amax);
}

// vitta_conv2d_wgrad_f16x3 that also returns the bias gradient dbias[co] = sum over the pixels of dY[., co] (nn.Linear /
// nn.Conv2d with bias): the kernel's dY-transposing threads add it up on the way, so no column-sum pass over dY is needed.
int vitta_conv2d_wgrad_f16x3_bias(const float* X, const float* x_amax, const float* dY, const float* dy_amax, int F, int H,
                                  int W, int Cin, int Cout, int KH, int KW, int stride, int pad, float* dW, float* dbias,
                                  int accumulate, float* ws, void* stream) {
  VITTA_CHECK_ARG(x_amax && dy_amax && dbias, VITTA_E_BADARG, "conv2d_wgrad_f16x3_bias: amax scalars and dbias are required");
  return wgrad_impl(X, dY, F, H, W, Cin, Cout, KH, KW, stride, pad, dW, accumulate, ws, stream, x_amax, dy_amax, dbias);
}

}  // extern "C"
