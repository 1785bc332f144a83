// K6/K8: fp32-accurate GEMM / implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05, kind::tf32) with the
// 3xTF32 operand split:   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,   x_hi = rna_tf32(x), x_lo = x - x_hi
// (error ~2^-21 relative per product, fp32 accumulation in TMEM), i.e. fp32-grade results at tensor-core speed.
//
//   D[m, n] = sum_{tap, k} A[row(m) shifted by tap, k] * B[n, tap*Kc + k]   (+ bias[n]) (+ residual) (GELU)
//
// * A (activations, channels-last) arrives RAW through TMA (4-D tiled map {C, W, H, F}, 128B swizzle, OOB = 0
//   gives the convolution padding for free); four "split" warps rewrite each landed stage in place as a_hi and emit
//   a_lo into a second buffer -- an elementwise pass, so the swizzled layout never has to be decoded.
// * B (weights) is pre-split in global memory (vitta_split_tf32, once per optimizer step) and arrives as two TMA tiles.
// * One elected thread issues 3 x (BK/8) tcgen05.mma per stage into a double-buffered TMEM accumulator; four
//   epilogue warps drain the other buffer (tcgen05.ld 32x32b) with the fused bias / GELU / residual epilogue.
// * Persistent: grid = min(#tiles, #SMs); tiles are walked n-fastest so CTAs running together share an A panel in L2.
//
// Warp roles (576 threads): 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2-9 = operand split, 10-17 = epilogue.
// (Eight split warps: profiling showed the four-warp split ~80 % busy per stage at BN = 64 -- LDS latency under
// tensor-core smem traffic, the rounding ALU work and the proxy fence -- i.e. it, not the tensor pipe, set the stage time.)
//
// F16 = true (vitta_*_f16x3 entry points, DESIGN.md section 3 "fp16 split"): the same pipeline on kind::f16.  Operands are
// split as x*s = hi + lo with hi = fp16(x*s), lo = fp16(x*s - hi), s a per-tensor power of two putting amax just below
// 2^14 (the caller passes a device pointer to an upper bound of max|x|); hi*hi + hi*lo + lo*hi has the same ~2^-21 error
// as the tf32 split at twice the tensor-pipe rate and half the weight bytes.  A stage carries 64 K elements: two raw
// fp32 A boxes (32 channels each) that the split warps convert IN PLACE into the fp16 hi tile (first box area) and lo
// tile (second box area), both [128][64] fp16 in exactly the SWIZZLE_128B K-major layout of the tf32 tiles, so shared
// memory plan, descriptors, k-step advance and MMA count per stage are those of the shared-memory-A tf32 form.
#include <cuda_fp16.h>

#include "tc05.cuh"

namespace vitta {

constexpr int kBM = 128;         // UMMA M (one TMEM lane per accumulator row)
constexpr int kBK = 32;          // fp32 elements per stage row = 128 B = one swizzle atom row
constexpr int kEpiWarps = 8;     // two per TMEM lane quadrant: warps q and q + 4 of the role split every 32-column chunk
constexpr int kSplitWarp0 = 2, kSplitWarps = 8, kEpiWarp0 = kSplitWarp0 + kSplitWarps;   // epilogue: warps 10-17
constexpr int kGemmThreads = (kEpiWarp0 + kEpiWarps) * 32;   // 576

struct GemmParams {
  float* C;
  const float* bias;       // [N] or null
  const float* residual;   // same indexing as C (row stride ldr) or null
  int64_t ldc, ldr;
  int M_total;             // rows of C (plain GEMM) -- unused for conv (validity from the box geometry)
  int N;
  int k_chunks;            // ceil(Kc / 32) per tap
  int Kc;                  // channels (K extent per tap)
  int taps_h, taps_w;      // filter extent (1,1 for a plain GEMM)
  int stride, pad;
  int stride_w;            // 0: same as stride; else the step of the TMA w coordinate per output column (stem: 1)
  int Ho, Wo, F;           // output geometry: rows of C = (f, ho, wo); plain GEMM: F = 1, Ho = 1, Wo = M
  int BW, BH, BF;          // output pixels covered by one M tile: BF frames x BH rows x BW cols  (BW*BH*BF <= 128)
  int tiles_w, tiles_h, tiles_f, tiles_n;
  int act;                 // 0 none, 1 exact GELU, 2 multiply by GELU'(residual[m, n]) (residual = saved pre-activation),
                           // 3 ReLU applied last, after bias and residual
  int vec_ok;              // C / bias / residual allow 128-bit accesses
  // optional tap table (strided data gradient): tap t reads the A box shifted by (tap_dh, tap_dw) and the B columns of
  // weight tap tap_wt; 0 taps = the plain taps_h x taps_w raster
  int ntaps;
  int tap_dh[9], tap_dw[9], tap_wt[9];
  // output pixel (f, ho, wo) -> row ((f*out_H + ho*out_sy + out_oy)*out_W + wo*out_sx + out_ox); out_sy == 0: dense
  int out_sy, out_sx, out_oy, out_ox, out_H, out_W;
  float* aux_out;          // optional second output: the pre-activation acc + bias (same indexing as C)
  const float* row_scale;  // optional per-row-group factor (DropPath): v *= row_scale[row / rows_per_group]
  int rows_per_group;
  const float* a_amax;     // F16 kernels: device scalars >= max|A|, >= max|B| (f16_split_scale)
  const float* b_amax;
  float* amax_out;         // optional: max|C| of everything this launch stores (operand range of the next fp16-split GEMM)
};

// ------------------------------------------------------------------------------------------------
// shared memory plan
// ------------------------------------------------------------------------------------------------
// TS = true (BN <= 128): the split warps move the A tile into TENSOR MEMORY (hi | lo, thread = row, tcgen05.st) and the
// MMAs take A from there ([d], [a_tmem], b_desc).  Shared memory then carries the raw A landing zone and the B tiles
// only: per stage the smem port sees TMA writes + one A read + the B operand reads, instead of additionally the
// hi/lo write-back and 12 A operand reads -- the SS form saturates the 128 B/clk shared-memory port long before the
// tensor pipe (BN = 64: ~152 KB per 384 MMA cycles), which is what bounds the narrow-N tiles.
template <int BN, bool TS, bool PAIR = false>
struct GemmSmem {
  static_assert(!TS || BN <= 128, "the TMEM-operand form needs 2*BN accumulator columns + 64 columns per stage");
  // PAIR (cta_group::2): a CTA stages only its half of the B tile's rows, which buys a third stage at BN = 256
  static constexpr int kStages = TS ? 4 : ((BN <= 128 || PAIR) ? 3 : 2);
  static constexpr int kABytes = kBM * kBK * 4;   // 16 KB
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kBK * 4;
  static constexpr int kBOff = TS ? kABytes : 2 * kABytes;   // B hi tile offset inside a stage (B lo follows)
  static constexpr int kStageBytes = kBOff + 2 * kBBytes;
  static constexpr int kBarBytes = 3072;          // barriers (first 256 bytes) + the epilogue's bias tile [2][BN <= 256] floats at +1024
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;   // + alignment slack
  // kCat: the hi and lo tiles of B are adjacent in shared memory, so  a_hi x [b_hi ; b_lo]  is ONE MMA of N = 2*BN whose
  // result lands in two column sets (hi*hi | hi*lo) that the epilogue adds; with a_lo x b_hi that is 2 MMAs per k-step
  // instead of 3 for the same flops.  Profiling (profiles/r01_gemm_issue.md) shows the single issuing thread ~65 % busy
  // at BN = 64 (~55 cycles per tcgen05.mma through the elect / uniform-register sequence), i.e. the instruction count,
  // not the tensor pipe, bounds the narrow-N tiles.  Needs 2 x 2*BN accumulator columns (double-buffered).
  static constexpr bool kCat = TS ? (BN == 64) : (BN <= 128);
  static constexpr int kChains = kCat ? 2 : 1;    // accumulator column sets summed by the epilogue
  static constexpr int kAccCols = kChains * BN;   // columns of one accumulator set
  static constexpr int kATmem = 2 * kAccCols;     // TS: first TMEM column of the A stages (64 columns each: hi | lo)
  static constexpr int kTmemNeed = TS ? (2 * kAccCols + kStages * 64) : 2 * kAccCols;
  static constexpr int kTmemCols = (kTmemNeed <= 32) ? 32 : (kTmemNeed <= 64) ? 64 : (kTmemNeed <= 128) ? 128 : (kTmemNeed <= 256) ? 256 : 512;
  static_assert(kTmemNeed <= 512, "tensor memory budget");
};

// PAIR = true (vitta_gemm_set_cta_pair(1), N tile 256 only): the kernel runs as clusters of two CTAs on one TPC and the
// MMAs are tcgen05.mma.cta_group::2 of M = 256: CTA r of the pair owns M tile 2*pm + r (its own A stages, its own 128
// accumulator lanes and epilogue) and loads rows [r*BN/2, +BN/2) of the B tile, so every weight byte is fetched from L2
// once per 256 output rows instead of once per 128 -- the operand feed, not the tensor pipe, bounds the wide layers
// (DESIGN.md section 9).  Only CTA 0 issues MMAs; the split warps and epilogue warps of CTA 1 arrive on CTA 0's
// barriers through the cluster address space, and CTA 0's commits are multicast to the barriers of both CTAs.
template <int BN, bool TS, bool F16 = false, bool PAIR = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                   const __grid_constant__ CUtensorMap tmBlo, const GemmParams p) {
  static_assert(!(TS && F16), "the fp16 split uses the shared-memory A form");
  static_assert(!PAIR || (!TS && BN == 256), "CTA pairs: shared-memory A form, N tile 256 (single accumulator chain)");
  using S = GemmSmem<BN, TS, PAIR>;
  constexpr int kStages = S::kStages;
  constexpr int kStageK = F16 ? 2 * kBK : kBK;   // K elements per stage (fp16: two raw A boxes)
  constexpr int kBRows = PAIR ? BN / 2 : BN;     // B rows this CTA loads per stage
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  // pointer arithmetic on the extern array (no integer round trip) keeps the shared address space: LDS/STS, not LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bars_mem = smem + kStages * S::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bars_mem);        // TMA landed          [kStages]
  uint64_t* split_bar = full_bar + kStages;                          // a_hi / a_lo written [kStages]
  uint64_t* empty_bar = split_bar + kStages;                         // MMAs retired        [kStages]
  uint64_t* acc_full = empty_bar + kStages;                          // accumulator ready   [2]
  uint64_t* acc_empty = acc_full + 2;                                // accumulator drained [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_f;
  // PAIR: the tile loop runs over pair tiles (two consecutive M tiles x one N tile); `tile_step` pairs work in parallel
  const int total_tiles = PAIR ? ((tiles_m + 1) / 2) * p.tiles_n : tiles_m * p.tiles_n;
// tile loop of a role: CTAs (PAIR: CTA pairs) stride over the tiles.  Spelled with blockIdx / gridDim at every use so that
// the compiler keeps the loop state warp-uniform (descriptors in uniform registers, see profiles/r01_gemm_issue.md).
#define VITTA_TILE_LOOP \
  for (int tile = PAIR ? (blockIdx.x >> 1) : blockIdx.x; tile < total_tiles; tile += PAIR ? (gridDim.x >> 1) : gridDim.x)
  const int taps = p.ntaps ? p.ntaps : p.taps_h * p.taps_w;
  const int k_iters = taps * p.k_chunks;
  const uint32_t a_box_bytes = (uint32_t)(p.BW * p.BH * p.BF) * kBK * 4;
  const uint32_t stage_tx = (F16 ? 2u : 1u) * a_box_bytes + 2u * S::kBBytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], (PAIR ? 2 : 1) * kSplitWarps);    // one elected arrive per split warp (of both CTAs)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], (PAIR ? 2 : 1) * kEpiWarps);    // one elected arrive per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBhi);
    tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) {
    if constexpr (PAIR) {   // one warp of EACH CTA of the pair performs the pair allocation
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(S::kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(S::kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers must be initialised before anyone arrives remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;


  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      VITTA_TILE_LOOP {
        const int nt = tile % p.tiles_n;
        int mt = tile / p.tiles_n;
        if constexpr (PAIR) mt = 2 * mt + (int)cta_rank;   // may equal tiles_m (odd count): an all-OOB tile, nothing stored
        const int wb = mt % p.tiles_w; mt /= p.tiles_w;
        const int hb = mt % p.tiles_h;
        const int fb = mt / p.tiles_h;
        const int w_in0 = wb * p.BW * (p.stride_w ? p.stride_w : p.stride) - p.pad;
        const int h_in0 = hb * p.BH * p.stride - p.pad;
        const int f0 = fb * p.BF;
        const int n0 = nt * BN + (PAIR ? (int)cta_rank * kBRows : 0);   // PAIR: this CTA's half of the B tile
        for (int it = 0; it < k_iters; ++it) {
          const int tap = it / p.k_chunks;
          const int kc = (it - tap * p.k_chunks) * kStageK;
          int th = tap / p.taps_w, tw = tap - th * p.taps_w, wt = tap;
          if (p.ntaps) { th = p.tap_dh[tap]; tw = p.tap_dw[tap]; wt = p.tap_wt[tap]; }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * S::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
          tma_load_4d(&tmA, &full_bar[stage], st, kc, w_in0 + tw, h_in0 + th, f0);
          if constexpr (F16)   // channels kc+32 .. kc+63 (zero fill past Kc)
            tma_load_4d(&tmA, &full_bar[stage], st + S::kABytes, kc + kBK, w_in0 + tw, h_in0 + th, f0);
          tma_load_2d(&tmBhi, &full_bar[stage], st + S::kBOff, wt * p.Kc + kc, n0);
          tma_load_2d(&tmBlo, &full_bar[stage], st + S::kBOff + S::kBBytes, wt * p.Kc + kc, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = F16 ? umma_idesc_f16(kBM, BN) : umma_idesc_tf32(kBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    VITTA_TILE_LOOP {
      if constexpr (PAIR) {
        if (cta_rank != 0) break;   // only the first CTA of a pair issues MMAs
      }
      if constexpr (PAIR) mbar_wait_cluster(&acc_empty[acc], acc_phase ^ 1);
      else mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * S::kAccCols);
      for (int it = 0; it < k_iters; ++it) {
        if constexpr (PAIR) {
          // split done in BOTH CTAs; a CTA's split warps only start after its own TMA landed (A boxes and its B half)
          mbar_wait_cluster(&split_bar[stage], phase);
        } else {
          mbar_wait(&full_bar[stage], phase);    // B tiles (and raw A) landed
          mbar_wait(&split_bar[stage], phase);   // a_hi / a_lo written
        }
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = smem_u32(smem + stage * S::kStageBytes);
          const uint64_t b_hi = umma_desc_sw128(st + S::kBOff);
          const uint64_t b_lo = umma_desc_sw128(st + S::kBOff + S::kBBytes);
          constexpr uint32_t idesc_cat = F16 ? umma_idesc_f16(kBM, S::kCat ? 2 * BN : BN)
                                             : umma_idesc_tf32(kBM, S::kCat ? 2 * BN : BN);   // B = [b_hi ; b_lo], N = 2*BN
          if constexpr (TS) {
            const uint32_t a_hi = tmem_base + (uint32_t)(S::kATmem + stage * 64);   // lane 0; 32 K columns
            const uint32_t a_lo = a_hi + 32u;
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);
              const uint32_t ka = (uint32_t)(k * 8);   // 8 tf32 of A = 8 TMEM columns
              if constexpr (S::kCat) {
                umma_tf32_ts(d_tmem, a_hi + ka, b_hi + adv, idesc_cat, (it | k) != 0);
                umma_tf32_ts(d_tmem, a_lo + ka, b_hi + adv, idesc, 1);
              } else {
                umma_tf32_ts(d_tmem, a_lo + ka, b_hi + adv, idesc, (it | k) != 0);
                umma_tf32_ts(d_tmem, a_hi + ka, b_lo + adv, idesc, 1);
                umma_tf32_ts(d_tmem, a_hi + ka, b_hi + adv, idesc, 1);
              }
            }
          } else {
            const uint64_t a_hi = umma_desc_sw128(st);
            const uint64_t a_lo = umma_desc_sw128(st + S::kABytes);
            if constexpr (PAIR) {
              // M = 256 over the pair, N = BN: the descriptors name this CTA's tiles, the peer uses the same offsets
              constexpr uint32_t idesc2 = F16 ? umma_idesc_f16(2 * kBM, BN) : umma_idesc_tf32(2 * kBM, BN);
#pragma unroll
              for (int k = 0; k < kBK / 8; ++k) {
                const uint64_t adv = (uint64_t)(k * 2);
                if constexpr (F16) {
                  umma_f16_2cta(d_tmem, a_lo + adv, b_hi + adv, idesc2, (it | k) != 0);
                  umma_f16_2cta(d_tmem, a_hi + adv, b_lo + adv, idesc2, 1);
                  umma_f16_2cta(d_tmem, a_hi + adv, b_hi + adv, idesc2, 1);
                } else {
                  umma_tf32_2cta(d_tmem, a_lo + adv, b_hi + adv, idesc2, (it | k) != 0);
                  umma_tf32_2cta(d_tmem, a_hi + adv, b_lo + adv, idesc2, 1);
                  umma_tf32_2cta(d_tmem, a_hi + adv, b_hi + adv, idesc2, 1);
                }
              }
            } else if constexpr (F16) {
#pragma unroll
              for (int k = 0; k < kBK / 8; ++k) {
                const uint64_t adv = (uint64_t)(k * 2);   // 16 fp16 = 32 B = 2 x 16 B inside the swizzle atom row
                if constexpr (S::kCat) {
                  umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc_cat, (it | k) != 0);
                  umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1);
                } else {
                  umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, (it | k) != 0);
                  umma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                  umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
                }
              }
            }
#pragma unroll
            for (int k = 0; k < ((F16 || PAIR) ? 0 : kBK / 8); ++k) {
              const uint64_t adv = (uint64_t)(k * 2);   // 8 tf32 = 32 B = 2 x 16 B inside the swizzle atom row
              if constexpr (S::kCat) {
                umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc_cat, (it | k) != 0);
                umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, 1);
              } else {
                // small terms first, the dominant hi*hi product last
                umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (it | k) != 0);
                umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
              }
            }
          }
          if constexpr (PAIR) {
            umma_commit_2cta(&empty_bar[stage]);                       // both CTAs' producers may refill the stage
            if (it == k_iters - 1) umma_commit_2cta(&acc_full[acc]);   // both CTAs' epilogues may drain their lanes
          } else {
            umma_commit(&empty_bar[stage]);                       // smem stage reusable once these MMAs retire
            if (it == k_iters - 1) umma_commit(&acc_full[acc]);   // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < kEpiWarp0) {
    // ===================== operand split: raw A -> a_hi + a_lo =====================
    const int t = threadIdx.x - kSplitWarp0 * 32;   // 0..255
    int stage = 0;
    uint32_t phase = 0;
    if constexpr (TS) {
      // thread = A row = TMEM lane (a warp may touch lanes 32*(warp%4) .. +31 only).  The row's eight 16-byte chunks
      // sit at chunk ^ (row % 8) inside its 128-byte line (TMA SWIZZLE_128B; stage bases are 1024-aligned).
      const int row = (warp & 3) * 32 + lane;
      const int half = (warp - kSplitWarp0) >> 2;   // two warps per lane quadrant: K columns [0,16) and [16,32)
      const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(S::kATmem + half * 16);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          const uint8_t* arow = smem + stage * S::kStageBytes + row * 128;
          float4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(arow + (((half * 4 + j) ^ (row & 7)) << 4));
          tc_fence_after();   // the MMAs that read this TMEM stage last retired before full_bar could complete
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 x = v[u];
            const float hx = tf32_rna_fast(x.x), hy = tf32_rna_fast(x.y), hz = tf32_rna_fast(x.z), hw = tf32_rna_fast(x.w);
            hi[u * 4] = __float_as_uint(hx); hi[u * 4 + 1] = __float_as_uint(hy);
            hi[u * 4 + 2] = __float_as_uint(hz); hi[u * 4 + 3] = __float_as_uint(hw);
            lo[u * 4] = __float_as_uint(x.x - hx); lo[u * 4 + 1] = __float_as_uint(x.y - hy);
            lo[u * 4 + 2] = __float_as_uint(x.z - hz); lo[u * 4 + 3] = __float_as_uint(x.w - hw);
          }
          tmem_st16(t_lane + (uint32_t)(stage * 64), hi);
          tmem_st16(t_lane + (uint32_t)(stage * 64 + 32), lo);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&split_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if constexpr (F16) {
      // fp16 split, in place.  Lanes l and l+16 of a warp share A row (warp-2)*16 + (l & 15): lane l owns the raw box of
      // channels [32*(l>>4), +32).  Both read their whole 128-byte line, __syncwarp, then each writes its four 16-byte
      // chunks (8 fp16 = K 8j..8j+7, chunk j = 4*(l>>4) + 0..3) of the hi tile (first box area) and of the lo tile (second
      // box area), at chunk ^ (row % 8) like every SWIZZLE_128B K-major row.  A line is only overwritten by the two lanes
      // that read it, so no block-level barrier is needed; loads and stores are bank-conflict free (8 consecutive rows per
      // quarter warp hit 8 distinct chunk positions).
      const int row = (warp - kSplitWarp0) * 16 + (lane & 15);
      const int half = lane >> 4;
      const uint32_t sw = (uint32_t)(row & 7);
      float sa, inv_unused;
      f16_split_scale(__ldg(p.a_amax), sa, inv_unused);
      VITTA_TILE_LOOP {
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          uint8_t* st = smem + stage * S::kStageBytes;
          const uint8_t* src = st + half * S::kABytes + row * 128;
          float4 v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(src + (((uint32_t)c ^ sw) << 4));
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x[8] = {v[2 * j].x, v[2 * j].y, v[2 * j].z, v[2 * j].w,
                                v[2 * j + 1].x, v[2 * j + 1].y, v[2 * j + 1].z, v[2 * j + 1].w};
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x0 = x[2 * e] * sa, x1 = x[2 * e + 1] * sa;
              const __half2 h = __floats2half2_rn(x0, x1);            // .x (low half) = lower K index
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
              hw[e] = *reinterpret_cast<const uint32_t*>(&h);
              lw[e] = *reinterpret_cast<const uint32_t*>(&l);
            }
            const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)(half * 4 + j)) ^ sw) << 4);
            *reinterpret_cast<uint4*>(st + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(st + S::kABytes + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
          fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(&split_bar[stage], 0);   // the issuing CTA's barrier
            else mbar_arrive(&split_bar[stage]);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else {
      VITTA_TILE_LOOP {
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          float4* a = reinterpret_cast<float4*>(smem + stage * S::kStageBytes);
          float4* lo = reinterpret_cast<float4*>(smem + stage * S::kStageBytes + S::kABytes);
#pragma unroll
          for (int j = 0; j < (kBM * kBK / 4) / (kSplitWarps * 32); ++j) {
            const int i = j * (kSplitWarps * 32) + t;
            const float4 v = a[i];
            float4 h, l;
            h.x = tf32_rna_fast(v.x); h.y = tf32_rna_fast(v.y); h.z = tf32_rna_fast(v.z); h.w = tf32_rna_fast(v.w);
            l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
            a[i] = h;
            lo[i] = l;
          }
          fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(&split_bar[stage], 0);   // the issuing CTA's barrier
            else mbar_arrive(&split_bar[stage]);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    // Eight warps: the narrow-K layers (layer-1 pointwise convolutions, the K = 96 ... 192 linear layers of Video-Swin) do
    // ~1.2 K cycles of MMA work per tile against 6-9 K cycles of epilogue in four warps (GELU / GELU' per element, bias,
    // residual, 64 KB of stores: profiles/r02_swin_gemm_*.md).  Warps w and w + 4 of the role share a lane quadrant and
    // take the two 16-column halves of every 32-column chunk.
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int half = (warp - kEpiWarp0) >> 2;     // 16-column half of every chunk
    const int row = q * 32 + lane;                // accumulator row == TMEM lane
    int acc = 0;
    uint32_t acc_phase = 0;
    float out_amax = 0.f;
    // bias of the tile's BN columns, staged in shared memory once per tile (two buffers by tile parity, one named barrier
    // of the four epilogue warps per tile).  Reading it with __ldg inside the store loop was the largest single stall of
    // the narrow-K linear layers (Video-Swin stage 1, K = 96: every group of 8 columns waited for an L2 round trip --
    // the streaming stores evict the 1 KB vector from L1 -- 16 times per row and tile: profiles/r02_swin_gemm_*.md).
    float* sbias_all = reinterpret_cast<float*>(bars_mem + 1024);
    uint32_t tile_par = 0;
    const int et = threadIdx.x - kEpiWarp0 * 32;           // 0..255 within the epilogue warps
    float inv_a = 1.f, inv_b = 1.f;   // F16: 1 / s_a, 1 / s_b -- exact powers of two, applied one after the other (their
    if constexpr (F16) {              // product alone could leave the fp32 range for tiny gradient tensors)
      float s_unused;
      f16_split_scale(__ldg(p.a_amax), s_unused, inv_a);
      f16_split_scale(__ldg(p.b_amax), s_unused, inv_b);
    }
    VITTA_TILE_LOOP {
      const int nt = tile % p.tiles_n;
      int mt = tile / p.tiles_n;
      if constexpr (PAIR) mt = 2 * mt + (int)cta_rank;
      const int wb = mt % p.tiles_w; mt /= p.tiles_w;
      const int hb = mt % p.tiles_h;
      const int fb = mt / p.tiles_h;
      // row -> (frame, ho, wo) inside the box
      const int bw = row % p.BW;
      const int bh = (row / p.BW) % p.BH;
      const int bf = row / (p.BW * p.BH);
      const int wo = wb * p.BW + bw, ho = hb * p.BH + bh, f = fb * p.BF + bf;
      bool row_ok = (bf < p.BF) && (wo < p.Wo) && (ho < p.Ho) && (f < p.F);
      int64_t out_row = ((int64_t)f * p.Ho + ho) * p.Wo + wo;
      if (p.out_sy) {
        const int oy = ho * p.out_sy + p.out_oy, ox = wo * p.out_sx + p.out_ox;
        row_ok = row_ok && oy < p.out_H && ox < p.out_W;
        out_row = ((int64_t)f * p.out_H + oy) * p.out_W + ox;
      }
      float* crow = p.C + out_row * p.ldc;
      const float* rrow = p.residual ? p.residual + out_row * p.ldr : nullptr;
      float* arow = p.aux_out ? p.aux_out + out_row * p.ldc : nullptr;
      const float rscale = (p.row_scale && row_ok) ? __ldg(p.row_scale + out_row / p.rows_per_group) : 1.f;
      const int n0 = nt * BN;
      float* sbias = sbias_all + tile_par * BN;
      if (p.bias) {
        if (et < BN) sbias[et] = (n0 + et < p.N) ? __ldg(p.bias + n0 + et) : 0.f;
        asm volatile("bar.sync 2, 256;" ::: "memory");   // the eight epilogue warps
      }
      tile_par ^= 1u;

      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * S::kAccCols);
#pragma unroll 1
      for (int c32 = 0; c32 < BN; c32 += 32) {
        if (n0 + c32 >= p.N) break;               // columns past N (N = 288 on 128-wide tiles: three of four chunks of the last tile)
        const int c0 = c32 + half * 16;           // this warp's 16 columns of the chunk
        uint32_t r[16], r2[16];
        tmem_ld16(t_addr + (uint32_t)c0, r);
        if constexpr (S::kChains == 2) tmem_ld16(t_addr + (uint32_t)(BN + c0), r2);   // both accumulator chains under one wait
        static_assert(S::kChains <= 2, "epilogue sums at most two accumulator chains");
        // the residual / pre-activation operand of this chunk: both 32-byte loads of the row go out now, under the
        // tensor-memory load, instead of one at a time right in front of their use (a DRAM round trip each)
        const bool res_pre = rrow != nullptr && row_ok && (n0 + c0 + 16 <= p.N) && p.vec_ok == 2;
        float4 rq[4];
        if (res_pre) {
#pragma unroll
          for (int u = 0; u < 2; ++u) ld8(rrow + n0 + c0 + u * 8, rq[2 * u], rq[2 * u + 1]);
        }
        tmem_ld_wait();
        if constexpr (S::kChains == 2) {   // fixed summation order over the accumulator chains
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        }
        if constexpr (F16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * inv_a * inv_b);
        }
        if (row_ok) {
          const int nbase = n0 + c0;
          if (nbase + 16 <= p.N && p.vec_ok == 2) {
            // 256-bit path: every store / load instruction moves one full 32-byte sector per thread
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              float4 v0 = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                      __uint_as_float(r[j + 3]));
              float4 v1 = make_float4(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]), __uint_as_float(r[j + 6]),
                                      __uint_as_float(r[j + 7]));
              if (p.bias) {
                const float4 b0 = *reinterpret_cast<const float4*>(sbias + c0 + j);
                const float4 b1 = *reinterpret_cast<const float4*>(sbias + c0 + j + 4);
                v0.x += b0.x; v0.y += b0.y; v0.z += b0.z; v0.w += b0.w;
                v1.x += b1.x; v1.y += b1.y; v1.z += b1.z; v1.w += b1.w;
              }
              if (p.act == 4) {   // GELU, its derivative to aux_out (instead of the pre-activation)
                float4 d0, d1;
                v0.x = gelu_with_grad(v0.x, d0.x); v0.y = gelu_with_grad(v0.y, d0.y);
                v0.z = gelu_with_grad(v0.z, d0.z); v0.w = gelu_with_grad(v0.w, d0.w);
                v1.x = gelu_with_grad(v1.x, d1.x); v1.y = gelu_with_grad(v1.y, d1.y);
                v1.z = gelu_with_grad(v1.z, d1.z); v1.w = gelu_with_grad(v1.w, d1.w);
                if (arow) st8(arow + nbase + j, d0, d1);
              } else if (arow) {
                st8(arow + nbase + j, v0, v1);
              }
              if (p.act == 1) {
                v0.x = gelu_exact(v0.x); v0.y = gelu_exact(v0.y); v0.z = gelu_exact(v0.z); v0.w = gelu_exact(v0.w);
                v1.x = gelu_exact(v1.x); v1.y = gelu_exact(v1.y); v1.z = gelu_exact(v1.z); v1.w = gelu_exact(v1.w);
              }
              if (rrow) {
                const float4 r0 = rq[j >> 2], r1 = rq[(j >> 2) + 1];
                if (p.act == 2) {
                  v0.x *= gelu_grad(r0.x) * rscale; v0.y *= gelu_grad(r0.y) * rscale;
                  v0.z *= gelu_grad(r0.z) * rscale; v0.w *= gelu_grad(r0.w) * rscale;
                  v1.x *= gelu_grad(r1.x) * rscale; v1.y *= gelu_grad(r1.y) * rscale;
                  v1.z *= gelu_grad(r1.z) * rscale; v1.w *= gelu_grad(r1.w) * rscale;
                } else if (p.act == 5) {   // multiply by the operand (the saved GELU derivative)
                  v0.x *= r0.x * rscale; v0.y *= r0.y * rscale; v0.z *= r0.z * rscale; v0.w *= r0.w * rscale;
                  v1.x *= r1.x * rscale; v1.y *= r1.y * rscale; v1.z *= r1.z * rscale; v1.w *= r1.w * rscale;
                } else {
                  v0.x = fmaf(v0.x, rscale, r0.x); v0.y = fmaf(v0.y, rscale, r0.y);
                  v0.z = fmaf(v0.z, rscale, r0.z); v0.w = fmaf(v0.w, rscale, r0.w);
                  v1.x = fmaf(v1.x, rscale, r1.x); v1.y = fmaf(v1.y, rscale, r1.y);
                  v1.z = fmaf(v1.z, rscale, r1.z); v1.w = fmaf(v1.w, rscale, r1.w);
                }
              } else {
                v0.x *= rscale; v0.y *= rscale; v0.z *= rscale; v0.w *= rscale;
                v1.x *= rscale; v1.y *= rscale; v1.z *= rscale; v1.w *= rscale;
              }
              if (p.act == 3) {   // ReLU after bias and residual (BN-folded inference convolution)
                v0.x = fmaxf(v0.x, 0.f); v0.y = fmaxf(v0.y, 0.f); v0.z = fmaxf(v0.z, 0.f); v0.w = fmaxf(v0.w, 0.f);
                v1.x = fmaxf(v1.x, 0.f); v1.y = fmaxf(v1.y, 0.f); v1.z = fmaxf(v1.z, 0.f); v1.w = fmaxf(v1.w, 0.f);
              }
              if (p.amax_out) {
                out_amax = fmaxf(out_amax, fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))));
                out_amax = fmaxf(out_amax, fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w))));
              }
              st8(crow + nbase + j, v0, v1);
            }
          } else if (nbase + 16 <= p.N && p.vec_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                     __uint_as_float(r[j + 3]));
              if (p.bias) {
                const float4 b = *reinterpret_cast<const float4*>(sbias + c0 + j);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
              }
              if (p.act == 4) {
                float4 d;
                v.x = gelu_with_grad(v.x, d.x); v.y = gelu_with_grad(v.y, d.y);
                v.z = gelu_with_grad(v.z, d.z); v.w = gelu_with_grad(v.w, d.w);
                if (arow) st4(arow + nbase + j, d);
              } else if (arow) {
                st4(arow + nbase + j, v);
              }
              if (p.act == 1) { v.x = gelu_exact(v.x); v.y = gelu_exact(v.y); v.z = gelu_exact(v.z); v.w = gelu_exact(v.w); }
              if (rrow) {
                const float4 rr = *reinterpret_cast<const float4*>(rrow + nbase + j);
                if (p.act == 2) {
                  v.x *= gelu_grad(rr.x); v.y *= gelu_grad(rr.y); v.z *= gelu_grad(rr.z); v.w *= gelu_grad(rr.w);
                  v.x *= rscale; v.y *= rscale; v.z *= rscale; v.w *= rscale;
                } else if (p.act == 5) {
                  v.x *= rr.x * rscale; v.y *= rr.y * rscale; v.z *= rr.z * rscale; v.w *= rr.w * rscale;
                } else {
                  v.x = fmaf(v.x, rscale, rr.x); v.y = fmaf(v.y, rscale, rr.y);
                  v.z = fmaf(v.z, rscale, rr.z); v.w = fmaf(v.w, rscale, rr.w);
                }
              } else {
                v.x *= rscale; v.y *= rscale; v.z *= rscale; v.w *= rscale;
              }
              if (p.act == 3) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              if (p.amax_out) out_amax = fmaxf(out_amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
              st4(crow + nbase + j, v);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = nbase + j;
              if (n < p.N) {
                float v = __uint_as_float(r[j]);
                if (p.bias) v += sbias[c0 + j];
                if (p.act == 4) {
                  float d;
                  v = gelu_with_grad(v, d);
                  if (arow) arow[n] = d;
                } else if (arow) {
                  arow[n] = v;
                }
                if (p.act == 1) v = gelu_exact(v);
                if (rrow) v = (p.act == 2) ? v * gelu_grad(rrow[n]) * rscale
                              : (p.act == 5) ? v * rrow[n] * rscale : fmaf(v, rscale, rrow[n]);
                else v *= rscale;
                if (p.act == 3) v = fmaxf(v, 0.f);
                if (p.amax_out) out_amax = fmaxf(out_amax, fabsf(v));
                crow[n] = v;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(&acc_empty[acc], 0);
        else mbar_arrive(&acc_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.amax_out) {   // one integer atomic per epilogue warp (non-negative floats order like their bit patterns)
      const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(out_amax));
      if (lane == 0 && wmax) atomicMax(reinterpret_cast<unsigned int*>(p.amax_out), wmax);
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // the peer's shared and tensor memory are operands until the last MMA retired
  else __syncthreads();
  if (warp == 1) {
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::kTmemCols));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::kTmemCols));
  }
#undef VITTA_TILE_LOOP
}

// ------------------------------------------------------------------------------------------------
// weight preparation: hi/lo split, optional transposition / filter rotation (for the data-gradient pass)
// ------------------------------------------------------------------------------------------------
// src: [R][T][Cc] (R = out channels, T = taps, Cc = in channels).  mode 0: dst[r][t][c] = src[r][t][c];
// mode 1 (dgrad operand): dst[c][T-1-t][r] = src[r][t][c].
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi,
                                                        float* __restrict__ lo, int R, int T, int Cc, int mode) {
  const int64_t n = (int64_t)R * T * Cc;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    int64_t s = i;
    if (mode == 1) {   // i indexes dst [c][t'][r]
      const int r = (int)(i % R);
      const int64_t q = i / R;
      const int tp = (int)(q % T);
      const int c = (int)(q / T);
      s = ((int64_t)r * T + (T - 1 - tp)) * Cc + c;
    }
    const float v = __ldg(src + s);
    const float h = tf32_rna(v);
    hi[i] = h;
    lo[i] = v - h;
  }
}

// max|x| over a tensor, accumulated into *amax with an integer atomic (non-negative floats order like their bit patterns);
// the caller zero-initialises *amax.  NaN / Inf inputs propagate as a huge bound (garbage in, garbage out).
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ amax) {
  uint32_t m = 0;
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? (n >> 2) : 0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = __ldg(x4 + i);
    m = max(max(m, __float_as_uint(fabsf(v.x))), __float_as_uint(fabsf(v.y)));
    m = max(max(m, __float_as_uint(fabsf(v.z))), __float_as_uint(fabsf(v.w)));
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    m = max(m, __float_as_uint(fabsf(__ldg(x + i))));
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(reinterpret_cast<unsigned int*>(amax), m);
}

// fp16 weight preparation: same index modes as split_tf32_kernel; hi = fp16(x*s), lo = fp16(x*s - hi), s from *amax
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ src, __half* __restrict__ hi,
                                                       __half* __restrict__ lo, const float* __restrict__ amax, int R,
                                                       int T, int Cc, int mode) {
  float sc, inv_unused;
  f16_split_scale(__ldg(amax), sc, inv_unused);
  const int64_t n = (int64_t)R * T * Cc;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    int64_t s = i;
    if (mode == 1) {   // i indexes dst [c][t'][r]
      const int r = (int)(i % R);
      const int64_t q = i / R;
      const int tp = (int)(q % T);
      const int c = (int)(q / T);
      s = ((int64_t)r * T + (T - 1 - tp)) * Cc + c;
    }
    const float v = __ldg(src + s) * sc;
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}

// ------------------------------------------------------------------------------------------------
// multi-tensor weight preparation: every weight of the model, both operand forms, in 2-3 launches per optimizer step
// (instead of one amax + one split launch per weight and form: 312 launches per TANet step).  Same arithmetic as
// split_tf32_kernel / amax_kernel / split_f16_kernel, so the operands are bit-identical to the per-tensor path.
// ------------------------------------------------------------------------------------------------
constexpr int kSplitBlock = 4096;   // elements of one tensor handled by one CTA

struct SplitEntry {
  const float* src;
  void* hi;
  void* lo;
  float* amax;
  int32_t R, T, Cc, mode;
  int32_t src_tap_inner, compute_amax;
  int64_t n;
  const float* fold_w;    // optional eval-mode BatchNorm folded into the rows: row r is scaled by
  const float* fold_rv;   //   fold_w[r] * (1 / sqrt(fold_rv[r] + fold_eps))   (null: no scaling)
  float fold_eps;
  int32_t reserved;
};

__device__ __forceinline__ float split_row_scale(const SplitEntry& t, int r) {
  return t.fold_w ? __ldg(t.fold_w + r) * (1.f / sqrtf(__ldg(t.fold_rv + r) + t.fold_eps)) : 1.f;
}

__device__ __forceinline__ int split_find(const int32_t* __restrict__ block_start, int n_tensors, int bid) {
  int lo = 0, hi = n_tensors - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (block_start[mid] <= bid) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void split_multi_zero_kernel(const SplitEntry* __restrict__ t, int n_tensors) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_tensors; i += gridDim.x * blockDim.x)
    if (t[i].compute_amax) *t[i].amax = 0.f;
}

__global__ void __launch_bounds__(256) split_multi_amax_kernel(const SplitEntry* __restrict__ tensors,
                                                              const int32_t* __restrict__ block_start, int n_tensors) {
  const int ti = split_find(block_start, n_tensors, blockIdx.x);
  const SplitEntry t = tensors[ti];
  if (!t.compute_amax) return;
  const int64_t e0 = (int64_t)(blockIdx.x - block_start[ti]) * kSplitBlock;
  const int64_t rem = t.n - e0;
  const int cnt = (int)(rem < kSplitBlock ? rem : kSplitBlock);
  uint32_t m = 0;
  const int64_t row_elems = (int64_t)t.T * t.Cc;   // both source layouts keep a row's elements together
  for (int i = threadIdx.x; i < cnt; i += 256) {
    float v = __ldg(t.src + e0 + i);
    if (t.fold_w) v *= split_row_scale(t, (int)((e0 + i) / row_elems));
    m = max(m, __float_as_uint(fabsf(v)));
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(reinterpret_cast<unsigned int*>(t.amax), m);
}

// dst index i -> source element; dst is [R][T][Cc] (mode 0) or [Cc][T'][R] with the taps reversed (mode 1); the source is
// [R][T][Cc] or, for src_tap_inner, the contiguous conv weight [R][Cc][T] (no channels_last copy of the weight needed)
template <bool F16>
__global__ void __launch_bounds__(256) split_multi_kernel(const SplitEntry* __restrict__ tensors,
                                                         const int32_t* __restrict__ block_start, int n_tensors) {
  const int ti = split_find(block_start, n_tensors, blockIdx.x);
  const SplitEntry t = tensors[ti];
  const int64_t e0 = (int64_t)(blockIdx.x - block_start[ti]) * kSplitBlock;
  const int64_t rem = t.n - e0;
  const int cnt = (int)(rem < kSplitBlock ? rem : kSplitBlock);
  float sc = 1.f, inv_unused;
  if constexpr (F16) f16_split_scale(__ldcg(t.amax), sc, inv_unused);
  // Fast path: the innermost destination extent (R in mode 1, Cc in mode 0) is a multiple of 8, so 8 consecutive
  // destination elements share their outer indices: one index decomposition (32-bit) and one 16-byte store per operand
  // half instead of eight of each; the source is read with stride 1 (two 16-byte loads), T (tap-inner conv weights) or
  // Cc * T (transposed form; the eight neighbours of a sector are fetched by adjacent threads' groups -> L1).  Same
  // arithmetic per element as the loop below: bit-identical operands.
  const int inner = (t.mode == 1) ? t.R : t.Cc;
  if ((inner & 7) == 0 && t.n < ((int64_t)1 << 31) && (reinterpret_cast<uintptr_t>(t.hi) & 15u) == 0 &&
      (reinterpret_cast<uintptr_t>(t.lo) & 15u) == 0) {
    for (int k = threadIdx.x * 8; k < cnt; k += 256 * 8) {
      const uint32_t i = (uint32_t)e0 + (uint32_t)k;
      uint32_t s0, stride;
      int r0, rstep;
      if (t.mode == 1) {   // dst [c][t'][r], r = r0 .. r0 + 7
        const uint32_t r = i % (uint32_t)t.R, q = i / (uint32_t)t.R;
        const uint32_t tp = (uint32_t)t.T - 1u - q % (uint32_t)t.T, c = q / (uint32_t)t.T;
        s0 = t.src_tap_inner ? (r * (uint32_t)t.Cc + c) * (uint32_t)t.T + tp : (r * (uint32_t)t.T + tp) * (uint32_t)t.Cc + c;
        stride = (uint32_t)t.Cc * (uint32_t)t.T;
        r0 = (int)r; rstep = 1;
      } else {             // dst [r][t][c], c = c0 .. c0 + 7
        const uint32_t c = i % (uint32_t)t.Cc, q = i / (uint32_t)t.Cc;
        const uint32_t tp = q % (uint32_t)t.T, r = q / (uint32_t)t.T;
        s0 = t.src_tap_inner ? (r * (uint32_t)t.Cc + c) * (uint32_t)t.T + tp : (r * (uint32_t)t.T + tp) * (uint32_t)t.Cc + c;
        stride = t.src_tap_inner ? (uint32_t)t.T : 1u;
        r0 = (int)r; rstep = 0;
      }
      float w[8];
      if (stride == 1u && (reinterpret_cast<uintptr_t>(t.src) & 15u) == 0) {
        const float4 a = ldg4(t.src + s0), b = ldg4(t.src + s0 + 4);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = __ldg(t.src + s0 + (uint32_t)j * stride);
      }
      if (t.fold_w) {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = w[j] * split_row_scale(t, r0 + j * rstep);
      }
      if constexpr (F16) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = w[2 * j] * sc, v1 = w[2 * j + 1] * sc;
          const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
          const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
          h[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          l[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint4*>(static_cast<__half*>(t.hi) + i) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(static_cast<__half*>(t.lo) + i) = make_uint4(l[0], l[1], l[2], l[3]);
      } else {
        float hh[8], ll[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hh[j] = tf32_rna(w[j]);
          ll[j] = w[j] - hh[j];
        }
        float* ph = static_cast<float*>(t.hi) + i;
        float* pl = static_cast<float*>(t.lo) + i;
        st4(ph, make_float4(hh[0], hh[1], hh[2], hh[3]));
        st4(ph + 4, make_float4(hh[4], hh[5], hh[6], hh[7]));
        st4(pl, make_float4(ll[0], ll[1], ll[2], ll[3]));
        st4(pl + 4, make_float4(ll[4], ll[5], ll[6], ll[7]));
      }
    }
    return;
  }
  for (int k = threadIdx.x; k < cnt; k += 256) {
    const int64_t i = e0 + k;
    int r, tp, c;
    if (t.mode == 1) {   // i indexes dst [c][t'][r]
      r = (int)(i % t.R);
      const int64_t q = i / t.R;
      tp = t.T - 1 - (int)(q % t.T);
      c = (int)(q / t.T);
    } else {             // i indexes dst [r][t][c]
      c = (int)(i % t.Cc);
      const int64_t q = i / t.Cc;
      tp = (int)(q % t.T);
      r = (int)(q / t.T);
    }
    const int64_t s = t.src_tap_inner ? ((int64_t)r * t.Cc + c) * t.T + tp : ((int64_t)r * t.T + tp) * t.Cc + c;
    const float w = t.fold_w ? __ldg(t.src + s) * split_row_scale(t, r) : __ldg(t.src + s);
    if constexpr (F16) {
      const float v = w * sc;
      const __half h = __float2half_rn(v);
      static_cast<__half*>(t.hi)[i] = h;
      static_cast<__half*>(t.lo)[i] = __float2half_rn(v - __half2float(h));
    } else {
      const float v = w;
      const float h = tf32_rna(v);
      static_cast<float*>(t.hi)[i] = h;
      static_cast<float*>(t.lo)[i] = v - h;
    }
  }
}

// folded bias of an eval-mode BatchNorm behind a convolution: b'[c] = beta[c] - running_mean[c] * k[c], all layers at once
struct FoldBiasEntry {
  const float *w, *b, *rm, *rv;
  float* out;
  float eps;
  int32_t C;
};

__global__ void __launch_bounds__(256) bn_fold_bias_multi_kernel(const FoldBiasEntry* __restrict__ tab) {
  const FoldBiasEntry e = tab[blockIdx.x];
  for (int c = threadIdx.x; c < e.C; c += 256) {
    const float k = __ldg(e.w + c) * (1.f / sqrtf(__ldg(e.rv + c) + e.eps));
    e.out[c] = __ldg(e.b + c) - __ldg(e.rm + c) * k;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

static int make_tensor_map_typed(CUtensorMap* m, CUtensorMapDataType dtype, const void* base, int rank,
                                 const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                 const uint32_t* estr, bool swizzle_atom_32b);

int make_tensor_map_f32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, const uint32_t* estr, bool swizzle_atom_32b) {
  return make_tensor_map_typed(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, estr,
                               swizzle_atom_32b);
}

int make_tensor_map_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, const uint32_t* estr) {
  return make_tensor_map_typed(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box, estr, false);
}

static int make_tensor_map_typed(CUtensorMap* m, CUtensorMapDataType dtype, const void* base, int rank,
                                 const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                 const uint32_t* estr, bool swizzle_atom_32b) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VITTA_E_UNSUPPORTED;
  }
  CUresult r = enc(m, dtype, (cuuint32_t)rank, const_cast<void*>(base),
                   reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                   reinterpret_cast<const cuuint32_t*>(box), reinterpret_cast<const cuuint32_t*>(estr),
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_atom_32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims %llu %llu box %u %u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return VITTA_E_BADARG;
  }
  return 0;
}

static int g_sms = 0;
int cached_sm_count() {
  if (g_sms <= 0) g_sms = vitta_sm_count();
  return g_sms > 0 ? g_sms : 148;
}

template <int BN, bool TS, bool F16 = false, bool PAIR = false>
static int launch_gemm(const CUtensorMap& a, const CUtensorMap& bh, const CUtensorMap& bl, GemmParams p,
                       cudaStream_t st) {
  using S = GemmSmem<BN, TS, PAIR>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, TS, F16, PAIR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) {
      set_error("gemm_tf32x3: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  p.tiles_n = (p.N + BN - 1) / BN;
  const int64_t tiles = (int64_t)p.tiles_w * p.tiles_h * p.tiles_f * p.tiles_n;
  if (tiles <= 0 || tiles >= (1ll << 31)) {
    set_error("gemm_tf32x3: bad tile count");
    return VITTA_E_BADARG;
  }
  cudaError_t e;
  if constexpr (PAIR) {
    // clusters of two CTAs (one TPC); a pair works on two consecutive M tiles of one N tile
    const int64_t tiles_m = (int64_t)p.tiles_w * p.tiles_h * p.tiles_f;
    const int64_t pair_tiles = ((tiles_m + 1) / 2) * p.tiles_n;
    const int64_t max_pairs = cached_sm_count() / 2;
    const int pairs = (int)(pair_tiles < max_pairs ? pair_tiles : max_pairs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = S::kTotal;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<BN, TS, F16, PAIR>, a, bh, bl, p);
    if (e != cudaSuccess) {
      set_error("gemm_tf32x3 (CTA pairs) launch: %s", cudaGetErrorString(e));
      return (int)e;
    }
    return 0;
  }
  const int grid = (int)(tiles < cached_sm_count() ? tiles : cached_sm_count());
  gemm_tf32x3_kernel<BN, TS, F16><<<grid, kGemmThreads, S::kTotal, st>>>(a, bh, bl, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gemm_tf32x3 launch: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// force_bn: 0 = automatic, 64 / 128 / 256 = that N tile; | kForceSS / kForceTS = shared-memory / tensor-memory A operands
// (both forms stay selectable so that they are tested and timed against each other)
constexpr int kForceSS = 0x1000, kForceTS = 0x2000;
constexpr int kConvRelu = 0x4000;   // conv entry points: apply ReLU last (after bias and residual)

static int pick_bn(int N, int forced, int64_t tiles_m = 0) {
  forced &= ~(kForceSS | kForceTS | kConvRelu);
  if (forced == 64 || forced == 128 || forced == 256) return forced;
  if (N <= 64) return 64;
  if (N <= 128 || N % 256 != 0) return 128;
  if (tiles_m > 0) {
    // wave quantisation: the persistent grid has one CTA per SM, so a tile count just above a multiple of the SM count
    // wastes most of the last wave.  Compare the 256-wide tile with the 128-wide one (more, smaller tiles; its MMAs and
    // epilogues are somewhat less efficient per flop).
    const int sms = cached_sm_count();
    auto wave_eff = [&](int64_t tiles) { return (double)tiles / (double)(((tiles + sms - 1) / sms) * sms); };
    const double e256 = wave_eff(tiles_m * (N / 256));
    const double e128 = wave_eff(tiles_m * (N / 128)) * 0.88;
    if (e128 > e256) return 128;
  }
  return 256;
}

int g_gemm_operand_form = 0;   // vitta_gemm_set_operand_form: 0 automatic, 1 shared-memory A, 2 tensor-memory A
int g_gemm_cta_pair = 0;       // vitta_gemm_set_cta_pair: N = 256 tiles as cta_group::2 pairs (opt-in until validated on hardware)

static bool use_pair(int bn, int force) { return g_gemm_cta_pair && bn == 256 && !(force & kForceTS); }

// form: bits of force_bn (kForceSS / kForceTS), else the process-wide setting, else automatic
static int dispatch(const CUtensorMap& a, const CUtensorMap& bh, const CUtensorMap& bl, const GemmParams& p, int bn,
                    cudaStream_t st, int force = 0) {
  if (p.a_amax) {   // fp16 split: shared-memory A form only
    if (bn == 64) return launch_gemm<64, false, true>(a, bh, bl, p, st);
    if (bn == 128) return launch_gemm<128, false, true>(a, bh, bl, p, st);
    if (use_pair(bn, force)) return launch_gemm<256, false, true, true>(a, bh, bl, p, st);
    return launch_gemm<256, false, true>(a, bh, bl, p, st);
  }
  if (use_pair(bn, force)) return launch_gemm<256, false, false, true>(a, bh, bl, p, st);
  int form = (force & kForceSS) ? 1 : (force & kForceTS) ? 2 : g_gemm_operand_form;
  if (form == 0) form = (bn == 64) ? 2 : 1;   // measured per tile width: profiles/r01_conv_shapes.md
  const bool ss = form == 1;
  if (bn == 64) return ss ? launch_gemm<64, false>(a, bh, bl, p, st) : launch_gemm<64, true>(a, bh, bl, p, st);
  if (bn == 128) return ss ? launch_gemm<128, false>(a, bh, bl, p, st) : launch_gemm<128, true>(a, bh, bl, p, st);
  return launch_gemm<256, false>(a, bh, bl, p, st);
}

static int make_b_maps(CUtensorMap* bh, CUtensorMap* bl, const void* Bhi, const void* Blo, int64_t ldb, int N,
                       int Ktot, int bn, bool f16 = false, int force = 0) {
  if (use_pair(bn, force)) bn /= 2;   // CTA pairs: each CTA loads half of the B tile's rows
  const uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)N};
  const uint32_t es[2] = {1, 1};
  if (f16) {   // [N][Ktot] fp16, 64 K elements (128 B) per stage row
    const uint64_t str[1] = {(uint64_t)ldb * 2};
    const uint32_t box[2] = {(uint32_t)(2 * kBK), (uint32_t)bn};
    int rc = make_tensor_map_f16(bh, Bhi, 2, dims, str, box, es);
    if (rc) return rc;
    return make_tensor_map_f16(bl, Blo, 2, dims, str, box, es);
  }
  const uint64_t str[1] = {(uint64_t)ldb * 4};
  const uint32_t box[2] = {(uint32_t)kBK, (uint32_t)bn};
  int rc = make_tensor_map_f32(bh, Bhi, 2, dims, str, box, es);
  if (rc) return rc;
  return make_tensor_map_f32(bl, Blo, 2, dims, str, box, es);
}

// M tile = BF frames x BH rows x BW cols of output pixels, at most 128
static void pick_boxes(int Wo, int Ho, int F, int* pBW, int* pBH, int* pBF) {
  const int BW = Wo < kBM ? Wo : kBM;
  int BH = 1, BF = 1;
  if (BW == Wo) {
    int bh_max = kBM / BW;
    if (bh_max > Ho) bh_max = Ho;
    BH = bh_max;
    for (int d = bh_max; d >= 1; --d) {   // largest divisor of Ho that fits; keep it unless it halves the tile
      if (Ho % d == 0) {
        if (2 * d > bh_max) BH = d;
        break;
      }
    }
    if (BH == Ho) {
      BF = kBM / (BW * BH);
      if (BF > F) BF = F;
      if (BF < 1) BF = 1;
    }
  }
  *pBW = BW; *pBH = BH; *pBF = BF;
}

}  // namespace vitta

using namespace vitta;

extern "C" {

int vitta_gemm_set_cta_pair(int on) {
  g_gemm_cta_pair = on ? 1 : 0;
  return 0;
}

int vitta_gemm_set_operand_form(int form) {
  VITTA_CHECK_ARG(form >= 0 && form <= 2, VITTA_E_BADARG,
                  "gemm_set_operand_form: 0 (automatic), 1 (shared-memory A) or 2 (tensor-memory A)");
  g_gemm_operand_form = form;
  return 0;
}

int vitta_split_tf32(const float* src, float* hi, float* lo, int R, int T, int Cc, int mode, void* stream) {
  VITTA_CHECK_ARG(src && hi && lo && R > 0 && T > 0 && Cc > 0 && (mode == 0 || mode == 1), VITTA_E_BADARG,
                  "split_tf32: bad arguments");
  const int64_t n = (int64_t)R * T * Cc;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, hi, lo, R, T, Cc, mode);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_gemm_tf32x3(const float* A, int64_t lda, const float* Bhi, const float* Blo, int64_t ldb, float* C,
                      int64_t ldc, int64_t M, int N, int K, const float* bias, const float* residual, int64_t ldr,
                      int act, int force_bn, void* stream) {
  return vitta_gemm_tf32x3_ex(A, lda, Bhi, Blo, ldb, C, ldc, M, N, K, bias, residual, ldr, act, nullptr, nullptr, 1,
                              force_bn, stream);
}

// shared by the tf32 (a_amax == null, B = fp32 hi/lo) and fp16 (a_amax / b_amax given, B = fp16 hi/lo) entry points
static int gemm_impl(const float* A, int64_t lda, const void* Bhi, const void* Blo, int64_t ldb, float* C,
                     int64_t ldc, int64_t M, int N, int K, const float* bias, const float* residual, int64_t ldr,
                     int act, float* aux_out, const float* row_scale, int64_t rows_per_group, int force_bn,
                     void* stream, const float* a_amax, const float* b_amax, float* c_amax = nullptr) {
  const bool f16 = a_amax != nullptr;
  const int stage_k = f16 ? 2 * kBK : kBK;
  VITTA_CHECK_ARG(A && Bhi && Blo && C && M > 0 && N > 0 && K > 0, VITTA_E_BADARG, "gemm_tf32x3: bad arguments");
  VITTA_CHECK_ARG(!f16 || (b_amax && ldb % 8 == 0), VITTA_E_BADARG,
                  "gemm_f16x3: needs both amax scalars and fp16 weight rows that are multiples of 8 elements");
  VITTA_CHECK_ARG((act >= 0 && act <= 2 || act == 4 || act == 5) && !((act == 2 || act == 5) && !residual), VITTA_E_BADARG,
                  "gemm_tf32x3: act must be 0/1/2 and act 2 needs the pre-activation in `residual`");
  VITTA_CHECK_ARG(!row_scale || (rows_per_group > 0 && rows_per_group < (1ll << 31)), VITTA_E_BADARG,
                  "gemm_tf32x3: rows_per_group");
  VITTA_CHECK_ARG(M < (1ll << 31), VITTA_E_UNSUPPORTED, "gemm_tf32x3: M too large");
  VITTA_CHECK_ARG((lda % 4) == 0 && (ldb % 4) == 0 && aligned16(A) && aligned16(Bhi) && aligned16(Blo), VITTA_E_ALIGN,
                  "gemm_tf32x3: operands need 16-byte aligned rows (lda, ldb multiples of 4 floats)");
  VITTA_CHECK_ARG(lda >= K && ldb >= K && ldc >= N, VITTA_E_BADARG, "gemm_tf32x3: leading dimension too small");
  const int bn = pick_bn(N, force_bn, (M + kBM - 1) / kBM);
  CUtensorMap ta, tbh, tbl;
  {
    const uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, 1, 1};
    const uint64_t str[3] = {(uint64_t)lda * 4, (uint64_t)lda * 4 * (uint64_t)M, (uint64_t)lda * 4 * (uint64_t)M};
    const uint32_t box[4] = {(uint32_t)kBK, (uint32_t)kBM, 1, 1};
    const uint32_t es[4] = {1, 1, 1, 1};
    int rc = make_tensor_map_f32(&ta, A, 4, dims, str, box, es);
    if (rc) return rc;
  }
  int rc = make_b_maps(&tbh, &tbl, Bhi, Blo, ldb, N, K, bn, f16, force_bn);
  if (rc) return rc;
  GemmParams p{};
  p.a_amax = a_amax; p.b_amax = b_amax; p.amax_out = c_amax;
  p.C = C; p.bias = bias; p.residual = residual; p.ldc = ldc; p.ldr = ldr;
  p.M_total = (int)M; p.N = N; p.Kc = K; p.k_chunks = (K + stage_k - 1) / stage_k;
  p.taps_h = p.taps_w = 1; p.stride = 1; p.pad = 0;
  p.Ho = 1; p.Wo = (int)M; p.F = 1;
  p.BW = kBM; p.BH = 1; p.BF = 1;
  p.tiles_w = (int)((M + kBM - 1) / kBM); p.tiles_h = 1; p.tiles_f = 1;
  p.act = act;
  p.aux_out = aux_out; p.row_scale = row_scale; p.rows_per_group = row_scale ? (int)rows_per_group : 1;
  p.vec_ok = aligned16(C) && (ldc % 4 == 0) && (!bias || aligned16(bias)) &&
             (!residual || (aligned16(residual) && ldr % 4 == 0)) && (!aux_out || aligned16(aux_out));
  if (p.vec_ok && aligned32(C) && (ldc % 8 == 0) && (!residual || (aligned32(residual) && ldr % 8 == 0)) &&
      (!aux_out || aligned32(aux_out)))
    p.vec_ok = 2;
  return dispatch(ta, tbh, tbl, p, bn, (cudaStream_t)stream, force_bn);
}

int vitta_conv2d_tf32x3(const float* X, int F, int H, int W, int Cin, const float* Whi, const float* Wlo, int Cout,
                        int KH, int KW, int stride, int pad, float* Y, const float* bias, int force_bn, void* stream) {
  return vitta_conv2d_tf32x3_ex(X, F, H, W, Cin, Whi, Wlo, Cout, KH, KW, stride, pad, Y, bias, nullptr, force_bn, stream);
}

static int conv2d_impl(const float* X, int F, int H, int W, int Cin, const void* Whi, const void* Wlo, int Cout,
                       int KH, int KW, int stride, int pad, float* Y, const float* bias, const float* residual,
                       int force_bn, void* stream, const float* a_amax, const float* b_amax, float* amax_out = nullptr) {
  const bool f16 = a_amax != nullptr;
  const int stage_k = f16 ? 2 * kBK : kBK;
  VITTA_CHECK_ARG(X && Whi && Wlo && Y && F > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, VITTA_E_BADARG,
                  "conv2d_tf32x3: bad arguments");
  VITTA_CHECK_ARG(!f16 || (b_amax && (KH * KW * Cin) % 8 == 0), VITTA_E_BADARG,
                  "conv2d_f16x3: needs both amax scalars and KH*KW*Cin a multiple of 8");
  VITTA_CHECK_ARG(KH > 0 && KW > 0 && stride >= 1 && stride <= 8 && pad >= 0, VITTA_E_BADARG, "conv2d_tf32x3: bad filter");
  VITTA_CHECK_ARG(Cin % 4 == 0 && aligned16(X) && aligned16(Whi) && aligned16(Wlo), VITTA_E_ALIGN,
                  "conv2d_tf32x3: Cin must be a multiple of 4 and tensors 16-byte aligned");
  if (KH == 1 && KW == 1 && stride == 1 && pad == 0 && (int64_t)F * H * W < (1ll << 31)) {
    // pointwise convolution: the pixels form one dense row range, so M tiles need not respect the image geometry
    // (a 14x14 or 7x7 map only fills 98 of the 128 rows of a {W, H, F} box)
    W = F * H * W;
    H = 1;
    F = 1;
  }
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  VITTA_CHECK_ARG(Ho > 0 && Wo > 0, VITTA_E_BADARG, "conv2d_tf32x3: empty output");
  int BW, BH, BF;
  pick_boxes(Wo, Ho, F, &BW, &BH, &BF);
  VITTA_CHECK_ARG((int64_t)BW * stride <= 256 && (int64_t)BH * stride <= 256, VITTA_E_UNSUPPORTED,
                  "conv2d_tf32x3: box exceeds the TMA limit");
  const int bn = pick_bn(Cout, force_bn,
                         (int64_t)((Wo + BW - 1) / BW) * ((Ho + BH - 1) / BH) * ((F + BF - 1) / BF));
  CUtensorMap ta, tbh, tbl;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)F};
    const uint64_t str[3] = {(uint64_t)Cin * 4, (uint64_t)Cin * 4 * W, (uint64_t)Cin * 4 * W * H};
    const uint32_t box[4] = {(uint32_t)kBK, (uint32_t)(BW * stride), (uint32_t)(BH * stride), (uint32_t)BF};
    const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = make_tensor_map_f32(&ta, X, 4, dims, str, box, es);
    if (rc) return rc;
  }
  const int Ktot = KH * KW * Cin;
  int rc = make_b_maps(&tbh, &tbl, Whi, Wlo, Ktot, Cout, Ktot, bn, f16, force_bn);
  if (rc) return rc;
  GemmParams p{};
  p.a_amax = a_amax; p.b_amax = b_amax; p.amax_out = amax_out;
  p.C = Y; p.bias = bias; p.residual = residual; p.ldc = Cout; p.ldr = Cout;
  p.M_total = 0; p.N = Cout; p.Kc = Cin; p.k_chunks = (Cin + stage_k - 1) / stage_k;
  p.taps_h = KH; p.taps_w = KW; p.stride = stride; p.pad = pad;
  p.Ho = Ho; p.Wo = Wo; p.F = F;
  p.BW = BW; p.BH = BH; p.BF = BF;
  p.tiles_w = (Wo + BW - 1) / BW; p.tiles_h = (Ho + BH - 1) / BH; p.tiles_f = (F + BF - 1) / BF;
  p.act = (force_bn & kConvRelu) ? 3 : 0;
  p.vec_ok = aligned16(Y) && (Cout % 4 == 0) && (!bias || aligned16(bias)) && (!residual || aligned16(residual));
  if (p.vec_ok && aligned32(Y) && (Cout % 8 == 0) && (!residual || aligned32(residual))) p.vec_ok = 2;
  return dispatch(ta, tbh, tbl, p, bn, (cudaStream_t)stream, force_bn);
}

static int dgrad_impl(const float* dY, int F, int Ho, int Wo, int Cout, const void* Wthi, const void* Wtlo,
                      int Cin, int KH, int KW, int stride, int pad, int H, int W, float* dX, void* stream,
                      const float* a_amax, const float* b_amax) {
  const bool f16 = a_amax != nullptr;
  const int stage_k = f16 ? 2 * kBK : kBK;
  VITTA_CHECK_ARG(!f16 || (b_amax && (KH * KW * Cout) % 8 == 0), VITTA_E_BADARG,
                  "conv2d_dgrad_f16x3: needs both amax scalars and KH*KW*Cout a multiple of 8");
  VITTA_CHECK_ARG(dY && Wthi && Wtlo && dX && F > 0 && Ho > 0 && Wo > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0,
                  VITTA_E_BADARG, "conv2d_dgrad: bad arguments");
  VITTA_CHECK_ARG(KH > 0 && KW > 0 && KH * KW <= 9 && stride >= 1 && stride <= 4 && pad >= 0, VITTA_E_UNSUPPORTED,
                  "conv2d_dgrad: filters up to 3x3, stride up to 4");
  VITTA_CHECK_ARG(Ho == (H + 2 * pad - KH) / stride + 1 && Wo == (W + 2 * pad - KW) / stride + 1, VITTA_E_BADARG,
                  "conv2d_dgrad: output geometry does not match");
  VITTA_CHECK_ARG(Cout % 4 == 0 && aligned16(dY) && aligned16(Wthi) && aligned16(Wtlo), VITTA_E_ALIGN,
                  "conv2d_dgrad: Cout must be a multiple of 4 and tensors 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ktot = KH * KW * Cout;
  const int bn = pick_bn(Cin, 0);
  CUtensorMap tbh, tbl;
  int rc = make_b_maps(&tbh, &tbl, Wthi, Wtlo, Ktot, Cin, Ktot, bn, f16);
  if (rc) return rc;
  // residue classes no filter tap reaches (e.g. 1x1 stride 2) have zero gradient: clear dX first if there are any
  bool any_empty = false;
  for (int a = 0; a < stride; ++a) {
    bool hit = false;
    for (int kh = 0; kh < KH; ++kh) hit = hit || ((a + pad - kh) % stride == 0);
    any_empty = any_empty || !hit;
  }
  for (int b = 0; b < stride; ++b) {
    bool hit = false;
    for (int kw = 0; kw < KW; ++kw) hit = hit || ((b + pad - kw) % stride == 0);
    any_empty = any_empty || !hit;
  }
  if (any_empty) {
    cudaError_t e = cudaMemsetAsync(dX, 0, (size_t)F * H * W * Cin * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("conv2d_dgrad: memset: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  // one launch per residue class (a, b) of the input-pixel coordinates modulo the stride
  for (int a = 0; a < stride && a < H; ++a) {
    for (int b = 0; b < stride && b < W; ++b) {
      GemmParams p{};
      int nt = 0;
      for (int kh = 0; kh < KH; ++kh) {
        if ((a + pad - kh) % stride != 0) continue;
        for (int kw = 0; kw < KW; ++kw) {
          if ((b + pad - kw) % stride != 0) continue;
          // floor division is exact here; negative offsets fall into the TMA zero fill
          p.tap_dh[nt] = (a + pad - kh) / stride;
          p.tap_dw[nt] = (b + pad - kw) / stride;
          p.tap_wt[nt] = (KH - 1 - kh) * KW + (KW - 1 - kw);   // split mode 1 stores the filter rotated by 180 degrees
          ++nt;
        }
      }
      if (nt == 0) continue;   // no filter tap reaches this class: already zeroed
      const int Hc = (H - a + stride - 1) / stride, Wc = (W - b + stride - 1) / stride;
      int BW, BH, BF;
      pick_boxes(Wc, Hc, F, &BW, &BH, &BF);
      CUtensorMap ta;
      {
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)F};
        const uint64_t str[3] = {(uint64_t)Cout * 4, (uint64_t)Cout * 4 * Wo, (uint64_t)Cout * 4 * Wo * Ho};
        const uint32_t box[4] = {(uint32_t)kBK, (uint32_t)BW, (uint32_t)BH, (uint32_t)BF};
        const uint32_t es[4] = {1, 1, 1, 1};
        rc = make_tensor_map_f32(&ta, dY, 4, dims, str, box, es);
        if (rc) return rc;
      }
      p.a_amax = a_amax; p.b_amax = b_amax;
      p.C = dX; p.ldc = Cin; p.N = Cin; p.Kc = Cout; p.k_chunks = (Cout + stage_k - 1) / stage_k;
      p.taps_h = p.taps_w = 1; p.stride = 1; p.pad = 0; p.ntaps = nt;
      p.Ho = Hc; p.Wo = Wc; p.F = F;
      p.BW = BW; p.BH = BH; p.BF = BF;
      p.tiles_w = (Wc + BW - 1) / BW; p.tiles_h = (Hc + BH - 1) / BH; p.tiles_f = (F + BF - 1) / BF;
      p.out_sy = stride; p.out_sx = stride; p.out_oy = a; p.out_ox = b; p.out_H = H; p.out_W = W;
      p.vec_ok = aligned16(dX) && (Cin % 4 == 0);
      if (p.vec_ok && aligned32(dX) && (Cin % 8 == 0)) p.vec_ok = 2;
      rc = dispatch(ta, tbh, tbl, p, bn, st, 0);
      if (rc) return rc;
    }
  }
  return 0;
}

// ResNet stem: 7x7 / stride 2 / pad 3 convolution of a 3-channel image, on the same kernel.  The image arrives
// zero-padded and widened to 4 channels (vitta_stem_pack: XP[F][H+6][W+6][4]), so that the 8 pixels x 4 channels = 32
// floats under one filter ROW of an output pixel are contiguous: the A operand of filter row kh is a [pixels x 32] matrix
// whose rows OVERLAP in memory (row wo starts 2 pixels = 32 bytes after row wo-1).  TMA expresses that directly -- a
// tensor map whose second dimension (the output column) has a 32-byte stride under a 128-byte inner extent -- so the
// convolution is 7 "taps" of K = 32 (kw padded 7 -> 8, c padded 3 -> 4; the weight operand holds zeros there) with no
// im2col buffer.  Replaces the cuDNN fp32 call for torchvision's conv1 (reference models/tanet_models/tanet.py:129).
int vitta_stem_conv_tf32x3(const float* XP, int F, int H, int W, const float* Whi, const float* Wlo, float* Y,
                           void* stream) {
  VITTA_CHECK_ARG(XP && Whi && Wlo && Y && F > 0 && H >= 7 && W >= 7 && H % 2 == 0 && W % 2 == 0, VITTA_E_BADARG,
                  "stem_conv: bad arguments (even H, W >= 7)");
  VITTA_CHECK_ARG(aligned16(XP) && aligned16(Whi) && aligned16(Wlo) && aligned32(Y), VITTA_E_ALIGN,
                  "stem_conv: tensors must be 16-byte aligned (output 32)");
  const int Hp = H + 6, Wp = W + 6, Ho = H / 2, Wo = W / 2, Cout = 64;
  int BW, BH, BF;
  pick_boxes(Wo, Ho, F, &BW, &BH, &BF);
  VITTA_CHECK_ARG(BH * 2 <= 256, VITTA_E_UNSUPPORTED, "stem_conv: box exceeds the TMA limit");
  CUtensorMap ta, tbh, tbl;
  {
    // dims {32 floats of a window row, output column, padded input row, frame}
    const uint64_t dims[4] = {32, (uint64_t)Wo, (uint64_t)Hp, (uint64_t)F};
    const uint64_t str[3] = {32, (uint64_t)Wp * 16, (uint64_t)Wp * 16 * Hp};
    const uint32_t box[4] = {32, (uint32_t)BW, (uint32_t)(BH * 2), (uint32_t)BF};
    const uint32_t es[4] = {1, 1, 2, 1};
    int rc = make_tensor_map_f32(&ta, XP, 4, dims, str, box, es);
    if (rc) return rc;
  }
  int rc = make_b_maps(&tbh, &tbl, Whi, Wlo, 224, Cout, 224, 64);
  if (rc) return rc;
  GemmParams p{};
  p.C = Y; p.ldc = Cout; p.ldr = Cout;
  p.N = Cout; p.Kc = 32; p.k_chunks = 1;
  p.taps_h = 7; p.taps_w = 1; p.stride = 2; p.stride_w = 1; p.pad = 0;
  p.Ho = Ho; p.Wo = Wo; p.F = F;
  p.BW = BW; p.BH = BH; p.BF = BF;
  p.tiles_w = (Wo + BW - 1) / BW; p.tiles_h = (Ho + BH - 1) / BH; p.tiles_f = (F + BF - 1) / BF;
  p.vec_ok = 2;
  return dispatch(ta, tbh, tbl, p, 64, (cudaStream_t)stream, 0);
}

int vitta_gemm_tf32x3_ex(const float* A, int64_t lda, const float* Bhi, const float* Blo, int64_t ldb, float* C,
                         int64_t ldc, int64_t M, int N, int K, const float* bias, const float* residual, int64_t ldr,
                         int act, float* aux_out, const float* row_scale, int64_t rows_per_group, int force_bn,
                         void* stream) {
  return gemm_impl(A, lda, Bhi, Blo, ldb, C, ldc, M, N, K, bias, residual, ldr, act, aux_out, row_scale, rows_per_group,
                   force_bn, stream, nullptr, nullptr);
}

int vitta_conv2d_tf32x3_ex(const float* X, int F, int H, int W, int Cin, const float* Whi, const float* Wlo, int Cout,
                           int KH, int KW, int stride, int pad, float* Y, const float* bias, const float* residual,
                           int force_bn, void* stream) {
  return conv2d_impl(X, F, H, W, Cin, Whi, Wlo, Cout, KH, KW, stride, pad, Y, bias, residual, force_bn, stream, nullptr,
                     nullptr);
}

int vitta_conv2d_dgrad_tf32x3(const float* dY, int F, int Ho, int Wo, int Cout, const float* Wthi, const float* Wtlo,
                              int Cin, int KH, int KW, int stride, int pad, int H, int W, float* dX, void* stream) {
  return dgrad_impl(dY, F, Ho, Wo, Cout, Wthi, Wtlo, Cin, KH, KW, stride, pad, H, W, dX, stream, nullptr, nullptr);
}

// ---- fp16 split (experimental in round 1: compiled and exported, not yet used by the step) ----
int vitta_amax_f32(const float* x, int64_t n, float* amax, void* stream) {
  VITTA_CHECK_ARG(x && amax && n > 0, VITTA_E_BADARG, "amax_f32: bad arguments");
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  amax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, amax);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_split_f16(const float* src, void* hi, void* lo, const float* amax, int R, int T, int Cc, int mode,
                    void* stream) {
  VITTA_CHECK_ARG(src && hi && lo && amax && R > 0 && T > 0 && Cc > 0 && (mode == 0 || mode == 1), VITTA_E_BADARG,
                  "split_f16: bad arguments");
  const int64_t n = (int64_t)R * T * Cc;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  split_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, static_cast<__half*>(hi),
                                                                       static_cast<__half*>(lo), amax, R, T, Cc, mode);
  VITTA_CHECK_LAUNCH();
  return 0;
}

int vitta_gemm_f16x3_ex(const float* A, int64_t lda, const float* a_amax, const void* Bhi, const void* Blo,
                        const float* b_amax, int64_t ldb, float* C, int64_t ldc, int64_t M, int N, int K,
                        const float* bias, const float* residual, int64_t ldr, int act, float* aux_out,
                        const float* row_scale, int64_t rows_per_group, int force_bn, void* stream) {
  VITTA_CHECK_ARG(a_amax && b_amax, VITTA_E_BADARG, "gemm_f16x3: amax scalars are required");
  return gemm_impl(A, lda, Bhi, Blo, ldb, C, ldc, M, N, K, bias, residual, ldr, act, aux_out, row_scale, rows_per_group,
                   force_bn, stream, a_amax, b_amax);
}

int vitta_conv2d_f16x3_ex(const float* X, const float* x_amax, int F, int H, int W, int Cin, const void* Whi,
                          const void* Wlo, const float* w_amax, int Cout, int KH, int KW, int stride, int pad, float* Y,
                          const float* bias, const float* residual, int force_bn, void* stream) {
  VITTA_CHECK_ARG(x_amax && w_amax, VITTA_E_BADARG, "conv2d_f16x3: amax scalars are required");
  return conv2d_impl(X, F, H, W, Cin, Whi, Wlo, Cout, KH, KW, stride, pad, Y, bias, residual, force_bn, stream, x_amax,
                     w_amax);
}

int vitta_conv2d_dgrad_f16x3(const float* dY, const float* dy_amax, int F, int Ho, int Wo, int Cout, const void* Wthi,
                             const void* Wtlo, const float* w_amax, int Cin, int KH, int KW, int stride, int pad, int H,
                             int W, float* dX, void* stream) {
  VITTA_CHECK_ARG(dy_amax && w_amax, VITTA_E_BADARG, "conv2d_dgrad_f16x3: amax scalars are required");
  return dgrad_impl(dY, F, Ho, Wo, Cout, Wthi, Wtlo, Cin, KH, KW, stride, pad, H, W, dX, stream, dy_amax, w_amax);
}

int vitta_split_block_elems(void) { return kSplitBlock; }

int vitta_split_multi(const VittaSplitTensor* tensors, const int32_t* block_start, int n_tensors, int total_blocks,
                      int f16, void* stream) {
  VITTA_CHECK_ARG(tensors && block_start && n_tensors > 0 && total_blocks > 0, VITTA_E_BADARG,
                  "split_multi: bad arguments");
  static_assert(sizeof(SplitEntry) == sizeof(VittaSplitTensor), "layout");
  const SplitEntry* t = reinterpret_cast<const SplitEntry*>(tensors);
  cudaStream_t st = (cudaStream_t)stream;
  if (f16) {
    split_multi_zero_kernel<<<(n_tensors + 255) / 256, 256, 0, st>>>(t, n_tensors);
    VITTA_CHECK_LAUNCH();
    split_multi_amax_kernel<<<(unsigned)total_blocks, 256, 0, st>>>(t, block_start, n_tensors);
    VITTA_CHECK_LAUNCH();
    split_multi_kernel<true><<<(unsigned)total_blocks, 256, 0, st>>>(t, block_start, n_tensors);
  } else {
    split_multi_kernel<false><<<(unsigned)total_blocks, 256, 0, st>>>(t, block_start, n_tensors);
  }
  VITTA_CHECK_LAUNCH();
  return 0;
}

// BN-folded inference convolution on the fp16 split (the per-step evaluation forward): Y = [relu](conv(X, W') + bias
// [+ residual]) with W' = k * W and bias = beta - mean * k prepared by vitta_split_multi / vitta_bn_fold_bias_multi;
// optionally accumulates max|Y| into *y_amax (zero-initialised by the caller) for the convolution that consumes Y.
int vitta_conv2d_f16x3_infer(const float* X, const float* x_amax, int F, int H, int W, int Cin, const void* Whi,
                             const void* Wlo, const float* w_amax, int Cout, int KH, int KW, int stride, int pad,
                             float* Y, const float* bias, const float* residual, int relu, float* y_amax, void* stream) {
  VITTA_CHECK_ARG(x_amax && w_amax, VITTA_E_BADARG, "conv2d_f16x3_infer: amax scalars are required");
  return conv2d_impl(X, F, H, W, Cin, Whi, Wlo, Cout, KH, KW, stride, pad, Y, bias, residual, relu ? kConvRelu : 0, stream,
                     x_amax, w_amax, y_amax);
}

int vitta_bn_fold_bias_multi(const VittaFoldBias* table, int n, void* stream) {
  VITTA_CHECK_ARG(table && n > 0, VITTA_E_BADARG, "bn_fold_bias_multi: bad arguments");
  static_assert(sizeof(FoldBiasEntry) == sizeof(VittaFoldBias), "layout");
  bn_fold_bias_multi_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const FoldBiasEntry*>(table));
  VITTA_CHECK_LAUNCH();
  return 0;
}

// vitta_gemm_f16x3_ex that also accumulates max|C| into *c_amax (zero-initialised by the caller): the operand range of
// the fp16-split GEMM that consumes C (fc1 -> GELU -> fc2, and the GELU' data gradient -> fc1's gradients), for free.
int vitta_gemm_f16x3_amax(const float* A, int64_t lda, const float* a_amax, const void* Bhi, const void* Blo,
                          const float* b_amax, int64_t ldb, float* C, int64_t ldc, int64_t M, int N, int K,
                          const float* bias, const float* residual, int64_t ldr, int act, float* aux_out,
                          const float* row_scale, int64_t rows_per_group, float* c_amax, void* stream) {
  VITTA_CHECK_ARG(a_amax && b_amax && c_amax, VITTA_E_BADARG, "gemm_f16x3_amax: amax scalars are required");
  return gemm_impl(A, lda, Bhi, Blo, ldb, C, ldc, M, N, K, bias, residual, ldr, act, aux_out, row_scale, rows_per_group, 0,
                   stream, a_amax, b_amax, c_amax);
}

}  // extern "C"
