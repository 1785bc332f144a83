/*
 * vitta_b200 -- C-ABI of the B200 (sm_100a) kernels behind ViTTA's test-time-adaptation inner loop.
 *
 * The reference (wlin-at/ViTTA) is pure Python and has no FFI today; every entry point below cites the
 * reference code whose arithmetic it replaces.  The Python host side (vitta_b200/_lib.py) binds these
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; buffers are owned by the caller
 *     (PyTorch allocations) and only borrowed for the duration of the call;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 on success, a positive cudaError_t value on a CUDA failure, a negative VITTA_E_*
 *     on bad arguments.  Nothing throws or exits across the ABI.  vitta_last_error() gives a message;
 *   - all arithmetic is IEEE fp32 ("f32"), statistics are merged with Chan's parallel formula;
 *   - "rows x C, channels-last" means element (r, c) at r*C + c.  TANet activations are stored NHWC
 *     (rows = frame*H*W + h*W + w); Video-Swin tokens are (B, D, H, W, C) already.
 */
#ifndef VITTA_B200_H
#define VITTA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITTA_E_BADARG (-1)
#define VITTA_E_ALIGN (-2)
#define VITTA_E_UNSUPPORTED (-3)

#define VITTA_REG_L1 0  /* reg_type 'l1_loss'  utils/norm_stats_utils.py:538 */
#define VITTA_REG_MSE 1 /* reg_type 'mse_loss' utils/norm_stats_utils.py:536 */
#define VITTA_REG_KLD 2 /* reg_type 'kld'      utils/norm_stats_utils.py:8-16,540 */

int vitta_version(void);
const char* vitta_last_error(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int vitta_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * K1  per-channel spatio-temporal statistics (partials)
 *   replaces: feature.view().permute().contiguous(); output.mean((0,2,3,4));
 *             output.permute(1,0,2,3,4).contiguous().view(c,-1).var(1, unbiased=False)
 *             utils/norm_stats_utils.py:188-193, 222-236, 242-243 (and :93-95 in ComputeNormStatsHook)
 *
 * A feature is described as (outer O, channels C, inner I): element (o,c,i) at (o*C + c)*I + i.
 *   BN2d output (N*T, C, H, W) contiguous:      O = N*T, I = H*W
 *   BN3d output (N, C, T, H, W) contiguous:     O = N,   I = T*H*W
 *   channels-last / LayerNorm output (rows, C): O = rows, I = 1
 * The kernel reads the feature exactly once and writes, per chunk e of the reduction domain and per
 * channel, a (mean, M2) pair: part[(e*C + c)*2 + {0,1}].  vitta_stats_chunking() gives the number of
 * chunks and the element count of each.  Chunks never straddle a "frame" of
 * `frame_rows` reduction elements (frame_rows = H*W for per-frame layouts, or O*I for none) so the same
 * partials can also give per-frame pooled sums.
 * ---------------------------------------------------------------------------------------------- */
typedef struct VittaChunking {
  int32_t chunk_rows;       /* elements per channel in a full chunk */
  int32_t chunks_per_frame; /* chunks per frame; the last chunk of every frame may be ragged */
  int64_t frame_rows;       /* reduction elements per channel per frame */
  int32_t n_entries;        /* total number of chunks = (mean, M2) entries per channel */
  int32_t reserved;
} VittaChunking;
/* `frames` only matters for I == 1 (must divide O; pass 1 when there is no frame structure). */
int vitta_stats_chunking(int64_t O, int C, int64_t I, int64_t frames, VittaChunking* out);
int vitta_stats_partial(const float* x, int64_t O, int C, int64_t I, int64_t frames, float* part, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  merge + EMA + alignment loss + backward coefficients, all hooked layers in ONE launch
 *   replaces: MovingAverageTensor.update / AverageMeterTensor.update   utils/utils_.py:190-211
 *             compute_regularization / compute_kld                     utils/norm_stats_utils.py:8-16,531-542
 *             and prepares the closed-form backward of both (autograd in the reference).
 *
 * Layer table entry (host fills it once, uploads it; all offsets are in floats into the arenas).
 * For entry e of layer l the per-channel pair is  part[part_off + e*entry_stride + c*2 + {0,1}]  with element
 * count  counts ? counts[cnt_off + e*cnt_stride] : min(chunk_rows, frame_rows - (e % chunks_per_frame)*chunk_rows).
 * ---------------------------------------------------------------------------------------------- */
typedef struct VittaLayerDesc {
  int32_t C;                /* channels */
  int32_t n_entries;        /* partial entries to merge (chunks, or ranks after the all-gather) */
  int32_t chunk_rows;       /* elements per channel in a full chunk */
  int32_t chunks_per_frame; /* chunks per frame (ragged last chunk of each frame) */
  int64_t frame_rows;       /* reduction elements per channel per frame */
  int64_t part_off;         /* offset of this layer's (mean,M2) pairs in `part` */
  int64_t entry_stride;     /* floats between consecutive entries (2*C locally; 2*sum(C) after an all-gather) */
  int64_t cnt_off;          /* counts[cnt_off + e*cnt_stride] (only read when counts != NULL) */
  int64_t cnt_stride;
  int64_t ch_off;           /* offset of this layer's channel vectors in every per-channel arena */
  int32_t reg_type;         /* VITTA_REG_* */
  int32_t has_source;       /* 0: statistics only (ComputeNormStatsHook), no EMA/loss/coefficients */
  float w_new;              /* meter: avg = w_new*stat + w_old*avg.detach()  (EMA: alpha, 1-alpha;      */
  float w_old;              /*        AverageMeterTensor: n/count_new, count_old/count_new)              */
} VittaLayerDesc;

/* Per-channel arenas (length = sum of C over layers): src_mean, src_var (read), ema_mean, ema_var
 * (read-modify-write), batch_mean, batch_var (write), coef_a, coef_b (write):
 *     dLoss_l/dy[.., c, ..] = coef_a[c] + coef_b[c] * (y - batch_mean[c])     (SURVEY.md section 8a row a5;
 *     the centred form keeps fp32 accuracy when |mean| >> std)
 * loss[l] receives r_feature of layer l; loss[n_layers] their sum (summed in layer order by the last CTA
 * to finish); loss[n_layers+1] is an int32 ticket the caller zero-initialises once (self-resetting); the rest of the
 * buffer (vitta_stats_finalize_loss_floats(n_layers, max_channels) floats in total) holds per-CTA loss partials.
 * max_channels = the largest C in the table (sizes the grid: one CTA per 32 channels per layer).
 * merge_only != 0: only merge entries and write (mean, M2) pairs to merged[(ch_off + c)*2 + {0,1}] and
 * the count (as int32) to merged_counts[l] -- the per-rank payload of the multi-GPU all-gather. */
int vitta_stats_finalize(const VittaLayerDesc* descs, int n_layers, const float* part, const int32_t* counts,
                         const float* src_mean, const float* src_var, float* ema_mean, float* ema_var,
                         float* batch_mean, float* batch_var, float* coef_a, float* coef_b, float* loss,
                         int merge_only, float* merged, int32_t* merged_counts, int max_channels, void* stream);
int64_t vitta_stats_finalize_loss_floats(int n_layers, int max_channels);

/* K3  standalone backward of the alignment loss of one layer (used by hooks on stock torch modules):
 *     gy[o,c,i] = (*gscale) * (coef_a[c] + coef_b[c] * (y[o,c,i] - mean[c])),  y = yscale[c]*x + yshift[c] when
 *     the two affine vectors are given (BatchNorm eval output recomputed from its saved input), else y = x.
 *     mean = the layer's slice of batch_mean.
 *   replaces: autograd of var/mean/permute/contiguous, utils/norm_stats_utils.py:242-253 */
int vitta_stats_inject(const float* x, const float* yscale, const float* yshift, const float* coef_a,
                       const float* coef_b, const float* mean, const float* gscale, float* gy, int64_t O, int C,
                       int64_t I, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4  fused BatchNorm(eval) [+ statistics partials] [+ residual (optionally through a second BN, with its
 *     own statistics)] [+ ReLU] [+ per-chunk column sums of the result], channels-last.
 *   replaces: bn -> forward hook -> relu / add -> relu in TemporalBottleneck.forward,
 *             models/tanet_models/temporal_module.py:88-104; adaptive_avg_pool2d in TAM.forward :50 and
 *             the ResNet avgpool (tanet.py:145).
 *   y  = (x - rm) * (w * rsqrt(rv + eps)) + b                      main branch, statistics on y
 *   r  = res (raw)  or  BN2(res)  (statistics on BN2(res))         optional
 *   out = relu?(y + r)
 *   pool_part[(e*C + c)] = sum over the rows of chunk e of out     optional (same chunking as K1), then
 *   pool_out[f*C + c]    = mean over the frame_rows rows of frame f of out
 * bn = {weight, bias, running_mean, running_var} each (C,).
 * ---------------------------------------------------------------------------------------------- */
typedef struct VittaBN {
  const float* weight;
  const float* bias;
  const float* running_mean;
  const float* running_var;
  float eps;
} VittaBN;

int vitta_bn_act_fwd(const float* x, VittaBN bn, const float* res, const VittaBN* res_bn, int relu, float* out,
                     float* part_main, float* part_res, float* pool_part, float* pool_out, int64_t frames,
                     int64_t frame_rows, int C, void* stream);

/* Backward of K4 in one pass.  Reads gout, x (and res if res_bn), recomputes y; writes gx (and gres):
 *   gpre = gout * (out > 0)           [out recomputed]   (+ gpool[frame, c] / frame_rows: gpool is the gradient of
 *                                                         pool_out, spread over every row of the frame; applied
 *                                                         to `out`, i.e. before the ReLU mask)
 *   gy   = gpre + gs_main * (a[c] + b[c]*(y - mean[c]))      gx   = gy * k[c]
 *   gr   = gpre (+ gs_res*(a2[c] + b2[c]*(r - mean2[c])) and gres = gr*k2[c] when res_bn)
 *   gw[c] += sum gy * xhat, gb[c] += sum gy  (and the same for the residual BN), accumulated into the
 *   given gradient vectors deterministically (per-chunk partials in `ws`, last CTA reduces).
 * ws: workspace of vitta_bn_act_bwd_ws_floats() floats, zero-initialised ONCE by the caller (self-resetting).
 *   replaces: autograd of the chain above + of the hook statistics. */
int64_t vitta_bn_act_bwd_ws_floats(int64_t frames, int64_t frame_rows, int C);
int vitta_bn_act_bwd(const float* gout, const float* gpool, const float* x, VittaBN bn, const float* res,
                     const VittaBN* res_bn, int relu, const float* coef_a, const float* coef_b, const float* mean_main,
                     const float* gs_main, const float* coef_a2, const float* coef_b2, const float* mean_res,
                     const float* gs_res, float* gx, float* gres, float* gw, float* gb, float* gw2, float* gb2,
                     float* ws, int64_t frames, int64_t frame_rows, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5  TAM temporal stencil, channels-last.  x, out: (N, T, HW, C); kern: (N, 3, C); act: (N, T, C)
 *   out[n,t,p,c] = sum_k kern[n,k,c] * act[n,t+k-1,c] * x[n,t+k-1,p,c]   (zero padded in t)
 *   replaces: new_x * local_activation; F.conv2d(groups=N*C) and both permute().contiguous() copies,
 *             models/tanet_models/temporal_module.py:47-48,56-63
 * Backward: gx, and per-chunk partial correlations dpart[n, chunk, t, k, c] = sum_p gout[n,t-k+1,p,c]*x[n,t,p,c]
 *   (chunks = vitta_tam_num_chunks) from which d kern and d act follow on (N,C)-sized data. */
int vitta_tam_fwd(const float* x, const float* kern, const float* act, float* out, int N, int T, int64_t HW, int C,
                  void* stream);
int vitta_tam_num_chunks(int64_t HW, int C);
/* *_amax variants (opt-in f16x3 path): the same kernels, additionally *amax = max(*amax, max|tensor written|) for the
 * tensors the next fp16-split GEMM / convolution consumes (out; gx and gres), through one integer atomic per warp --
 * the producer touches every element anyway, so the separate vitta_amax_f32 pass over the operand disappears.
 * The caller zero-initialises the scalars.  Default kernels are untouched (separate template instantiations). */
int vitta_bn_act_fwd_amax(const float* x, VittaBN bn, const float* res, const VittaBN* res_bn, int relu, float* out,
                          float* part_main, float* part_res, float* pool_part, float* pool_out, int64_t frames,
                          int64_t frame_rows, int C, float* amax_out, void* stream);
int vitta_bn_act_bwd_amax(const float* gout, const float* gpool, const float* x, VittaBN bn, const float* res,
                          const VittaBN* res_bn, int relu, const float* coef_a, const float* coef_b,
                          const float* mean_main, const float* gs_main, const float* coef_a2, const float* coef_b2,
                          const float* mean_res, const float* gs_res, float* gx, float* gres, float* gw, float* gb,
                          float* gw2, float* gb2, float* ws, int64_t frames, int64_t frame_rows, int C, float* amax_gx,
                          float* amax_gres, void* stream);
int vitta_tam_fwd_amax(const float* x, const float* kern, const float* act, float* out, int N, int T, int64_t HW, int C,
                       float* amax_out, void* stream);
int vitta_tam_bwd(const float* gout, const float* x, const float* kern, const float* act, float* gx, float* dpart,
                  int N, int T, int64_t HW, int C, void* stream);
/* After vitta_tam_bwd: sums dpart over its row chunks (fixed order) and contracts it with act / kern:
 *   gkern[n,k,c] = sum_t act[n,t,c] * D[n,t,k,c],  gact[n,t,c] = sum_k kern[n,k,c] * D[n,t,k,c].
 * One launch, parallel over (channel tile, frame, video) x 8 chunk slices; dpart is consumed (its chunk-0 slots are
 * reused as scratch for act * D); tickets: vitta_tam_bwd_finish_tickets() ints, zeroed once by the caller (every launch
 * leaves them at zero), one per (video, tile of 128 channels) -- the last frame-CTA of a pair adds over the frames. */
int vitta_tam_bwd_finish_tickets(void);
int vitta_tam_bwd_finish(float* dpart, const float* kern, const float* act, float* gkern, float* gact, int* tickets, int N,
                         int T, int nch, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K10 prediction consistency, forward + gradient in one launch.  preds (B, V, K) logits.
 *   loss = (1/V) sum_v sum_{b,k} | softmax(preds[b,v]) - mean_v softmax |, mean NOT detached
 *   replaces: compute_pred_consis, utils/pred_consistency_utils.py:15-31 */
int vitta_pred_consis(const float* preds, int B, int V, int K, float* loss, float* grad, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K11 multi-tensor SGD with momentum and weight decay (torch.optim.SGD semantics, dampening 0, no nesterov)
 *   d = g + wd*p;  buf = first ? d : mom*buf + d;  p -= lr*buf
 *   replaces: torch.optim.SGD(model.parameters(), lr, momentum, weight_decay).step(), corpus/basics.py:559-560,671
 * Tensor table entries live in device memory.  Every CTA updates one block of vitta_sgd_block_elems()
 * elements of one tensor: block_start[i] (device, int32, ascending) is the first block of tensor i and
 * total_blocks the grid size.  first_step != 0 initialises the momentum buffers (buf = d).
 * grad_scale multiplies g before use (1/world_size after a summed all-reduce). */
typedef struct VittaSgdTensor {
  float* p;
  const float* g;
  float* buf;
  int64_t n;
} VittaSgdTensor;
int vitta_sgd_block_elems(void);
int vitta_sgd_step(const VittaSgdTensor* tensors, const int32_t* block_start, int n_tensors, int total_blocks,
                   float lr, float momentum, float weight_decay, int first_step, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5b  TAM gate networks (csrc/tam_gate.cu): the G and L branches of the Temporal Adaptive Module on the pooled
 *      activation p (N, T, C), BatchNorm1d layers in eval mode -- replaces ~45 eager launches per TAM and direction
 *      (reference models/tanet_models/temporal_module.py:27-41,49-55).
 *   G: kern[n, :, c] = softmax(W2 . relu(BN1(W1 . p[n, :, c])))       W1 (2T, T), W2 (3, 2T)            -> (N, 3, C)
 *   L: act[n, t, :]  = sigmoid(Wb . relu(BN2(conv1d_k3_pad1(Wa, p)[n, t, :])))   Wa (C/4, C, 3), Wb (C, C/4) -> (N, T, C)
 *   fwd also writes pre (N*T, C/4), the L hidden layer before BN2 (saved for the backward).
 *   bwd: given gkern, gact returns gp (N, T, C) and the gradients of W1, BN1 (w, b), W2, Wa, BN2 (w, b), Wb (assigned);
 *        gpre / ghm (N*T, C/4) are scratch; ws: vitta_tam_gate_bwd_ws_floats floats, zeroed once by the caller.
 *   2 <= T <= 16, C % 4 == 0, p and Wa 16-byte aligned; every reduction runs in a fixed order.
 *   Launches: 2 forward (L hidden layer | G branch;  L output layer) and 2 backward (G branch | dWb | hidden-layer
 *   gradient;  dWa | gradient of p | BN2 gradients): stages whose inputs are ready run as CTA roles of one grid. */
int vitta_tam_gate_fwd(const float* p, const float* W1, VittaBN bn1, const float* W2, const float* Wa, VittaBN bn2,
                       const float* Wb, float* kern, float* act, float* pre, int N, int T, int C, void* stream);
int64_t vitta_tam_gate_bwd_ws_floats(int N, int T, int C);
int vitta_tam_gate_bwd(const float* p, const float* W1, VittaBN bn1, const float* W2, const float* Wa, VittaBN bn2,
                       const float* Wb, const float* act, const float* pre, const float* gkern, const float* gact,
                       float* gp, float* gW1, float* gbn1w, float* gbn1b, float* gW2, float* gWa, float* gbn2w,
                       float* gbn2b, float* gWb, float* gpre, float* ghm, float* ws, int N, int T, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stem of the ResNet-50 trunk (reference models/tanet_models/tanet.py:129: torchvision conv1 7x7/2 pad 3 on the 3-channel
 * frames, bn1, relu, maxpool 3x3/2 pad 1) -- csrc/stem.cu, csrc/gemm_tf32.cu.
 *   vitta_stem_pack: image X (F, 3, H, W) contiguous -> XP (F, H+6, W+6, 4): zero border of 3 pixels, 4th channel 0.
 *   vitta_stem_pack_weight: conv1 weight (64, 3, 7, 7) -> tf32 hi / lo operands [64][7][8][4] (224 floats per filter).
 *   vitta_stem_conv_tf32x3: Y (F, H/2, W/2, 64) channels-last = conv1(X) on the tcgen05 3xTF32 kernel; the A operand is
 *     read straight from XP through a TMA tensor map with overlapping rows (no im2col buffer).  H, W even.
 *   vitta_bn_relu_pool_fwd: out (F, Ho, Wo, C) = maxpool3x3s2p1(relu(BN_eval(x))), x (F, H, W, C) channels-last, plus one
 *     byte per output element naming the window position (3*dh + dw) of the FIRST maximum in scan order (PyTorch's
 *     max_pool2d_with_indices rule).  C % 4 == 0.
 *   vitta_bn_relu_pool_bwd: gx = dL/dx given gpool = dL/dout; gw / gb = BN weight / bias gradients (assigned, reduced in a
 *     fixed order).  ws: vitta_bn_relu_pool_bwd_ws_floats(C) floats, zero-initialised once by the caller. */
int vitta_stem_pack(const float* x, float* xp, int F, int H, int W, void* stream);
int vitta_stem_pack_weight(const float* w, float* hi, float* lo, void* stream);
int vitta_stem_conv_tf32x3(const float* XP, int F, int H, int W, const float* Whi, const float* Wlo, float* Y,
                           void* stream);
/* weight gradient of the stem convolution: dW (64, 3, 7, 7) from the packed image XP and dY (F, H/2, W/2, 64) channels-last;
 * exact fp32 FFMA (no operand split), per-CTA partials in ws (vitta_stem_wgrad_ws_floats floats), summed in a fixed order */
int64_t vitta_stem_wgrad_ws_floats(void);
int vitta_stem_wgrad(const float* XP, const float* dY, float* dW, float* ws, int F, int H, int W, void* stream);
int vitta_bn_relu_pool_fwd(const float* x, VittaBN bn, float* out, uint8_t* code, int F, int H, int W, int C,
                           void* stream);
/* ... that also accumulates max|out| into the device scalar *amax_out (zeroed by the caller): the operand range of the
 * fp16-split convolutions that consume `out` (layer1.0's conv1 and downsample), so no separate range pass reads it again */
int vitta_bn_relu_pool_fwd_amax(const float* x, VittaBN bn, float* out, uint8_t* code, int F, int H, int W, int C,
                                float* amax_out, void* stream);
int64_t vitta_bn_relu_pool_bwd_ws_floats(int C);
int vitta_bn_relu_pool_bwd(const float* gpool, const uint8_t* code, const float* x, VittaBN bn, float* gx, float* gw,
                           float* gb, float* ws, int F, int H, int W, int C, void* stream);

/* Multi-tensor weight preparation (once per optimizer step, all weights, both operand forms): the same hi/lo pieces as
 * vitta_split_tf32 / vitta_split_f16 (bit-identical), in 1 (tf32) or 3 (fp16: zero + amax + split) launches.  tensors /
 * block_start are device arrays; block_start[i] is the first CTA of tensor i when every CTA takes
 * vitta_split_block_elems() destination elements, total_blocks the grid size.  src_tap_inner: the source is the
 * contiguous conv weight [R][Cc][T] instead of [R][T][Cc].  f16: entries with compute_amax write *amax = max|src| first
 * (entries of the same weight share the scalar; exactly one of them computes it); hi / lo are fp16 then, fp32 otherwise. */
typedef struct VittaSplitTensor {
  const float* src;
  void* hi;
  void* lo;
  float* amax;
  int32_t R, T, Cc, mode;
  int32_t src_tap_inner, compute_amax;
  int64_t n;
  const float* fold_w;   /* optional eval-mode BatchNorm folded into the rows: row r scaled by fold_w[r] / sqrt(fold_rv[r] + */
  const float* fold_rv;  /*   fold_eps) before the split (null: none) -- the weights of the BN-folded inference convolutions */
  float fold_eps;
  int32_t reserved;
} VittaSplitTensor;
typedef struct VittaFoldBias {
  const float *w, *b, *rm, *rv;   /* BatchNorm weight, bias, running_mean, running_var */
  float* out;                     /* out[c] = b[c] - rm[c] * w[c] / sqrt(rv[c] + eps) */
  float eps;
  int32_t C;
} VittaFoldBias;
/* folded biases of all eval-mode BatchNorm layers behind convolutions, one launch (table in device memory) */
int vitta_bn_fold_bias_multi(const VittaFoldBias* table, int n, void* stream);
/* BN-folded inference convolution (fp16 split): Y = [relu](conv(X, W') + bias [+ residual]); *y_amax (optional, zeroed by
 * the caller) receives max|Y|.  Used by the per-step evaluation forward (reference corpus/basics.py:691-713). */
int vitta_conv2d_f16x3_infer(const float* X, const float* x_amax, int F, int H, int W, int Cin, const void* Whi,
                             const void* Wlo, const float* w_amax, int Cout, int KH, int KW, int stride, int pad,
                             float* Y, const float* bias, const float* residual, int relu, float* y_amax, void* stream);
/* vitta_gemm_f16x3_ex + max|C| accumulated into *c_amax (zeroed by the caller): the range of the next fp16-split GEMM. */
int vitta_gemm_f16x3_amax(const float* A, int64_t lda, const float* a_amax, const void* Bhi, const void* Blo,
                          const float* b_amax, int64_t ldb, float* C, int64_t ldc, int64_t M, int N, int K,
                          const float* bias, const float* residual, int64_t ldr, int act, float* aux_out,
                          const float* row_scale, int64_t rows_per_group, float* c_amax, void* stream);
int vitta_split_block_elems(void);
int vitta_split_multi(const VittaSplitTensor* tensors, const int32_t* block_start, int n_tensors, int total_blocks,
                      int f16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K6/K8  fp32-accurate GEMM and implicit-GEMM convolution on the tcgen05 tensor cores (3xTF32 split, fp32
 *        accumulation in TMEM, TMA-fed, persistent).  See csrc/gemm_tf32.cu.
 *   replaces: cuDNN / cuBLAS fp32 calls behind nn.Conv2d (torchvision Bottleneck convs inside TemporalBottleneck,
 *             models/tanet_models/temporal_module.py:85-106) and nn.Linear (qkv / proj / fc1+GELU / fc2 / reduction,
 *             models/videoswintransformer_models/swin_transformer.py:24-35,130-132,160-167,287,311).
 *
 * vitta_split_tf32: weight preparation, once per optimizer step.  src is [R][T][Cc] (R output channels, T filter
 *   taps, Cc input channels -- i.e. a conv weight in channels_last memory, or a Linear weight with T = 1).
 *   mode 0: hi/lo in the same order (forward operand).  mode 1: hi/lo[c][T-1-t][r] = split(src[r][t][c]) (the operand
 *   of the data-gradient pass: transposed, filter rotated by 180 degrees).   hi = rna_tf32(x), lo = x - hi.
 * vitta_gemm_tf32x3: C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) then act (0 none, 1 exact GELU) then (+ residual[M,N]).
 *   Row-major, leading dimensions in floats; lda, ldb multiples of 4; K tail and ragged M/N handled (TMA zero fill).
 *   force_bn: 0 = choose the N tile automatically, else 64 / 128 / 256; | VITTA_GEMM_FORCE_SS keeps the split A tile in
 *   shared memory, | VITTA_GEMM_FORCE_TS holds it in tensor memory (N tiles of 64 / 128: `tcgen05.mma [d], [a_tmem], b`).
 * vitta_gemm_set_operand_form: process-wide version of those flags (0 automatic, 1 shared memory, 2 tensor memory), also
 *   honoured by the weight-gradient kernel; for timing / cross-checking the forms (they differ only in the fp32
 *   summation order of the accumulator chains).
 * vitta_conv2d_tf32x3: Y[F,Ho,Wo,Cout] = conv2d(X[F,H,W,Cin], W[Cout][KH][KW][Cin], stride, pad) (+ bias), NHWC,
 *   Cin multiple of 4.  Padding comes from the TMA out-of-bounds zero fill; no im2col buffer exists. */
#define VITTA_GEMM_FORCE_SS 0x1000
#define VITTA_GEMM_FORCE_TS 0x2000
/* vitta_conv2d_*_ex only, or-ed into force_bn: the epilogue ends with ReLU, applied after bias and residual -- the
 * BN-folded inference convolution  relu(conv(x, k*W) + b' [+ shortcut])  of the per-step evaluation forward. */
#define VITTA_CONV_RELU 0x4000
int vitta_gemm_set_operand_form(int form);
/* vitta_gemm_set_cta_pair(1): N tiles of 256 run as clusters of two CTAs issuing tcgen05.mma.cta_group::2 (M = 256 per
 * pair; each CTA loads half of the weight tile, halving the L2 -> SM weight traffic that bounds the wide layers).
 * Applies to the tf32 and fp16 kernels (forward, data gradient, Linear).  Default 0; opt-in until validated on hardware
 * (round 1 wrote it after the GPU budget was spent -- DESIGN.md section 9). */
int vitta_gemm_set_cta_pair(int on);
int vitta_split_tf32(const float* src, float* hi, float* lo, int R, int T, int Cc, int mode, void* stream);
int vitta_gemm_tf32x3(const float* A, int64_t lda, const float* Bhi, const float* Blo, int64_t ldb, float* C,
                      int64_t ldc, int64_t M, int N, int K, const float* bias, const float* residual, int64_t ldr,
                      int act, int force_bn, void* stream);
/* Extended epilogue (Swin MLP / attention projections):
 *   v = acc + bias;  aux_out[m,n] = v (optional: the pre-activation, leading dimension ldc);
 *   act 1: v = GELU(v);  act 2: v = v * GELU'(residual[m,n]) (residual = saved pre-activation; backward of fc1's GELU);
 *   act 4: v = GELU(v) and aux_out[m,n] = GELU'(pre-activation) INSTEAD of the pre-activation (one erf serves both);
 *   act 5: v = v * residual[m,n] (residual = the derivative saved by act 4: the backward multiplies, no erf / exp);
 *   v *= row_scale[m / rows_per_group] (optional: DropPath's per-sample factor, swin_transformer.py:210,246,252);
 *   act 0, 1, 4: v += residual[m,n] (the block's shortcut, swin_transformer.py:266,272). */
int vitta_gemm_tf32x3_ex(const float* A, int64_t lda, const float* Bhi, const float* Blo, int64_t ldb, float* C,
                         int64_t ldc, int64_t M, int N, int K, const float* bias, const float* residual, int64_t ldr,
                         int act, float* aux_out, const float* row_scale, int64_t rows_per_group, int force_bn,
                         void* stream);
int vitta_conv2d_tf32x3(const float* X, int F, int H, int W, int Cin, const float* Whi, const float* Wlo, int Cout,
                        int KH, int KW, int stride, int pad, float* Y, const float* bias, int force_bn, void* stream);
/* Same, plus `residual` (shape of Y) added in the epilogue.  Used by the data-gradient pass of a convolution whose input
 * also feeds the block's shortcut: dX = dgrad(dY) + dShortcut in one store instead of a separate accumulation pass
 * (autograd's gradient sum of the residual split in TemporalBottleneck.forward, temporal_module.py:85-104). */
int vitta_conv2d_tf32x3_ex(const float* X, int F, int H, int W, int Cin, const float* Whi, const float* Wlo, int Cout,
                           int KH, int KW, int stride, int pad, float* Y, const float* bias, const float* residual,
                           int force_bn, void* stream);

/* Data gradient of a STRIDED convolution: dX[F,H,W,Cin] = conv2d_backward_input(dY[F,Ho,Wo,Cout], W), stride >= 1.
 * Wthi / Wtlo: the weight split with mode 1 ([Cin][rotated tap][Cout]).  Input pixels are processed per residue class
 * modulo the stride: each class is a small dense stride-1 convolution over dY with only the filter taps that reach it
 * (no zero-insertion, no wasted MMAs), stored straight to its strided positions of dX; classes no tap reaches are zeroed.
 *   replaces: autograd's convolution_backward input branch (cuDNN dgrad) for the stride-2 3x3 / 1x1 convolutions of
 *             layer2-4.0 (torchvision Bottleneck v1.5 inside TemporalBottleneck, temporal_module.py:85-106). */
int vitta_conv2d_dgrad_tf32x3(const float* dY, int F, int Ho, int Wo, int Cout, const float* Wthi, const float* Wtlo,
                              int Cin, int KH, int KW, int stride, int pad, int H, int W, float* dX, void* stream);

/* ---- fp16 operand split (kind::f16) -- same contractions, same ~2^-21 per-product error, half the tensor-pipe work ----
 * x*s = hi + lo with hi = fp16(x*s), lo = fp16(x*s - hi); s = the power of two that puts `amax` just below 2^14, where
 * `amax` is a DEVICE scalar holding any upper bound of max|x| over the tensor (exact or up to ~2^10 too large: the
 * residual keeps full precision for elements within 2^17 of the bound).  Results are rescaled by 1/(s_a*s_b) in the
 * epilogue (exact).  Parity budget: tools/split_numerics.py emulates the scheme through forward, data and weight
 * gradients of the whole adaptation step on the CPU; bf16 pieces fail the 1e-4 budget, these pass
 * with the margin of the tf32 split.
 *   vitta_amax_f32: *amax = max(*amax, max|x|) (integer atomic on the bit pattern; zero-initialise *amax first).
 *   vitta_split_f16: weight preparation like vitta_split_tf32 (same modes) into two fp16 arrays, scaled by *amax's s.
 *   vitta_gemm_f16x3_ex / vitta_conv2d_f16x3_ex / vitta_conv2d_dgrad_f16x3: the tf32 entry points with fp16 weight pieces
 *     (ldb in fp16 elements, multiple of 8) and the two amax scalars.  Shared-memory A form; a stage holds 64 K elements.
 *   vitta_conv2d_wgrad_f16x3: the weight gradient with both activations split in-kernel (dY^T as packed fp16 pairs in
 *     tensor memory, X converted in place into 16-bit MN-major atoms); amax scalars of X and dY.
 * Round-1 status: compiled, exported and covered by opt-in GPU tests (VITTA_TEST_F16X3=1); the adaptation step still
 * runs the tf32 kernels until the producers emit amax. */
int vitta_amax_f32(const float* x, int64_t n, float* amax, void* stream);
int vitta_split_f16(const float* src, void* hi, void* lo, const float* amax, int R, int T, int Cc, int mode,
                    void* stream);
int vitta_gemm_f16x3_ex(const float* A, int64_t lda, const float* a_amax, const void* Bhi, const void* Blo,
                        const float* b_amax, int64_t ldb, float* C, int64_t ldc, int64_t M, int N, int K,
                        const float* bias, const float* residual, int64_t ldr, int act, float* aux_out,
                        const float* row_scale, int64_t rows_per_group, int force_bn, void* stream);
int vitta_conv2d_f16x3_ex(const float* X, const float* x_amax, int F, int H, int W, int Cin, const void* Whi,
                          const void* Wlo, const float* w_amax, int Cout, int KH, int KW, int stride, int pad, float* Y,
                          const float* bias, const float* residual, int force_bn, void* stream);
int vitta_conv2d_wgrad_f16x3(const float* X, const float* x_amax, const float* dY, const float* dy_amax, int F, int H,
                             int W, int Cin, int Cout, int KH, int KW, int stride, int pad, float* dW, int accumulate,
                             float* ws, void* stream);
/* ... and with the bias gradient dbias[co] = sum_pixels dY[., co] as a by-product (no column-sum pass over dY) */
int vitta_conv2d_wgrad_f16x3_bias(const float* X, const float* x_amax, const float* dY, const float* dy_amax, int F, int H,
                                  int W, int Cin, int Cout, int KH, int KW, int stride, int pad, float* dW, float* dbias,
                                  int accumulate, float* ws, void* stream);   /* same ws size and reduction as vitta_conv2d_wgrad_tf32x3 */
int vitta_conv2d_dgrad_f16x3(const float* dY, const float* dy_amax, int F, int Ho, int Wo, int Cout, const void* Wthi,
                             const void* Wtlo, const float* w_amax, int Cin, int KH, int KW, int stride, int pad, int H,
                             int W, float* dX, void* stream);

/* Weight gradient of the same convolution: dW[Cout][Cin][KH][KW] (contiguous NCHW, the layout of nn.Conv2d.weight)
 *   (+)= sum over output pixels of dY[F,Ho,Wo,Cout] (x) X[F,H,W,Cin] shifted by the filter tap.
 * Split-K over pixel ranges on the tcgen05 tensor cores (both operands MN-major, split hi/lo in-kernel); the K splits
 * are summed in a fixed order (deterministic).  ws: vitta_conv2d_wgrad_ws_floats() floats of scratch (no init needed).
 *   replaces: autograd's convolution_backward weight branch (cuDNN wgrad) for the layers above. */
int64_t vitta_conv2d_wgrad_ws_floats(int F, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
/* Host-only: the split-K plan (out[4] = N tile, base items, K splits, 32-pixel stages).  The persistent grid walks
 * base items x splits work items round-robin over the SMs; the split count minimises ceil(items / SMs) * (stages per item +
 * a fixed per-item cost) -- never a partial extra pass with most SMs idle. */
int vitta_conv2d_wgrad_plan(int F, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int f16,
                            int* out);
int vitta_conv2d_wgrad_tf32x3(const float* X, const float* dY, int F, int H, int W, int Cin, int Cout, int KH, int KW,
                              int stride, int pad, float* dW, int accumulate, float* ws, void* stream);


/* ------------------------------------------------------------------------------------------------
 * K9  LayerNorm over the channel axis of a (rows, C) token matrix with the statistics hook fused in.
 *   replaces: nn.LayerNorm (norm1 / norm2 / downsample.norm / backbone.norm,
 *             models/videoswintransformer_models/swin_transformer.py:204,212,288,545) followed by the forward hook
 *             (utils/norm_stats_utils.py:222-243), and in "merge" mode the PatchMerging gather that feeds
 *             downsample.norm (swin_transformer.py:293-310).
 * Forward: y = (x - mean_r) * rstd_r * gamma + beta, mean/rstd saved per row; when `part` is given every chunk of
 *   vitta_ln_chunking(rows, C, 1) consecutive rows writes the per-channel (mean, M2) of y to part[(e*C + c)*2 + {0,1}]
 *   (same format as K1, consumed by vitta_stats_finalize).
 * gather != NULL: the logical row (b, d, h2, w2) of 4*Cin channels is read from x = (B, D, H, W, Cin) as
 *   [x(2h2,2w2), x(2h2+1,2w2), x(2h2,2w2+1), x(2h2+1,2w2+1)] (zeros outside H x W); rows = B*D*ceil(H/2)*ceil(W/2).
 * Backward: gy' = gy + gscale*(coef_a + coef_b*(y - coef_mean))   [hook term, optional, y recomputed]
 *   gx = rstd*(gy'*gamma - mean_C(gy'*gamma) - xhat*mean_C(gy'*gamma*xhat)) (+ gadd), dgamma += sum_r gy'*xhat,
 *   dbeta += sum_r gy'.  In merge mode gx is scattered back into the (B, D, H, W, Cin) layout.
 * ws: vitta_ln_bwd_ws_floats() floats, zero-initialised ONCE by the caller (self-resetting tickets).
 * ---------------------------------------------------------------------------------------------- */
typedef struct VittaLnGather {
  int32_t B, D, H, W, Cin;
} VittaLnGather;
int vitta_ln_chunking(int64_t rows, int C, int want_stats, VittaChunking* out);
int vitta_ln_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean, float* rstd,
                 float* part, int64_t rows, int C, const VittaLnGather* gather, void* stream);
int64_t vitta_ln_bwd_ws_floats(int64_t rows, int C);
int vitta_ln_bwd(const float* gy, const float* x, const float* gamma, const float* beta, const float* mean,
                 const float* rstd, const float* gadd, const float* coef_a, const float* coef_b, const float* coef_mean,
                 const float* gscale, float* gx, float* dgamma, float* dbeta, float* ws, int64_t rows, int C,
                 const VittaLnGather* gather, void* stream);
/* The same two entry points with max|y| / max|gx| accumulated into a zero-initialised device scalar (null: off): the
 * operand range of the fp16-split GEMMs that consume the LayerNorm output / the residual-stream gradient. */
int vitta_ln_fwd_amax(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean, float* rstd,
                      float* part, int64_t rows, int C, const VittaLnGather* gather, float* y_amax, void* stream);
int vitta_ln_bwd_amax(const float* gy, const float* x, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, const float* gadd, const float* coef_a, const float* coef_b,
                      const float* coef_mean, const float* gscale, float* gx, float* dgamma, float* dbeta, float* ws,
                      int64_t rows, int C, const VittaLnGather* gather, float* gx_amax, void* stream);

/* Small token-matrix helpers of the Swin path.
 *   vitta_colsum:      out[c] (+)= sum_r x[r, c]   (bias gradients of nn.Linear; deterministic two-stage sum;
 *                      ws: vitta_colsum_ws_floats() floats zero-initialised once)
 *   vitta_frame_mean:  out[f, c] = mean over the `rows` rows of frame f   (AdaptiveAvgPool3d of I3DHead,
 *                      models/videoswintransformer_models/i3d_head.py:66-67) and its backward
 *   vitta_patchify3d:  video (B, 3, T, H, W) -> (B*D*Hp*Wp, 3*pt*ph*pw) patch rows in nn.Conv3d weight order, so that
 *                      PatchEmbed3D.proj (swin_transformer.py:432,446) becomes one GEMM */
int64_t vitta_colsum_ws_floats(int64_t rows, int C);
int vitta_colsum(const float* x, int64_t rows, int C, float* out, int accumulate, float* ws, void* stream);
int vitta_frame_mean(const float* x, int64_t frames, int rows, int C, float* out, void* stream);
int vitta_frame_mean_bwd(const float* g, int64_t frames, int rows, int C, float* gx, void* stream);
int vitta_patchify3d(const float* video, int B, int T, int H, int W, int pt, int ph, int pw, float* out, void* stream);
/* out[r, :] = x[r, :] * scale[r / rows_per_group]: DropPath's per-sample factor applied to a gradient before the
 * weight-gradient GEMM (timm DropPath, call site swin_transformer.py:210). */
int vitta_row_scale(const float* x, const float* scale, int64_t rows, int64_t rows_per_group, int C, float* out,
                    void* stream);
/* ... with max|out| accumulated into a zero-initialised device scalar (null: off) */
int vitta_row_scale_amax(const float* x, const float* scale, int64_t rows, int64_t rows_per_group, int C, float* out,
                         float* out_amax, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7  Video-Swin 3-D (shifted-)window multi-head self-attention, head_dim 32.
 *   replaces: WindowAttention3D.forward (q*scale, q@k^T, + relative_position_bias_table[relative_position_index[:N,:N]],
 *             + shift mask (0 / -100), softmax, attn@v; models/videoswintransformer_models/swin_transformer.py:145-166)
 *             and the roll / window_partition / window_reverse / roll-back copies around it (:229-248), and
 *             compute_mask (:316-329) -- the mask is evaluated from region ids inside the kernel.
 *   qkv: (B, D, H, W, 3, heads, 32) = the output of nn.Linear(dim, 3*dim) on the normalised tokens, natural token order;
 *   bias_table: (prod(2*window-1), heads);  out: (B, D, H, W, heads*32) = the proj input, natural token order;
 *   lse: (B*windows, heads, N) log-sum-exp per query row (saved for the backward);
 *   window / shift: the CONFIGURED 3-vectors (e.g. {8,7,7}, {4,3,3} or {0,0,0}); they are clamped per dimension exactly
 *   as get_window_size (:71-84) does, and the bias index uses the configured window (relative_position_index[:N,:N]).
 *   D, H, W must be multiples of the clamped window (true for every 224x224 configuration); other volumes are zero-padded
 *   by the caller exactly as the reference pads them (:222-227: padding tokens carry the qkv bias, the shift mask is the
 *   padded volume's, the result is cropped) -- vitta_b200.ops_swin.SwinAttentionFn does this; the raw entry points
 *   return VITTA_E_UNSUPPORTED for a volume that is not a multiple.
 * Forward: tcgen05 -- S = QK^T (3xTF32) accumulates in TMEM, softmax in place, O += P V on kind::f16 with P and V as fp16
 *   hi / lo pairs (22 mantissa bits; K = 16 per MMA halves the MMA issue count that bounds the kernel).  qkv_amax: device
 *   scalar >= max|qkv| (as emitted by the qkv GEMM's epilogue, or vitta_amax_f32): the power-of-two scale of V's split.
 * Backward: dqkv has the layout of qkv (every element written once); dbias_table is ACCUMULATED into (zero it first).
 *   dS = P o (dO V^T - rowsum(dO o O)), dQ = scale dS K, dK = dS^T (scale Q), dV = P^T dO, dTable[rel(i,j)] += dS_ij.
 * ---------------------------------------------------------------------------------------------- */
int vitta_wmsa3d_fwd(const float* qkv, const float* qkv_amax, const float* bias_table, float* out, float* lse, int B, int D,
                     int H, int W, int heads, int head_dim, const int* window_host, const int* shift_host, float scale,
                     void* stream);
/* ws: vitta_wmsa3d_bwd_ws_floats() floats of scratch (D_i = dO_i . O_i per token and head; no init needed).
 * impl 0: tcgen05, every product on kind::f16 with fp16 hi / lo operand pairs (qkv_amax, dout_amax: device scalars >= max|qkv|,
 *         >= max|dout|: the power-of-two operand scales) -- a query-outer launch (dQ, dTable) and a key-outer launch (dK, dV);
 * impl 1: the exact-fp32 FFMA2 kernel (one CTA per window and head), kept as an on-device cross-check. */
int64_t vitta_wmsa3d_bwd_ws_floats(int B, int D, int H, int W, int heads);
int vitta_wmsa3d_bwd(const float* qkv, const float* qkv_amax, const float* bias_table, const float* out, const float* dout,
                     const float* dout_amax, const float* lse, float* dqkv, float* dbias_table, float* ws, int B, int D, int H,
                     int W, int heads, int head_dim, const int* window_host, const int* shift_host, float scale, int impl,
                     void* stream);
/* The attention entry points with max|out| / max|dqkv| accumulated into a zero-initialised device scalar (null: off). */
int vitta_wmsa3d_fwd_amax(const float* qkv, const float* qkv_amax, const float* bias_table, float* out, float* lse, int B,
                          int D, int H, int W, int heads, int head_dim, const int* window, const int* shift, float scale,
                          float* out_amax, void* stream);
/* Profiling aid: the forward kernel with time stamps.  trace = 15 x trace_cap records, zeroed by the caller: lane 0 of warp
 * w of CTA 0 writes (clock << 16 | warp << 8 | event id) into trace[w * trace_cap ...] at the hand-over points of the
 * pipeline (event ids: csrc/wmsa3d.cu, WMSA_TR; tools/wmsa_trace.py prints the timeline).  Results = vitta_wmsa3d_fwd. */
int vitta_wmsa3d_fwd_trace(const float* qkv, const float* qkv_amax, const float* bias_table, float* out, float* lse, int B,
                           int D, int H, int W, int heads, int head_dim, const int* window, const int* shift, float scale,
                           unsigned long long* trace, int trace_cap, void* stream);
int vitta_wmsa3d_bwd_amax(const float* qkv, const float* qkv_amax, const float* bias_table, const float* out,
                          const float* dout, const float* dout_amax, const float* lse, float* dqkv, float* dbias_table,
                          float* ws, int B, int D, int H, int W, int heads, int head_dim, const int* window, const int* shift,
                          float scale, int impl, float* dqkv_amax, void* stream);
/* Profiling aid for the backward: trace != null makes the following vitta_wmsa3d_bwd calls (impl 0) run their traced
 * instantiation -- time stamps of CTA 0 as in vitta_wmsa3d_fwd_trace, 2 launches x 16 warps x trace_cap zeroed records
 * (tools/wmsa_trace.py --bwd); null switches it off.  Results are unchanged. */
int vitta_wmsa3d_bwd_set_trace(unsigned long long* trace, int trace_cap);

/* ------------------------------------------------------------------------------------------------
 * View gathering + normalisation (the step before the hot path).
 *   replaces, for decoded frames at the target scale: container.get_batch(frame_indices) -> crop -> Stack ->
 *             ToTorchFormatTensor (/255) -> GroupNormalize (models/tanet_models/video_dataset.py:318-345,
 *             models/tanet_models/transforms.py:627-690) and the Swin pipeline's Normalize + FormatShape('NCTHW').
 *   frames: (F, H, W, 3) uint8 device; idx: (n_idx = V*T) int32 device, clamped to [0, F-1];
 *   mean3_host / std3_host: three HOST floats each on the [0, 1] scale (utils/opts.py:4-5);
 *   layout 0: out (V*T*3, out_h, out_w) (TANet loader);  layout 1: out (V, 3, T, out_h, out_w) (Swin loader).
 * ---------------------------------------------------------------------------------------------- */
int vitta_gather_normalize_u8(const uint8_t* frames, int F, int H, int W, const int32_t* idx, int n_idx, int crop_y,
                              int crop_x, int out_h, int out_w, const float* mean3_host, const float* std3_host, int layout,
                              int T, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-view random multi-scale crop + bilinear resize + normalisation.
 *   replaces: SubgroupWise_MultiScaleCrop_TANet.crop_scale_subgroup -- img.crop(box).resize((S, S), Image.BILINEAR) per
 *             frame of a temporal view (models/tanet_models/transforms.py:312-323; selected by corpus/basics.py:1238-1245)
 *             -> Stack -> ToTorchFormatTensor (/255) -> GroupNormalize (transforms.py:627-690).
 *   The resize is Pillow's 8-bit two-pass fixed-point resampler (Pillow 8.4.0, requirements.txt:37, libImaging/Resample.c):
 *   results are bit-exact with PIL before the float normalisation.
 *
 * vitta_resample_ksize / vitta_resample_coeffs_u8 are HOST-ONLY (no CUDA call): Pillow's precompute_coeffs +
 *   normalize_coeffs_8bpc for one axis.  bounds_host: [out_size][2] = (first source index + in_offset, count);
 *   kk_host: [out_size][slots] 22-bit fixed-point weights, zero padded; slots >= vitta_resample_ksize(in, out).
 * vitta_gather_crop_resize_normalize_u8: frames (F, H, W, 3) uint8 device; idx (n_idx = n_views * T) int32 device;
 *   boxes_host: n_views x (crop_w, crop_h, offset_w, offset_h) HOST ints, the tuple _sample_crop_size returns
 *   (validated against the frame); hbounds / hk: device tables [n_views][out_w][2] / [n_views][out_w][slots] built with
 *   in_offset = offset_w, vbounds / vk likewise for the rows; mean / std / layout / T / out as vitta_gather_normalize_u8.
 *   A table row describes ONE output position, so any subset of the rows of a resize is "resize, then crop": the
 *   reference's GroupScale + GroupCenterCrop path (transforms.py:46-52,170-183) uses the same entry point with the rows
 *   [left, left + S) of a whole-frame resize.  The kernel clamps every source coordinate into the frame, so foreign
 *   tables cannot read outside it.
 * ---------------------------------------------------------------------------------------------- */
int vitta_resample_ksize(int in_size, int out_size);
int vitta_resample_coeffs_u8(int in_size, int out_size, int in_offset, int slots, int32_t* bounds_host, int32_t* kk_host);
int vitta_gather_crop_resize_normalize_u8(const uint8_t* frames, int F, int H, int W, const int32_t* idx, int n_idx,
                                          const int32_t* boxes_host, int n_views, const int32_t* hbounds, const int32_t* hk,
                                          const int32_t* vbounds, const int32_t* vk, int slots, int out_h, int out_w,
                                          const float* mean3_host, const float* std3_host, int layout, int T, float* out,
                                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * The Video-Swin loader's resize: OpenCV's 8-bit INTER_LINEAR, bit exact (Swin side).
 *   replaces: Resize._resize_imgs -> mmcv.imresize -> cv2.resize(INTER_LINEAR) (models/videoswintransformer_models/
 *             transforms_backup.py:794-798, pipeline video_dataset.py:66-101) and Normalize -> mmcv.imnormalize_ (:1151-1166).
 * vitta_cv_linear_tables is HOST-ONLY: per output position the first tap and the two 11-bit weights of one axis
 *   (ofs_host [dst], w_host [dst][2]); horizontal != 0 applies OpenCV's left / right border rule, the vertical pass clips
 *   row indices in the kernel instead.
 * vitta_cv_resize_u8: region (x0, y0, cw, ch) of frame idx[k] (idx NULL: frame k) of src (F, H, W, 3) uint8 ->
 *   out (n, out_h, out_w, 3) uint8; tables built for cw -> out_w and ch -> out_h.
 * vitta_cv_resize_normalize_u8: the same resize, then (x - mean) * (1 / std) with HOST mean / std on the 0..255 scale
 *   (utils/opts.py:8-9), written as layout 1 (V, 3, T, h, w) (FormatShape 'NCTHW') or layout 0 (TANet planes).
 * ---------------------------------------------------------------------------------------------- */
int vitta_cv_linear_tables(int src, int dst, int horizontal, int32_t* ofs_host, int32_t* w_host);
int vitta_cv_resize_u8(const uint8_t* src, int F, int H, int W, const int32_t* idx, int n, int x0, int y0, int cw, int ch,
                       const int32_t* xofs, const int32_t* xw, const int32_t* yofs, const int32_t* yw, int out_h, int out_w,
                       uint8_t* out, void* stream);
int vitta_cv_resize_normalize_u8(const uint8_t* src, int F, int H, int W, const int32_t* idx, int n, int x0, int y0, int cw,
                                 int ch, const int32_t* xofs, const int32_t* xw, const int32_t* yofs, const int32_t* yw,
                                 int out_h, int out_w, const float* mean3_host, const float* std3_host, int layout, int T,
                                 float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITTA_B200_H */
