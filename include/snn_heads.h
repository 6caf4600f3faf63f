/* snn_heads.h -- C ABI of the B200-native spiking detection heads (libsnn_heads_b200.so).
 *
 * The reference (aitor-martinez-seras/SNN-Automotive-Object-Detection) has no FFI: its hot path is two
 * torch.nn.Module.forward methods built on Norse.  Each entry point below is what a binding for that
 * path binds; the reference interface it replaces is cited as file:line into the reference repo.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory (torch tensors); the library never
 *     allocates or frees device memory.  Scratch is a caller-provided workspace whose size the
 *     matching *_workspace_bytes() call returns.
 *   - kernels are enqueued on `stream` (a cudaStream_t) and the call returns without synchronising.
 *   - return value 0 = success, negative SNN_E_* otherwise; snn_last_error() gives the message
 *     (thread-local).  There is no CPU fallback: a non-sm_100 device is SNN_E_ARCH.
 *   - `mode` selects how fp32 weights are fed to the 16-bit tensor cores.  The other operand of every
 *     contraction on this path is an exact {0,1} spike, so a weight kept as k 16-bit pieces
 *     (hi + mid + lo) gives exact products; only the fp32 accumulation order differs from the reference.
 *     bf16 pieces carry 8 significant bits each (3 pieces = the fp32 weight exactly); fp16 pieces carry
 *     11 bits each and are taken from the weight row scaled by a power of two (so no piece is subnormal),
 *     the accumulator being scaled back in the epilogue: 2 fp16 pieces reproduce every weight to within
 *     one fp32 ulp (>= 23 of its 24 significant bits) at 2/3 of the tensor work of 3 bf16 pieces.
 */
#ifndef SNN_HEADS_H_
#define SNN_HEADS_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNN_ABI_VERSION 8

#define SNN_MODE_FP32_EXACT 0 /* 3 bf16 pieces per weight (24 mantissa bits): the parity mode          */
#define SNN_MODE_BF16 1       /* 1 piece: weights rounded to bf16, the throughput mode                 */
#define SNN_MODE_BF16X2 2     /* 2 pieces (16 mantissa bits)                                           */
#define SNN_MODE_FP16X2 3     /* 2 fp16 pieces of the row-scaled weight (>= 23 bits): fast fp32-grade mode */
#define SNN_MODE_FP16 4       /* 1 fp16 piece of the row-scaled weight (11 bits)                        */

#define SNN_OK 0
#define SNN_E_ARG (-1)       /* bad argument / unsupported shape                                        */
#define SNN_E_ARCH (-2)      /* device is not sm_100 (B200)                                             */
#define SNN_E_WORKSPACE (-3) /* workspace too small                                                     */
#define SNN_E_CUDA (-4)      /* a CUDA runtime/driver call failed                                       */

typedef void* snn_stream_t; /* cudaStream_t */

int snn_version(void);
const char* snn_last_error(void);

/* Bytes per neuron of a time-packed spike train for T steps: 1 (T<=8), 2 (T<=16), 4 (T<=32). */
int snn_train_word_bytes(int T);
/* Number of 16-bit pieces per weight for `mode`. */
int snn_mode_pieces(int mode);

/* ---- one-time weight preparation (re-run only when the fp32 weights change) ------------------------- */
/* bytes of a prepared weight with `rows` outputs and `cols` inputs: [pieces][rows][cols] 16-bit pieces
 * followed by float scale[rows] (the power of two the accumulator row is multiplied with; 1 for bf16 modes) */
size_t snn_prepared_weight_bytes(int rows, int cols, int mode);
/* RPNHeadSNN.shared_conv.weight [O][C][3][3] fp32 (rpn.py:65-66) -> [pieces][O][9*C] bf16, k = (ky*3+kx)*C + c */
int snn_prepare_conv3x3_weights(const float* w, int O, int C, int mode, void* out, snn_stream_t stream);
/* nn.Linear weight [O][K] fp32 (fc6 / fc7, faster_rcnn.py:448,451) -> [pieces][O][K] bf16 */
int snn_prepare_fc_weights(const float* w, int O, int K, int mode, void* out, snn_stream_t stream);

/* ---- RPNHeadSNN.forward (rpn.py:84-121) ---------------------------------------------------------------
 * feat_ptrs[l]   : level l features, fp32 NCHW [N][C_in][H[l]][W[l]]            (input `x`, rpn.py:84)
 * w_shared_prep  : snn_prepare_conv3x3_weights() of shared_conv.weight           (rpn.py:105)
 * w_cls, w_bbox  : conv_cls.weight [A][C_in], conv_bbox.weight [4A][C_in] fp32   (rpn.py:110,114)
 * logits_out[l]  : fp32 NCHW [N][A][H][W]   = last-step membrane of lif_obj      (rpn.py:118)
 * bbox_out[l]    : fp32 NCHW [N][4A][H][W]  = last-step membrane of lif_bbox     (rpn.py:119)
 * spike_trains_out[l] (nullable array / entries): shared_lif spikes, NHWC [N][H][W][C_in] words of
 *                  snn_train_word_bytes(T) bytes, bit t = spk_shared at step t   (rpn.py:106)
 * spike_counts_out (nullable): [n_levels][N] total shared_lif spikes per level and image (must be zeroed
 *                  by the caller; accumulated with integer atomics -> deterministic)
 * T = num_steps (rpn.py:52, 1 <= T <= 32); C_in multiple of 128; every level processed independently
 * with fresh state exactly as rpn.py:90-96. */
size_t snn_rpn_head_workspace_bytes(const int* H, const int* W, int n_levels, int N, int C_in, int T, int mode);
int snn_rpn_head_forward(const void* const* feat_ptrs, const int* H, const int* W, int n_levels, int N, int C_in,
                         int A, int T, int mode, const void* w_shared_prep, const float* w_cls, const float* w_bbox,
                         void* const* logits_out, void* const* bbox_out, void* const* spike_trains_out,
                         unsigned long long* spike_counts_out, void* workspace, size_t workspace_bytes,
                         snn_stream_t stream);

/* ---- FastRCNNPredictorSNNFull.forward (faster_rcnn.py:470-516) -----------------------------------------
 * x          : RoI features fp32 [R][K] (= [R][256][7][7] flattened, k = c*49 + h*7 + w; faster_rcnn.py:474)
 * w6_prep    : snn_prepare_fc_weights() of fc6.weight [Hdim][K];  w7_prep: of fc7.weight [Hdim][Hdim]
 * w_cls      : cls_score.weight [C][Hdim] fp32; w_bbox: bbox_pred.weight [n_box_out][Hdim] fp32
 *              (n_box_out = 4C, or 4 with only_one_bbox; faster_rcnn.py:455-467)
 * cls_out    : fp32 [R][C] last-step membrane of lif_cls;  bbox_out: fp32 [R][n_box_out] (faster_rcnn.py:513-516)
 * spk6_trains, spk7_trains (nullable): [R][Hdim] spike-train words of lif6 / lif7 (faster_rcnn.py:499,501)
 * spike_counts_out (nullable): unsigned [2][R] spikes per RoI of lif6 and lif7 over all T steps.  When
 *              it or spk6_trains is given, fc6 is also evaluated for the one step whose spikes no output needs.
 * K multiple of 64, Hdim multiple of 256, 1 <= T <= 32, R >= 1 (ragged R is handled by TMA zero fill).
 * T = 1 or 2: no fc6 current reaches lif_cls / lif_bbox before the last step, the outputs are exactly zero as in the
 * reference (faster_rcnn.py:492-516).  T > 32 is not supported (the reference's own sweeps stop at 12 / 16,
 * metrics_for_different_timesteps.py:30-33). */
size_t snn_box_head_workspace_bytes(int R, int K, int Hdim, int T, int mode);
int snn_box_head_forward(const void* x, int R, int K, int Hdim, int C, int n_box_out, int T, int mode,
                         const void* w6_prep, const void* w7_prep, const float* w_cls, const float* w_bbox,
                         float* cls_out, float* bbox_out, void* spk6_trains, void* spk7_trains,
                         unsigned int* spike_counts_out, void* workspace, size_t workspace_bytes,
                         snn_stream_t stream);

/* ---- building blocks exposed for tests / profiling ---------------------------------------------------
 * Spikes travel between kernels only as time-packed spike-train WORDS: one word per neuron, bit t = spike
 * at step t, 1 / 2 / 4 bytes for up to 8 / 16 / 32 steps.
 *
 * encoder only: x [R][K] fp32 (K a multiple of 16) -> z_words [R][K], words of 1/2/4 bytes for T_live <= 8/16/32, bit t = z_t
 * (Norse lif_current_encoder, faster_rcnn.py:494). */
int snn_encode_rows(const float* x, int R, int K, int T_live, void* z_words, snn_stream_t stream);
/* The encoders evaluate lif_current_encoder as a comparator bank: the input current is constant and a spike resets
 * the membrane to its initial value, so a neuron's train is periodic with period n(x) = its first-spike step, and
 * n(x) = min{ n : x >= thresholds[n] }.  HOST call: copies the 33-entry tables (index n = 1..32; thresholds[n] =
 * the smallest fp32 input whose first spike comes at step <= n; deltas[n] = train(n) ^ train(n+1) over 32 steps). */
void snn_encoder_table(float* thresholds33, unsigned int* deltas33);
/* Exhaustive device self-test: counts, over ALL 2^32 fp32 bit patterns, the inputs whose comparator-bank word differs
 * from the step-by-step simulation of lif_current_encoder for T_live steps; *mismatches (device, zeroed by the caller)
 * must stay 0. */
int snn_encoder_selftest(int T_live, unsigned long long* mismatches, snn_stream_t stream);
/* one fully-connected spiking layer: z_words [R][K] input spike-train words of `in_word_bytes` bytes whose
 * bit (in_bit0 + i) is the input spike injected at step t0 + i (i < T_live); w_prep [pieces][M][K] + scales;
 * runs the LIF recurrence for steps 0..T-1 and writes trains [R][M] (words of snn_train_word_bytes(T));
 * optional raw currents dump [T_live][R][M] fp32.  cta_group 1 or 2 (0 = auto). */
int snn_fc_lif_layer(const void* z_words, int in_word_bytes, int in_bit0, int R, int K, int M, int T, int t0,
                     int T_live, int mode, const void* w_prep, void* trains, float* dump, int cta_group,
                     snn_stream_t stream);

/* ---- "next" row 8f-2: MultiScaleRoIAlign fused with the box head's encoder --------------------------------------
 * (roi_heads.py:1217 box_roi_pool -> faster_rcnn.py:474-494 flatten + lif_current_encoder).  feat_ptrs[l]: FPN level l,
 * fp32 NCHW [N][C][H[l]][W[l]], scales[l] = its spatial scale (2^-k); rois [R][5] = (batch index, x1, y1, x2, y2) in
 * image coordinates, roi_level [R] = the level torchvision's LevelMapper assigns; torchvision roi_align arithmetic
 * (aligned = False, sampling_ratio samples per bin).  words_out [R][C*P*P] spike-train words of 1/2/4 bytes for
 * T_live <= 8/16/32 (k = c*P*P + ph*P + pw, faster_rcnn.py:474); optional pooled_out [R][C*P*P] fp32 (tests). */
int snn_roi_align_encode(const void* const* feat_ptrs, const int* H, const int* W, const float* scales, int n_levels,
                         int C, const float* rois, const int* roi_level, int R, int pooled_size, int sampling_ratio,
                         int T_live, void* words_out, float* pooled_out, snn_stream_t stream);
/* Experiment (off by default): 1 = the RPN conv runs in clusters of two CTA pairs and every weight tile is fetched from L2
 * once per cluster (each CTA loads half of its tile and multicasts it to its counterpart, `.multicast::cluster`); the
 * results are bit-identical to the default schedule.  0 = one CTA pair per cluster, unicast weight tiles. */
void snn_set_conv_multicast(int on);
/* A-B timing: 0 = two channel planes (32 loads) in flight per thread, 1 = four (64 loads, one block fewer per SM) */
void snn_set_roi_kernel(int which);
/* snn_box_head_forward on input that is already encoded: words [R][K] of snn_train_word_bytes-like size for T - 1 steps
 * (1/2/4 bytes for T - 1 <= 8/16/32), e.g. from snn_roi_align_encode(..., T_live = T - 1, ...). */
int snn_box_head_forward_encoded(const void* words, int R, int K, int Hdim, int C, int n_box_out, int T, int mode,
                                 const void* w6_prep, const void* w7_prep, const float* w_cls, const float* w_bbox,
                                 float* cls_out, float* bbox_out, void* spk6_trains, void* spk7_trains,
                                 unsigned int* spike_counts_out, void* workspace, size_t workspace_bytes,
                                 snn_stream_t stream);

/* ---- "next" row 8f-1: the proposal selection that follows RPNHeadSNN in RegionProposalNetwork.forward ----------
 * (rpn.py:636-670: concat_box_prediction_layers + AnchorGenerator + BoxCoder.decode + sigmoid, applied by the
 * reference to ALL A*H*W anchors of every level before the per-level top-k).  The top-k is taken on the head's
 * native NCHW logits; this call decodes ONLY the selected anchors.
 * logits[l] / deltas[l]: the head outputs of level l, fp32 NCHW [N][A][H][W] / [N][4A][H][W];
 * base_anchors[l]: [A][4] fp32 cell anchors of level l (AnchorGenerator.cell_anchors); stride_h/w[l] = image size
 * // feature size; idx [N][K] (K = sum k_per_level): per image and level the selected positions inside that
 * level's [A][H][W] logits; boxes_out [N][K][4] (x1,y1,x2,y2, BoxCoder weights (1,1,1,1), dw/dh clamped to
 * log(1000/16)), scores_out [N][K] = sigmoid(logit); optional logits_out [N][K] and ref_index_out [N][K] = the
 * anchor's index in the reference's concatenated (level, h, w, a) order. */
int snn_rpn_decode_selected(const void* const* logits, const void* const* deltas, const float* const* base_anchors,
                            const int* H, const int* W, const int* stride_h, const int* stride_w, const int* k_per_level,
                            int n_levels, int N, int A, const long long* idx, float* boxes_out, float* scores_out,
                            float* logits_out, long long* ref_index_out, snn_stream_t stream);

/* Sort keys for that per-level top-k: keys_out[l] [N][A*H*W] int64, key = (order-preserving integer image of the fp32
 * logit) << 32 | (2^32 - 1 - the anchor's index in the reference's (H, W, A) flattening, rpn.py:248-259).  top-k on the
 * keys selects the largest logits and breaks ties -- e.g. the exactly-zero membranes of pixels without a spike -- by
 * the lowest reference index, whatever the top-k implementation does with equal elements. */
int snn_rpn_topk_keys(const void* const* logits, const int* H, const int* W, int n_levels, int N, int A,
                      long long* const* keys_out, snn_stream_t stream);

/* The per-level top-k itself, all levels and images of a batch in eight launches: exact radix select on those keys
 * (recomputed from the logits on every pass; 2048-bin histograms, six passes) + a shared-memory sort of the k selected
 * keys per (level, image).  idx_out [N][K] int64, K = sum_l min(k, A*H[l]*W[l]), level-major: positions inside the
 * level's [A][H][W] logits, largest logit first, ties by the lowest index in the reference's (H, W, A) order -- the input
 * of snn_rpn_decode_selected.  k <= 2048 (pre_nms_top_n: 1000 at test time, 2000 in training, model.py:50-53). */
size_t snn_rpn_topk_workspace_bytes(int n_levels, int N);
int snn_rpn_topk_select(const void* const* logits, const int* H, const int* W, int n_levels, int N, int A, int k,
                        long long* idx_out, void* workspace, size_t workspace_bytes, snn_stream_t stream);

/* ---- "next" row 8f-3: linear statistics of the spike trains the heads emit (the spike-rate / energy report the
 * reference obtains from hand-edited forwards, rpn.py:126-200, faster_rcnn.py:520-618; train.py:426-517).
 * out[o] = sum_c w[o][c] * sum_t step_weights[t] * spk_t[c], the two leaky-integrator readout kernels with a caller-given
 * weight per step (32 doubles; entries beyond the word's bits are ignored): kappa_{T-1-t} gives the last LI membrane,
 * the cumulative K_t = sum_{n <= T-1-t} kappa_n the sum over time of the LI membrane (its time mean, the reference's LI
 * "rate", is K_t / T), 1 the spike count.
 * nhwc: trains [N][HW][C] words (RPNHeadSNN spike_trains_out) -> out_a [N][n_a][HW], out_b [N][n_b][HW];
 * rows: trains [R][Hd] words (spk6 / spk7 trains) -> out_a [R][n_a], out_b [R][n_b]. */
int snn_li_readout_nhwc(const void* trains, int train_bytes, int N, int HW, int C, const double* step_weights,
                        const float* w_a, int n_a, const float* w_b, int n_b, float* out_a, float* out_b,
                        snn_stream_t stream);
int snn_li_readout_rows(const void* trains, int train_bytes, int R, int Hd, const double* step_weights, const float* w_a,
                        int n_a, const float* w_b, int n_b, float* out_a, float* out_b, snn_stream_t stream);

/* ---- "next" row 8f-1, second half: the tail of RegionProposalNetwork.filter_proposals (rpn.py:493-525) on the entries
 * snn_rpn_topk_select / snn_rpn_decode_selected produced -- clip to the image, remove_small_boxes(min_size), score >=
 * score_thresh, batched NMS over the levels, the post_nms_top_n best -- in one launch of n_levels x N blocks (every
 * (image, level) is its own NMS problem; the last block of an image merges the levels), no host synchronisation.  The
 * NMS repeats torchvision's CUDA path operation for operation (see snn_det_postprocess), including batched_nms's shift
 * of the boxes by level * (largest coordinate of the image + 1) up to 5000 candidates per image.
 * proposals [N][K][4] (decoded, NOT clipped), probs [N][K], level-major with level_sizes[l] entries of level l per image
 * (HOST array; K = their sum; each <= 2048; sum_l min(post_nms_top_n, level_sizes[l]) <= 8192); img_h, img_w HOST [N].
 * Outputs (device): out_boxes [N][post_nms_top_n][4] clipped, out_scores [N][post_nms_top_n] -- score descending, ties by
 * the lower position --, out_counts [N] int32 rows written per image. */
size_t snn_rpn_nms_workspace_bytes(int n_levels, int N);
int snn_rpn_nms(const float* proposals, const float* probs, const int* level_sizes, const int* img_h, const int* img_w,
                int n_levels, int N, float min_size, float score_thresh, float nms_thresh, int post_nms_top_n,
                float* out_boxes, float* out_scores, int* out_counts, void* workspace, size_t workspace_bytes,
                snn_stream_t stream);

/* ---- "next" row 8f-4: the detector's post-processing, RoIHeadsSNN.postprocess_detections (roi_heads.py:1075-1176)
 * after its softmax and box decode: clip (l.1105), score threshold on classes >= 1 (l.1127-1128), the background box of
 * every RoI none of whose classes passed (l.1139-1150), remove_small_boxes (l.1153-1158), batched NMS of both sets
 * (l.1161-1162), the detections_per_img best objects (l.1164), objects then background (l.1170-1172) -- one block per
 * image, no host synchronisation.  The NMS repeats torchvision's CUDA path operation for operation (stable descending
 * sort, devIoU as compiled for sm_100, the coordinate trick up to 5000 boxes and per-class NMS above), so the kept set
 * and its order are the ones the reference gets on the same device.
 * scores [R_total][C] (softmax), boxes [R_total][C][4] (decoded, NOT clipped), both fp32 device; rois_per_image, img_h,
 * img_w: HOST arrays [N].  Outputs (device, caller-owned): all_boxes [R_total][C][4] clipped; out_boxes [N][cap][4],
 * out_scores [N][cap], out_labels [N][cap] int64 -- image b's rows [0, counts[b][0]) are its objects, the next
 * counts[b][1] rows its background boxes; out_counts [N][2] int32.  cap >= detections_per_img + max RoIs per image.
 * Limits: C >= 2, RoIs per image * (C - 1) <= snn_det_postprocess_max_candidates() (8192). */
int snn_det_postprocess_max_candidates(void);
int snn_det_postprocess(const float* scores, const float* boxes, const int* rois_per_image, const int* img_h,
                        const int* img_w, int N, int C, float score_thresh, float nms_thresh, float min_size,
                        int detections_per_img, int cap, float* all_boxes, float* out_boxes, float* out_scores,
                        long long* out_labels, int* out_counts, snn_stream_t stream);

/* number of kernels the last forward call on this thread enqueued (for bench accounting) */
int snn_last_launch_count(void);
/* force cta_group (1 or 2; 0 = auto) for subsequent forward calls on this thread -- tests/profiling only */
void snn_set_cta_group(int cta_group);
/* fc tiling of subsequent calls on this thread -- tests/profiling only.  dual: 0 = auto (dual tiles -- 2J units whose
 * two accumulators share every weight tile -- for K >= 4096), 1 = never, 2 = whenever the tile shape allows it;
 * max_units > 0 caps the units per accumulator tile (0 = the widest tile without a padding step, else the widest);
 * tail_split: 0 = auto (the units a last, partly filled wave of dual tiles would cover run in a second launch as one
 * wave of single tiles), 1 = never, 2 = as one wave of narrower dual tiles when such a shape exists. */
void snn_set_fc_tiling(int dual, int max_units, int tail_split);
/* Profiling only: the MMA-issuing thread of every CTA pair of the next spike-GEMM launches of `phase` (0 = RPN conv,
 * 1 = fc with K >= 4096 in dual tiles, 3 = the same in single tiles, 2 = other fc) on this thread writes counters in SM clock cycles to device_counters
 * [pairs][12]: [0] whole role, [1] waiting for a free accumulator, [2] for this CTA's half of a spike tile, [3] for the
 * peer CTA's half, [4] for weight tiles, [5] number of tiles, [6] the LIF-epilogue role (one warp of the leader CTA),
 * [7] of which waiting for a full accumulator, [8] kernel entry to the MMA role's start, [9] kernel entry to exit of the
 * leader CTA, [10] the same in nanoseconds (globaltimer).  NULL switches it off. */
void snn_set_role_timers(unsigned long long* device_counters, int phase);

/* Measurement only: every RPN conv+LIF launch of this thread ADDS, per CTA pair, its kernel entry-to-exit time in SM
 * cycles and in globaltimer nanoseconds to device_counters [pairs][2] (zeroed by the caller; NULL = off): the SM clock
 * the kernel really ran at, measured on the same launches a bench times with CUDA events. */
void snn_set_clock_probe(unsigned long long* device_counters);
/* Host-side descriptor cache of this thread: TMA tensor maps served from the cache / encoded (immutable-
 * after-init cache keyed by address, shape and box). */
void snn_host_cache_stats(unsigned long long* hits, unsigned long long* misses);

/* Per-phase device timing for bench.py: when enabled, every forward records CUDA events on its stream
 * around each phase (up to 256 forwards).  snn_profile_read() waits for them, writes the summed
 * milliseconds and the number of timed forwards per phase, and resets.  Phase order:
 * 0 rpn encoder, 1 rpn conv+LIF GEMM, 2 rpn readout, 3 box encoder, 4 fc6+LIF GEMM, 5 fc7+LIF GEMM, 6 box readout.
 * on: 0 = off, 1 = every phase, else a mask with bit (1 + phase) set for each phase to time (4 = the conv GEMM only). */
#define SNN_PHASES 7
void snn_profile_enable(int on);
int snn_profile_read(float* ms_out, int* counts_out);

#ifdef __cplusplus
}
#endif
#endif /* SNN_HEADS_H_ */
