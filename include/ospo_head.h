/*
 * ospo_head.h -- C ABI of the B200-native (sm_100a) Janus-Pro image-token head.
 *
 * This is the drop-in boundary for OSPO's one data-parallel hot path.  The reference has no FFI for
 * it (pure Python); the entry points below are what a binding for that path attaches to, each citing
 * the reference lines it replaces (paths relative to the OSPO repository root):
 *
 *   ospo_head_logits          janus/models/modeling_vlm.py:47-51   vision_head.forward
 *   ospo_head_logps_fwd/_bwd  ospo/wrapper/train.py:357 + 375-396  gen_head + get_batch_logps (+ autograd)
 *   ospo_head_simpo_fwd/_bwd  ospo/wrapper/train.py:317-342, 345-372, 399-445  SimPO loss fwd + bwd
 *   ospo_head_cfg_sample      ospo/wrapper/image_generation.py:156-164 (== ospo/inference.py:147-155)
 *   ospo_head_cfg_merge_sample   the merge/softmax/sample tail of the same lines, on supplied logits
 *   ospo_head_pack_weight     one-time re-layout of W1 / W2 for the decode step's weight stream (no reference counterpart)
 *   ospo_head_gen_img_embeds  janus/models/modeling_vlm.py:263-264 + projector.py:39-45 (next row N1)
 *   ospo_head_grad_sqnorm / ospo_head_adamw_step   ospo/utils/train.py:30,50 + ospo/wrapper/train.py:108-115 (next row N3)
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless the name ends in _host.  The caller owns all memory,
 *     including the workspace.  The library itself owns one 16 KB block of device flag words (decode kernel).
 *   - All work is enqueued asynchronously on `stream`; no call synchronises.
 *   - Weights use the nn.Linear layout [out, in], row-major, bf16.  Biases are fp32.
 *   - x / hidden states: bf16 [rows, hidden] row-major contiguous, one row per image token whose label
 *     is not masked (the caller drops the label == -100 rows: they carry no work and no gradient).
 *   - Devices and threads: a call works on the calling thread's CURRENT device (cudaGetDevice); all pointers and the
 *     stream must belong to it.  Any number of sm_100 devices may be used from one process (the library keeps its
 *     few pieces of state -- SM count, flag words, kernel attributes -- per device).  Calls on different devices or
 *     streams may come from different threads; ospo_head_profile_* and the ospo_head_set_* / ospo_head_trace knobs are
 *     process-wide and must not race with running calls.
 *   - Return value 0 = success, negative = ospo_head_status; no exception crosses this boundary.
 *   - Requires an sm_100a device: there is no fallback path.
 */
#ifndef OSPO_HEAD_H_
#define OSPO_HEAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* ospo_stream_t; /* == cudaStream_t */

#if defined(__GNUC__)
#define OSPO_API __attribute__((visibility("default")))
#else
#define OSPO_API
#endif

typedef enum {
  OSPO_OK = 0,
  OSPO_ERR_BAD_SHAPE = -1,      /* non-positive or unsupported dimension */
  OSPO_ERR_ALIGNMENT = -2,      /* pointer not 16-byte aligned / dimension not a multiple of 8 */
  OSPO_ERR_NULL = -3,           /* required pointer is NULL */
  OSPO_ERR_WORKSPACE = -4,      /* workspace too small */
  OSPO_ERR_ARCH = -5,           /* device is not sm_100 */
  OSPO_ERR_TENSORMAP = -6,      /* cuTensorMapEncodeTiled failed / unavailable */
  OSPO_ERR_LAUNCH = -7,         /* kernel launch failed */
  OSPO_ERR_CUDA = -8,           /* other CUDA runtime error */
  OSPO_ERR_UNSUPPORTED = -9     /* argument combination not supported */
} ospo_head_status;

typedef struct {
  int32_t rows;     /* N: image-token rows given to the head (sum over sequences)          */
  int32_t hidden;   /* H: n_embed                                                          */
  int32_t embed;    /* E: image_token_embed                                                */
  int32_t vocab;    /* V: image_token_size (16384 for the sampler)                         */
  int32_t num_seqs; /* S: sequences (SimPO: S = 2B, chosen [0,B) then rejected [B,2B))     */
} ospo_head_shape;

typedef struct {
  const void* w1;   /* bf16 [E, H]  output_mlp_projector.weight */
  const float* b1;  /* fp32 [E]     output_mlp_projector.bias   */
  const void* w2;   /* bf16 [V, E]  vision_head.weight          */
  const float* b2;  /* fp32 [V]     vision_head.bias            */
} ospo_head_weights;

/* ---- plain logits: vision_head.forward ---------------------------------------------------- */
typedef struct {
  ospo_head_shape shape;       /* num_seqs ignored */
  ospo_head_weights w;
  const void* x;               /* bf16 [rows, H] */
  void* logits;                /* bf16 [rows, V] out */
  void* workspace;             /* >= ospo_head_workspace_bytes() */
  size_t workspace_bytes;
} ospo_head_args;

/* ---- log-probs / SimPO -------------------------------------------------------------------- */
#define OSPO_LOSS_SIGMOID 0
#define OSPO_LOSS_HINGE 1

/* indices into ospo_simpo_args.scalars (fp32[16]) */
#define OSPO_SC_LOSS 0            /* losses.mean() + sft_weight * sft_loss          train.py:419,428 */
#define OSPO_SC_SIMPO_LOSS 1      /* losses.mean()                                  train.py:419     */
#define OSPO_SC_SFT_LOSS 2        /* CE over unmasked chosen rows                   train.py:425-427 */
#define OSPO_SC_REWARD_CHOSEN 3   /* chosen_rewards.mean()                          train.py:435     */
#define OSPO_SC_REWARD_REJECTED 4
#define OSPO_SC_REWARD_ACC 5      /* (chosen_rewards > rejected_rewards).mean()     train.py:432     */
#define OSPO_SC_REWARD_MARGIN 6
#define OSPO_SC_LOGPS_CHOSEN 7
#define OSPO_SC_LOGPS_REJECTED 8
#define OSPO_SC_LOGITS_CHOSEN 9   /* mean logit over the unmasked chosen rows       train.py:442     */
#define OSPO_SC_LOGITS_REJECTED 10
#define OSPO_SC_COUNT 16

/* ---- data-parallel gradient exchange over NVLink peer memory ---------------------------------------
 * Replaces DDP's NCCL all-reduce of the head-weight gradients (ospo/utils/train.py:26-28) for up to 8 ranks of one
 * NVSwitch domain.  The flat gradient dW2 | dW1 | db2 | db1 is partitioned by rows (rank r owns rows
 * [r V/N, (r+1) V/N) of dW2 and db2, [r E/N, (r+1) E/N) of dW1 and db1; V/N and E/N must be multiples of 256).
 * Step 1 (inside ospo_head_*_bwd when ospo_simpo_args.dp is set): the weight-gradient GEMM epilogues and the bias
 * reductions write their results, times wgrad_scale, straight into the OWNER's inbox -- slot [this rank] -- with
 * plain peer stores; nothing is written to flat_grads.  Step 2 (ospo_head_dp_reduce_broadcast, after every rank has
 * finished step 1: the caller puts a barrier between them): each owner adds the N slots of its inbox in rank order
 * (bit-reproducible) and stores the sums into the flat gradient buffer of every rank (one multimem.st through the
 * NVSwitch multicast mapping if given, else one store per peer).  After a second barrier every rank holds the same
 * bits.  All pointers are mappings valid in THIS process (e.g. torch symmetric memory / cuMemMap of peer handles). */
typedef struct {
  int32_t world, rank;         /* 2 <= world <= 8 */
  float* inbox[8];             /* [world] inbox of every rank, fp32 [world slots][shard_elems]; [rank] is the local one */
  float* flat[8];              /* [world] flat gradient buffer of every rank, fp32 [V*E + E*H + V + E] */
  float* flat_multicast;       /* multicast mapping of the flat buffers, or NULL */
} ospo_dp_exchange;

typedef struct {
  ospo_head_shape shape;
  ospo_head_weights w;
  const void* x;               /* bf16 [rows, H], or the base of a [S, x_seg_pitch, H] tensor (see x_seg_*) */
  /* Row-segmented hidden states (all 0 = x is a contiguous [rows, H] matrix).  With x_seg_rows = T > 0 (a multiple
     of 64), x and dx point at [S, x_seg_pitch, H] tensors -- the backbone's last hidden state as it lies in memory,
     ospo/wrapper/train.py:356 -- and the head's rows are rows [x_seg_off, x_seg_off + T) of every sequence; rows
     must equal S * T.  Nothing is copied; dx rows outside the span are left untouched (the caller zeroes them). */
  int32_t x_seg_rows, x_seg_pitch, x_seg_off;
  const int64_t* labels;       /* [rows] target code of each row, in [0, V)  (labels[:,1:] after masking) */
  const int64_t* seq_offsets;  /* [S+1] row range of each sequence: rows [off[s], off[s+1]) */
  int32_t average_log_prob;    /* get_batch_logps(average_log_prob=...)  train.py:393-396 */

  /* SimPO hyper-parameters (ignored by the logps_* entry points) */
  float beta, gamma_beta_ratio, label_smoothing, sft_weight;
  int32_t loss_type;           /* OSPO_LOSS_SIGMOID | OSPO_LOSS_HINGE */

  /* forward outputs */
  float* row_logps;            /* [rows] per-token log-prob of the label */
  float* seq_logps;            /* [S]    get_batch_logps result */
  float* losses;               /* [B]    simpo only */
  float* chosen_rewards;       /* [B]    simpo only */
  float* rejected_rewards;     /* [B]    simpo only */
  float* scalars;              /* [OSPO_SC_COUNT] simpo only */

  /* tensors saved by the forward for the backward (caller-allocated; may all be NULL for a
     forward-only call, in which case nothing is spilled) */
  void* pre;                   /* bf16 [rows, E]  pre-activation */
  void* act;                   /* bf16 [rows, E]  required even forward-only */
  void* logits;                /* bf16 [rows, V]  spill of g = exp(logit - row_ref) - onehot * exp(row_lse - row_ref), the
                                  unscaled softmax-minus-onehot numerator (dlogits = w_row * g); written once by the
                                  forward GEMM2 epilogue, READ-ONLY in the backward */
  float* row_lse;              /* [rows] */
  float* row_ref;              /* [rows] exponent reference of each spill row (0 unless the row was repaired); required
                                  whenever `logits` is given */
  float* grad_seq;             /* [S] d loss / d seq_logps: written by simpo_fwd, read by *_bwd
                                  (for logps_bwd the caller fills it with the upstream gradient) */

  /* backward */
  const float* grad_loss;      /* device scalar multiplying the whole gradient, NULL = 1 */
  void* dx;                    /* bf16 [rows, H] out, NULL = skip */
  float* flat_grads;           /* fp32 [V*E + E*H + V + E] = dW2 | dW1 | db2 | db1, NULL = head frozen */

  void* workspace;
  size_t workspace_bytes;

  /* staged backward (data-parallel overlap): 0 = whole backward in one call; otherwise a bit mask of the
     parts to run now -- 1 = row weights, dpre, db2, dW2;  2 = db1, dW1;  4 = dX -- so the caller can
     start the all-reduce of dW2 after part 1 and of the remainder after part 2 while the later parts still run.  All
     calls must pass the same workspace (dpre lives there).  reserve_sms (parts 2 / 4): leave this many SMs free for
     the collective kernel running beside the GEMMs. */
  int32_t bwd_stage;
  int32_t reserve_sms;
  float wgrad_scale;           /* multiplies dW2 | dW1 | db2 | db1 as they are stored (dx is not scaled): a data-parallel
                                  caller passes 1 / world_size and all-reduces with SUM (ospo/utils/train.py:26-28);
                                  0 means 1 */
  const ospo_dp_exchange* dp;  /* backward only, optional: fuse the reduce-scatter of the data-parallel exchange into the
                                  weight-gradient stores (see ospo_dp_exchange); flat_grads must still be given (it
                                  marks the head as trainable) but is not written by the backward */
} ospo_simpo_args;

/* ---- next row: sampled ids -> next-step input embeddings ---------------------------------
 * prepare_gen_img_embeds = gen_aligner(gen_embed(ids))  (janus/models/modeling_vlm.py:263-264; MlpProjector
 * "mlp_gelu" depth 2, janus/models/projector.py:39-45,77-86), called right after sampling at
 * ospo/wrapper/image_generation.py:166-168.  All bf16 like the generation path; the second Linear streams its
 * D x D weight once (swap-AB tcgen05 GEMM, cluster split-K). */
typedef struct {
  int32_t rows;             /* n output rows (the reference passes the 2P duplicated ids; n <= 32 per call) */
  int32_t embed;            /* D: n_embed of the language model */
  int32_t codebook;         /* rows of gen_embed (16384) */
  int32_t code_dim;         /* columns of gen_embed; must be 8 */
  const int64_t* ids;       /* [n] */
  const void* gen_embed;    /* bf16 [codebook, 8] */
  const void* wa;           /* bf16 [D, 8]   gen_aligner.layers.0.weight */
  const float* ba;          /* fp32 [D]      gen_aligner.layers.0.bias   */
  const void* wb;           /* bf16 [D, D]   gen_aligner.layers.2.weight */
  const float* bb;          /* fp32 [D]      gen_aligner.layers.2.bias   */
  void* out;                /* bf16 [n, D] */
  void* workspace;          /* >= n * D * 2 bytes, 16-byte aligned */
  size_t workspace_bytes;
  int32_t id_repeat;        /* 0 / 1: ids has n entries.  r > 1: ids has n / r entries and entry i feeds rows
                               [i*r, (i+1)*r) -- r = 2 is the cond/uncond duplication of image_generation.py:166,
                               so the sampler's ids[P] can be passed as they are */
  const void* table;        /* optional bf16 [codebook, D]: table[id] = gen_aligner(gen_embed(id)) for EVERY code, built
                               once by the caller with this same entry point (the module is a pure function of the
                               id and generation runs with frozen weights).  When given, out rows are copied from it:
                               no weight is streamed and, behind ospo_head_cfg_sample, no extra kernel runs (the
                               sampler's finish kernel writes the rows).  wa / ba / wb / bb / workspace may then be
                               NULL.  The caller rebuilds the table when the aligner's weights change. */
} ospo_aligner_args;

/* ---- CFG decode step ---------------------------------------------------------------------- */
#define OSPO_MERGE_BF16 0  /* reference semantics: bf16 rounding after every op (image_generation.py:160-161) */
#define OSPO_MERGE_FP32 1

typedef struct {
  ospo_head_shape shape;       /* rows = 2P (row 2k conditional, 2k+1 unconditional); vocab must be 16384 */
  ospo_head_weights w;
  const void* h;               /* bf16 [2P, H] last hidden state of every CFG row; NULL for merge_sample */
  void* logits;                /* bf16 [2P, V]: read by cfg_merge_sample; optional dump for cfg_sample (NULL =
                                  the logits never leave the chip) */
  float cfg_weight, temperature;
  int32_t merge_mode;          /* OSPO_MERGE_BF16 | OSPO_MERGE_FP32 */
  int32_t greedy;              /* 1 = argmax (lowest index on ties), uniforms ignored */
  int32_t num_steps;           /* cfg_merge_sample only: independent steps batched in one launch
                                  (logits [steps, 2P, V], uniforms / ids [steps, P]); 0 or 1 = one step */
  const float* uniforms;       /* fp32 [P] in [0,1) */
  int64_t* ids;                /* [P] out */
  float* merged;               /* optional fp32 [P, V] dump of the merged, temperature-scaled logits */
  void* workspace;
  size_t workspace_bytes;
  const void* w1_packed;       /* cfg_sample only, optional: W1 / W2 pre-packed by ospo_head_pack_weight.  The decode */
  const void* w2_packed;       /* kernel then streams each 16 KB weight tile with one contiguous bulk copy.  The caller
                                  re-packs when the weights change; either may be NULL. */
  const ospo_aligner_args* next_embeds;  /* cfg_sample only, optional: also produce the next step's input embeddings
                                  (image_generation.py:166-168) in the same launch chain.  rows must be 2P,
                                  id_repeat 2; its `ids` field is ignored (the sampled ids are used).  The first
                                  aligner layer runs inside the sampler's finish kernel. */
} ospo_cfg_args;

OSPO_API int ospo_head_gen_img_embeds(const ospo_aligner_args* args, ospo_stream_t stream);

/* ---- next row: gradient-norm clip + AdamW on the flat gradient buffer -----------------------
 * Lightning gradient_clip_val (torch.nn.utils.clip_grad_norm_, ospo/utils/train.py:30,50) and
 * torch.optim.AdamW (ospo/wrapper/train.py:108-115, configs/step5.yaml:37-43) for the head's parameters, which
 * live in the layout of ospo_simpo_args.flat_grads (W2 | W1 | b2 | b1, fp32).  Two streaming passes. */
typedef struct {
  int64_t numel;              /* elements of the flat buffers */
  const float* grads;         /* fp32 [numel]  (after the all-reduce) */
  float* params;              /* fp32 [numel]  master parameters, updated in place */
  float* exp_avg;             /* fp32 [numel]  first moment, updated in place */
  float* exp_avg_sq;          /* fp32 [numel]  second moment, updated in place */
  void* params_bf16;          /* optional bf16 [shadow_numel]: refreshed copy of params[0 : shadow_numel] (the GEMM
                                 operands W2 | W1), written in the same pass */
  int64_t shadow_numel;       /* multiple of 4, <= numel */
  double lr, beta1, beta2, eps, weight_decay;   /* as torch takes them: Python floats */
  int32_t step;               /* 1-based count of optimizer steps including this one (bias correction) */
  float max_norm;             /* gradient clipping threshold; <= 0 disables clipping */
  const float* total_sqnorm;  /* device scalar: squared L2 norm over ALL parameters being clipped together (the
                                 caller adds the other modules' share to ospo_head_grad_sqnorm's result);
                                 required when max_norm > 0 */
} ospo_adamw_args;
/* out_sq[0] (device) = sum of grads^2, summed in a fixed order (bit-reproducible run to run).
   workspace: >= 4096 bytes of device scratch. */
OSPO_API int ospo_head_grad_sqnorm(const float* grads, int64_t numel, float* out_sq, void* workspace,
                                   size_t workspace_bytes, ospo_stream_t stream);
OSPO_API int ospo_head_adamw_step(const ospo_adamw_args* args, ospo_stream_t stream);

/* scratch bytes needed by any entry point for this shape */
OSPO_API int ospo_head_workspace_bytes(const ospo_head_shape* shape, size_t* out_bytes);

OSPO_API int ospo_head_logits(const ospo_head_args* args, ospo_stream_t stream);

OSPO_API int ospo_head_logps_fwd(const ospo_simpo_args* args, ospo_stream_t stream);
OSPO_API int ospo_head_logps_bwd(const ospo_simpo_args* args, ospo_stream_t stream);
OSPO_API int ospo_head_simpo_fwd(const ospo_simpo_args* args, ospo_stream_t stream);
OSPO_API int ospo_head_simpo_bwd(const ospo_simpo_args* args, ospo_stream_t stream);
/* step 2 of the peer-memory gradient exchange: this rank's inbox -> its shard of every rank's flat gradient.
   regions: 1 = the dW2 rows of the shard (complete after backward part 1, so a caller can exchange them on a side
   stream beside the remaining GEMMs), 2 = the rest (dW1 rows, db2, db1), 0 or 3 = everything.  max_blocks > 0 bounds
   the grid (a small grid shares the SMs with GEMMs running beside it). */
OSPO_API int ospo_head_dp_reduce_broadcast(const ospo_head_shape* shape, const ospo_dp_exchange* dp, int32_t regions,
                                           int32_t max_blocks, ospo_stream_t stream);

/* Pre-pack a [rows, cols] bf16 row-major weight for the decode step: [rows/128][cols/64] tiles of 16 KB, each the
   128-byte-swizzled shared-memory image of a {64 x 128} box.  packed: ospo_head_packed_weight_bytes() bytes. */
OSPO_API int ospo_head_packed_weight_bytes(int32_t rows, int32_t cols, size_t* out_bytes);
OSPO_API int ospo_head_pack_weight(const void* w, int32_t rows, int32_t cols, void* packed, ospo_stream_t stream);
OSPO_API int ospo_head_cfg_sample(const ospo_cfg_args* args, ospo_stream_t stream);
OSPO_API int ospo_head_cfg_merge_sample(const ospo_cfg_args* args, ospo_stream_t stream);

OSPO_API const char* ospo_head_strerror(int status);

/* ---- introspection / tuning ---------------------------------------------------------------- */
/* tcgen05 cta_group used by the training GEMMs: 1 (one CTA per 128-row tile) or 2 (CTA pair per
   256-row tile).  Returns the value now in effect; pass 0 to query. */
OSPO_API int ospo_head_set_cta_group(int cta_group);
/* decode step: fused = CFG tail inside the GEMM2 epilogue (1, default) or separate sampler pass (0);
   pdl = programmatic dependent launch along the decode kernel chain (1, default).  -1 leaves a setting
   unchanged.  Returns fused | pdl << 1. */
OSPO_API int ospo_head_set_decode_mode(int fused, int pdl);
/* decode step, fused form only: 1 (default) = the whole step (both GEMMs + CFG epilogue) is one persistent
   kernel; 0 = GEMM1 and GEMM2 are separate launches.  -1 queries.  Returns the value in effect. */
OSPO_API int ospo_head_set_decode_merged(int merged);
/* one-kernel decode step: number of 16 KB W2 tiles per CTA requested into L2 while the activation flag is still
   closed (keeps HBM streaming through that wait); 0 = off, -1 queries.  Returns the value in effect. */
OSPO_API int ospo_head_set_decode_l2_ahead(int kblocks);
/* rasterisation group size (M-blocks walked together); pass 0 to query */
OSPO_API int ospo_head_set_group_m(int group_m);
/* k-splits of the two weight-gradient GEMMs: 0 (default) = chosen per shape so that the last wave of the persistent
   grid is not left mostly empty (dW1 of the 7B head: 512 half-length work items instead of 256 tiles, both halves
   added into the zeroed output -- two addends, order-independent bits), 1 = never split, 2 = always; -1 queries.
   Returns the value in effect. */
OSPO_API int ospo_head_set_wgrad_splitk(int splits);
/* per training GEMM (kernel: 0 gemm1, 1 gemm2, 2 dact, 3 wgrad W2, 4 wgrad W1, 5 dgrad X): rasterisation group
   (0 = the global group_m) and the L2 eviction hints of its A / B operand loads (0 normal, 1 evict-first,
   2 evict-last); a negative argument leaves that setting unchanged */
OSPO_API int ospo_head_set_kernel_tune(int kernel, int group_m, int a_evict, int b_evict);
/* Per-kernel timing with CUDA events recorded on the caller's stream around each launch group (off by
   default).  profile_read synchronises on the recorded events and returns, per OSPO_K_* id, the summed
   milliseconds and the number of spans since the previous read. */
#define OSPO_K_GEMM1 0          /* x W1^T + b1, GELU                      (tcgen05 GEMM, K/K)   */
#define OSPO_K_GEMM2_LSE 1      /* act W2^T + b2, softmax numerator spill, LSE partials, gather (tcgen05 GEMM, K/K) */
#define OSPO_K_SCALAR_STAGE 2   /* lse merge (+ repair pass, normally empty), one-hot fix-up, per-sequence reduce, SimPO scalars */
#define OSPO_K_ROW_WEIGHTS 3    /* per-row weights of the backward GEMM pair                      */
#define OSPO_K_DACT 4           /* row_w * (g W2), GELU', act_w           (tcgen05 GEMM, K/MN)  */
#define OSPO_K_WGRAD2 5         /* g^T act_w                              (tcgen05 GEMM, MN/MN) */
#define OSPO_K_COLSUM 6         /* db2, db1: fixed-order column sums                              */
#define OSPO_K_WGRAD1 7         /* dpre^T x                               (tcgen05 GEMM, MN/MN) */
#define OSPO_K_DGRAD 8          /* dpre W1                                (tcgen05 GEMM, K/MN)  */
#define OSPO_K_GEMM2_PLAIN 9    /* act W2^T + b2 -> logits                                         */
#define OSPO_K_DECODE_GEMM1 10  /* swap-AB W1 h^T, GELU                                            */
#define OSPO_K_DECODE_GEMM2 11  /* swap-AB W2 act^T                                                */
#define OSPO_K_SAMPLER 12       /* CFG merge + softmax + inverse-CDF sample                        */
#define OSPO_K_ALIGNER 13       /* gen_embed lookup + Linear(8->D) + GELU, then swap-AB Linear(D->D)   */
#define OSPO_K_OPTIMIZER 14     /* squared-norm reduction + clip + AdamW on the flat buffer           */
#define OSPO_K_DP_EXCHANGE 15   /* peer-memory gradient exchange: inbox -> every rank's flat gradient  */
#define OSPO_K_COUNT 16
OSPO_API int ospo_head_profile_enable(int enable);
OSPO_API int ospo_head_profile_read(float* total_ms, int32_t* counts, int32_t n);
/* tuning aid: CTA timeline of the decode chain.  device_buf = u64[5][160][8] (kernel: 1 GEMM1, 2 GEMM2, 3 finalize,
   4 finish; per CTA: 0 entry, 1 weights prefetched, 2 dependency wait over, 3 producer done, 4 first accumulator
   ready, 5 epilogue done), NULL = off */
OSPO_API int ospo_head_trace(void* device_buf);
/* number of kernels launched by this library since load (the bench's gpu_launches counter) */
OSPO_API uint64_t ospo_head_launch_count(void);
/* 6 x u32 watchdog record written by a kernel whose mbarrier wait expired (host pointer, may be NULL) */
OSPO_API const uint32_t* ospo_head_watchdog_record_host(void);
/* validation hook: out[M,N] fp32 = A B^T through one GEMM-engine variant
   (variant = cta_group*100 + majors*10 + tile; majors 0 K/K, 1 K/MN, 2 MN/MN; tile 0 BN256, 1 BN32, 2 BN128).
   K-major operand: [rows, K] pitch ld; MN-major operand: [K, rows] pitch ld. */
OSPO_API int ospo_head_gemm_debug(int variant, const void* a, int64_t lda, const void* b, int64_t ldb, float* out,
                         int64_t ldo, int32_t M, int32_t N, int32_t K, ospo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OSPO_HEAD_H_ */
