/*
 * lf_fusion.h — C ABI of the B200 (sm_100a) late-fusion training-step library.
 *
 * The reference (Nano1337/multimodal-clinical) has no FFI: its boundary for this path is the Python
 * module API.  The entry points below are what a ctypes binding underneath that API
 * needs; each cites the reference code it replaces (paths relative to the reference tree).
 * All pointers are DEVICE pointers unless stated otherwise; all matrices are dense row-major fp32.
 * `stream` is a cudaStream_t passed as void*.  Every call is asynchronous w.r.t. the host, returns
 * 0 on success or a negative LF_ERR_* code, never throws, and keeps no hidden state between calls:
 * outputs, state (History, EMA) and workspace are caller-owned.
 */
#ifndef LF_FUSION_H_
#define LF_FUSION_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LF_ABI_VERSION 12

/* error codes */
#define LF_OK 0
#define LF_ERR_BAD_ARG (-1)
#define LF_ERR_WORKSPACE (-2)
#define LF_ERR_CUDA (-3)
#define LF_ERR_UNSUPPORTED (-4)

/* head type: which loss the step assembles */
#define LF_MODE_JLOGITS 0 /* avg=(z1+z2)/2, L=CE(avg): cremad/joint_model_ogm_ge.py:50-58, enrico/joint_model.py:78-86 */
#define LF_MODE_QMF 1     /* L=CE(z_df)+CE(z1)+CE(z2)+L_reg: cremad/joint_model_qmf.py:57-75 */

/* arithmetic of the three head GEMMs */
#define LF_PREC_FP32 0 /* exact fp32; parity 1e-5.  Narrow heads: FMA kernels.  Wide heads (C >= 32) whose ld_dlogits is a non-zero
                          multiple of 4: the tensor-pipe kernels through the 3xTF32 operand split (ABI v11); else FMA GEMMs */
#define LF_PREC_TF32 1 /* tcgen05 kind::tf32 tensor pipe (wide heads); parity 2e-2 like the reference's bf16-mixed */
#define LF_PREC_BF16 2 /* the reference's own mode (Trainer precision "bf16-mixed", utils/run_trainer.py:47): features, dfeat
                          and the dL/dlogits scratch are bf16 in HBM (half the bytes), heads are cast to bf16 per step like
                          autocast does, tcgen05 kind::f16 with fp32 accumulation; logits, losses, statistics and head
                          gradients stay fp32.  Wide heads only (C >= 32); dim and ld_dlogits multiples of 8.  Parity 2e-2 */

/* OGM-GE modulation (existing_algos/OGM_GE.py:48-54) */
#define LF_MOD_OGM_GE 0 /* g <- k g + N(0, std(g)+1e-8) */
#define LF_MOD_OGM 1    /* g <- k g */
#define LF_MOD_NOISE 2  /* g <- g + N(0, std(g)+1e-8) */

/*
 * Packed per-step statistics, fp64, LF_STATS_HEADER + 2*C entries.  lf_heads_forward writes the LOCAL
 * (this shard's) sums; with several GPUs the host all-reduces (sum) the buffer before the calls that
 * consume it, which makes every mean a global-batch mean.
 */
#define LF_STAT_CE_JOINT 0  /* sum_b CE(avg_b) (JLOGITS) or sum_b CE(z_df_b) (QMF) */
#define LF_STAT_CE_X1 1     /* sum_b CE(z1_b)      cremad/joint_model_qmf.py:64 */
#define LF_STAT_CE_X2 2
#define LF_STAT_SCORE_X1 3  /* sum_b softmax(z1)[b,y_b]   existing_algos/OGM_GE.py:21 */
#define LF_STAT_SCORE_X2 4  /*                           existing_algos/OGM_GE.py:22 */
#define LF_STAT_CNT_X1 5    /* #argmax(z1)==y      utils/BaseModel.py:78 */
#define LF_STAT_CNT_X2 6    /*                     utils/BaseModel.py:79 */
#define LF_STAT_CNT_JOINT 7 /* #argmax(avg)==y     utils/BaseModel.py:92 */
#define LF_STAT_CNT_DF 8    /* #argmax(z_df)==y    utils/BaseModel.py:961 */
#define LF_STAT_CNT_X1_CAL 9  /* #argmax(z1+off1)==y  utils/BaseModel.py:88 (written by lf_heads_backward) */
#define LF_STAT_CNT_X2_CAL 10 /*                      utils/BaseModel.py:89 */
#define LF_STAT_REG_SUM 11  /* sum of the two ranking-loss relu sums (written by lf_qmf_history_step) */
#define LF_STATS_HEADER 16  /* [16, 16+C): sum_b z1[b,:]   [16+C, 16+2C): sum_b z2[b,:]  (utils/BaseModel.py:82-83) */

/* QMF loss-term ablations (cremad/joint_model_qmf_ablate_Ljoint.py:68, cremad/joint_model_qmf_ablate_Lunimodal.py:70):
   bits of LfHeadsArgs.loss_terms / LfMidArgs.loss_terms.  0 = the full loss CE(z_df) + sum CE(z_m) + L_reg.
   Dropped terms leave the forward quantities (logits, History update with the unimodal losses, statistics)
   untouched; they vanish from the reported loss and from dL/dz. */
#define LF_LOSS_NO_JOINT 1 /* loss_joint = 0 */
#define LF_LOSS_NO_UNI 2   /* the sum of the unimodal CE terms is dropped */
#define LF_LOSS_NO_REG 4   /* no ranking regulariser and no History: with LF_LOSS_NO_JOINT this is the ENSEMBLE loss
                              CE(z1) + CE(z2) of cremad/ensemble_model_noised.py:52-53 (one CE per modality); lf_step_mid
                              then needs neither idx nor the History arrays and leaves qmf_g (all zeros) alone */

/*
 * Optional SGD(momentum, weight decay) update of the head parameters fused into the tail of lf_heads_backward
 * (utils/BaseModel.py:275-285: torch.optim.SGD(lr, momentum=0.9, weight_decay=1e-4), dampening 0, no nesterov):
 *   d = g + wd*p;  buf = momentum*buf + d;  p -= lr*buf        (zero-initialised buf == torch's first-step buf = d)
 * applied in place to weight[] / bias[] right after dW / db have been reduced, in the same launch.  Single-GPU
 * steps only (batch_global == batch): a sharded step must all-reduce the gradients first (LfPeerReduceArgs.sgd).
 */
typedef struct LfSgdFused {
  const float* hyper;        /* device: [lr, momentum, weight_decay, unused], read at run time so a captured graph follows StepLR */
  float* momentum_buf[4];    /* W1, b1, W2, b2 momentum buffers (device), zero-initialised by the caller */
  void* weight_bf16_out[2];  /* optional: bf16 copies of the UPDATED weights (what LfHeadsArgs.weight_bf16 consumes) */
} LfSgdFused;

typedef struct LfHeadsArgs {
  int32_t batch;        /* B: samples in this shard */
  int32_t batch_global; /* denominator of every batch mean (== batch on one GPU) */
  int32_t dim;          /* D: feature width (multiple of 4) */
  int32_t classes;      /* C */
  int32_t mode;         /* LF_MODE_* */
  int32_t precision;    /* LF_PREC_* */
  int32_t need_dfeat;   /* 0: encoders frozen (enrico/joint_model.py:36-38), dfeat not produced */
  int32_t ld_dlogits;   /* row pitch (elements) of dlogits; 0 = classes.  LF_PREC_TF32 needs a multiple of 4 (TMA) */
  const float* feat[2];   /* (B,D) pooled encoder features  cremad/joint_model_qmf.py:48-55.  LF_PREC_BF16: bf16 data */
  const float* weight[2]; /* (C,D) x{1,2}_classifier.weight cremad/joint_model_qmf.py:26,28 */
  const float* bias[2];   /* (C)   x{1,2}_classifier.bias */
  const int64_t* label;   /* (B) */
  float* logits[2];       /* out (B,C) x1_logits, x2_logits, row pitch ld_logits */
  float* avg_logits;      /* out (B,C) (x1+x2)/2            cremad/joint_model_qmf.py:73 */
  float* logits_df;       /* out (B,C) QMF only             existing_algos/QMF.py:115-117 */
  float* conf;            /* out (2,B) QMF only: log(sum(exp z))/10  existing_algos/QMF.py:113-114 */
  float* dlogits[2];      /* scratch (B,ld_dlogits) each: dL/dz_m.  JLOGITS uses dlogits[0] only (dz1 == dz2).  LF_PREC_BF16: bf16 */
  float* dfeat[2];        /* out (B,D) dL/df_m, or NULL when need_dfeat == 0.  LF_PREC_BF16: bf16 data */
  float* dweight[2];      /* out (C,D) LOCAL-shard dL/dW_m (host all-reduces across GPUs) */
  float* dbias[2];        /* out (C) */
  const float* qmf_g;     /* in  (2,B) dL_reg/dconf from lf_qmf_history_step (QMF backward only) */
  const float* ema_offset;/* in  (2,C) offsets from lf_ema_update, for the calibrated counts */
  double* stats;          /* in/out packed statistics, see LF_STAT_* */
  void* workspace;        /* >= lf_workspace_bytes(batch, dim, classes) */
  size_t workspace_bytes;
  int32_t fwd_only;       /* 1: validation / test forward, lf_heads_backward will not be called for this step.
                             Lets the narrow-head path (C <= 32) fuse forward and backward into one pass otherwise. */
  int32_t bwd_phase;      /* lf_heads_backward: 0 = everything; 1 = dL/dz, dW, db, calibrated counts (all but dfeat);
                             2 = dfeat only (after a phase-1 call).  Lets the caller start the gradient exchange of
                             dW/db on a second stream while dfeat, which no other rank needs, is still being written.
                             3 = the row kernel only (QMF: dL/dz + db partials; both modes: calibrated counts, which need
                             this step's EMA offsets); 4 = dW GEMM + finalisation only (after 3, and 2 if dfeat is wanted).
                             With mean fusion dL/dz is final after the forward (cremad/joint_model_ogm_ge.py:54-56 has no
                             grid-wide term), so a caller may run lf_step_mid + phase 3 on a second stream beside phase 2
                             and join before phase 4.  Only where lf_heads_backward_splits_rows() says so. */
  int32_t ld_logits;      /* row pitch (elements) of logits[0], logits[1]; 0 = classes (dense).  A multiple of 4 lets the
                             tensor-pipe GEMM write them with TMA stores (1236-byte rows of a dense C = 309 cannot be). */
  int32_t ld_fused;       /* row pitch (elements) of avg_logits and logits_df; 0 = classes (dense).  A multiple of 4 (with 16-byte
                             aligned bases) lets the row kernels write them with 128-bit stores. */
  int32_t loss_terms;     /* QMF: LF_LOSS_* bits */
  int32_t reserved3;
  const void* weight_bf16[2]; /* LF_PREC_BF16, optional: bf16 copies of weight[0/1] that the caller keeps current
                                 (lf_cast_heads_bf16 after an optimizer step, or LfSgdFused.weight_bf16_out).  NULL:
                                 lf_heads_forward casts the heads itself on every call, like autocast does */
  const LfSgdFused* sgd;      /* optional (host pointer): fused SGD update in lf_heads_backward, see LfSgdFused */
  const struct LfPeerComm* grad_comm; /* optional (host pointer), sharded steps: lf_heads_backward all-reduces
                                 [dW1|dW2|db1|db2|calibrated counts|ranking-loss partial] over peer memory INSIDE the tail of
                                 the dW kernel (every CTA stores the chunk it has reduced into all peers' receive slots and
                                 polls the same chunk of every rank against the sentinel -- no fence, flag or barrier -- then
                                 sums in rank order), so dweight / dbias / stats come back as global sums and the fused SGD
                                 step may follow.  Only honoured when lf_heads_backward_fuses_allreduce() says so. */
  const float* reg_partial;   /* with grad_comm, QMF: this rank's ranking-loss partial (LfMidArgs.reg_partial_out) */
  float* loss_out;            /* with grad_comm, QMF: the loss lf_step_mid wrote without the ranking term; the term is added here */
  uint64_t* stats_rows_out;   /* optional (HOST pointer to 2 words): when the forward leaves its statistics as per-CTA
                                 partial rows (the fused tensor-pipe forward), lf_heads_forward skips the launch that sums
                                 them into `stats` and returns {device pointer of the float rows, number of rows} here for
                                 LfMidArgs.stats_rows; otherwise it writes {0, 0} and `stats` holds the sums as usual.  NULL: the
                                 forward always finishes the sums itself (what an NCCL all-gather of the statistics needs) */
} LfHeadsArgs;

/* 1 when lf_heads_backward(args) would run the gradient all-reduce inside the dW kernel (tensor-pipe heads whose dW
   tiles fit one wave), i.e. when LfHeadsArgs.grad_comm is honoured; 0: the caller exchanges dweight / dbias itself. */
int lf_heads_backward_fuses_allreduce(const LfHeadsArgs* args);

/* 1 when lf_heads_backward accepts bwd_phase 3 / 4 for these arguments (every shape whose dL/dz comes from the stand-alone
   row kernels: wide heads outside the fused QMF backward, exact-fp32 fallbacks); 0 for narrow heads and the fused backward. */
int lf_heads_backward_splits_rows(const LfHeadsArgs* args);

/* Floats per rank slot LfPeerComm.recv_grad must provide for that fused all-reduce ([dW1|dW2|db1|db2|cal x2|reg], 16-byte padded). */
size_t lf_grad_exchange_floats(int32_t dim, int32_t classes);

/* bf16 copies of two (n_each)-element fp32 tensors: out16 = [bf16(w0) | bf16(w1)] (the cast autocast does per step). */
int lf_cast_heads_bf16(const float* w0, const float* w1, void* out16, size_t n_each, void* stream);

/* Bytes of caller-provided scratch the heads calls need.  The workspace starts with LF_WS_SYNC_BYTES of inter-CTA
   counters: zero-fill the workspace ONCE after allocating it; every call leaves the counters zero. */
#define LF_WS_SYNC_BYTES 256
size_t lf_workspace_bytes(int32_t batch, int32_t dim, int32_t classes);

/*
 * Forward half: logits of both heads (a1), mean fusion (a2) or QMF energy fusion (a3), per-sample CE
 * terms, OGM-GE scores (a8), accuracy counts (a12) and the logit sums the EMA needs (a11), reduced into
 * `stats`.  In JLOGITS mode dL/dz is final here and is left in dlogits[0].
 * Replaces nn.Linear x2 + torch.stack/QMF.df + nn.CrossEntropyLoss x1..3 + the argmax/mean metrics.
 */
int lf_heads_forward(const LfHeadsArgs* args, void* stream);

/*
 * Backward half: dL/dz_m (QMF: needs stats with the global sums and qmf_g), dfeat = dz W, dW = dz^T f,
 * db = sum dz, plus the calibrated-accuracy counts (z_m + ema_offset[m]).
 * Replaces loss.backward() through the heads (autograd of the lines above).
 */
int lf_heads_backward(const LfHeadsArgs* args, void* stream);

/* loss_out[0] = total loss of the step from (all-reduced) stats.  cremad/joint_model_qmf.py:70 */
int lf_loss_finalize(const double* stats, int32_t mode, int32_t batch_global, float* loss_out, void* stream);

/*
 * EMA of the batch-mean logits and its offsets (utils/EMA.py:29-38; call site utils/BaseModel.py:82-85):
 * x <- smoothing*mean_b(z_m) + (1-smoothing)*x ; offset = mean_m(x) - x.  ema_x, ema_offset: (2,C).
 */
int lf_ema_update(float* ema_x, float* ema_offset, const double* stats, int32_t classes,
                  int32_t batch_global, float smoothing, void* stream);

/*
 * OGM-GE coefficients from the global score sums (existing_algos/OGM_GE.py:24-40).
 * coeff_out[0] scales x1_model's conv grads, coeff_out[1] x2_model's.
 */
int lf_ogm_coeff(const double* stats, float alpha, float* coeff_out, void* stream);

typedef struct LfQmfArgs {
  int32_t batch_global;  /* Bg >= 2: length of idx / conf rows (the gathered global batch) */
  int32_t n_data;        /* N: length of the History arrays (args.num_samples) */
  const int64_t* idx;    /* (Bg) dataset indices of the batch, global batch order */
  const float* conf;     /* (2,Bg) */
  double* correctness;   /* in/out (2,N) History.correctness  existing_algos/QMF.py:13 */
  double* confidence;    /* in/out (2,N) History.confidence   existing_algos/QMF.py:14 */
  int64_t* last_writer;  /* in/out (N [+1]) scratch owned by the History, zero-initialised once */
  int64_t step_base;     /* strictly increasing by >= Bg per call, starting at 1; 0 = use the device-resident
                            counter kept at last_writer[N] (then last_writer has N+1 entries) */
  double* stats;         /* in: global CE sums (LF_STAT_CE_X1/X2); out: LF_STAT_REG_SUM */
  float* qmf_g;          /* out (2,Bg) dL_reg/dconf (already divided by Bg) */
  float* target_out;     /* out (2,Bg) ranking targets in {-1,0,1}, or NULL */
  int32_t g_begin;       /* [g_begin, g_begin+g_count) slice of the global batch whose qmf_g rows are */
  int32_t g_count;       /* written at qmf_g[m*g_count + (j-g_begin)]; use 0,Bg for everything */
  void* workspace;       /* >= lf_qmf_workspace_bytes(n_data) */
  size_t workspace_bytes;
  int32_t flags;         /* LF_QMF_* bits: which parts run (the fused step passes LF_QMF_ALL) */
  int32_t reserved;
  const float* loss_uni[2]; /* optional device scalars: batch-mean CE of modality m handed to
                               History.correctness_update (QMF.py:20-29); NULL = stats[CE_Xm]/Bg */
} LfQmfArgs;

#define LF_QMF_UPDATE_X1 1 /* history[0].correctness_update   cremad/joint_model_qmf.py:65 */
#define LF_QMF_UPDATE_X2 2 /* history[1].correctness_update */
#define LF_QMF_REG 4       /* reg_loss + dL_reg/dconf         existing_algos/QMF.py:119-141 */
#define LF_QMF_ALL 7

size_t lf_qmf_workspace_bytes(int32_t n_data);

/*
 * QMF History update + ranking regulariser on the (global) batch:
 *   History.correctness_update  existing_algos/QMF.py:20-29   (scalar batch-mean loss, alpha = 0.1)
 *   History.get_target_margin   existing_algos/QMF.py:37-68   (global min/max over N)
 *   QMF.reg_loss                existing_algos/QMF.py:119-141 (closed form, incl. the flattened roll)
 */
int lf_qmf_history_step(const LfQmfArgs* args, void* stream);

/*
 * Peer-memory communicator for the batch-sharded step (lf_peer.cu): one process per GPU, every rank's
 * symmetric buffer (lf_comm_alloc) is mapped into every peer through CUDA IPC over NVLink / NVSwitch.
 * All pointers below are device pointers valid on THIS rank; index r addresses rank r's buffer.
 */
#define LF_MAX_RANKS 8
#define LF_PEER_FLAGS_BYTES (2 * LF_MAX_RANKS * 8)
typedef struct LfPeerComm {
  int32_t n_ranks;
  int32_t rank;
  void* flags[LF_MAX_RANKS];        /* rank r's flag array, LF_PEER_FLAGS_BYTES: 2 sets x LF_MAX_RANKS int64 (set 1: the stand-alone
                                       lf_peer_allreduce; the exchanges fused into lf_step_mid and the dW kernel need no flags) */
  void* recv_payload[LF_MAX_RANKS]; /* rank r's payload receive area: [2 parities][n_ranks][payload bytes]; filled with 0xFF bytes by
                                       the caller once (lf_comm_fill): the exchanges validate every word against that sentinel
                                       and re-arm the slots themselves */
  void* recv_grad[LF_MAX_RANKS];    /* rank r's gradient receive area: [2 parities][n_ranks][n_padded floats]; 0xFF-filled likewise */
  int64_t* epoch;                   /* local, device-resident: epoch[0] payload exchanges done, epoch[1] gradient exchanges */
  int32_t* error;                   /* device-visible (ideally pinned host) flag: set to 1 right before the kernel traps
                                       because a peer did not arrive within minutes */
} LfPeerComm;

typedef struct LfPeerReduceArgs {
  LfPeerComm comm;
  float* buf;          /* in: this rank's n floats; out: the sum over ranks (rank order, identical on every rank) */
  int32_t n;
  int32_t n_padded;    /* slot size in floats, multiple of 4, >= n */
  double* tail_dst;    /* optional: the last tail_n sums are also written here as doubles (calibrated counts) */
  int32_t tail_n;
  int32_t reserved;
} LfPeerReduceArgs;

int lf_comm_alloc(size_t bytes, void** ptr);                 /* cudaMalloc + zero fill (host call) */
int lf_comm_fill(void* ptr, int32_t byte_value, size_t bytes); /* cudaMemset + sync (host call): the receive areas start as 0xFF */
int lf_comm_free(void* ptr);
int lf_comm_ipc_handle(void* ptr, void* handle64);           /* 64-byte CUDA IPC handle of an lf_comm_alloc buffer */
int lf_comm_ipc_open(const void* handle64, void** ptr);      /* map a peer's buffer (enables peer access lazily) */
int lf_comm_ipc_close(void* ptr);
/* One-shot all-reduce(sum) of the packed head gradients [dW1|db1|dW2|db2|calibrated counts] over peer memory. */
int lf_peer_allreduce(const LfPeerReduceArgs* args, void* stream);

/*
 * The middle of the step in ONE launch (lf_mid.cu): sums the per-rank partial statistics in rank order,
 * updates the EMA (utils/EMA.py:29-38), computes the OGM-GE coefficients (existing_algos/OGM_GE.py:24-40),
 * and -- QMF -- runs History.correctness_update, the global min/max, the ranking targets, the ranking
 * loss and dL_reg/dconf (existing_algos/QMF.py:20-68, 119-141), then the total loss
 * (cremad/joint_model_qmf.py:70).  Inputs are rank-major: the buffer a single all-gather of
 * [stats | idx | conf] produces; with one GPU n_ranks = 1 and the pointers are the local buffers.
 * Supersedes the sequence lf_ema_update, lf_ogm_coeff, lf_qmf_history_step, lf_loss_finalize.
 */
typedef struct LfMidArgs {
  int32_t mode;            /* LF_MODE_* */
  int32_t classes;
  int32_t batch_global;    /* n_ranks * batch_local */
  int32_t n_ranks;
  int32_t batch_local;
  int32_t rank;            /* qmf_g is produced for samples [rank*batch_local, (rank+1)*batch_local) */
  int32_t n_data;          /* QMF: length N of the History arrays */
  int32_t update_ema;      /* 0 on validation / test steps (utils/BaseModel.py:133-160) */
  const double* stats_parts; /* rank r's LF_STATS_HEADER + 2C partial statistics at stats_parts + r*stats_stride */
  int64_t stats_stride;      /* in doubles */
  const int64_t* idx_parts;  /* QMF: rank r's idx (batch_local) at idx_parts + r*idx_stride (elements) */
  int64_t idx_stride;
  const float* conf_parts;   /* QMF: rank r's conf (2, batch_local) at conf_parts + r*conf_stride (elements) */
  int64_t conf_stride;
  double* stats;           /* out: global statistics (CNT_*_CAL entries are left alone; REG_SUM is written) */
  float* ema_x;            /* in/out (2,C) */
  float* ema_offset;       /* out (2,C) */
  float smoothing;
  float alpha;             /* OGM-GE alpha; used when coeff_out != NULL */
  float* coeff_out;        /* out (2) or NULL */
  double* correctness;     /* QMF in/out (2,N) */
  double* confidence;      /* QMF in/out (2,N) */
  int64_t* last_writer;    /* QMF in/out (N) tickets, zero-initialised once */
  int64_t step_base;       /* QMF: as in LfQmfArgs; 0 = device-resident counter at last_writer[N] (graph-replayable) */
  float* qmf_g;            /* QMF out (2, batch_local) dL_reg/dconf of this rank's samples, or NULL (forward only) */
  float* loss_out;         /* out (1) total loss, or NULL */
  void* workspace;         /* QMF: >= lf_mid_workspace_bytes(batch_global), ZERO-INITIALISED once by the caller (holds an
                              inter-CTA arrival counter that every launch leaves at zero) */
  size_t workspace_bytes;
  /* optional fused exchange: when use_peer != 0 the kernel first pushes this rank's payload (payload_bytes,
     multiple of 16, laid out [stats | idx | conf] like the gathered buffer) into every peer's receive area and
     waits for all peers, then consumes the LOCAL receive area; stats_parts / idx_parts / conf_parts are
     ignored (their byte offsets inside the payload are off_idx / off_conf). */
  int32_t use_peer;
  int32_t loss_terms;      /* QMF: LF_LOSS_* bits, applied to loss_out */
  const void* payload_local;
  int64_t payload_bytes;
  int64_t off_idx;
  int64_t off_conf;
  LfPeerComm comm;
  float* reg_partial_out;    /* optional, QMF sharded steps: the ranking terms are evaluated for this rank's slice only and
                                their sum is written here (a later exchange adds the ranks' partials, LfHeadsArgs.reg_partial);
                                stats[LF_STAT_REG_SUM] is left alone and loss_out gets the loss WITHOUT the ranking term.
                                NULL: every rank walks the pairs of the whole global batch itself */
  const int64_t* payload_idx_src; /* optional with use_peer: the idx part of the payload is pushed from here (the caller's
                                index tensor) instead of payload_local + off_idx */
  const float* stats_rows;   /* optional (QMF; one rank, or use_peer): per-CTA partial rows [n_stats_rows][LF_STATS_HEADER + 2C] of THIS */
  int64_t n_stats_rows;      /* rank's forward (LfHeadsArgs.stats_rows_out), summed in row order instead of reading stats_parts;
                                with use_peer the ranks' column sums are exchanged inside the kernel and added in rank order */
} LfMidArgs;

size_t lf_mid_workspace_bytes(int32_t batch_global);

/*
 * Feature-side pooling in front of the heads (cremad/joint_model_qmf.py:48-55):
 *   pooled[b, c] = mean over t < frames, k < hw of maps[(b*frames + t), c, k]       maps: (batch*frames, channels, hw) contiguous
 * i.e. F.adaptive_avg_pool2d(a, 1) for frames = 1 and view(B, T, C, H, W).permute(0, 2, 1, 3, 4) + adaptive_avg_pool3d(v, 1)
 * for the frame-stacked visual stream.  elem_bytes = 4 (fp32) or 2 (bf16); fp32 accumulation, output in the input's type.
 * lf_pool_mean_backward writes dmaps[(b*frames + t), c, k] = dpooled[b, c] / (frames * hw).
 */
int lf_pool_mean(const void* maps, void* pooled, int32_t batch, int32_t frames, int32_t channels, int32_t hw, int32_t elem_bytes, void* stream);
int lf_pool_mean_backward(const void* dpooled, void* dmaps, int32_t batch, int32_t frames, int32_t channels, int32_t hw, int32_t elem_bytes, void* stream);

/*
 * Epoch-end unimodal offset correction over all logits collected during a validation / test epoch
 * (utils/BaseModel.py:168-185): logits (n, 2, classes) fp32 contiguous, labels (n) int64 ->
 *   offset_out (2, classes) = mean_m(mean_n logits) - mean_n logits
 *   acc_out[4] (fp64)       = accuracies of x1 / x2 uncorrected, x1 / x2 with the offset added
 * No host synchronisation.  workspace: >= lf_epoch_workspace_bytes(classes), zero-initialised once by the caller.
 */
size_t lf_epoch_workspace_bytes(int32_t classes);
int lf_epoch_offset_correction(const float* logits, const int64_t* labels, int64_t n, int32_t classes, float* offset_out,
                               double* acc_out, void* workspace, size_t workspace_bytes, void* stream);


int lf_step_mid(const LfMidArgs* args, void* stream);

typedef struct LfTensorList {
  int32_t count;          /* number of gradient tensors (<= LF_MAX_TENSORS) */
  int32_t reserved;
  float* data[64];        /* device pointers of the 4-D .grad tensors of one encoder */
  int64_t numel[64];
} LfTensorList;
#define LF_MAX_TENSORS 64

size_t lf_modulate_workspace_bytes(void);

/*
 * OGM-GE add_factor over the 4-D gradients of one encoder (existing_algos/OGM_GE.py:42-54):
 * sigma_t = unbiased std of tensor t (before scaling) + 1e-8; g <- coeff*g + sigma_t*xi (OGM_GE),
 * coeff*g (OGM), g + sigma_t*xi (NOISE); xi ~ N(0,1) from Philox4x32-10 keyed by (seed, offset).
 * coeff_dev points at ONE float on the device (lf_ogm_coeff output), so no host sync is needed.
 */
int lf_ogm_modulate(const LfTensorList* list, const float* coeff_dev, int32_t mode, uint64_t seed,
                    uint64_t offset, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Evidence hooks for bench.py (no reference counterpart): number of kernels this library has launched
 * since it was loaded, and optional CUDA-event timing of every launch (events recorded on the launch
 * stream around each kernel).  lf_profile_report synchronises the device and writes one
 * "name count total_ms" line per kernel name into buf, clearing the records.
 */
int64_t lf_launch_count(void);
void lf_profile_enable(int32_t on);
int32_t lf_profile_report(char* buf, int32_t buf_bytes);

/*
 * Test hook (no reference counterpart): ONE tensor-pipe (tcgen05 kind::tf32 + TMA) GEMM,
 * out(M,N) = A*B (+bias), so the kernel behind LF_PREC_TF32 can be checked in isolation.
 *   a_mn_major = 0: A is (M,K) row-major, pitch lda.   1: A is stored (K,M) row-major (A = stored^T)
 *   b_mn_major = 0: B is stored (N,K) row-major, pitch ldb (out = A*stored^T).   1: B is (K,N) row-major
 * block_n: N tile (multiple of 16; 32 if b_mn_major), splits: split-K factor writing partial outputs
 * split_stride elements apart.
 */
int lf_debug_tc_gemm(const float* A, const float* B, const float* bias, float* out, int32_t M, int32_t N,
                     int32_t K, int32_t lda, int32_t ldb, int32_t ld_out, int32_t a_mn_major,
                     int32_t b_mn_major, int32_t block_n, int32_t splits, int64_t split_stride, void* stream);
/* Same hook through the 3xTF32 split (hi/lo tf32 halves formed in shared memory, three MMAs per k-step): the kernel
 * behind LF_PREC_FP32 on wide heads (ABI v11).  Parity class 1e-5. */
int lf_debug_tc_gemm_x3(const float* A, const float* B, const float* bias, float* out, int32_t M, int32_t N,
                        int32_t K, int32_t lda, int32_t ldb, int32_t ld_out, int32_t a_mn_major,
                        int32_t b_mn_major, int32_t block_n, int32_t splits, int64_t split_stride, void* stream);

/*
 * Stand-alone pieces of the reference's algorithm API, for callers that use existing_algos/ directly
 * instead of the fused step.
 *   lf_qmf_df      QMF.df (existing_algos/QMF.py:109-117): conf_m = log(sum(exp z_m))/10 (non-stabilised,
 *                  like the reference), z_df = sum_m z_m * conf_m.   z1,z2,zdf: (B,C); conf: (2,B).
 *   lf_ogm_scores  the two score sums of ogm_ge (existing_algos/OGM_GE.py:21-22) written to
 *                  stats[LF_STAT_SCORE_X1/X2]; deterministic two-stage reduction.
 *                  workspace >= lf_ogm_scores_workspace_bytes(), zero-initialised ONCE by the caller.
 */
int lf_qmf_df(const float* z1, const float* z2, int32_t batch, int32_t classes, float* zdf, float* conf, void* stream);
size_t lf_ogm_scores_workspace_bytes(void);
int lf_ogm_scores(const float* z1, const float* z2, const int64_t* label, int32_t batch, int32_t classes,
                  double* stats, void* workspace, size_t workspace_bytes, void* stream);

/* Same test hook for bf16 operands (kind::f16): A, B bf16; out fp32, or bf16 when out_bf16 != 0. */
int lf_debug_tc_gemm16(const void* A, const void* B, void* out, int32_t M, int32_t N, int32_t K, int32_t lda,
                       int32_t ldb, int32_t ld_out, int32_t a_mn_major, int32_t b_mn_major, int32_t block_n,
                       int32_t splits, int64_t split_stride, int32_t out_bf16, void* stream);

/*
 * SGD(momentum, weight decay) on the head parameters in one launch (utils/BaseModel.py:275-285:
 * torch.optim.SGD(lr, momentum=0.9, weight_decay=1e-4), dampening 0, no nesterov):
 *   d = g + wd*p;  buf = first_step ? d : momentum*buf + d;  p -= lr*buf.   Up to 8 tensors (W1,b1,W2,b2).
 */
typedef struct LfSgdArgs {
  int32_t count;
  int32_t first_step;      /* 1: momentum buffers are initialised with d (torch's first-step behaviour) */
  float lr, momentum, weight_decay;
  int32_t reserved;
  float* param[8];
  const float* grad[8];
  float* momentum_buf[8];
  int64_t numel[8];
} LfSgdArgs;
int lf_sgd_heads(const LfSgdArgs* args, void* stream);

/*
 * Mean fusion of M = 2..4 narrow heads with PER-MODALITY feature widths, forward and backward in one pass (ABI v11;
 * SURVEY.md 8f rank 4).  Replaces, after the encoders:
 *   mustard/joint_model.py:72-83   three heads (LstmClassifier.fc3, 100 -> C), avg = (z1 + z2 + z3) / 3, CE(avg, y)
 *   avmnist/joint_model.py:128-138 two heads of different widths (48 -> C, 192 -> C), avg = (z1 + z2) / 2, CE(avg, y)
 * and autograd's backward through them.  Exact fp32 (parity 1e-5), classes <= 32, bit-reproducible.
 *   stats[0] = sum of the per-sample CE, stats[1] = #(argmax avg == y), stats[2 + m] = #(argmax z_m == y)
 */
#define LF_MAX_MODALITIES 4
typedef struct LfMultiHeadsArgs {
  int32_t modalities;                       /* M */
  int32_t batch, classes;
  int32_t need_dfeat;
  int32_t dim[LF_MAX_MODALITIES];           /* D_m */
  const float* feat[LF_MAX_MODALITIES];     /* (B, D_m) row-major, dense */
  const float* weight[LF_MAX_MODALITIES];   /* (C, D_m) */
  const float* bias[LF_MAX_MODALITIES];     /* (C) */
  const int64_t* label;                     /* (B) */
  float* logits[LF_MAX_MODALITIES];         /* out (B, C) */
  float* avg_logits;                        /* out (B, C) */
  float* dweight[LF_MAX_MODALITIES];        /* out (C, D_m) */
  float* dbias[LF_MAX_MODALITIES];          /* out (C) */
  float* dfeat[LF_MAX_MODALITIES];          /* out (B, D_m), or NULL when need_dfeat == 0 */
  float* loss_out;                          /* out: batch-mean CE */
  double* stats;                            /* out [2 + M] */
  void* workspace;                          /* >= lf_multi_heads_workspace_bytes(M, C, sum D_m); contents irrelevant */
  size_t workspace_bytes;
} LfMultiHeadsArgs;
size_t lf_multi_heads_workspace_bytes(int32_t modalities, int32_t classes, int32_t dim_total);
int lf_multi_heads_step(const LfMultiHeadsArgs* args, void* stream);

/*
 * Hidden layers of the Food101 per-modality MLPs (ABI v11): the two modalities' layers of one shape,
 *   forward   h = dropout_p(relu(x W^T + b))        food101/joint_model_qmf.py:15-20 (nn.Linear, nn.ReLU, nn.Dropout(0.2))
 *   backward  dP = dh * [h > 0] / (1 - p);  dx = dP W;  dW = dP^T x;  db = sum_rows dP      (autograd of the same)
 * on the tensor-pipe GEMM kernel (precision: LF_PREC_BF16 = x, h, dh, dpre, dx bf16 with fp32 accumulation, what the
 * bf16-mixed trainer runs; LF_PREC_TF32 / LF_PREC_FP32 = fp32 tensors, single-pass TF32 / 3xTF32).  Bias, ReLU and the
 * dropout mask (Philox4x32-10: element e = row * dim_out + col draws 16-bit half e & 7 of the block with counter
 * (e >> 3, offset) and key seed; kept when the half >= round(p * 65536)) are applied in the forward GEMM's epilogue; no mask is stored.
 * training == 0: dropout is the identity (nn.Dropout in eval mode).  dx may be NULL for both layers (frozen input).
 */
typedef struct LfHiddenArgs {
  int32_t batch, dim_in, dim_out;   /* dims multiples of 8 */
  int32_t precision;                /* LF_PREC_* */
  int32_t training;
  float drop_p;
  uint64_t seed, offset;
  const void* x[2];                 /* (B, dim_in) */
  const float* weight[2];           /* (dim_out, dim_in) fp32 master weights */
  const float* bias[2];             /* (dim_out) */
  void* h[2];                       /* out (B, dim_out) */
  const void* dh[2];                /* backward: (B, dim_out) */
  void* dpre[2];                    /* backward scratch (B, dim_out), same element type as h */
  void* dx[2];                      /* backward out (B, dim_in) or NULL */
  float* dweight[2];                /* backward out (dim_out, dim_in) fp32 */
  float* dbias[2];                  /* backward out (dim_out) fp32 */
  void* workspace;                  /* >= lf_hidden_workspace_bytes(batch, dim_in, dim_out) */
  size_t workspace_bytes;
} LfHiddenArgs;
size_t lf_hidden_workspace_bytes(int32_t batch, int32_t dim_in, int32_t dim_out);
int lf_hidden_forward(const LfHiddenArgs* args, void* stream);
int lf_hidden_backward(const LfHiddenArgs* args, void* stream);

/* Last error message of the calling thread (host string). */
const char* lf_last_error(void);
int32_t lf_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LF_FUSION_H_ */
