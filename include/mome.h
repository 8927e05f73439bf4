/* libmome — C ABI of the B200-native VLMo MoME hot path.
 *
 * The reference (fanzhongyi/ExploreMultiModal) is pure Python: its "FFI" for this path is the set of
 * ATen / cuBLAS / NCCL calls its nn.Modules make. Each entry point below replaces one of those call
 * groups; the citation names the reference lines (relative to the reference root) it stands in for.
 * Host code (exploremultimodal_b200/*.py) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions: plain pointers (device memory unless noted), explicit sizes, the CUDA stream as an
 * opaque pointer (cudaStream_t). No allocation, no host synchronisation, no implicit stream.
 * Return 0 on success, otherwise a MomeStatus; mome_last_error() gives the text (thread local).
 * dtype codes: 0 = fp32, 1 = bf16. fp32 runs on CUDA cores (validation mode, 1e-4 parity);
 * bf16 runs the tcgen05/TMEM/TMA kernels.
 */
#ifndef MOME_H_
#define MOME_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOME_ABI_VERSION 6
#define MOME_MAX_GROUPS 4

enum MomeStatus { MOME_OK = 0, MOME_ERR_ARG = 1, MOME_ERR_CUDA = 2, MOME_ERR_UNSUPPORTED = 3 };
enum MomeDtype { MOME_F32 = 0, MOME_BF16 = 1 };

int mome_version(void);
const char* mome_last_error(void);
/* number of SMs of the current device (grid sizing; 148 on B200) */
int mome_sm_count(void);

/* Dropout of a branch in the backward kernels: `branch` is the stored, already dropped forward value; the
 * mask of element (row0 + r, c) is regenerated (csrc/dropout.cuh). row_scale: stochastic-depth multipliers
 * indexed by global row, or NULL. Pass drop = NULL for no dropout and no row scaling. */
typedef struct {
  const uint32_t* seed; /* device scalar, NULL = no element dropout */
  const float* row_scale;
  int64_t row0;
  uint32_t salt;
  float p;
} MomeDropout;

/* ---- K3: LayerNorm and the HBM-bound epilogues -------------------------------------------------
 * reference: vlmo.py:26-36 (LayerNorm factory), vlmo.py:188,192,196 (norm1/norm2), vlmo.py:355,376,
 * 386,413 (final norm); eps 1e-12 from vlmo_module.py:21-23. */
int mome_ln_fwd(const float* x, const float* weight, const float* bias, void* y, int y_dtype, float* mean,
                float* rstd, int64_t rows, int64_t d, float eps, void* stream);
/* dx_out = (dres ? dres : 0) + LN'(dy); dweight/dbias are ACCUMULATED (+=) in fp32. `ws`: see
 * mome_reduce_ws_bytes below. */
int mome_ln_bwd(const void* dy, int dy_dtype, const float* x, const float* mean, const float* rstd,
                const float* weight, const float* dres, float* dx_out, float* dweight, float* dbias,
                int64_t rows, int64_t d, void* ws, size_t ws_bytes, void* stream);
/* mome_ln_bwd fused with the LayerScale backward of the branch that produced this residual stream
 * (x1 = x + gamma * branch, reference vlmo.py:194): besides dx_out it writes dbranch = gamma * dx_out
 * (branch dtype) and accumulates dgamma += sum_rows dx_out * branch, dbias_branch += sum_rows dbranch.
 * `dtype` is the dtype of dy, branch and dbranch. gamma / dgamma / dbias_branch may be NULL. */
int mome_ln_bwd_scale(const void* dy, int dtype, const float* x, const float* mean, const float* rstd,
                      const float* weight, const float* dres, float* dx_out, float* dweight, float* dbias,
                      const void* branch, const float* gamma, void* dbranch, float* dgamma, float* dbias_branch,
                      int64_t rows, int64_t d, const MomeDropout* drop, void* ws, size_t ws_bytes, void* stream);
/* LayerScale backward (reference vlmo.py:194-196, `x + gamma * branch`):
 *   dbranch = gamma * dx (cast to dbranch_dtype); dgamma += sum_rows dx * branch;
 *   dbias += sum_rows dbranch (bias of the Linear that produced `branch`). gamma may be NULL (=1). */
int mome_scale_bwd(const float* dx, const void* branch, int branch_dtype, const float* gamma, void* dbranch,
                   int dbranch_dtype, float* dgamma, float* dbias, int64_t rows, int64_t d, const MomeDropout* drop,
                   void* ws, size_t ws_bytes, void* stream);
/* out[row] = stochastic-depth multiplier of row's sample: 0 with probability p, else 1 / (1 - p); one draw per
 * (sample, salt). row_sample[row] = index of the sequence / sample the row belongs to. reference: timm DropPath. */
int mome_droppath_scales(const int32_t* row_sample, int64_t rows, const uint32_t* seed, uint32_t salt, float p,
                         float* out, void* stream);
/* out[j] += sum_rows x[r, j]  (bias gradients of qkv / fc1) */
int mome_colsum(const void* x, int dtype, int64_t rows, int64_t cols, int64_t ld, float* out, void* ws, size_t ws_bytes,
                void* stream);
/* out[j] += sum_p partials[p * cols + j] (second stage of a GEMM epilogue's fused column sums) */
int mome_colreduce(const float* partials, int64_t nparts, int64_t cols, float* out, void* stream);
/* Column reductions (dweight, dbias, dgamma, colsum) are two-stage and atomic-free: stage 1 writes per-CTA
 * partial sums into the caller's workspace `ws`, stage 2 adds them into the outputs, both inside the
 * call. mome_reduce_ws_bytes(cols) is enough for any of the calls above with d (or cols) <= `cols`. */
size_t mome_reduce_ws_bytes(int64_t cols);
/* fp32 -> bf16 cast (weights, once per optimizer step) */
int mome_cast_bf16(const float* src, void* dst, int64_t n, void* stream);

/* ---- K2: grouped GEMM with fused epilogues --------------------------------------------------------
 * reference: F.linear at vlmo.py:76-78 (qkv), vlmo.py:96 (proj), timm Mlp fc1/fc2 selected by
 * `self.mlp[route]` at vlmo.py:192/196; the residual / LayerScale adds at vlmo.py:190-196; and the
 * autograd backward of those (dgrad, wgrad).
 *
 *   out[m, n] = epilogue( sum_k A(m, k) * B(n, k) )      for each group g independently.
 *
 * Operand storage: major 0 = "K-major": element (mn, k) at base[mn * ld + k];
 *                  major 1 = "MN-major": element (mn, k) at base[k * ld + mn].
 * forward  y = x W^T      : A = x  (K-major),  B = W  (K-major)
 * dgrad    dx = dy W      : A = dy (K-major),  B = W  (MN-major)
 * wgrad    dW = dy^T x    : A = dy (MN-major), B = x  (MN-major), contraction over the group's rows.
 * Groups are the expert segments of the packed token buffer (text rows -> 'l', image rows -> 'v',
 * fused rows -> 'vl'): same N/K/epilogue, per-group pointers and row counts. */
enum MomeEpilogue {
  MOME_EPI_STORE = 0,    /* out = acc + bias                                   (qkv, dgrad) */
  MOME_EPI_GELU = 1,     /* z = acc + bias; out = gelu_erf(z), out2 = gelu_erf'(z)      (fc1) */
  MOME_EPI_RESIDUAL = 2, /* out2 = b = acc + bias; out(fp32) = res + gamma * b  (proj, fc2) */
  MOME_EPI_DGELU = 3,    /* out = acc * aux, aux = the saved gelu_erf'(z)  (fc2 dgrad -> dz) */
  MOME_EPI_ATOMIC = 4    /* out(fp32) += acc via red.add (split-K wgrad)                     */
};

typedef struct {
  const void* a;
  const void* b;
  int64_t M;         /* rows of out for this group */
  int64_t K;         /* contraction length for this group */
  void* out;
  void* out2;        /* may be NULL */
  const float* bias; /* [N] or NULL */
  const float* res;  /* fp32 [M, ldres] (RESIDUAL) */
  const void* aux;   /* operand dtype [M, ldaux] (DGELU) */
  int64_t row0;      /* row of the packed token buffer this group starts at (dropout masks and row_scale are
                        indexed by that global row, so forward and backward agree however rows are grouped) */
  float* colsum;     /* optional (STORE / DGELU), fp32 [ceil(M/32), N]: row i receives the column sums of the stored out
                        rows [32 i, 32 i + 32); mome_colreduce adds the rows (bias gradient). The bf16 path writes
                        every element once; the fp32 validation path accumulates, so ZERO it there */
} MomeGemmGroup;

typedef struct {
  int32_t dtype;     /* operand dtype of A, B, out2, aux */
  int32_t a_major, b_major;
  int32_t epilogue;
  int32_t out_dtype; /* dtype of out */
  int32_t num_groups;
  int32_t split_k;   /* 0 = choose automatically (only with MOME_EPI_ATOMIC) */
  int32_t reserved;
  int64_t N;
  int64_t lda, ldb, ldo, ldo2, ldres, ldaux; /* in elements */
  const float* gamma; /* [N] or NULL (RESIDUAL) */
  MomeGemmGroup group[MOME_MAX_GROUPS];
  /* Dropout (bf16 path; reference vlmo.py:97 proj_drop, timm Mlp drop, DropPath at vlmo.py:194-196). NULL seed = off.
   *   GELU    : out = drop(gelu(z)), out2 = gelu'(z) * mask            (timm Mlp's dropout after the activation)
   *   RESIDUAL: out2 = b = drop(acc + bias); out = res + gamma * row_scale[row] * b
   * Masks are a pure function of (*drop_seed, drop_salt, global row, column): see csrc/dropout.cuh. */
  const uint32_t* drop_seed; /* device scalar */
  const float* row_scale;    /* RESIDUAL: per-row multiplier of the branch (stochastic depth), [tokens] or NULL */
  uint32_t drop_salt;
  float drop_p;
} MomeGemmArgs;

int mome_gemm(const MomeGemmArgs* args, void* stream);

/* ---- K1: masked multi-head self-attention over packed segments -----------------------------------
 * reference: vlmo.py:79-95 (split heads, q k^T * scale, masked_fill(~mask, -inf), softmax, @ v,
 * merge heads). qkv is the [tokens, 3*d] output of the qkv GEMM (column = s*d + h*64 + e).
 * A sequence is up to two row ranges of the packed buffer ([text | image] after the fusion layer,
 * one range before it): seq_desc[4*s + {0,1,2,3}] = {start0, len0, start1, len1}, len0 + len1 >= 1 (no empty
 * sequences) and <= max_seq_len.
 * key_mask[row] = 1 keeps the key, 0 excludes it; query rows are never masked. head_dim is 64.
 * drop_seed (device scalar, NULL = off) / drop_salt / drop_p: dropout on the attention probabilities (vlmo.py:93),
 * bf16 path only; the backward must be given the same three values (masks are regenerated, csrc/dropout.cuh). */
int mome_attn_fwd(const void* qkv, int dtype, const int32_t* seq_desc, const uint8_t* key_mask, void* out,
                  float* lse, int64_t tokens, int32_t num_seqs, int32_t max_seq_len, int32_t num_heads,
                  float scale, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, void* stream);
/* dqkv gets every element of its [tokens, 3*d] rows written. delta_ws: scratch of mome_attn_bwd_ws_floats(...) floats
 * (rowsum(dO o O) per (sequence, head, query); for sequences longer than 256 tokens also the fp32 dK / dV accumulators
 * the query-tile pairs of a sequence add into). */
int64_t mome_attn_bwd_ws_floats(int64_t tokens, int32_t num_seqs, int32_t max_seq_len, int32_t num_heads);
int mome_attn_bwd(const void* qkv, const void* out, const void* dout, int dtype, const int32_t* seq_desc,
                  const uint8_t* key_mask, const float* lse, void* dqkv, float* delta_ws, int64_t tokens,
                  int32_t num_seqs, int32_t max_seq_len, int32_t num_heads, float scale, const uint32_t* drop_seed,
                  uint32_t drop_salt, float drop_p, void* stream);

/* ---- K4: ITC head — similarity GEMM fused with softmax cross-entropy -------------------------------
 * reference: objectives.py:99-108 (global-reduce logits), 166-171 (naive), 173-180 (CE + accuracy),
 * heads.py:125-126 (L2 normalise). feats are fp32 [rows, dim]; `all_*` hold the gathered features of
 * every rank in rank order ([world*bs, dim]); this rank's rows start at rank*bs. Outputs per
 * direction (0 = i2t, 1 = t2i): loss_sum[dir] (sum over local rows of lse - logit[target]),
 * correct[dir] (argmax over the local block == target), lse[dir*bs + r], and the local [bs, bs]
 * logit block sim_local[dir] that ITM's hard-negative mining consumes (objectives.py:251-255).
 * `temp` (= exp(itc_temp)) and `gscale` are DEVICE scalars so that no host synchronisation is needed. */
int mome_l2norm_fwd(const void* x, int x_dtype, float* y, float* inv_norm, int64_t rows, int64_t dim,
                    void* stream);
int mome_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int64_t rows,
                    int64_t dim, void* stream);
int mome_itc_fwd(const float* i_feat, const float* t_feat, const float* all_i, const float* all_t,
                 const float* temp, int32_t bs, int32_t world, int32_t rank, int32_t dim, float* loss_sum, int32_t* correct,
                 float* lse, float* sim_local, void* stream);
/* Gradients of mean-CE ((i2t + t2i) / 2, each a mean over bs rows) scaled by gscale:
 *   d_i_feat, d_t_feat [bs, dim] (local-row terms), d_all_i, d_all_t [world*bs, dim] (column terms, to
 *   be reduce-scattered across ranks by the caller), d_temp (scalar, accumulated). */
int mome_itc_bwd(const float* i_feat, const float* t_feat, const float* all_i, const float* all_t,
                 const float* temp, int32_t bs, int32_t world, int32_t rank, int32_t dim, const float* lse,
                 const float* gscale,
                 float* d_i_feat, float* d_t_feat, float* d_all_i, float* d_all_t, float* d_temp, void* stream);

/* In-kernel gather variants of the two calls above: instead of gathered arrays they take `peers`, a DEVICE
 * array of `world` pointers, peers[r] = rank r's fp32 [2, bs, dim] feature buffer (0 = image, 1 = text)
 * mapped into this process (NVLink peer access / symmetric memory). The kernels load the remote rows
 * directly, which replaces GatherLayer.forward's all_gather + cat + roll (objectives.py:401-414, 102-105);
 * the caller separates the ranks' buffer writes from these reads with barriers. */
int mome_itc_fwd_peer(const float* i_feat, const float* t_feat, const void* peers, const float* temp, int32_t bs,
                      int32_t world, int32_t rank, int32_t dim, float* loss_sum, int32_t* correct, float* lse,
                      float* sim_local, void* stream);
int mome_itc_bwd_peer(const float* i_feat, const float* t_feat, const void* peers, const float* temp, int32_t bs,
                      int32_t world, int32_t rank, int32_t dim, const float* lse, const float* gscale, float* d_i_feat,
                      float* d_t_feat, float* d_all_i, float* d_all_t, float* d_temp, void* stream);

/* The cross-rank gather on its own: all_i / all_t [world*bs, dim] <- peers[r] (rank r's [2, bs, dim] buffer), 128-bit
 * peer loads over NVLink, every remote element read once (reference GatherLayer.forward, objectives.py:401-414:
 * all_gather + cat). Followed by mome_itc_fwd / mome_itc_bwd on the gathered arrays. */
int mome_itc_gather_peer(const void* peers, int32_t bs, int32_t world, int32_t dim, float* all_i, float* all_t, void* stream);

/* ---- One whole MoME block per call ------------------------------------------------------------------
 * reference: Block.forward, vlmo.py:187-197, and its autograd backward. These two entry points only
 * sequence the kernels above (7 launches forward, 14 backward) in native code, so that a host language
 * with expensive FFI calls (Python: ~20 ctypes calls + as many tensor allocations per block otherwise)
 * issues one call per block and direction. All buffers are caller-owned; shapes: tokens x d unless noted.
 *
 * forward : x -> LN1 -> qkv GEMM -> attention -> proj GEMM (+gamma_1, residual) = x1
 *             -> LN2 -> fc1 GEMM (+GELU, per expert group) -> fc2 GEMM (+gamma_2, residual) = x2
 * backward: dx2 -> ... -> dx; every parameter gradient is ACCUMULATED (+=) into the given fp32 buffer. */
typedef struct {
  int64_t first_row, rows;   /* expert segment of the packed token buffer */
  const void* w1;            /* fc1.weight [hid, d]  (compute dtype) */
  const float* b1;           /* fc1.bias   [hid] */
  const void* w2;            /* fc2.weight [d, hid]  (compute dtype) */
  const float* b2;           /* fc2.bias   [d] */
  float *dw1, *db1, *dw2, *db2; /* backward: gradient accumulators */
  float* colsum_part;        /* backward scratch: fp32 [ceil(rows / 32), hid] (zeroed on the fp32 path, see MomeGemmGroup.colsum) */
} MomeBlockGroup;

typedef struct {
  int32_t dtype;             /* MomeDtype of activations / weights fed to the GEMMs */
  int32_t num_heads, num_groups, num_seqs, max_seq_len, reserved;
  int64_t tokens, d, hid;
  float eps, scale;          /* LayerNorm eps; attention scale (head_dim^-0.5) */
  const int32_t* seq_desc;   /* [num_seqs, 4] (see mome_attn_fwd) */
  const uint8_t* key_mask;   /* [tokens] or NULL */
  /* parameters (fp32) and compute-dtype weights */
  const float *gamma_1, *gamma_2, *n1w, *n1b, *n2w, *n2b, *qkv_bias /* [3d] = [q_bias, 0, v_bias] or NULL */, *proj_b;
  const void *w_qkv /* [3d, d] */, *w_proj /* [d, d] */;
  MomeBlockGroup group[MOME_MAX_GROUPS];
  /* activations: written by the forward, read by the backward */
  const float* x;            /* block input (fp32 residual stream) */
  void* h;  float* mean1; float* rstd1;   /* LN1 output (compute dtype) and statistics [tokens] */
  void* qkv;                 /* [tokens, 3d] */
  void* o;  float* lse;      /* attention output; log-sum-exp [num_seqs * num_heads * max_seq_len] */
  void* br1; float* x1;      /* proj output before LayerScale; residual stream after the attention branch */
  void* h2; float* mean2; float* rstd2;
  void* gp; void* u;         /* gelu'(z) and gelu(z), [tokens, hid] */
  void* br2; float* x2;      /* fc2 output before LayerScale; block output */
  /* backward only */
  const float* dx2;          /* gradient of the block output */
  float* dx;                 /* gradient of the block input */
  float *dgamma_1, *dgamma_2, *dn1w, *dn1b, *dn2w, *dn2b, *dq_bias /* [d] or NULL */, *dv_bias /* [d] or NULL */, *dproj_b, *dw_qkv, *dw_proj;
  void *s_dbr2, *s_dh2, *s_dbr1, *s_do, *s_dh;  /* scratch [tokens, d] (compute dtype) */
  void* s_dz;                /* scratch [tokens, hid] */
  void* s_dqkv;              /* scratch [tokens, 3d] */
  float* s_dx1;              /* scratch fp32 [tokens, d] */
  float* s_delta;            /* scratch fp32, same size as lse */
  void* ws; size_t ws_bytes; /* mome_reduce_ws_bytes(max(hid, 3d)) */
  /* Dropout, bf16 path only (reference Block: attn_drop vlmo.py:93, proj_drop :97, timm Mlp drop, DropPath
   * :194-196). drop_seed NULL or all p == 0: off. Call sites use drop_salt + {0: attention probabilities,
   * 1: proj output, 2: after GELU, 3: fc2 output, 4 / 5: stochastic depth of the attention / FFN branch}. */
  const uint32_t* drop_seed;   /* device scalar, bumped by the caller once per step */
  const int32_t* row_sample;   /* [tokens] index of the sample a row belongs to (needed when p_path > 0) */
  float* row_scale1;           /* [tokens] stochastic-depth multipliers: written by the forward, read by the backward */
  float* row_scale2;
  uint32_t drop_salt;
  float p_attn, p_hidden, p_branch, p_path;
} MomeBlockArgs;

int mome_block_fwd(const MomeBlockArgs* args, void* stream);
int mome_block_bwd(const MomeBlockArgs* args, void* stream);

/* ---- MLM head tail: softmax cross-entropy over the vocabulary on bf16 logits ----------------
 * reference: F.cross_entropy(mlm_logits, mlm_labels, ignore_index=-100) + compute_accuracy, objectives.py:52-66, 24-37.
 * logits: bf16 [rows, ld], columns [cols, ld) are padding (ld % 8 == 0). mome_ce_fwd writes lse[row] for every row and
 * ACCUMULATES loss_sum (sum over valid rows of lse - logit[target]), count (valid rows) and correct (argmax == target);
 * mome_ce_bwd overwrites logits IN PLACE with d loss_sum / d logits * *gscale (0 for ignored rows and padding), so the
 * [rows, vocab] matrix exists once, in bf16; gscale is a DEVICE scalar (upstream gradient / count). */
int mome_ce_fwd(const void* logits, int64_t ld, int32_t rows, int32_t cols, const int64_t* targets, int64_t ignore_index,
                float* lse, float* loss_sum, int32_t* count, int32_t* correct, void* stream);
int mome_ce_bwd(void* logits, int64_t ld, int32_t rows, int32_t cols, const int64_t* targets, int64_t ignore_index,
                const float* lse, const float* gscale, void* stream);

/* ---- Text input embedding ---------------------------------------------------------------------
 * reference: transformers BertEmbeddings as used at vlmo.py:259 (word + position + token_type(0) -> LayerNorm(eps 1e-12)
 * -> dropout) plus `+ token_type_embeddings(zeros)` of embed_txt, vlmo.py:321-324. rows = B * T tokens in row-major (b, t)
 * order; all tables fp32. y [rows, d] fp32; xhat [rows, d] and rstd [rows] are saved for the backward. The backward
 * ACCUMULATES into dword [vocab, d], dpos [T.., d] (red.add scatter) and dtype0 / dln_w / dln_b / dmodal0 [d] (may be NULL). */
int mome_text_embed_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0, const float* ln_w,
                        const float* ln_b, const float* modal0, float* y, float* xhat, float* rstd, int64_t rows, int32_t T,
                        int32_t d, float eps, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, void* stream);
size_t mome_text_embed_ws_bytes(int32_t d);
int mome_text_embed_bwd(const float* dy, const int64_t* ids, const float* xhat, const float* rstd, const float* ln_w,
                        float* dword, float* dpos, float* dtype0, float* dln_w, float* dln_b, float* dmodal0, int64_t rows,
                        int32_t T, int32_t d, int64_t padding_idx /* row of `word` that receives no gradient, -1 = none */,
                        const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, void* ws, size_t ws_bytes, void* stream);

/* ---- Optimizer step on flat buffers -------------------------------------------------------
 * reference: apex FusedAdam / DeepSpeed Adam(adam_w_mode) built by utils/optim_factory.py:93-199 over the three-tier
 * parameter groups of get_parameter_groups (:22-90), and the gradient clipping of train/pretrain/multimodal.py:311-330.
 * One launch updates n consecutive fp32 parameters (a whole flat buffer or a rank's ZeRO shard of it):
 *   g = grad * *grad_scale;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
 *   p = p (1 - lr wd) - lr (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps),   lr = lr_table[group_id[i]], wd likewise.
 * `step` (t, >= 1) and `grad_scale` (may be NULL = 1) are DEVICE scalars. */
int mome_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const uint8_t* group_id,
                    const float* lr_table, const float* wd_table, const float* step, const float* grad_scale, float beta1,
                    float beta2, float eps, int64_t n, void* stream);
/* out[0] += sum_i x[i]^2 (global gradient norm for clipping) */
int mome_sumsq(const float* x, int64_t n, float* out, void* stream);

/* ---- measurement hooks (bench.py): CUDA-event timing of every mome_gemm launch on its own stream */
int mome_prof_enable(int on);
/* Synchronises the recorded events; returns launches, summed milliseconds and summed FLOPs. */
int mome_prof_read(int64_t* launches, double* ms, double* flops, int reset);
/* total kernel launches issued by this library since load (gpu_launches claim) */
int64_t mome_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MOME_H_ */
