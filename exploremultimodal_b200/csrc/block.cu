// One MoME block per call: native sequencing of the kernels of Block.forward (reference vlmo.py:187-197)
// and of its backward. No computation of its own — it exists so that the host pays one FFI call per block
// and direction instead of ~20 (the Python host was launch-bound in eager mode without it).
#include "common.cuh"

namespace {

inline size_t esize(int dtype) { return dtype == MOME_BF16 ? 2 : 4; }
inline const char* at(const void* p, int64_t row, int64_t ld, size_t es) { return static_cast<const char*>(p) + row * ld * es; }
inline char* at(void* p, int64_t row, int64_t ld, size_t es) { return static_cast<char*>(p) + row * ld * es; }

MomeGemmArgs gemm_args(int dtype, int a_major, int b_major, int epilogue, int out_dtype, int groups, int64_t N, int64_t lda,
                       int64_t ldb, int64_t ldo) {
  MomeGemmArgs g;
  memset(&g, 0, sizeof(g));
  g.dtype = dtype; g.a_major = a_major; g.b_major = b_major; g.epilogue = epilogue; g.out_dtype = out_dtype;
  g.num_groups = groups; g.N = N; g.lda = lda; g.ldb = ldb; g.ldo = ldo;
  return g;
}

// dropout is active for probability p?
inline bool on(const MomeBlockArgs* a, float p) { return a->drop_seed != nullptr && p > 0.f; }

#define MOME_TRY(call)          \
  do {                          \
    int rc_ = (call);           \
    if (rc_ != MOME_OK) return rc_; \
  } while (0)

// Side stream of the backward: the four weight-gradient GEMMs of a block depend on the dgrad chain but nothing in the block depends on
// them, so they are forked onto a second (lowest-priority) stream and joined before the call returns. Two persistent GEMMs cannot share
// an SM (shared memory), but the HBM-bound row kernels of the chain (LayerNorm / LayerScale backward, column sums) can run next to a
// weight-gradient GEMM: tensor pipe and memory pipe are busy at the same time. Fork / join are event record / wait pairs, which a
// CUDA-graph capture turns into parallel branches. Opt-in (MOME_BWD_SIDE_STREAM=1): measured on a power-capped B200 the graphed
// pretraining step gains 0.4 % (122.06 against 122.50 ms; the SM clock drops from 1597 to 1515 MHz as more units are busy at once),
// the eager step 3 % (122.3 against 126.2 ms).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t join = nullptr;
  bool ok = false;
};
SideStream* side_stream() {
  static const bool enabled = [] { const char* e = getenv("MOME_BWD_SIDE_STREAM"); return e != nullptr && e[0] == '1'; }();
  if (!enabled || mome::gemm_prof_active()) return nullptr;
  static SideStream per_dev[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& s = per_dev[dev];
  if (s.stream == nullptr) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = numerically greatest = lowest priority
    bool ok = cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, lo) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreateWithFlags(&s.fork[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess;
    s.ok = ok;
    if (!ok) cudaGetLastError();
  }
  return s.ok ? &s : nullptr;
}
// everything issued on `main` so far happens before what is issued on the side stream from now on
inline int fork_to(SideStream* ss, int i, cudaStream_t main) {
  if (cudaEventRecord(ss->fork[i], main) != cudaSuccess || cudaStreamWaitEvent(ss->stream, ss->fork[i], 0) != cudaSuccess) {
    mome::set_error("block_bwd: fork to the side stream failed: %s", cudaGetErrorString(cudaGetLastError()));
    return MOME_ERR_CUDA;
  }
  return MOME_OK;
}
inline int join_from(SideStream* ss, cudaStream_t main) {
  if (cudaEventRecord(ss->join, ss->stream) != cudaSuccess || cudaStreamWaitEvent(main, ss->join, 0) != cudaSuccess) {
    mome::set_error("block_bwd: join of the side stream failed: %s", cudaGetErrorString(cudaGetLastError()));
    return MOME_ERR_CUDA;
  }
  return MOME_OK;
}

}  // namespace

extern "C" int mome_block_fwd(const MomeBlockArgs* a, void* stream) {
  MOME_REQUIRE(a != nullptr && a->num_groups >= 1 && a->num_groups <= MOME_MAX_GROUPS, "block_fwd: bad group count");
  const int dt = a->dtype;
  const size_t es = esize(dt);
  const int64_t T = a->tokens, d = a->d, hid = a->hid;
  if (T == 0) return MOME_OK;
  MOME_REQUIRE(dt == MOME_BF16 || !(on(a, a->p_attn) || on(a, a->p_hidden) || on(a, a->p_branch) || on(a, a->p_path)),
               "block_fwd: dropout exists on the bf16 path only");
  const bool path = on(a, a->p_path);
  if (path) {
    MOME_REQUIRE(a->row_sample != nullptr && a->row_scale1 != nullptr && a->row_scale2 != nullptr, "block_fwd: stochastic depth needs row_sample / row_scale buffers");
    MOME_TRY(mome::droppath_scales2(a->row_sample, T, a->drop_seed, a->drop_salt + 4, a->drop_salt + 5, a->p_path, a->row_scale1, a->row_scale2,
                                    static_cast<cudaStream_t>(stream)));
  }
  MOME_TRY(mome_ln_fwd(a->x, a->n1w, a->n1b, a->h, dt, a->mean1, a->rstd1, T, d, a->eps, stream));
  {
    MomeGemmArgs g = gemm_args(dt, 0, 0, MOME_EPI_STORE, dt, 1, 3 * d, d, d, 3 * d);
    g.group[0].a = a->h; g.group[0].b = a->w_qkv; g.group[0].M = T; g.group[0].K = d; g.group[0].out = a->qkv;
    g.group[0].bias = a->qkv_bias;
    MOME_TRY(mome_gemm(&g, stream));
  }
  MOME_TRY(mome_attn_fwd(a->qkv, dt, a->seq_desc, a->key_mask, a->o, a->lse, T, a->num_seqs, a->max_seq_len, a->num_heads, a->scale,
                         on(a, a->p_attn) ? a->drop_seed : nullptr, a->drop_salt, a->p_attn, stream));
  {
    MomeGemmArgs g = gemm_args(dt, 0, 0, MOME_EPI_RESIDUAL, MOME_F32, 1, d, d, d, d);
    g.ldo2 = d; g.ldres = d; g.gamma = a->gamma_1;
    g.group[0].a = a->o; g.group[0].b = a->w_proj; g.group[0].M = T; g.group[0].K = d; g.group[0].out = a->x1;
    g.group[0].out2 = a->br1; g.group[0].bias = a->proj_b; g.group[0].res = a->x;
    if (on(a, a->p_branch)) { g.drop_seed = a->drop_seed; g.drop_salt = a->drop_salt + 1; g.drop_p = a->p_branch; }
    if (path) g.row_scale = a->row_scale1;
    MOME_TRY(mome_gemm(&g, stream));
  }
  MOME_TRY(mome_ln_fwd(a->x1, a->n2w, a->n2b, a->h2, dt, a->mean2, a->rstd2, T, d, a->eps, stream));
  {
    MomeGemmArgs g1 = gemm_args(dt, 0, 0, MOME_EPI_GELU, dt, a->num_groups, hid, d, d, hid);
    g1.ldo2 = hid;
    MomeGemmArgs g2 = gemm_args(dt, 0, 0, MOME_EPI_RESIDUAL, MOME_F32, a->num_groups, d, hid, hid, d);
    g2.ldo2 = d; g2.ldres = d; g2.gamma = a->gamma_2;
    if (on(a, a->p_hidden)) { g1.drop_seed = a->drop_seed; g1.drop_salt = a->drop_salt + 2; g1.drop_p = a->p_hidden; }
    if (on(a, a->p_branch)) { g2.drop_seed = a->drop_seed; g2.drop_salt = a->drop_salt + 3; g2.drop_p = a->p_branch; }
    if (path) g2.row_scale = a->row_scale2;
    for (int i = 0; i < a->num_groups; ++i) {
      const MomeBlockGroup& s = a->group[i];
      g1.group[i].row0 = s.first_row;
      g2.group[i].row0 = s.first_row;
      g1.group[i].a = at(a->h2, s.first_row, d, es); g1.group[i].b = s.w1; g1.group[i].M = s.rows; g1.group[i].K = d;
      g1.group[i].out = at(a->u, s.first_row, hid, es); g1.group[i].out2 = at(a->gp, s.first_row, hid, es); g1.group[i].bias = s.b1;
      g2.group[i].a = at(a->u, s.first_row, hid, es); g2.group[i].b = s.w2; g2.group[i].M = s.rows; g2.group[i].K = hid;
      g2.group[i].out = at(a->x2, s.first_row, d, 4); g2.group[i].out2 = at(a->br2, s.first_row, d, es); g2.group[i].bias = s.b2;
      g2.group[i].res = reinterpret_cast<const float*>(at(a->x1, s.first_row, d, 4));
    }
    MOME_TRY(mome_gemm(&g1, stream));
    MOME_TRY(mome_gemm(&g2, stream));
  }
  return MOME_OK;
}

extern "C" int mome_block_bwd(const MomeBlockArgs* a, void* stream) {
  MOME_REQUIRE(a != nullptr && a->num_groups >= 1 && a->num_groups <= MOME_MAX_GROUPS, "block_bwd: bad group count");
  const int dt = a->dtype;
  const size_t es = esize(dt);
  const int64_t T = a->tokens, d = a->d, hid = a->hid;
  if (T == 0) return MOME_OK;
  const bool path = on(a, a->p_path);
  // ---- expert FFN branch: x2 = x1 + gamma_2 * fc2(gelu(fc1(LN2(x1))))
  MomeGemmArgs dgrad2 = gemm_args(dt, 0, 1, MOME_EPI_DGELU, dt, a->num_groups, hid, d, hid, hid);
  dgrad2.ldaux = hid;
  MomeGemmArgs wgrad2 = gemm_args(dt, 1, 1, MOME_EPI_ATOMIC, MOME_F32, a->num_groups, hid, d, hid, hid);
  MomeGemmArgs wgrad1 = gemm_args(dt, 1, 1, MOME_EPI_ATOMIC, MOME_F32, a->num_groups, d, hid, d, d);
  MomeGemmArgs dgrad1 = gemm_args(dt, 0, 1, MOME_EPI_STORE, dt, a->num_groups, d, hid, d, d);
  for (int i = 0; i < a->num_groups; ++i) {
    const MomeBlockGroup& s = a->group[i];
    const MomeDropout drop2{on(a, a->p_branch) ? a->drop_seed : nullptr, path ? a->row_scale2 : nullptr, s.first_row, a->drop_salt + 3, a->p_branch};
    MOME_TRY(mome_scale_bwd(reinterpret_cast<const float*>(at(a->dx2, s.first_row, d, 4)), at(a->br2, s.first_row, d, es), dt, a->gamma_2,
                            at(a->s_dbr2, s.first_row, d, es), dt, a->dgamma_2, s.db2, s.rows, d, &drop2, a->ws, a->ws_bytes, stream));
    dgrad2.group[i].a = at(a->s_dbr2, s.first_row, d, es); dgrad2.group[i].b = s.w2; dgrad2.group[i].M = s.rows; dgrad2.group[i].K = d;
    dgrad2.group[i].out = at(a->s_dz, s.first_row, hid, es); dgrad2.group[i].aux = at(a->gp, s.first_row, hid, es);
    dgrad2.group[i].colsum = s.colsum_part;
    wgrad2.group[i].a = at(a->s_dbr2, s.first_row, d, es); wgrad2.group[i].b = at(a->u, s.first_row, hid, es); wgrad2.group[i].M = d;
    wgrad2.group[i].K = s.rows; wgrad2.group[i].out = s.dw2;
    wgrad1.group[i].a = at(a->s_dz, s.first_row, hid, es); wgrad1.group[i].b = at(a->h2, s.first_row, d, es); wgrad1.group[i].M = hid;
    wgrad1.group[i].K = s.rows; wgrad1.group[i].out = s.dw1;
    dgrad1.group[i].a = at(a->s_dz, s.first_row, hid, es); dgrad1.group[i].b = s.w1; dgrad1.group[i].M = s.rows; dgrad1.group[i].K = hid;
    dgrad1.group[i].out = at(a->s_dh2, s.first_row, d, es);
  }
  SideStream* ss = side_stream();
  cudaStream_t main_s = static_cast<cudaStream_t>(stream);
  void* wstream = ss != nullptr ? static_cast<void*>(ss->stream) : stream;  // where the weight-gradient GEMMs go
  if (ss != nullptr) MOME_TRY(fork_to(ss, 0, main_s));  // dbr2 is there
  MOME_TRY(mome_gemm(&wgrad2, wstream));  // dW2 += dbr2^T u
  MOME_TRY(mome_gemm(&dgrad2, stream));   // dz = (dbr2 W2) * gelu'(z); per-32-row column sums -> colsum_part
  for (int i = 0; i < a->num_groups; ++i) {
    const MomeBlockGroup& s = a->group[i];
    MOME_TRY(mome_colreduce(s.colsum_part, (s.rows + 31) / 32, hid, s.db1, stream));
  }
  if (ss != nullptr) MOME_TRY(fork_to(ss, 1, main_s));  // dz is there
  MOME_TRY(mome_gemm(&wgrad1, wstream));  // dW1 += dz^T h2
  MOME_TRY(mome_gemm(&dgrad1, stream));   // dh2 = dz W1
  // LN2 backward (+ dx2) fused with the LayerScale backward of the attention branch
  const MomeDropout drop1{on(a, a->p_branch) ? a->drop_seed : nullptr, path ? a->row_scale1 : nullptr, 0, a->drop_salt + 1, a->p_branch};
  MOME_TRY(mome_ln_bwd_scale(a->s_dh2, dt, a->x1, a->mean2, a->rstd2, a->n2w, a->dx2, a->s_dx1, a->dn2w, a->dn2b, a->br1, a->gamma_1,
                             a->s_dbr1, a->dgamma_1, a->dproj_b, T, d, &drop1, a->ws, a->ws_bytes, stream));
  // ---- attention branch: x1 = x + gamma_1 * proj(attn(qkv(LN1(x))))
  {
    MomeGemmArgs g = gemm_args(dt, 1, 1, MOME_EPI_ATOMIC, MOME_F32, 1, d, d, d, d);
    g.group[0].a = a->s_dbr1; g.group[0].b = a->o; g.group[0].M = d; g.group[0].K = T; g.group[0].out = a->dw_proj;
    if (ss != nullptr) MOME_TRY(fork_to(ss, 2, main_s));  // dbr1 is there
    MOME_TRY(mome_gemm(&g, wstream));
  }
  {
    MomeGemmArgs g = gemm_args(dt, 0, 1, MOME_EPI_STORE, dt, 1, d, d, d, d);
    g.group[0].a = a->s_dbr1; g.group[0].b = a->w_proj; g.group[0].M = T; g.group[0].K = d; g.group[0].out = a->s_do;
    MOME_TRY(mome_gemm(&g, stream));
  }
  MOME_TRY(mome_attn_bwd(a->qkv, a->o, a->s_do, dt, a->seq_desc, a->key_mask, a->lse, a->s_dqkv, a->s_delta, T, a->num_seqs,
                         a->max_seq_len, a->num_heads, a->scale, on(a, a->p_attn) ? a->drop_seed : nullptr, a->drop_salt, a->p_attn, stream));
  // q_bias / v_bias gradients: column sums of the q and v thirds of dqkv (the k third has no bias, vlmo.py:72-75)
  if (a->dq_bias != nullptr && a->dv_bias != nullptr)
    MOME_TRY(mome::colsum_qv(a->s_dqkv, dt, T, d, a->dq_bias, a->dv_bias, a->ws, a->ws_bytes, static_cast<cudaStream_t>(stream)));
  {
    MomeGemmArgs g = gemm_args(dt, 1, 1, MOME_EPI_ATOMIC, MOME_F32, 1, d, 3 * d, d, d);
    g.group[0].a = a->s_dqkv; g.group[0].b = a->h; g.group[0].M = 3 * d; g.group[0].K = T; g.group[0].out = a->dw_qkv;
    if (ss != nullptr) MOME_TRY(fork_to(ss, 3, main_s));  // dqkv is there
    MOME_TRY(mome_gemm(&g, wstream));
  }
  {
    MomeGemmArgs g = gemm_args(dt, 0, 1, MOME_EPI_STORE, dt, 1, d, 3 * d, d, d);
    g.group[0].a = a->s_dqkv; g.group[0].b = a->w_qkv; g.group[0].M = T; g.group[0].K = 3 * d; g.group[0].out = a->s_dh;
    MOME_TRY(mome_gemm(&g, stream));
  }
  MOME_TRY(mome_ln_bwd(a->s_dh, dt, a->x, a->mean1, a->rstd1, a->n1w, a->s_dx1, a->dx, a->dn1w, a->dn1b, T, d, a->ws, a->ws_bytes, stream));
  // the scratch tensors the weight-gradient GEMMs read are released by the caller when this call returns: join first
  if (ss != nullptr) MOME_TRY(join_from(ss, main_s));
  return MOME_OK;
}
