// C-ABI entry points of K1 (see include/mome.h): argument checks and dispatch between the fp32
// CUDA-core kernels (attention_simt.cu) and the bf16 tensor-core kernels (attention_mma.cu).
#include <stdlib.h>

#include "common.cuh"

namespace mome {
int attn_fwd_simt_dispatch(const void* qkv, int dtype, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse,
                           int num_seqs, int max_seq_len, int H, float scale, cudaStream_t stream);
int attn_bwd_simt_dispatch(const void* qkv, const void* out, const void* dout, int dtype, const int32_t* seq_desc,
                           const uint8_t* key_mask, const float* lse, void* dqkv, float* delta_ws, int num_seqs, int max_seq_len,
                           int H, float scale, cudaStream_t stream);
int attn_fwd_mma(const void* qkv, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse, int num_seqs,
                 int max_seq_len, int H, float scale, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, cudaStream_t stream);
int attn_bwd_mma(const void* qkv, const void* out, const void* dout, const int32_t* seq_desc, const uint8_t* key_mask,
                 const float* lse, void* dqkv, float* delta_ws, int num_seqs, int max_seq_len, int H, float scale,
                 const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, cudaStream_t stream);

int attn_fwd_tc(const void* qkv, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse, int64_t tokens, int num_seqs,
                int max_seq_len, int H, float scale, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, cudaStream_t stream);

// The tcgen05 forward (attention_tc.cu) takes the layouts whose sequences fit one MMA (max_seq_len <= 256) and are
// long enough to fill its 128-row tiles; shorter ones (text-only passes) and longer ones (VQA at 480 px) stay on
// the mma.sync kernels. MOME_ATTN_TC=0 turns it off (read per call: the tests compare the two paths).
int attn_fwd_tc_long(const void* qkv, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse, int64_t tokens, int num_seqs,
                     int max_seq_len, int H, float scale, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, cudaStream_t stream);
int attn_fwd_tc_long_max_seq_len();
// longer sequences (VQA at 480 / 384 px): the key-blocked tcgen05 forward with an online softmax (attention_tc_long.cu)
static bool use_tc_long(int max_seq_len) {
  const char* e = getenv("MOME_ATTN_TC");
  return (e == nullptr || e[0] != '0') && max_seq_len > 256 && max_seq_len <= attn_fwd_tc_long_max_seq_len();
}
// shortest layout the tcgen05 kernels take (text-only layouts of <= 64 tokens fill a third of a 128-row tile: the mma.sync
// kernels measured faster there, tools/attn_bench.py 'text 40'); MOME_ATTN_TC_MIN overrides it for that measurement
static int tc_min_len() {
  const char* e = getenv("MOME_ATTN_TC_MIN");
  return e != nullptr ? atoi(e) : 64;
}
static bool use_tc(int max_seq_len) {
  const char* e = getenv("MOME_ATTN_TC");
  return (e == nullptr || e[0] != '0') && max_seq_len > tc_min_len() && max_seq_len <= 256;
}

int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const int32_t* seq_desc, const uint8_t* key_mask, const float* lse,
                void* dqkv, float* delta_ws, int64_t tokens, int num_seqs, int max_seq_len, int H, float scale, const uint32_t* drop_seed,
                uint32_t drop_salt, float drop_p, cudaStream_t stream);

int attn_bwd_tc_pipe(const void* qkv, const void* out, const void* dout, const int32_t* seq_desc, const uint8_t* key_mask, const float* lse,
                     void* dqkv, float* delta_ws, int64_t tokens, int num_seqs, int max_seq_len, int H, float scale, const uint32_t* drop_seed,
                     uint32_t drop_salt, float drop_p, cudaStream_t stream);

int attn_bwd_tc_pipe_max_seq_len();
int64_t attn_bwd_tc_pipe_extra_ws_floats(int64_t tokens, int max_seq_len, int H);

// The tcgen05 backward. MOME_ATTN_TC_BWD: unset = the software-pipelined kernel (attention_tc_bwd_pipe.cu; sequences of
// 65 .. 1024 tokens, i.e. also VQA at 480 px), 1 / 2 / 3 = the first kernel and its variants (attention_tc_bwd.cu, up to
// 256 tokens), 0 = mma.sync.
static bool use_tc_bwd_pipe(int max_seq_len) {
  const char* e = getenv("MOME_ATTN_TC_BWD");
  return (e == nullptr || e[0] == '\0' || e[0] == 'p') && max_seq_len > tc_min_len() && max_seq_len <= attn_bwd_tc_pipe_max_seq_len();
}
static bool use_tc_bwd(int max_seq_len) {
  const char* e = getenv("MOME_ATTN_TC_BWD");
  return e != nullptr && e[0] >= '1' && e[0] <= '3' && max_seq_len > 64 && max_seq_len <= 256;
}

// MOME_ATTN_SIMT=1 forces the CUDA-core kernels for bf16 too (cross-check in tests / debugging).
static bool force_simt() {
  static const bool v = [] {
    const char* e = getenv("MOME_ATTN_SIMT");
    return e != nullptr && e[0] == '1';
  }();
  return v;
}
}  // namespace mome

using namespace mome;

extern "C" int mome_attn_fwd(const void* qkv, int dtype, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse,
                             int64_t tokens, int32_t num_seqs, int32_t max_seq_len, int32_t num_heads, float scale,
                             const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, void* stream) {
  const bool drop = drop_seed != nullptr && drop_p > 0.f;
  MOME_REQUIRE(!drop || (dtype == MOME_BF16 && !force_simt()), "attn_fwd: dropout exists on the bf16 tensor-core path only");
  MOME_REQUIRE(dtype == MOME_F32 || dtype == MOME_BF16, "attn_fwd: unknown dtype %d", dtype);
  MOME_REQUIRE(num_heads > 0 && max_seq_len > 0 && tokens >= 0, "attn_fwd: bad shape heads=%d max_seq_len=%d", num_heads, max_seq_len);
  if (num_seqs == 0 || tokens == 0) return MOME_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MOME_BF16 && !force_simt() && use_tc_long(max_seq_len))
    return attn_fwd_tc_long(qkv, seq_desc, key_mask, out, lse, tokens, num_seqs, max_seq_len, num_heads, scale, drop_seed, drop_salt, drop_p, s);
  if (dtype == MOME_BF16 && !force_simt() && use_tc(max_seq_len))
    return attn_fwd_tc(qkv, seq_desc, key_mask, out, lse, tokens, num_seqs, max_seq_len, num_heads, scale, drop_seed, drop_salt, drop_p, s);
  if (dtype == MOME_BF16 && !force_simt())
    return attn_fwd_mma(qkv, seq_desc, key_mask, out, lse, num_seqs, max_seq_len, num_heads, scale, drop_seed, drop_salt, drop_p, s);
  return attn_fwd_simt_dispatch(qkv, dtype, seq_desc, key_mask, out, lse, num_seqs, max_seq_len, num_heads, scale, s);
}

extern "C" int64_t mome_attn_bwd_ws_floats(int64_t tokens, int32_t num_seqs, int32_t max_seq_len, int32_t num_heads) {
  const int64_t delta = (static_cast<int64_t>(num_seqs) * num_heads * max_seq_len + 3) & ~int64_t(3);
  return delta + attn_bwd_tc_pipe_extra_ws_floats(tokens, max_seq_len, num_heads);
}

extern "C" int mome_attn_bwd(const void* qkv, const void* out, const void* dout, int dtype, const int32_t* seq_desc,
                             const uint8_t* key_mask, const float* lse, void* dqkv, float* delta_ws, int64_t tokens, int32_t num_seqs,
                             int32_t max_seq_len, int32_t num_heads, float scale, const uint32_t* drop_seed, uint32_t drop_salt,
                             float drop_p, void* stream) {
  const bool drop = drop_seed != nullptr && drop_p > 0.f;
  MOME_REQUIRE(!drop || (dtype == MOME_BF16 && !force_simt()), "attn_bwd: dropout exists on the bf16 tensor-core path only");
  MOME_REQUIRE(dtype == MOME_F32 || dtype == MOME_BF16, "attn_bwd: unknown dtype %d", dtype);
  MOME_REQUIRE(num_heads > 0 && max_seq_len > 0 && tokens >= 0, "attn_bwd: bad shape heads=%d max_seq_len=%d", num_heads, max_seq_len);
  if (num_seqs == 0 || tokens == 0) return MOME_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MOME_BF16 && !force_simt() && use_tc_bwd_pipe(max_seq_len))
    return attn_bwd_tc_pipe(qkv, out, dout, seq_desc, key_mask, lse, dqkv, delta_ws, tokens, num_seqs, max_seq_len, num_heads, scale,
                            drop_seed, drop_salt, drop_p, s);
  if (dtype == MOME_BF16 && !force_simt() && use_tc_bwd(max_seq_len))
    return attn_bwd_tc(qkv, out, dout, seq_desc, key_mask, lse, dqkv, delta_ws, tokens, num_seqs, max_seq_len, num_heads, scale, drop_seed,
                       drop_salt, drop_p, s);
  if (dtype == MOME_BF16 && !force_simt())
    return attn_bwd_mma(qkv, out, dout, seq_desc, key_mask, lse, dqkv, delta_ws, num_seqs, max_seq_len, num_heads, scale, drop_seed,
                        drop_salt, drop_p, s);
  return attn_bwd_simt_dispatch(qkv, out, dout, dtype, seq_desc, key_mask, lse, dqkv, delta_ws, num_seqs, max_seq_len, num_heads, scale, s);
}
