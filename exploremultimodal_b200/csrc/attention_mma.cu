// K1 (bf16 path): tensor-core attention. Placeholder revision: forwards to the CUDA-core kernels
// until the mma kernels land (same results, lower throughput).
#include "common.cuh"

namespace mome {
int attn_fwd_simt_dispatch(const void* qkv, int dtype, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse,
                           int num_seqs, int max_seq_len, int H, float scale, cudaStream_t stream);
int attn_bwd_simt_dispatch(const void* qkv, const void* out, const void* dout, int dtype, const int32_t* seq_desc,
                           const uint8_t* key_mask, const float* lse, void* dqkv, float* delta_ws, int num_seqs, int max_seq_len,
                           int H, float scale, cudaStream_t stream);

int attn_fwd_mma(const void* qkv, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse, int num_seqs,
                 int max_seq_len, int H, float scale, cudaStream_t stream) {
  return attn_fwd_simt_dispatch(qkv, MOME_BF16, seq_desc, key_mask, out, lse, num_seqs, max_seq_len, H, scale, stream);
}
int attn_bwd_mma(const void* qkv, const void* out, const void* dout, const int32_t* seq_desc, const uint8_t* key_mask,
                 const float* lse, void* dqkv, float* delta_ws, int num_seqs, int max_seq_len, int H, float scale,
                 cudaStream_t stream) {
  return attn_bwd_simt_dispatch(qkv, out, dout, MOME_BF16, seq_desc, key_mask, lse, dqkv, delta_ws, num_seqs, max_seq_len, H, scale, stream);
}
}  // namespace mome
