// MLM head tail: softmax cross-entropy over the vocabulary fused with its backward, on bf16 logits
// that are read once per direction and never converted to an fp32 [rows, vocab] tensor.
//
// Replaces: F.cross_entropy(mlm_logits.view(-1, vocab), mlm_labels.view(-1), ignore_index=-100) and compute_accuracy
// (reference objectives.py:52-66, 24-37) and their autograd backward. The decoder GEMM in front and the dgrad / wgrad
// GEMMs behind are libmome's grouped tcgen05 GEMM (mome_gemm).
//
// One CTA (256 threads) per row; the row (vocab 30522 -> 61 KB of bf16) lives in registers as 128-bit chunks.
#include <algorithm>

#include "common.cuh"
#include "vec.cuh"

namespace mome {

constexpr int kCeThreads = 256;
constexpr int kCeMaxChunks = 16;  // 128-bit chunks (8 bf16) per thread: up to 16 * 256 * 8 = 32768 columns

__device__ __forceinline__ float ce_block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kCeThreads / 32; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float ce_block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kCeThreads / 32; ++w) r += red[w];
  return r;
}

// forward: lse[row], loss_sum += lse - logit[target], count += 1, correct += (argmax == target) for rows with target != ignore
// backward (BWD): logits[row][c] <- (exp(logit - lse) - [c == target]) * gscale for valid rows, 0 for ignored rows (in place)
template <bool BWD>
__global__ void __launch_bounds__(kCeThreads) ce_kernel(__nv_bfloat16* __restrict__ logits, long long ld, int cols,
                                                        const long long* __restrict__ targets, long long ignore_index,
                                                        float* __restrict__ lse, float* __restrict__ loss_sum, int* __restrict__ count,
                                                        int* __restrict__ correct, const float* __restrict__ gscale) {
  __shared__ float red[kCeThreads / 32];
  __shared__ int red_i[kCeThreads / 32];
  const int row = blockIdx.x;
  __nv_bfloat16* lr = logits + static_cast<long long>(row) * ld;
  const long long tgt = targets[row];
  const bool valid = tgt != ignore_index;
  const int nchunk = (cols + 7) >> 3;
  uint4 v[kCeMaxChunks];
#pragma unroll
  for (int i = 0; i < kCeMaxChunks; ++i) {
    const int ch = i * kCeThreads + threadIdx.x;
    v[i] = ch < nchunk ? *reinterpret_cast<const uint4*>(lr + ch * 8) : make_uint4(0, 0, 0, 0);
  }
  if (!BWD) {
    float mx = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int i = 0; i < kCeMaxChunks; ++i) {
      const int c0 = (i * kCeThreads + threadIdx.x) * 8;
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        if (c0 + 2 * j < cols && f.x > mx) { mx = f.x; arg = c0 + 2 * j; }
        if (c0 + 2 * j + 1 < cols && f.y > mx) { mx = f.y; arg = c0 + 2 * j + 1; }
      }
    }
    const float m = ce_block_max(mx, red);
    // first column reaching the maximum (torch.argmax returns the first maximal index)
    int cand = (mx == m) ? arg : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    if ((threadIdx.x & 31) == 0) red_i[threadIdx.x >> 5] = cand;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kCeMaxChunks; ++i) {
      const int c0 = (i * kCeThreads + threadIdx.x) * 8;
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        if (c0 + 2 * j < cols) s += __expf(f.x - m);
        if (c0 + 2 * j + 1 < cols) s += __expf(f.y - m);
      }
    }
    const float total = ce_block_sum(s, red);  // contains the __syncthreads that publish red_i
    if (threadIdx.x == 0) {
      const float l = m + __logf(total);
      lse[row] = l;
      if (valid) {
        int best = red_i[0];
#pragma unroll
        for (int w = 1; w < kCeThreads / 32; ++w) best = min(best, red_i[w]);
        atomicAdd(loss_sum, l - __bfloat162float(lr[tgt]));
        atomicAdd(count, 1);
        if (best == static_cast<int>(tgt)) atomicAdd(correct, 1);
      }
    }
  } else {
    const float l = lse[row];
    const float g = valid ? __ldg(gscale) : 0.f;
#pragma unroll
    for (int i = 0; i < kCeMaxChunks; ++i) {
      const int ch = i * kCeThreads + threadIdx.x;
      if (ch >= nchunk) continue;
      const int c0 = ch * 8;
      uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        float a = c0 + 2 * j < cols ? (__expf(f.x - l) - (c0 + 2 * j == tgt ? 1.f : 0.f)) * g : 0.f;
        float b = c0 + 2 * j + 1 < cols ? (__expf(f.y - l) - (c0 + 2 * j + 1 == tgt ? 1.f : 0.f)) * g : 0.f;
        w[j] = pack_bf16(a, b);
      }
      *reinterpret_cast<uint4*>(lr + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

}  // namespace mome

using namespace mome;

#define MOME_CE_CHECK(name)                                                                                              \
  MOME_REQUIRE(logits != nullptr && targets != nullptr && lse != nullptr, name ": null argument");                        \
  MOME_REQUIRE(cols >= 1 && cols <= kCeMaxChunks * kCeThreads * 8, name ": cols=%d unsupported (<= %d)", cols, kCeMaxChunks * kCeThreads * 8); \
  MOME_REQUIRE(ld % 8 == 0 && ld >= ((cols + 7) / 8) * 8 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0,             \
               name ": rows must be 16-byte aligned and padded to a multiple of 8 columns (ld=%lld)", (long long)ld)

extern "C" int mome_ce_fwd(const void* logits, int64_t ld, int32_t rows, int32_t cols, const int64_t* targets, int64_t ignore_index,
                           float* lse, float* loss_sum, int32_t* count, int32_t* correct, void* stream) {
  MOME_CE_CHECK("ce_fwd");
  MOME_REQUIRE(loss_sum != nullptr && count != nullptr && correct != nullptr, "ce_fwd: null output");
  if (rows == 0) return MOME_OK;
  ce_kernel<false><<<rows, kCeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(const_cast<void*>(logits)), ld, cols, reinterpret_cast<const long long*>(targets), ignore_index, lse, loss_sum,
      count, correct, nullptr);
  return check_launch("ce_fwd");
}

extern "C" int mome_ce_bwd(void* logits, int64_t ld, int32_t rows, int32_t cols, const int64_t* targets, int64_t ignore_index,
                           const float* lse, const float* gscale, void* stream) {
  MOME_CE_CHECK("ce_bwd");
  MOME_REQUIRE(gscale != nullptr, "ce_bwd: null gscale");
  if (rows == 0) return MOME_OK;
  ce_kernel<true><<<rows, kCeThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<__nv_bfloat16*>(logits), ld, cols,
                                                                              reinterpret_cast<const long long*>(targets), ignore_index,
                                                                              const_cast<float*>(lse), nullptr, nullptr, nullptr, gscale);
  return check_launch("ce_bwd");
}
