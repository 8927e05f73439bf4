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
#include "dropout.cuh"
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

// ------------------------------------------------------------------------------------------- text embedding
// BERT input embedding + VLMo's token-type add, one warp per token:
//   e = word[id] + pos[t] + type[0];  y = dropout(LayerNorm(e) * g + b) + modal_type[0]        (fp32 out, packed rows)
// Replaces: transformers BertEmbeddings (gather x3, add x2, LayerNorm, dropout) + `+ token_type_embeddings(zeros)`
// (reference vlmo.py:321-324) — ~8 eager launches forward — and, in the backward, torch's sort-based
// embedding_dense_backward (~12 launches per table) by red.add scatters.
namespace mome {

constexpr int kEmbMaxV4 = 8;  // d <= 8 * 128 = 1024

__global__ void __launch_bounds__(256) text_embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                                             const float* __restrict__ pos, const float* __restrict__ type0,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ modal0, float* __restrict__ y,
                                                             float* __restrict__ xhat, float* __restrict__ rstd, int rows, int T, int d,
                                                             float eps, const uint32_t* __restrict__ seed, uint32_t salt, uint32_t thr) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long long id = ids[row];
  const int t = row % T;
  const int nv = d >> 2;
  float4 v[kEmbMaxV4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kEmbMaxV4; ++i) {
    const int c = i * 32 + lane;
    if (c < nv) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(word + id * d) + c);
      const float4 b = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * d) + c);
      const float4 e = __ldg(reinterpret_cast<const float4*>(type0) + c);
      v[i] = make_float4(a.x + b.x + e.x, a.y + b.y + e.y, a.z + b.z + e.z, a.w + b.w + e.w);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mu = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kEmbMaxV4; ++i) {
    const int c = i * 32 + lane;
    if (c < nv) {
      v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
  const float rs = rsqrtf(warp_sum(q) / d + eps);
  if (lane == 0) rstd[row] = rs;
  const uint32_t key = seed != nullptr ? drop_mix(salt, __ldg(seed)) : 0u;
  const float dsc = drop_scale(thr);
#pragma unroll
  for (int i = 0; i < kEmbMaxV4; ++i) {
    const int c = i * 32 + lane;
    if (c < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(modal0) + c);
      const float4 xh = make_float4(v[i].x * rs, v[i].y * rs, v[i].z * rs, v[i].w * rs);
      reinterpret_cast<float4*>(xhat + static_cast<long long>(row) * d)[c] = xh;
      float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
      if (seed != nullptr) m = drop_mul4(drop_group(row, d, 4 * c), key, thr, dsc);
      reinterpret_cast<float4*>(y + static_cast<long long>(row) * d)[c] =
          make_float4((xh.x * g.x + b.x) * m.x + m0.x, (xh.y * g.y + b.y) * m.y + m0.y, (xh.z * g.z + b.z) * m.z + m0.z,
                      (xh.w * g.w + b.w) * m.w + m0.w);
    }
  }
}

// backward: dy -> (mask) -> LayerNorm backward -> de; scatter de into dword[id], dpos[t]; column sums (dgamma, dbeta, dtype0 = sum de,
// dmodal0 = sum dy) go to per-CTA partial rows in `ws` ([gridDim.x][4 d]), added up by mome_colreduce-style second stage.
__global__ void __launch_bounds__(256) text_embed_bwd_kernel(const float* __restrict__ dy, const long long* __restrict__ ids,
                                                             const float* __restrict__ xhat, const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma, float* __restrict__ dword,
                                                             float* __restrict__ dpos, float* __restrict__ ws, int rows, int T, int d,
                                                             const uint32_t* __restrict__ seed, uint32_t salt, uint32_t thr, long long pad_id) {
  extern __shared__ float sacc[];  // [4][d]: dgamma, dbeta, dtype0, dmodal0 of this CTA
  for (int i = threadIdx.x; i < 4 * d; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nv = d >> 2;
  const uint32_t key = seed != nullptr ? drop_mix(salt, __ldg(seed)) : 0u;
  const float dsc = drop_scale(thr);
  const int warps = gridDim.x * 8;
  // a lane owns the same columns in every row its warp visits: the four column sums accumulate in registers
  float4 a_g[kEmbMaxV4], a_b[kEmbMaxV4], a_t[kEmbMaxV4], a_m[kEmbMaxV4];
#pragma unroll
  for (int i = 0; i < kEmbMaxV4; ++i) a_g[i] = a_b[i] = a_t[i] = a_m[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += warps) {
    const float rs = rstd[row];
    float4 gy[kEmbMaxV4], xh[kEmbMaxV4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kEmbMaxV4; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) {
        const float4 g0 = reinterpret_cast<const float4*>(dy + static_cast<long long>(row) * d)[c];
        xh[i] = reinterpret_cast<const float4*>(xhat + static_cast<long long>(row) * d)[c];
        float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
        if (seed != nullptr) m = drop_mul4(drop_group(row, d, 4 * c), key, thr, dsc);
        const float4 dl = make_float4(g0.x * m.x, g0.y * m.y, g0.z * m.z, g0.w * m.w);  // gradient of LayerNorm's output
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        a_g[i].x += dl.x * xh[i].x; a_g[i].y += dl.y * xh[i].y; a_g[i].z += dl.z * xh[i].z; a_g[i].w += dl.w * xh[i].w;
        a_b[i].x += dl.x; a_b[i].y += dl.y; a_b[i].z += dl.z; a_b[i].w += dl.w;
        a_m[i].x += g0.x; a_m[i].y += g0.y; a_m[i].z += g0.z; a_m[i].w += g0.w;
        gy[i] = make_float4(dl.x * g.x, dl.y * g.y, dl.z * g.z, dl.w * g.w);
        s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
        s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
      }
    }
    const float m1 = warp_sum(s1) / d, m2 = warp_sum(s2) / d;
    const long long id = ids[row];
    const int t = row % T;
#pragma unroll
    for (int i = 0; i < kEmbMaxV4; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) {
        const float4 de = make_float4(rs * (gy[i].x - m1 - xh[i].x * m2), rs * (gy[i].y - m1 - xh[i].y * m2),
                                      rs * (gy[i].z - m1 - xh[i].z * m2), rs * (gy[i].w - m1 - xh[i].w * m2));
        if (id != pad_id) atomicAdd(reinterpret_cast<float4*>(dword + id * d) + c, de);  // nn.Embedding(padding_idx): no gradient
        atomicAdd(reinterpret_cast<float4*>(dpos + static_cast<long long>(t) * d) + c, de);
        a_t[i].x += de.x; a_t[i].y += de.y; a_t[i].z += de.z; a_t[i].w += de.w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kEmbMaxV4; ++i) {
    const int c = i * 32 + lane;
    if (c < nv) {
      const float* src[4] = {&a_g[i].x, &a_b[i].x, &a_t[i].x, &a_m[i].x};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicAdd(&sacc[a * d + 4 * c + e], src[a][e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * d; i += 256) ws[static_cast<long long>(blockIdx.x) * 4 * d + i] = sacc[i];
}

// out[a][j] += sum_p ws[p][a d + j], a < 4
__global__ void __launch_bounds__(256) embed_reduce_kernel(const float* __restrict__ ws, int parts, int d, float* o0, float* o1, float* o2,
                                                           float* o3) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= 4 * d) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += ws[static_cast<long long>(p) * 4 * d + j];
  float* dst = j < d ? o0 : j < 2 * d ? o1 : j < 3 * d ? o2 : o3;
  if (dst != nullptr) dst[j % d] += s;
}

}  // namespace mome

extern "C" int mome_text_embed_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0, const float* ln_w,
                                   const float* ln_b, const float* modal0, float* y, float* xhat, float* rstd, int64_t rows, int32_t T,
                                   int32_t d, float eps, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, void* stream) {
  MOME_REQUIRE(ids && word && pos && type0 && ln_w && ln_b && modal0 && y && xhat && rstd, "text_embed_fwd: null argument");
  MOME_REQUIRE(d % 4 == 0 && d <= kEmbMaxV4 * 128 && T >= 1, "text_embed_fwd: d=%d unsupported (multiple of 4, <= %d)", d, kEmbMaxV4 * 128);
  if (rows == 0) return MOME_OK;
  const bool drop = drop_seed != nullptr && drop_p > 0.f;
  text_embed_fwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), word, pos, type0, ln_w, ln_b, modal0, y, xhat, rstd, static_cast<int>(rows), T, d, eps,
      drop ? drop_seed : nullptr, drop_salt, drop_threshold(drop_p));
  return check_launch("text_embed_fwd");
}

extern "C" size_t mome_text_embed_ws_bytes(int32_t d) { return static_cast<size_t>(sm_count()) * 2 * 4 * d * sizeof(float); }

extern "C" int mome_text_embed_bwd(const float* dy, const int64_t* ids, const float* xhat, const float* rstd, const float* ln_w,
                                   float* dword, float* dpos, float* dtype0, float* dln_w, float* dln_b, float* dmodal0, int64_t rows,
                                   int32_t T, int32_t d, int64_t padding_idx, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p,
                                   void* ws, size_t ws_bytes, void* stream) {
  MOME_REQUIRE(dy && ids && xhat && rstd && ln_w && dword && dpos, "text_embed_bwd: null argument");
  MOME_REQUIRE(d % 4 == 0 && d <= kEmbMaxV4 * 128, "text_embed_bwd: d=%d unsupported", d);
  MOME_REQUIRE(ws != nullptr && ws_bytes >= mome_text_embed_ws_bytes(d), "text_embed_bwd: workspace of %zu bytes needed", mome_text_embed_ws_bytes(d));
  if (rows == 0) return MOME_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, 2LL * sm_count())));
  const bool drop = drop_seed != nullptr && drop_p > 0.f;
  text_embed_bwd_kernel<<<grid, 256, 4 * d * sizeof(float), s>>>(dy, reinterpret_cast<const long long*>(ids), xhat, rstd, ln_w, dword, dpos,
                                                                  static_cast<float*>(ws), static_cast<int>(rows), T, d,
                                                                  drop ? drop_seed : nullptr, drop_salt, drop_threshold(drop_p), padding_idx);
  int rc = check_launch("text_embed_bwd");
  if (rc != MOME_OK) return rc;
  embed_reduce_kernel<<<(4 * d + 255) / 256, 256, 0, s>>>(static_cast<const float*>(ws), grid, d, dln_w, dln_b, dtype0, dmodal0);
  return check_launch("text_embed_reduce");
}
