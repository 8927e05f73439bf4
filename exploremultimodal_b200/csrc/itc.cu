// K4: the ITC contrastive head — L2 normalisation, and the similarity product fused with
// softmax cross-entropy (both directions), its accuracy count and its backward. The [bs, world*bs]
// logit matrices of the reference are never written to HBM: a CTA keeps the logits of a few local
// rows in shared memory, reduces them to (lse, loss, argmax) and only emits the local [bs, bs] block
// that ITM's hard-negative mining reads. The stage is latency bound (bs <= 512, dim 256), so it runs
// on CUDA cores with fp32 accumulation; the gathered features stay L2 resident.
//
// Replaces: reference heads.py:125-126 (F.normalize), objectives.py:99-108 / 166-171 (logits),
// objectives.py:173-180 (cross-entropy, accuracy) and their autograd backward.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"
#include "vec.cuh"

namespace mome {

constexpr int kItcThreads = 256;
constexpr int kItcWarps = kItcThreads / 32;
constexpr int kItcRows = 4;  // local rows (forward / row-backward) or columns (column-backward) per CTA

// ------------------------------------------------------------------------------------------- L2 normalise
template <typename T>
__global__ void __launch_bounds__(kItcThreads) l2norm_fwd_kernel(const T* __restrict__ x, float* __restrict__ y, float* __restrict__ inv_norm,
                                                                 int rows, int dim) {
  const int warp = (blockIdx.x * kItcThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const T* xr = x + static_cast<long long>(warp) * dim;
  float ss = 0.f;
  for (int k = lane * 4; k < dim; k += 128) {
    const float4 v = load4(xr + k);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);  // F.normalize eps
  for (int k = lane * 4; k < dim; k += 128) {
    const float4 v = load4(xr + k);
    store4(y + static_cast<long long>(warp) * dim + k, make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv));
  }
  if (lane == 0) inv_norm[warp] = inv;
}

__global__ void __launch_bounds__(kItcThreads) l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                 const float* __restrict__ inv_norm, float* __restrict__ dx, int rows, int dim) {
  const int warp = (blockIdx.x * kItcThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long off = static_cast<long long>(warp) * dim;
  float dot = 0.f;
  for (int k = lane * 4; k < dim; k += 128) {
    const float4 a = load4(dy + off + k), b = load4(y + off + k);
    dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  dot = warp_sum(dot);
  const float inv = inv_norm[warp];
  for (int k = lane * 4; k < dim; k += 128) {
    const float4 a = load4(dy + off + k), b = load4(y + off + k);
    store4(dx + off + k, make_float4(inv * (a.x - b.x * dot), inv * (a.y - b.y * dot), inv * (a.z - b.z * dot), inv * (a.w - b.w * dot)));
  }
}

// ------------------------------------------------------------------------------------------- shared pieces
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kItcWarps; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// Where the gathered features of every rank live. Either one local [world*bs, dim] array per modality
// (filled by an NCCL all_gather), or — the in-kernel gather — `peers[r]` = rank r's symmetric
// [2, bs, dim] buffer (0 = image, 1 = text features), read directly over NVLink with plain loads.
struct ColSrc {
  const float* all_i;
  const float* all_t;
  const float* const* peers;
  int bs, dim;
  // features of global column c of modality sel (0 image, 1 text)
  __device__ __forceinline__ const float* col(int sel, int c) const {
    if (peers != nullptr) {
      const int r = c / bs;
      return peers[r] + (static_cast<long long>(sel) * bs + (c - r * bs)) * dim;
    }
    return (sel == 0 ? all_i : all_t) + static_cast<long long>(c) * dim;
  }
};

// logits[rr][c] = temp * <f[rr], all[c]> for the CTA's kItcRows local rows and every column c in [0, C)
__device__ __forceinline__ void row_logits(const ColSrc& src, int sel, const float* sf, float* slog, int C, int dim, float temp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < C; c += kItcWarps) {
    float acc[kItcRows];
#pragma unroll
    for (int rr = 0; rr < kItcRows; ++rr) acc[rr] = 0.f;
    for (int k = lane * 4; k < dim; k += 128) {
      const float4 a = *reinterpret_cast<const float4*>(src.col(sel, c) + k);
#pragma unroll
      for (int rr = 0; rr < kItcRows; ++rr) {
        const float4 f = *reinterpret_cast<const float4*>(sf + rr * dim + k);
        acc[rr] += a.x * f.x + a.y * f.y + a.z * f.z + a.w * f.w;
      }
    }
#pragma unroll
    for (int rr = 0; rr < kItcRows; ++rr) {
      const float v = warp_sum(acc[rr]);
      if (lane == 0) slog[rr * C + c] = v * temp;
    }
  }
}

__device__ __forceinline__ void load_local_rows(const float* __restrict__ feat, float* sf, int r0, int bs, int dim) {
  for (int idx = threadIdx.x; idx < kItcRows * dim; idx += kItcThreads) {
    const int rr = idx / dim, k = idx - rr * dim;
    sf[idx] = (r0 + rr < bs) ? feat[static_cast<long long>(r0 + rr) * dim + k] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------- forward
// grid (ceil(bs / kItcRows), 2 directions); dynamic smem: kItcRows * (dim + C) floats
__global__ void __launch_bounds__(kItcThreads) itc_fwd_kernel(const float* __restrict__ i_feat, const float* __restrict__ t_feat,
                                                              const ColSrc src, const float* __restrict__ temp_p,
                                                              int bs, int world, int rank, int dim, float* loss_sum, int32_t* correct,
                                                              float* __restrict__ lse, float* __restrict__ sim_local) {
  extern __shared__ __align__(16) float smem_itc[];
  __shared__ float red[kItcWarps];
  const int dir = blockIdx.y, r0 = blockIdx.x * kItcRows, C = world * bs;
  const float temp = __ldg(temp_p);
  float* sf = smem_itc;
  float* slog = smem_itc + kItcRows * dim;
  load_local_rows(dir == 0 ? i_feat : t_feat, sf, r0, bs, dim);
  __syncthreads();
  row_logits(src, dir == 0 ? 1 : 0, sf, slog, C, dim, temp);
  __syncthreads();
  for (int rr = 0; rr < kItcRows && r0 + rr < bs; ++rr) {
    const int r = r0 + rr, tcol = rank * bs + r;
    const float* lg = slog + rr * C;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < C; c += kItcThreads) m = fmaxf(m, lg[c]);
    m = block_reduce(m, red, true);
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += kItcThreads) s += __expf(lg[c] - m);
    s = block_reduce(s, red, false);
    // accuracy over the local block (reference objectives.py:177-180 uses sim[:, :bs] after the roll)
    float bm = -INFINITY;
    for (int c = threadIdx.x; c < bs; c += kItcThreads) bm = fmaxf(bm, lg[rank * bs + c]);
    bm = block_reduce(bm, red, true);
    float first = 3.0e38f;  // smallest local column index attaining the maximum (torch.argmax tie rule)
    for (int c = threadIdx.x; c < bs; c += kItcThreads)
      if (lg[rank * bs + c] == bm) first = fminf(first, static_cast<float>(c));
    first = -block_reduce(-first, red, true);
    for (int c = threadIdx.x; c < bs; c += kItcThreads)
      sim_local[(static_cast<long long>(dir) * bs + r) * bs + c] = lg[rank * bs + c];
    if (threadIdx.x == 0) {
      const float L = m + __logf(s);
      lse[dir * bs + r] = L;
      atomicAdd(loss_sum + dir, L - lg[tcol]);
      if (static_cast<int>(first) == r) atomicAdd(correct + dir, 1);
    }
  }
}

// ------------------------------------------------------------------------------------------- backward, local-row terms
// d_feat[r] = temp * sum_c g[r, c] * all[c];  d_temp += sum_{r, c} g[r, c] * <f_r, a_c>
// with g = coef * (softmax - onehot), coef = gscale * 0.5 / bs.
__global__ void __launch_bounds__(kItcThreads) itc_bwd_rows_kernel(const float* __restrict__ i_feat, const float* __restrict__ t_feat,
                                                                   const ColSrc src,
                                                                   const float* __restrict__ temp_p, int bs, int world, int rank, int dim,
                                                                   const float* __restrict__ lse, const float* __restrict__ gscale_p, float* __restrict__ d_i_feat,
                                                                   float* __restrict__ d_t_feat, float* d_temp) {
  extern __shared__ __align__(16) float smem_itc[];
  __shared__ float red[kItcWarps];
  const int dir = blockIdx.y, r0 = blockIdx.x * kItcRows, C = world * bs;
  const float temp = __ldg(temp_p);
  const float coef = __ldg(gscale_p) * 0.5f / bs;
  float* sf = smem_itc;
  float* slog = smem_itc + kItcRows * dim;
  const int sel = dir == 0 ? 1 : 0;
  load_local_rows(dir == 0 ? i_feat : t_feat, sf, r0, bs, dim);
  __syncthreads();
  row_logits(src, sel, sf, slog, C, dim, temp);
  __syncthreads();
  float dt = 0.f;
  for (int rr = 0; rr < kItcRows; ++rr) {
    const int r = r0 + rr;
    const float L = r < bs ? lse[dir * bs + r] : 0.f;
    for (int c = threadIdx.x; c < C; c += kItcThreads) {
      float g = 0.f;
      if (r < bs) {
        const float lg = slog[rr * C + c];
        g = coef * (__expf(lg - L) - (c == rank * bs + r ? 1.f : 0.f));
        dt += g * lg;
      }
      slog[rr * C + c] = g;
    }
  }
  dt = block_reduce(dt, red, false);  // also orders the slog writes before the reads below
  if (threadIdx.x == 0) atomicAdd(d_temp, dt / temp);
  float* dfeat = dir == 0 ? d_i_feat : d_t_feat;
  for (int k = threadIdx.x; k < dim; k += kItcThreads) {
    float acc[kItcRows];
#pragma unroll
    for (int rr = 0; rr < kItcRows; ++rr) acc[rr] = 0.f;
    for (int c = 0; c < C; ++c) {
      const float a = src.col(sel, c)[k];
#pragma unroll
      for (int rr = 0; rr < kItcRows; ++rr) acc[rr] = fmaf(slog[rr * C + c], a, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < kItcRows; ++rr)
      if (r0 + rr < bs) dfeat[static_cast<long long>(r0 + rr) * dim + k] = acc[rr] * temp;
  }
}

// ------------------------------------------------------------------------------------------- backward, column terms
// d_all[c] = temp * sum_r g[r, c] * f_r. grid (ceil(C / kItcRows), 2); dynamic smem: kItcRows * (dim + bs) floats
__global__ void __launch_bounds__(kItcThreads) itc_bwd_cols_kernel(const float* __restrict__ i_feat, const float* __restrict__ t_feat,
                                                                   const ColSrc src,
                                                                   const float* __restrict__ temp_p, int bs, int world, int rank, int dim,
                                                                   const float* __restrict__ lse, const float* __restrict__ gscale_p, float* __restrict__ d_all_i,
                                                                   float* __restrict__ d_all_t) {
  extern __shared__ __align__(16) float smem_itc[];
  const int dir = blockIdx.y, c0 = blockIdx.x * kItcRows, C = world * bs;
  const float temp = __ldg(temp_p);
  const float coef = __ldg(gscale_p) * 0.5f / bs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sa = smem_itc;                  // [kItcRows][dim] gathered rows (columns of the logit matrix)
  float* sg = smem_itc + kItcRows * dim;  // [kItcRows][bs]
  const int sel = dir == 0 ? 1 : 0;
  const float* feat = dir == 0 ? i_feat : t_feat;
  for (int idx = threadIdx.x; idx < kItcRows * dim; idx += kItcThreads) {
    const int cc = idx / dim, k = idx - cc * dim;
    sa[idx] = (c0 + cc < C) ? src.col(sel, c0 + cc)[k] : 0.f;
  }
  __syncthreads();
  for (int r = warp; r < bs; r += kItcWarps) {
    float acc[kItcRows];
#pragma unroll
    for (int cc = 0; cc < kItcRows; ++cc) acc[cc] = 0.f;
    for (int k = lane * 4; k < dim; k += 128) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(feat + static_cast<long long>(r) * dim + k));
#pragma unroll
      for (int cc = 0; cc < kItcRows; ++cc) {
        const float4 a = *reinterpret_cast<const float4*>(sa + cc * dim + k);
        acc[cc] += a.x * f.x + a.y * f.y + a.z * f.z + a.w * f.w;
      }
    }
    const float L = lse[dir * bs + r];
#pragma unroll
    for (int cc = 0; cc < kItcRows; ++cc) {
      const float lg = warp_sum(acc[cc]) * temp;
      if (lane == 0) sg[cc * bs + r] = coef * (__expf(lg - L) - (c0 + cc == rank * bs + r ? 1.f : 0.f));
    }
  }
  __syncthreads();
  float* dall = dir == 0 ? d_all_t : d_all_i;
  for (int k = threadIdx.x; k < dim; k += kItcThreads) {
    float acc[kItcRows];
#pragma unroll
    for (int cc = 0; cc < kItcRows; ++cc) acc[cc] = 0.f;
    for (int r = 0; r < bs; ++r) {
      const float f = __ldg(feat + static_cast<long long>(r) * dim + k);
#pragma unroll
      for (int cc = 0; cc < kItcRows; ++cc) acc[cc] = fmaf(sg[cc * bs + r], f, acc[cc]);
    }
#pragma unroll
    for (int cc = 0; cc < kItcRows; ++cc)
      if (c0 + cc < C) dall[static_cast<long long>(c0 + cc) * dim + k] = acc[cc] * temp;
  }
}

template <typename K>
static int opt_in_smem(K kern, size_t bytes, const char* what) {
  if (bytes <= 48 * 1024) return MOME_OK;
  if (bytes > 200 * 1024) {
    set_error("%s: %zu bytes of shared memory needed (world*bs too large)", what, bytes);
    return MOME_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return MOME_ERR_CUDA;
  }
  return MOME_OK;
}

}  // namespace mome

using namespace mome;

extern "C" int mome_l2norm_fwd(const void* x, int x_dtype, float* y, float* inv_norm, int64_t rows, int64_t dim, void* stream) {
  MOME_REQUIRE(dim % 4 == 0, "l2norm_fwd: dim=%lld must be a multiple of 4", (long long)dim);
  if (rows == 0) return MOME_OK;
  const int grid = static_cast<int>((rows + kItcWarps - 1) / kItcWarps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x_dtype == MOME_BF16)
    l2norm_fwd_kernel<__nv_bfloat16><<<grid, kItcThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), y, inv_norm, (int)rows, (int)dim);
  else
    l2norm_fwd_kernel<float><<<grid, kItcThreads, 0, s>>>(static_cast<const float*>(x), y, inv_norm, (int)rows, (int)dim);
  return check_launch("l2norm_fwd");
}

extern "C" int mome_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int64_t rows, int64_t dim, void* stream) {
  MOME_REQUIRE(dim % 4 == 0, "l2norm_bwd: dim=%lld must be a multiple of 4", (long long)dim);
  if (rows == 0) return MOME_OK;
  const int grid = static_cast<int>((rows + kItcWarps - 1) / kItcWarps);
  l2norm_bwd_kernel<<<grid, kItcThreads, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, inv_norm, dx, (int)rows, (int)dim);
  return check_launch("l2norm_bwd");
}

static int itc_fwd_launch(const float* i_feat, const float* t_feat, const ColSrc& src, const float* temp, int32_t bs, int32_t world,
                          int32_t rank, int32_t dim, float* loss_sum, int32_t* correct, float* lse, float* sim_local, cudaStream_t s) {
  const size_t smem = sizeof(float) * kItcRows * (static_cast<size_t>(dim) + static_cast<size_t>(world) * bs);
  int rc = opt_in_smem(itc_fwd_kernel, smem, "itc_fwd");
  if (rc != MOME_OK) return rc;
  cudaMemsetAsync(loss_sum, 0, 2 * sizeof(float), s);
  cudaMemsetAsync(correct, 0, 2 * sizeof(int32_t), s);
  dim3 grid((bs + kItcRows - 1) / kItcRows, 2);
  itc_fwd_kernel<<<grid, kItcThreads, smem, s>>>(i_feat, t_feat, src, temp, bs, world, rank, dim, loss_sum, correct, lse, sim_local);
  return check_launch("itc_fwd");
}

static int itc_bwd_launch(const float* i_feat, const float* t_feat, const ColSrc& src, const float* temp, int32_t bs, int32_t world,
                          int32_t rank, int32_t dim, const float* lse, const float* gscale, float* d_i_feat, float* d_t_feat,
                          float* d_all_i, float* d_all_t, float* d_temp, cudaStream_t s) {
  const int C = world * bs;
  const size_t smem_r = sizeof(float) * kItcRows * (static_cast<size_t>(dim) + C);
  int rc = opt_in_smem(itc_bwd_rows_kernel, smem_r, "itc_bwd(rows)");
  if (rc != MOME_OK) return rc;
  dim3 grid_r((bs + kItcRows - 1) / kItcRows, 2);
  itc_bwd_rows_kernel<<<grid_r, kItcThreads, smem_r, s>>>(i_feat, t_feat, src, temp, bs, world, rank, dim, lse, gscale, d_i_feat, d_t_feat, d_temp);
  rc = check_launch("itc_bwd_rows");
  if (rc != MOME_OK) return rc;
  const size_t smem_c = sizeof(float) * kItcRows * (static_cast<size_t>(dim) + bs);
  rc = opt_in_smem(itc_bwd_cols_kernel, smem_c, "itc_bwd(cols)");
  if (rc != MOME_OK) return rc;
  dim3 grid_c((C + kItcRows - 1) / kItcRows, 2);
  itc_bwd_cols_kernel<<<grid_c, kItcThreads, smem_c, s>>>(i_feat, t_feat, src, temp, bs, world, rank, dim, lse, gscale, d_all_i, d_all_t);
  return check_launch("itc_bwd_cols");
}

#define MOME_ITC_CHECK(name) \
  MOME_REQUIRE(bs > 0 && world > 0 && rank >= 0 && rank < world && dim > 0 && dim % 4 == 0, name ": bad shape bs=%d world=%d rank=%d dim=%d", bs, world, rank, dim)

extern "C" int mome_itc_fwd(const float* i_feat, const float* t_feat, const float* all_i, const float* all_t, const float* temp, int32_t bs,
                            int32_t world, int32_t rank, int32_t dim, float* loss_sum, int32_t* correct, float* lse, float* sim_local,
                            void* stream) {
  MOME_ITC_CHECK("itc_fwd");
  const ColSrc src{all_i, all_t, nullptr, bs, dim};
  return itc_fwd_launch(i_feat, t_feat, src, temp, bs, world, rank, dim, loss_sum, correct, lse, sim_local, static_cast<cudaStream_t>(stream));
}

extern "C" int mome_itc_bwd(const float* i_feat, const float* t_feat, const float* all_i, const float* all_t, const float* temp, int32_t bs,
                            int32_t world, int32_t rank, int32_t dim, const float* lse, const float* gscale, float* d_i_feat, float* d_t_feat,
                            float* d_all_i, float* d_all_t, float* d_temp, void* stream) {
  MOME_ITC_CHECK("itc_bwd");
  const ColSrc src{all_i, all_t, nullptr, bs, dim};
  return itc_bwd_launch(i_feat, t_feat, src, temp, bs, world, rank, dim, lse, gscale, d_i_feat, d_t_feat, d_all_i, d_all_t, d_temp,
                        static_cast<cudaStream_t>(stream));
}

// In-kernel gather variants: `peers` is a DEVICE array of `world` pointers; peers[r] is rank r's [2, bs, dim]
// fp32 feature buffer (0 = image, 1 = text) mapped into this process (NVLink peer / symmetric memory). The
// kernels read the remote rows directly; the caller orders the ranks' writes and reads with barriers.
extern "C" int mome_itc_fwd_peer(const float* i_feat, const float* t_feat, const void* peers, const float* temp, int32_t bs,
                                 int32_t world, int32_t rank, int32_t dim, float* loss_sum, int32_t* correct, float* lse,
                                 float* sim_local, void* stream) {
  MOME_ITC_CHECK("itc_fwd_peer");
  MOME_REQUIRE(peers != nullptr, "itc_fwd_peer: null peer table");
  const ColSrc src{nullptr, nullptr, static_cast<const float* const*>(peers), bs, dim};
  return itc_fwd_launch(i_feat, t_feat, src, temp, bs, world, rank, dim, loss_sum, correct, lse, sim_local, static_cast<cudaStream_t>(stream));
}

extern "C" int mome_itc_bwd_peer(const float* i_feat, const float* t_feat, const void* peers, const float* temp, int32_t bs,
                                 int32_t world, int32_t rank, int32_t dim, const float* lse, const float* gscale, float* d_i_feat,
                                 float* d_t_feat, float* d_all_i, float* d_all_t, float* d_temp, void* stream) {
  MOME_ITC_CHECK("itc_bwd_peer");
  MOME_REQUIRE(peers != nullptr, "itc_bwd_peer: null peer table");
  const ColSrc src{nullptr, nullptr, static_cast<const float* const*>(peers), bs, dim};
  return itc_bwd_launch(i_feat, t_feat, src, temp, bs, world, rank, dim, lse, gscale, d_i_feat, d_t_feat, d_all_i, d_all_t, d_temp,
                        static_cast<cudaStream_t>(stream));
}

// Cross-rank gather over NVLink, once per step: all_i / all_t [world * bs, dim] <- every rank's symmetric
// [2, bs, dim] buffer, read with 128-bit peer loads (GatherLayer.forward's all_gather + cat, objectives.py:401-414).
// Every remote element crosses NVLink exactly once (2 * (world - 1) * bs * dim * 4 bytes per rank), instead of once
// per CTA of the similarity kernel as in the *_peer variants above.
__global__ void __launch_bounds__(256) itc_gather_peer_kernel(const float* const* __restrict__ peers, float* __restrict__ all_i,
                                                              float* __restrict__ all_t, int bs, int dim, int world) {
  const long long per_rank = static_cast<long long>(bs) * dim / 4;  // float4 per modality per rank
  const long long total = 2LL * world * per_rank;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int sel = static_cast<int>(v / (world * per_rank));
    const long long w = v - sel * world * per_rank;
    const int r = static_cast<int>(w / per_rank);
    const long long e = w - r * per_rank;
    const float4 val = reinterpret_cast<const float4*>(peers[r])[sel * per_rank + e];
    reinterpret_cast<float4*>(sel == 0 ? all_i : all_t)[r * per_rank + e] = val;
  }
}

extern "C" int mome_itc_gather_peer(const void* peers, int32_t bs, int32_t world, int32_t dim, float* all_i, float* all_t, void* stream) {
  MOME_REQUIRE(peers != nullptr && all_i != nullptr && all_t != nullptr, "itc_gather_peer: null argument");
  MOME_REQUIRE(bs >= 1 && world >= 1 && dim % 4 == 0, "itc_gather_peer: need bs, world >= 1 and dim %% 4 == 0");
  const long long total = 2LL * world * bs * dim / 4;
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, 2LL * sm_count())));
  itc_gather_peer_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float* const*>(peers), all_i, all_t, bs, dim, world);
  return check_launch("itc_gather_peer");
}
