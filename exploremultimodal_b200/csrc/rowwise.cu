// K3: the HBM-bound row kernels — LayerNorm forward/backward, LayerScale backward, column sums,
// fp32->bf16 cast. One warp per token row, 128-bit loads/stores, grid sized in multiples of the SM
// count with a grid-stride loop; column reductions (dweight, dbias, dgamma) are accumulated in
// registers across the rows a warp visits, combined through shared memory and flushed with one
// red.add per column per CTA.
//
// Replaces: nn.LayerNorm / apex FusedLayerNorm (reference vlmo.py:26-36, 188-196, 413), the
// LayerScale multiply-adds (vlmo.py:194-196) and their autograd backward.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"
#include "dropout.cuh"
#include "vec.cuh"

namespace mome {

constexpr int kRowThreads = 256;  // 8 warps / CTA
constexpr int kMaxVec = 16;       // d <= 16 * 128 = 2048

// ------------------------------------------------------------------------------------------- LN fwd
template <typename OutT, int NV>
__global__ void __launch_bounds__(kRowThreads) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, OutT* __restrict__ y,
                                                             float* __restrict__ mean, float* __restrict__ rstd,
                                                             long long rows, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;  // float4 per row
  const long long warp0 = (static_cast<long long>(blockIdx.x) * kRowThreads + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * kRowThreads) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
    const float* xr = x + row * d;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nvec) {
        v[i] = load4(xr + 4 * c);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mu = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nvec) {
        const float a0 = v[i].x - mu, a1 = v[i].y - mu, a2 = v[i].z - mu, a3 = v[i].w - mu;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rs = rsqrtf(warp_sum(q) / d + eps);
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
    OutT* yr = y + row * d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nvec) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + c);
        float4 o;
        o.x = (v[i].x - mu) * rs * ww.x + bb.x;
        o.y = (v[i].y - mu) * rs * ww.y + bb.y;
        o.z = (v[i].z - mu) * rs * ww.z + bb.z;
        o.w = (v[i].w - mu) * rs * ww.w + bb.w;
        store4(yr + 4 * c, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- column-owner kernels
// Backward row kernels need per-column reductions over all rows (dweight, dbias, dgamma). Here a thread
// owns 4 consecutive columns for every row its CTA visits, so those reductions are 4 registers per
// output and one red.add per column per CTA at the end; row statistics (LayerNorm backward needs two
// sums over the row) are a block reduction batched over R rows per iteration. Few registers -> several
// CTAs per SM and R x 3 independent 128-bit loads in flight per thread, which is what an HBM-bound
// kernel needs.
// device-side view of MomeDropout (include/mome.h)
struct DropDev {
  const uint32_t* seed;
  const float* row_scale;
  long long row0;
  uint32_t salt, thr;
};
// mask * scale (and stochastic-depth multiplier) of the branch elements (row0 + r, c..c+3) of a d-wide matrix
__device__ __forceinline__ float4 branch_mul4(const DropDev& dd, uint32_t key, float scale, long long r, int d, int c) {
  float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
  if (dd.seed != nullptr) m = drop_mul4(drop_group(dd.row0 + r, d, c), key, dd.thr, scale);
  if (dd.row_scale != nullptr) {
    const float rs = __ldg(dd.row_scale + dd.row0 + r);
    m.x *= rs; m.y *= rs; m.z *= rs; m.w *= rs;
  }
  return m;
}

// same with the row's stochastic-depth multiplier already loaded (requested together with the row's other operands)
__device__ __forceinline__ float4 branch_mul4_rs(const DropDev& dd, uint32_t key, float scale, long long r, int d, int c, float rs) {
  float4 m = make_float4(rs, rs, rs, rs);
  if (dd.seed != nullptr) {
    const float4 k = drop_mul4(drop_group(dd.row0 + r, d, c), key, dd.thr, scale);
    m.x *= k.x; m.y *= k.y; m.z *= k.z; m.w *= k.w;
  }
  return m;
}

constexpr int kColRows = 4;   // rows per iteration (LayerScale backward)
constexpr int kLnRows = 2;    // rows per iteration (LayerNorm backward: keeps registers low enough for 5-6 CTAs / SM)

template <int NVAL>
__device__ __forceinline__ void block_sums(float (&v)[NVAL], float* red /* [nwarps][NVAL] */, int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NVAL; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NVAL; ++k) red[warp * NVAL + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NVAL; ++k) v[k] = 0.f;
  for (int w = 0; w < nwarps; ++w) {
#pragma unroll
    for (int k = 0; k < NVAL; ++k) v[k] += red[w * NVAL + k];
  }
}

// Column reductions are two-stage and atomic-free: every CTA writes its partial sums to a workspace row
// ([part][array * d + column]) and colreduce_kernel adds the parts into the outputs. (Contended red.add
// on a few thousand addresses serialises in L2 and cost more than the streaming itself.)
struct ColOuts {
  float* out[4];
  int d;
};
__device__ __forceinline__ float4 bf16_round4(float4 v) {
  return make_float4(__bfloat162float(__float2bfloat16_rn(v.x)), __bfloat162float(__float2bfloat16_rn(v.y)),
                     __bfloat162float(__float2bfloat16_rn(v.z)), __bfloat162float(__float2bfloat16_rn(v.w)));
}

// out[a][j] += sum_p ws[p][a * d + j]; CTA (bx, by) = 32 columns x every gridDim.y-th group of 8 parts
// (one part per warp, 8 independent loads in flight per lane); slabs combine with <= 8 red.add per address.
__global__ void __launch_bounds__(256) colreduce_kernel(const float* __restrict__ ws, int nparts, int ncols, ColOuts o) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int stride = 8 * gridDim.y;
  float s = 0.f;
  if (j < ncols) {
    int p = blockIdx.y * 8 + warp;
    for (; p + 7 * stride < nparts; p += 8 * stride) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ws[static_cast<long long>(p + u * stride) * ncols + j];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; p < nparts; p += stride) s += ws[static_cast<long long>(p) * ncols + j];
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && j < ncols) {
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][lane];
    float* dst = o.out[j / o.d];
    if (dst != nullptr) {
      if (gridDim.y == 1) dst[j % o.d] += s;
      else atomicAdd(dst + j % o.d, s);
    }
  }
}

// LayerNorm backward (+ residual gradient), optionally fused with the LayerScale backward of the branch
// that produced this residual stream (FUSE): dbranch = gamma * dx, dgamma += sum dx * branch,
// dbias_br += sum dbranch.
// HOIST: the operands of the second phase (residual gradient, branch) are requested together with dy / x, before the
// block reduction, so that an iteration costs one memory round trip instead of two.
template <typename InT, typename BrT, bool FUSE, bool HOIST>
__global__ void __launch_bounds__(256, 3) ln_bwd_cols_kernel(const InT* __restrict__ dy, const float* __restrict__ x,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ w, const float* __restrict__ dres,
                                                          float* __restrict__ dx,
                                                          const BrT* __restrict__ branch, const float* __restrict__ gamma,
                                                          BrT* __restrict__ dbranch, const DropDev dd, float* __restrict__ ws,
                                                          long long rows, int d) {
  __shared__ float red[2][8 * 2 * kLnRows];
  const uint32_t dkey = (FUSE && dd.seed != nullptr) ? drop_mix(dd.salt, __ldg(dd.seed)) : 0u;
  const float dscale = drop_scale(dd.thr);
  const int c = threadIdx.x * 4;
  const bool active = c < d;
  const int nwarps = blockDim.x >> 5;
  float4 ww = make_float4(0.f, 0.f, 0.f, 0.f), gm = make_float4(1.f, 1.f, 1.f, 1.f);
  if (active) {
    ww = __ldg(reinterpret_cast<const float4*>(w + c));
    if (FUSE && gamma != nullptr) gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
  }
  float4 aw = make_float4(0.f, 0.f, 0.f, 0.f), ab = aw, ag = aw, abb = aw;
  int buf = 0;
  for (long long r0 = static_cast<long long>(blockIdx.x) * kLnRows; r0 < rows; r0 += static_cast<long long>(gridDim.x) * kLnRows) {
    float4 g[kLnRows], xh[kLnRows], rsd[HOIST ? kLnRows : 1], brv[(HOIST && FUSE) ? kLnRows : 1];
    float rs[kLnRows], sums[2 * kLnRows], rowsc[(HOIST && FUSE) ? kLnRows : 1];
#pragma unroll
    for (int j = 0; j < kLnRows; ++j) {
      const long long row = r0 + j;
      g[j] = xh[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      rs[j] = 0.f;
      if (HOIST && row < rows && active) {
        rsd[j] = dres != nullptr ? load4(dres + row * d + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (FUSE) brv[j] = load4(branch + row * d + c);
        if (FUSE) rowsc[j] = dd.row_scale != nullptr ? __ldg(dd.row_scale + dd.row0 + row) : 1.f;
      }
      if (row < rows && active) {
        const float mu = __ldg(mean + row);
        rs[j] = __ldg(rstd + row);
        const float4 dyv = load4(dy + row * d + c);
        const float4 xv = load4(x + row * d + c);
        xh[j] = make_float4((xv.x - mu) * rs[j], (xv.y - mu) * rs[j], (xv.z - mu) * rs[j], (xv.w - mu) * rs[j]);
        g[j] = make_float4(dyv.x * ww.x, dyv.y * ww.y, dyv.z * ww.z, dyv.w * ww.w);
        aw.x += dyv.x * xh[j].x; aw.y += dyv.y * xh[j].y; aw.z += dyv.z * xh[j].z; aw.w += dyv.w * xh[j].w;
        ab.x += dyv.x; ab.y += dyv.y; ab.z += dyv.z; ab.w += dyv.w;
      }
      sums[2 * j] = (g[j].x + g[j].y) + (g[j].z + g[j].w);
      sums[2 * j + 1] = (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
    }
    block_sums<2 * kLnRows>(sums, red[buf], nwarps);
    buf ^= 1;
    const float invd = 1.f / d;
#pragma unroll
    for (int j = 0; j < kLnRows; ++j) {
      const long long row = r0 + j;
      if (row < rows && active) {
        const float m1 = sums[2 * j] * invd, m2 = sums[2 * j + 1] * invd;
        float4 o;
        o.x = rs[j] * (g[j].x - m1 - xh[j].x * m2);
        o.y = rs[j] * (g[j].y - m1 - xh[j].y * m2);
        o.z = rs[j] * (g[j].z - m1 - xh[j].z * m2);
        o.w = rs[j] * (g[j].w - m1 - xh[j].w * m2);
        if (HOIST) {
          o.x += rsd[j].x; o.y += rsd[j].y; o.z += rsd[j].z; o.w += rsd[j].w;
        } else if (dres != nullptr) {
          const float4 r = load4(dres + row * d + c);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        store4(dx + row * d + c, o);
        if (FUSE) {
          const float4 br = HOIST ? brv[j] : load4(branch + row * d + c);
          // x1 = x + gamma * row_scale * drop(branch): `br` is the stored dropped branch, bm = mask * scale * row_scale
          float rs = 1.f;
          if (HOIST) rs = rowsc[j];
          else if (dd.row_scale != nullptr) rs = __ldg(dd.row_scale + dd.row0 + row);
          const float4 bm = branch_mul4_rs(dd, dkey, dscale, row, d, c, rs);
          float4 dbr = make_float4(o.x * gm.x * bm.x, o.y * gm.y * bm.y, o.z * gm.z * bm.z, o.w * gm.w * bm.w);
          store4(dbranch + row * d + c, dbr);
          if (sizeof(BrT) == 2) dbr = bf16_round4(dbr);  // the bias gradient sums what the bf16 consumer sees
          ag.x += o.x * br.x * rs; ag.y += o.y * br.y * rs; ag.z += o.z * br.z * rs; ag.w += o.w * br.w * rs;
          abb.x += dbr.x; abb.y += dbr.y; abb.z += dbr.z; abb.w += dbr.w;
        }
      }
    }
  }
  if (active) {
    float* part = ws + static_cast<long long>(blockIdx.x) * ((FUSE ? 4 : 2) * d) + c;
    *reinterpret_cast<float4*>(part) = aw;
    *reinterpret_cast<float4*>(part + d) = ab;
    if (FUSE) {
      *reinterpret_cast<float4*>(part + 2 * d) = ag;
      *reinterpret_cast<float4*>(part + 3 * d) = abb;
    }
  }
}

// ------------------------------------------------------------------------------------------- LayerScale bwd
template <typename BrT>
__global__ void __launch_bounds__(256) scale_bwd_cols_kernel(const float* __restrict__ dx, const BrT* __restrict__ branch,
                                                             const float* __restrict__ gamma, BrT* __restrict__ dbranch,
                                                             bool need_dgamma, const DropDev dd, float* __restrict__ ws, long long rows,
                                                             int d) {
  const int c = threadIdx.x * 4;
  if (c >= d) return;
  const uint32_t dkey = dd.seed != nullptr ? drop_mix(dd.salt, __ldg(dd.seed)) : 0u;
  const float dscale = drop_scale(dd.thr);
  float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
  if (gamma != nullptr) gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
  for (long long r0 = static_cast<long long>(blockIdx.x) * kColRows; r0 < rows; r0 += static_cast<long long>(gridDim.x) * kColRows) {
    float4 g[kColRows], br[kColRows];
    float rsv[kColRows];
#pragma unroll
    for (int j = 0; j < kColRows; ++j) {
      g[j] = br[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      rsv[j] = 1.f;
      if (r0 + j < rows) {
        g[j] = load4(dx + (r0 + j) * d + c);
        if (need_dgamma) br[j] = load4(branch + (r0 + j) * d + c);
        if (dd.row_scale != nullptr) rsv[j] = __ldg(dd.row_scale + dd.row0 + r0 + j);
      }
    }
#pragma unroll
    for (int j = 0; j < kColRows; ++j) {
      if (r0 + j < rows) {
        const float rs = rsv[j];
        const float4 bm = branch_mul4_rs(dd, dkey, dscale, r0 + j, d, c, rs);
        float4 o = make_float4(g[j].x * gm.x * bm.x, g[j].y * gm.y * bm.y, g[j].z * gm.z * bm.z, g[j].w * gm.w * bm.w);
        store4(dbranch + (r0 + j) * d + c, o);
        if (sizeof(BrT) == 2) o = bf16_round4(o);
        ag.x += g[j].x * br[j].x * rs; ag.y += g[j].y * br[j].y * rs; ag.z += g[j].z * br[j].z * rs; ag.w += g[j].w * br[j].w * rs;
        ab.x += o.x; ab.y += o.y; ab.z += o.z; ab.w += o.w;
      }
    }
  }
  float* part = ws + static_cast<long long>(blockIdx.x) * (2 * d) + c;
  *reinterpret_cast<float4*>(part) = ag;
  *reinterpret_cast<float4*>(part + d) = ab;
}

// ------------------------------------------------------------------------------------------- column sums
// ws[by][j] = sum over the CTA's rows of x[r, j]; CTA (bx, by): 1024 columns x every gridDim.y-th group of 8 rows.
// `gap_at` / `gap`: columns >= gap_at are read `gap` columns further right (q and v thirds of dqkv in one launch).
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long rows, int cols, long long ld,
                                                     float* __restrict__ ws, int gap_at, int gap) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= cols) return;
  x += (c >= gap_at ? gap : 0);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r0 = static_cast<long long>(blockIdx.y) * 8; r0 < rows; r0 += static_cast<long long>(gridDim.y) * 8) {
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (r0 + j < rows) ? load4(x + (r0 + j) * ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; }
  }
  *reinterpret_cast<float4*>(ws + static_cast<long long>(blockIdx.y) * cols + c) = s;
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      store4(dst + i, load4(src + i));
    } else {
      for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
    }
  }
}

static int row_grid(long long rows) {
  const long long want = (rows + (kRowThreads / 32) - 1) / (kRowThreads / 32);
  const long long cap = static_cast<long long>(sm_count()) * 8;
  return static_cast<int>(std::max<long long>(1, std::min(want, cap)));
}


// NV = float4 chunks per lane (ceil(d / 128)); instantiated for the model widths in use.
#define MOME_DISPATCH_NV(d, MAXNV, ...)                                   \
  do {                                                                    \
    const int nv_ = static_cast<int>(((d) + 127) / 128);                  \
    if (nv_ <= 1) { constexpr int NV = 1; __VA_ARGS__; }                  \
    else if (nv_ <= 2) { constexpr int NV = 2; __VA_ARGS__; }             \
    else if (nv_ <= 4) { constexpr int NV = 4; __VA_ARGS__; }             \
    else if (nv_ <= 6) { constexpr int NV = 6; __VA_ARGS__; }             \
    else if (nv_ <= 8) { constexpr int NV = 8; __VA_ARGS__; }             \
    else { constexpr int NV = MAXNV; __VA_ARGS__; }                       \
  } while (0)

}  // namespace mome

using namespace mome;

extern "C" int mome_ln_fwd(const float* x, const float* weight, const float* bias, void* y, int y_dtype, float* mean,
                           float* rstd, int64_t rows, int64_t d, float eps, void* stream) {
  MOME_REQUIRE(d % 4 == 0 && d <= kMaxVec * 128, "ln_fwd: d=%lld must be a multiple of 4 and <= %d", (long long)d, kMaxVec * 128);
  if (rows == 0) return MOME_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (y_dtype == MOME_BF16)
    MOME_DISPATCH_NV(d, 16, (ln_fwd_kernel<__nv_bfloat16, NV><<<row_grid(rows), kRowThreads, 0, s>>>(x, weight, bias, static_cast<__nv_bfloat16*>(y), mean, rstd, rows, (int)d, eps)));
  else
    MOME_DISPATCH_NV(d, 16, (ln_fwd_kernel<float, NV><<<row_grid(rows), kRowThreads, 0, s>>>(x, weight, bias, static_cast<float*>(y), mean, rstd, rows, (int)d, eps)));
  return check_launch("ln_fwd");
}

static DropDev drop_dev(const MomeDropout* drop) {
  DropDev dd{nullptr, nullptr, 0, 0u, 0u};
  if (drop != nullptr) {
    dd.seed = (drop->seed != nullptr && drop->p > 0.f) ? drop->seed : nullptr;
    dd.row_scale = drop->row_scale;
    dd.row0 = drop->row0;
    dd.salt = drop->salt;
    dd.thr = drop_threshold(drop->p);
  }
  return dd;
}

// stochastic-depth multipliers: one draw per (sample, salt), 24-bit resolution
__global__ void droppath_scales_kernel(const int32_t* __restrict__ row_sample, long long rows, const uint32_t* __restrict__ seed,
                                       uint32_t salt, float p, float* __restrict__ out) {
  const uint32_t key = drop_mix(salt, __ldg(seed));
  const uint32_t thr = static_cast<uint32_t>(p * 16777216.f);
  const float keep = 1.f / (1.f - p);
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows; r += static_cast<long long>(gridDim.x) * blockDim.x)
    out[r] = (drop_mix(static_cast<uint32_t>(row_sample[r]), key) >> 8) >= thr ? keep : 0.f;
}

// both branches of a block (salts differ) in one launch: internal to block.cu
__global__ void droppath_scales2_kernel(const int32_t* __restrict__ row_sample, long long rows, const uint32_t* __restrict__ seed,
                                        uint32_t salt1, uint32_t salt2, float p, float* __restrict__ out1, float* __restrict__ out2) {
  const uint32_t s = __ldg(seed);
  const uint32_t key1 = drop_mix(salt1, s), key2 = drop_mix(salt2, s);
  const uint32_t thr = static_cast<uint32_t>(p * 16777216.f);
  const float keep = 1.f / (1.f - p);
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows; r += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t smp = static_cast<uint32_t>(row_sample[r]);
    out1[r] = (drop_mix(smp, key1) >> 8) >= thr ? keep : 0.f;
    out2[r] = (drop_mix(smp, key2) >> 8) >= thr ? keep : 0.f;
  }
}
namespace mome {
int droppath_scales2(const int32_t* row_sample, int64_t rows, const uint32_t* seed, uint32_t salt1, uint32_t salt2, float p, float* out1,
                     float* out2, cudaStream_t stream) {
  MOME_REQUIRE(p >= 0.f && p < 1.f && seed != nullptr, "droppath_scales: need 0 <= p < 1 and a seed");
  if (rows == 0) return MOME_OK;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((rows + 255) / 256, sm_count() * 4LL)));
  droppath_scales2_kernel<<<grid, 256, 0, stream>>>(row_sample, rows, seed, salt1, salt2, p, out1, out2);
  return check_launch("droppath_scales2");
}
}  // namespace mome

static int col_threads(int64_t d) { return static_cast<int>(((d / 4) + 31) / 32 * 32); }
static int col_grid(int64_t rows, int ctas_per_sm, int rows_per_iter = kColRows) {
  const long long groups = (rows + rows_per_iter - 1) / rows_per_iter;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(groups, static_cast<long long>(sm_count()) * ctas_per_sm)));
}
static int colreduce_launch(const float* ws, int nparts, int ncols, const ColOuts& o, cudaStream_t s) {
  dim3 grid((ncols + 31) / 32, std::max(1, std::min(8, nparts / 64)));
  colreduce_kernel<<<grid, 256, 0, s>>>(ws, nparts, ncols, o);
  return check_launch("colreduce");
}
#define MOME_REQUIRE_WS(name, need)                                                                             \
  MOME_REQUIRE(ws != nullptr && ws_bytes >= (need), "%s: workspace of %zu bytes needed (mome_reduce_ws_bytes), got %zu", \
               name, static_cast<size_t>(need), static_cast<size_t>(ws_bytes))

extern "C" size_t mome_reduce_ws_bytes(int64_t cols) {
  // the largest stage-1 grid any row kernel uses (6 CTAs per SM) x 4 arrays of `cols` floats
  return static_cast<size_t>(sm_count()) * 6 * 4 * static_cast<size_t>(cols) * sizeof(float);
}

template <bool FUSE>
static int ln_bwd_launch(const void* dy, int dy_dtype, const float* x, const float* mean, const float* rstd, const float* weight,
                         const float* dres, float* dx_out, float* dweight, float* dbias, const void* branch, const float* gamma,
                         void* dbranch, float* dgamma, float* dbias_br, int64_t rows, int64_t d, const DropDev& dd, float* ws,
                         cudaStream_t s) {
  static const int variant = [] { const char* e = getenv("MOME_LN_BWD_VARIANT"); return e ? atoi(e) : 1; }();  // 0: two-phase loads
  const int threads = col_threads(d);
  // persistent grid = exactly the CTAs that are resident at once (a partial second wave would start late and finish last)
  int per_sm = 4;
  {
    const void* fn = dy_dtype != MOME_BF16 ? reinterpret_cast<const void*>(&ln_bwd_cols_kernel<float, float, FUSE, false>)
                     : variant == 0      ? reinterpret_cast<const void*>(&ln_bwd_cols_kernel<__nv_bfloat16, __nv_bfloat16, FUSE, false>)
                                         : reinterpret_cast<const void*>(&ln_bwd_cols_kernel<__nv_bfloat16, __nv_bfloat16, FUSE, true>);
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, 0) == cudaSuccess && occ > 0) per_sm = std::min(occ, 6);
  }
  const int grid = col_grid(rows, per_sm, kLnRows);
  if (dy_dtype == MOME_BF16 && variant == 0)
    ln_bwd_cols_kernel<__nv_bfloat16, __nv_bfloat16, FUSE, false><<<grid, threads, 0, s>>>(
        static_cast<const __nv_bfloat16*>(dy), x, mean, rstd, weight, dres, dx_out, static_cast<const __nv_bfloat16*>(branch), gamma,
        static_cast<__nv_bfloat16*>(dbranch), dd, ws, rows, (int)d);
  else if (dy_dtype == MOME_BF16)
    ln_bwd_cols_kernel<__nv_bfloat16, __nv_bfloat16, FUSE, true><<<grid, threads, 0, s>>>(
        static_cast<const __nv_bfloat16*>(dy), x, mean, rstd, weight, dres, dx_out, static_cast<const __nv_bfloat16*>(branch), gamma,
        static_cast<__nv_bfloat16*>(dbranch), dd, ws, rows, (int)d);
  else
    ln_bwd_cols_kernel<float, float, FUSE, false><<<grid, threads, 0, s>>>(static_cast<const float*>(dy), x, mean, rstd, weight, dres, dx_out,
                                                                     static_cast<const float*>(branch), gamma,
                                                                     static_cast<float*>(dbranch), dd, ws, rows, (int)d);
  int rc = check_launch(FUSE ? "ln_bwd_scale" : "ln_bwd");
  if (rc != MOME_OK) return rc;
  ColOuts o{{dweight, dbias, dgamma, dbias_br}, (int)d};
  return colreduce_launch(ws, grid, (FUSE ? 4 : 2) * (int)d, o, s);
}

extern "C" int mome_ln_bwd(const void* dy, int dy_dtype, const float* x, const float* mean, const float* rstd,
                           const float* weight, const float* dres, float* dx_out, float* dweight, float* dbias,
                           int64_t rows, int64_t d, void* ws, size_t ws_bytes, void* stream) {
  MOME_REQUIRE(d % 4 == 0 && d <= 1024, "ln_bwd: d=%lld unsupported (multiple of 4, <= 1024)", (long long)d);
  MOME_REQUIRE_WS("ln_bwd", mome_reduce_ws_bytes(d));
  if (rows == 0) return MOME_OK;
  return ln_bwd_launch<false>(dy, dy_dtype, x, mean, rstd, weight, dres, dx_out, dweight, dbias, nullptr, nullptr, nullptr, nullptr,
                              nullptr, rows, d, drop_dev(nullptr), static_cast<float*>(ws), static_cast<cudaStream_t>(stream));
}

extern "C" int mome_ln_bwd_scale(const void* dy, int dtype, const float* x, const float* mean, const float* rstd,
                                 const float* weight, const float* dres, float* dx_out, float* dweight, float* dbias,
                                 const void* branch, const float* gamma, void* dbranch, float* dgamma, float* dbias_branch,
                                 int64_t rows, int64_t d, const MomeDropout* drop, void* ws, size_t ws_bytes, void* stream) {
  MOME_REQUIRE(d % 4 == 0 && d <= 1024, "ln_bwd_scale: d=%lld unsupported (multiple of 4, <= 1024)", (long long)d);
  MOME_REQUIRE(branch != nullptr && dbranch != nullptr, "ln_bwd_scale: branch / dbranch must be given");
  MOME_REQUIRE_WS("ln_bwd_scale", mome_reduce_ws_bytes(d));
  if (rows == 0) return MOME_OK;
  return ln_bwd_launch<true>(dy, dtype, x, mean, rstd, weight, dres, dx_out, dweight, dbias, branch, gamma, dbranch, dgamma,
                             dbias_branch, rows, d, drop_dev(drop), static_cast<float*>(ws), static_cast<cudaStream_t>(stream));
}

extern "C" int mome_scale_bwd(const float* dx, const void* branch, int branch_dtype, const float* gamma, void* dbranch,
                              int dbranch_dtype, float* dgamma, float* dbias, int64_t rows, int64_t d, const MomeDropout* drop,
                              void* ws, size_t ws_bytes, void* stream) {
  MOME_REQUIRE(d % 4 == 0 && d <= 1024, "scale_bwd: d=%lld unsupported (multiple of 4, <= 1024)", (long long)d);
  MOME_REQUIRE(branch_dtype == dbranch_dtype, "scale_bwd: branch and dbranch dtypes must match");
  MOME_REQUIRE_WS("scale_bwd", mome_reduce_ws_bytes(d));
  if (rows == 0) return MOME_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = col_threads(d), grid = col_grid(rows, 6);
  float* w = static_cast<float*>(ws);
  const DropDev dd = drop_dev(drop);
  if (branch_dtype == MOME_BF16)
    scale_bwd_cols_kernel<__nv_bfloat16><<<grid, threads, 0, s>>>(dx, static_cast<const __nv_bfloat16*>(branch), gamma,
                                                                   static_cast<__nv_bfloat16*>(dbranch), dgamma != nullptr, dd, w, rows, (int)d);
  else
    scale_bwd_cols_kernel<float><<<grid, threads, 0, s>>>(dx, static_cast<const float*>(branch), gamma, static_cast<float*>(dbranch),
                                                          dgamma != nullptr, dd, w, rows, (int)d);
  int rc = check_launch("scale_bwd");
  if (rc != MOME_OK) return rc;
  ColOuts o{{dgamma, dbias, nullptr, nullptr}, (int)d};
  return colreduce_launch(w, grid, 2 * (int)d, o, s);
}

extern "C" int mome_colsum(const void* x, int dtype, int64_t rows, int64_t cols, int64_t ld, float* out, void* ws, size_t ws_bytes,
                           void* stream) {
  MOME_REQUIRE(cols % 4 == 0 && ld % 4 == 0, "colsum: cols/ld must be multiples of 4");
  if (rows == 0) return MOME_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>((cols + 1023) / 1024), 1);
  const long long slabs = std::max<long long>(1, std::min<long long>((rows + 7) / 8, (6LL * sm_count() + grid.x - 1) / grid.x));
  grid.y = static_cast<unsigned>(slabs);
  MOME_REQUIRE_WS("colsum", static_cast<size_t>(slabs) * cols * sizeof(float));
  float* w = static_cast<float*>(ws);
  if (dtype == MOME_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), rows, (int)cols, ld, w, (int)cols, 0);
  else
    colsum_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), rows, (int)cols, ld, w, (int)cols, 0);
  int rc = check_launch("colsum");
  if (rc != MOME_OK) return rc;
  ColOuts o{{out, nullptr, nullptr, nullptr}, (int)cols};
  return colreduce_launch(w, (int)slabs, (int)cols, o, s);
}

// dq_bias[j] += sum_r dqkv[r, j], dv_bias[j] += sum_r dqkv[r, 2 d + j]: the q and v thirds of dqkv [rows, 3 d] in one pass
// (the k third has no bias, reference vlmo.py:72-75). Internal to the block sequencer (csrc/block.cu).
namespace mome {
int colsum_qv(const void* dqkv, int dtype, int64_t rows, int64_t d, float* dq_bias, float* dv_bias, void* ws, size_t ws_bytes,
              cudaStream_t s) {
  MOME_REQUIRE(d % 4 == 0, "colsum_qv: d must be a multiple of 4");
  if (rows == 0) return MOME_OK;
  const int cols = static_cast<int>(2 * d);
  dim3 grid(static_cast<unsigned>((cols + 1023) / 1024), 1);
  const long long slabs = std::max<long long>(1, std::min<long long>((rows + 7) / 8, (6LL * sm_count() + grid.x - 1) / grid.x));
  grid.y = static_cast<unsigned>(slabs);
  MOME_REQUIRE_WS("colsum_qv", static_cast<size_t>(slabs) * cols * sizeof(float));
  float* w = static_cast<float*>(ws);
  if (dtype == MOME_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(dqkv), rows, cols, 3 * d, w, (int)d, (int)d);
  else
    colsum_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(dqkv), rows, cols, 3 * d, w, (int)d, (int)d);
  int rc = check_launch("colsum_qv");
  if (rc != MOME_OK) return rc;
  ColOuts o{{dq_bias, dv_bias, nullptr, nullptr}, (int)d};
  return colreduce_launch(w, (int)slabs, cols, o, s);
}
}  // namespace mome

extern "C" int mome_droppath_scales(const int32_t* row_sample, int64_t rows, const uint32_t* seed, uint32_t salt, float p, float* out,
                                    void* stream) {
  MOME_REQUIRE(p >= 0.f && p < 1.f && seed != nullptr, "droppath_scales: need 0 <= p < 1 and a seed");
  if (rows == 0) return MOME_OK;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((rows + 255) / 256, sm_count() * 4LL)));
  droppath_scales_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(row_sample, rows, seed, salt, p, out);
  return check_launch("droppath_scales");
}

// out[j] += sum_p partials[p][j] (second stage for partial column sums written by a GEMM epilogue)
extern "C" int mome_colreduce(const float* partials, int64_t nparts, int64_t cols, float* out, void* stream) {
  if (nparts == 0 || cols == 0) return MOME_OK;
  ColOuts o{{out, nullptr, nullptr, nullptr}, (int)cols};
  return colreduce_launch(partials, (int)nparts, (int)cols, o, static_cast<cudaStream_t>(stream));
}

extern "C" int mome_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n == 0) return MOME_OK;
  MOME_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "cast_bf16: unaligned");
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, sm_count() * 8LL)));
  cast_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), n);
  return check_launch("cast_bf16");
}
