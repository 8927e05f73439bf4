// K3: the HBM-bound row kernels — LayerNorm forward/backward, LayerScale backward, column sums,
// fp32->bf16 cast. One warp per token row, 128-bit loads/stores, grid sized in multiples of the SM
// count with a grid-stride loop; column reductions (dweight, dbias, dgamma) are accumulated in
// registers across the rows a warp visits, combined through shared memory and flushed with one
// red.add per column per CTA.
//
// Replaces: nn.LayerNorm / apex FusedLayerNorm (reference vlmo.py:26-36, 188-196, 413), the
// LayerScale multiply-adds (vlmo.py:194-196) and their autograd backward.
#include "common.cuh"
#include "ptx.cuh"
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

// Flush per-lane column partials: combine the 8 warps of the CTA in smem, one red.add per column.
template <int NACC, int NV>
__device__ __forceinline__ void flush_columns(float4 (&acc)[NACC][NV], float* const (&dst)[NACC], int nvec,
                                              float* smem /* [8][d] */, int d) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    if (dst[a] == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nvec) *reinterpret_cast<float4*>(smem + warp * d + 4 * c) = acc[a][i];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += kRowThreads) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < kRowThreads / 32; ++wv) s += smem[wv * d + j];
      atomicAdd(dst[a] + j, s);
    }
  }
}

// ------------------------------------------------------------------------------------------- LN bwd
template <typename InT, int NV>
__global__ void __launch_bounds__(kRowThreads) ln_bwd_kernel(const InT* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ w, const float* dres,
                                                             float* dx, float* dw, float* db, long long rows,
                                                             int d) {
  extern __shared__ float smem_cols[];
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * kRowThreads + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * kRowThreads) >> 5;
  float4 acc[2][NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[0][i] = acc[1][i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = warp0; row < rows; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float4 g[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nvec) {
        const float4 dyv = load4(dy + row * d + 4 * c);
        const float4 xv = load4(x + row * d + 4 * c);
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[i] = make_float4(dyv.x * ww.x, dyv.y * ww.y, dyv.z * ww.z, dyv.w * ww.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
        acc[0][i].x += dyv.x * xh[i].x; acc[0][i].y += dyv.y * xh[i].y;
        acc[0][i].z += dyv.z * xh[i].z; acc[0][i].w += dyv.w * xh[i].w;
        acc[1][i].x += dyv.x; acc[1][i].y += dyv.y; acc[1][i].z += dyv.z; acc[1][i].w += dyv.w;
      }
    }
    const float m1 = warp_sum(s1) / d, m2 = warp_sum(s2) / d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nvec) {
        float4 o;
        o.x = rs * (g[i].x - m1 - xh[i].x * m2);
        o.y = rs * (g[i].y - m1 - xh[i].y * m2);
        o.z = rs * (g[i].z - m1 - xh[i].z * m2);
        o.w = rs * (g[i].w - m1 - xh[i].w * m2);
        if (dres != nullptr) {
          const float4 r = load4(dres + row * d + 4 * c);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        store4(dx + row * d + 4 * c, o);
      }
    }
  }
  float* const dst[2] = {dw, db};
  flush_columns<2, NV>(acc, dst, nvec, smem_cols, d);
}

// ------------------------------------------------------------------------------------------- LayerScale bwd
template <typename BrT, typename OutT, int NV>
__global__ void __launch_bounds__(kRowThreads) scale_bwd_kernel(const float* __restrict__ dx, const BrT* __restrict__ branch,
                                                                const float* __restrict__ gamma, OutT* __restrict__ dbranch,
                                                                float* dgamma, float* dbias, long long rows, int d) {
  extern __shared__ float smem_cols[];
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * kRowThreads + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * kRowThreads) >> 5;
  float4 acc[2][NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[0][i] = acc[1][i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = warp0; row < rows; row += nwarps) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nvec) {
        const float4 g = load4(dx + row * d + 4 * c);
        float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
        if (gamma != nullptr) gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        float4 o = make_float4(g.x * gm.x, g.y * gm.y, g.z * gm.z, g.w * gm.w);
        store4(dbranch + row * d + 4 * c, o);
        if (sizeof(OutT) == 2) {  // bias gradient sums what the bf16 consumer will actually see
          o.x = __bfloat162float(__float2bfloat16_rn(o.x)); o.y = __bfloat162float(__float2bfloat16_rn(o.y));
          o.z = __bfloat162float(__float2bfloat16_rn(o.z)); o.w = __bfloat162float(__float2bfloat16_rn(o.w));
        }
        if (dgamma != nullptr) {
          const float4 br = load4(branch + row * d + 4 * c);
          acc[0][i].x += g.x * br.x; acc[0][i].y += g.y * br.y; acc[0][i].z += g.z * br.z; acc[0][i].w += g.w * br.w;
        }
        acc[1][i].x += o.x; acc[1][i].y += o.y; acc[1][i].z += o.z; acc[1][i].w += o.w;
      }
    }
  }
  float* const dst[2] = {dgamma, dbias};
  flush_columns<2, NV>(acc, dst, nvec, smem_cols, d);
}

// ------------------------------------------------------------------------------------------- column sums
// out[j] += sum_r x[r, j]; CTA (bx, by) covers 128 columns x a slab of rows; thread = 4 columns x 1/8 of the slab.
template <typename T>
__global__ void __launch_bounds__(kRowThreads) colsum_kernel(const T* __restrict__ x, long long rows, int cols, long long ld,
                                                             float* out) {
  __shared__ float4 part[kRowThreads];
  const int cv = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cv) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < cols) {
    for (long long r = static_cast<long long>(blockIdx.y) * 8 + ry; r < rows; r += static_cast<long long>(gridDim.y) * 8) {
      const float4 v = load4(x + r * ld + col);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  part[threadIdx.x] = s;
  __syncthreads();
  if (ry == 0 && col < cols) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = part[k * 32 + cv];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    atomicAdd(out + col, s.x); atomicAdd(out + col + 1, s.y); atomicAdd(out + col + 2, s.z); atomicAdd(out + col + 3, s.w);
  }
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

extern "C" int mome_ln_bwd(const void* dy, int dy_dtype, const float* x, const float* mean, const float* rstd,
                           const float* weight, const float* dres, float* dx_out, float* dweight, float* dbias,
                           int64_t rows, int64_t d, void* stream) {
  MOME_REQUIRE(d % 4 == 0 && d <= 1024, "ln_bwd: d=%lld unsupported (multiple of 4, <= 1024)", (long long)d);
  if (rows == 0) return MOME_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = sizeof(float) * (kRowThreads / 32) * d;
  const int grid = std::min(row_grid(rows), sm_count() * 2);
  if (dy_dtype == MOME_BF16)
    MOME_DISPATCH_NV(d, 8, (ln_bwd_kernel<__nv_bfloat16, NV><<<grid, kRowThreads, smem, s>>>(static_cast<const __nv_bfloat16*>(dy), x, mean, rstd, weight, dres, dx_out, dweight, dbias, rows, (int)d)));
  else
    MOME_DISPATCH_NV(d, 8, (ln_bwd_kernel<float, NV><<<grid, kRowThreads, smem, s>>>(static_cast<const float*>(dy), x, mean, rstd, weight, dres, dx_out, dweight, dbias, rows, (int)d)));
  return check_launch("ln_bwd");
}

extern "C" int mome_scale_bwd(const float* dx, const void* branch, int branch_dtype, const float* gamma, void* dbranch,
                              int dbranch_dtype, float* dgamma, float* dbias, int64_t rows, int64_t d, void* stream) {
  MOME_REQUIRE(d % 4 == 0 && d <= 1024, "scale_bwd: d=%lld unsupported (multiple of 4, <= 1024)", (long long)d);
  MOME_REQUIRE(branch_dtype == dbranch_dtype, "scale_bwd: branch and dbranch dtypes must match");
  if (rows == 0) return MOME_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = sizeof(float) * (kRowThreads / 32) * d;
  const int grid = std::min(row_grid(rows), sm_count() * 2);
  if (branch_dtype == MOME_BF16)
    MOME_DISPATCH_NV(d, 8, (scale_bwd_kernel<__nv_bfloat16, __nv_bfloat16, NV><<<grid, kRowThreads, smem, s>>>(dx, static_cast<const __nv_bfloat16*>(branch), gamma, static_cast<__nv_bfloat16*>(dbranch), dgamma, dbias, rows, (int)d)));
  else
    MOME_DISPATCH_NV(d, 8, (scale_bwd_kernel<float, float, NV><<<grid, kRowThreads, smem, s>>>(dx, static_cast<const float*>(branch), gamma, static_cast<float*>(dbranch), dgamma, dbias, rows, (int)d)));
  return check_launch("scale_bwd");
}

extern "C" int mome_colsum(const void* x, int dtype, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream) {
  MOME_REQUIRE(cols % 4 == 0 && ld % 4 == 0, "colsum: cols/ld must be multiples of 4");
  if (rows == 0) return MOME_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>((cols + 127) / 128), 1);
  const long long slabs = std::max<long long>(1, std::min<long long>((rows + 63) / 64, (2LL * sm_count() + grid.x - 1) / grid.x));
  grid.y = static_cast<unsigned>(slabs);
  if (dtype == MOME_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, kRowThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), rows, (int)cols, ld, out);
  else
    colsum_kernel<float><<<grid, kRowThreads, 0, s>>>(static_cast<const float*>(x), rows, (int)cols, ld, out);
  return check_launch("colsum");
}

extern "C" int mome_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n == 0) return MOME_OK;
  MOME_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "cast_bf16: unaligned");
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, sm_count() * 8LL)));
  cast_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), n);
  return check_launch("cast_bf16");
}
