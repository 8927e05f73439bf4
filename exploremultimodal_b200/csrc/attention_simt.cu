// K1 (validation / fp32 path): masked multi-head self-attention over packed segments on CUDA cores.
//
// One thread owns one query row (forward, dq) or one key row (dk, dv) of a (sequence, head) pair and
// keeps its 64-wide vectors in registers; the other side of the product streams through shared
// memory in 64-row tiles read with warp-wide broadcasts. Softmax is the online (running max / sum)
// form, so the [B,H,N,N] probability tensor of the reference is never materialised; the backward
// recomputes probabilities from the saved log-sum-exp.
//
// This is the fp32 path (1e-4 parity against the oracle) and the cross-check for the tensor-core
// kernel in attention_mma.cu, which serves bf16.
//
// Replaces: reference vlmo.py:79-95 (split heads, q k^T * scale, masked_fill(~mask, -inf), softmax,
// @ v, merge heads) and its autograd backward.
#include "common.cuh"
#include "ptx.cuh"
#include "vec.cuh"

namespace mome {

constexpr int kHd = 64;        // head_dim of every VLMo size (768/12, 1024/16)
constexpr int kAttThreads = 128;
constexpr int kAttTile = 64;   // rows of the streamed operand per shared-memory tile

struct SeqDesc {
  int start0, len0, start1, len1;
};
__device__ __forceinline__ SeqDesc load_seq(const int32_t* seq_desc, int s) {
  const int4 v = *reinterpret_cast<const int4*>(seq_desc + 4 * s);
  return SeqDesc{v.x, v.y, v.z, v.w};
}
// packed-buffer row of logical token i of the sequence ([range0 | range1])
__device__ __forceinline__ long long seq_row(const SeqDesc& sd, int i) {
  return i < sd.len0 ? static_cast<long long>(sd.start0) + i : static_cast<long long>(sd.start1) + (i - sd.len0);
}

template <typename T>
__device__ __forceinline__ void load_row64(const T* p, float (&r)[kHd]) {
#pragma unroll
  for (int e = 0; e < kHd; e += 4) {
    const float4 v = load4(p + e);
    r[e] = v.x; r[e + 1] = v.y; r[e + 2] = v.z; r[e + 3] = v.w;
  }
}
template <typename T>
__device__ __forceinline__ void store_row64(T* p, const float (&r)[kHd], float mul) {
#pragma unroll
  for (int e = 0; e < kHd; e += 4) store4(p + e, make_float4(r[e] * mul, r[e + 1] * mul, r[e + 2] * mul, r[e + 3] * mul));
}

// Cooperative load of up to kAttTile rows (logical tokens [t0, t0 + kAttTile) of the sequence) of one
// 64-wide head slice into shared memory as fp32; rows past the sequence end are zero.
template <typename T>
__device__ __forceinline__ void load_tile(float (*dst)[kHd], const T* base, long long ld, const SeqDesc& sd, int n, int t0) {
  for (int idx = threadIdx.x; idx < kAttTile * (kHd / 4); idx += kAttThreads) {
    const int j = idx >> 4, c = (idx & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t0 + j < n) v = load4(base + seq_row(sd, t0 + j) * ld + c);
    *reinterpret_cast<float4*>(&dst[j][c]) = v;
  }
}

__device__ __forceinline__ float dot64(const float (&a)[kHd], const float* b) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int e = 0; e < kHd; e += 4) {
    const float4 v = *reinterpret_cast<const float4*>(b + e);
    s0 = fmaf(a[e], v.x, s0); s1 = fmaf(a[e + 1], v.y, s1); s2 = fmaf(a[e + 2], v.z, s2); s3 = fmaf(a[e + 3], v.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ void axpy64(float (&acc)[kHd], float a, const float* b) {
#pragma unroll
  for (int e = 0; e < kHd; e += 4) {
    const float4 v = *reinterpret_cast<const float4*>(b + e);
    acc[e] = fmaf(a, v.x, acc[e]); acc[e + 1] = fmaf(a, v.y, acc[e + 1]);
    acc[e + 2] = fmaf(a, v.z, acc[e + 2]); acc[e + 3] = fmaf(a, v.w, acc[e + 3]);
  }
}

// ------------------------------------------------------------------------------------------- forward
// grid (query tiles of 128, heads, sequences)
template <typename T>
__global__ void __launch_bounds__(kAttThreads) attn_fwd_simt_kernel(const T* __restrict__ qkv, const int32_t* __restrict__ seq_desc,
                                                                    const uint8_t* __restrict__ key_mask, T* __restrict__ out,
                                                                    float* __restrict__ lse, int H, int max_seq_len, float scale) {
  __shared__ __align__(16) float Ks[kAttTile][kHd];
  __shared__ __align__(16) float Vs[kAttTile][kHd];
  __shared__ uint8_t keep[kAttTile];
  const int s = blockIdx.z, h = blockIdx.y;
  const SeqDesc sd = load_seq(seq_desc, s);
  const int n = sd.len0 + sd.len1;
  if (blockIdx.x * kAttThreads >= n) return;
  const int d = H * kHd;
  const long long ld = 3LL * d;
  const int i = blockIdx.x * kAttThreads + threadIdx.x;
  const bool valid = i < n;
  const long long row = valid ? seq_row(sd, i) : 0;

  float q[kHd], acc[kHd];
  if (valid) load_row64(qkv + row * ld + h * kHd, q);
#pragma unroll
  for (int e = 0; e < kHd; ++e) acc[e] = 0.f;
  float m = -INFINITY, l = 0.f;

  for (int t0 = 0; t0 < n; t0 += kAttTile) {
    __syncthreads();
    load_tile(Ks, qkv + d + h * kHd, ld, sd, n, t0);
    load_tile(Vs, qkv + 2 * d + h * kHd, ld, sd, n, t0);
    if (threadIdx.x < kAttTile) {
      const int j = t0 + threadIdx.x;
      keep[threadIdx.x] = (j < n) && (key_mask == nullptr || key_mask[seq_row(sd, j)] != 0);
    }
    __syncthreads();
    if (!valid) continue;
    const int jn = min(kAttTile, n - t0);
    for (int j = 0; j < jn; ++j) {
      if (!keep[j]) continue;  // warp-uniform
      const float sc = dot64(q, Ks[j]) * scale;
      if (sc > m) {
        const float corr = __expf(m - sc);  // exp(-inf) = 0 on the first kept key
        l *= corr;
#pragma unroll
        for (int e = 0; e < kHd; ++e) acc[e] *= corr;
        m = sc;
      }
      const float p = __expf(sc - m);
      l += p;
      axpy64(acc, p, Vs[j]);
    }
  }
  if (!valid) return;
  const float inv = l > 0.f ? 1.f / l : 0.f;
  store_row64(out + row * d + h * kHd, acc, inv);
  lse[(static_cast<long long>(s) * H + h) * max_seq_len + i] = l > 0.f ? m + __logf(l) : -INFINITY;
}

// ------------------------------------------------------------------------------------------- backward: dq (+ delta)
template <typename T>
__global__ void __launch_bounds__(kAttThreads) attn_bwd_dq_simt_kernel(const T* __restrict__ qkv, const T* __restrict__ out,
                                                                       const T* __restrict__ dout, const int32_t* __restrict__ seq_desc,
                                                                       const uint8_t* __restrict__ key_mask, const float* __restrict__ lse,
                                                                       T* __restrict__ dqkv, float* __restrict__ delta_ws, int H,
                                                                       int max_seq_len, float scale) {
  __shared__ __align__(16) float Ks[kAttTile][kHd];
  __shared__ __align__(16) float Vs[kAttTile][kHd];
  __shared__ uint8_t keep[kAttTile];
  const int s = blockIdx.z, h = blockIdx.y;
  const SeqDesc sd = load_seq(seq_desc, s);
  const int n = sd.len0 + sd.len1;
  if (blockIdx.x * kAttThreads >= n) return;
  const int d = H * kHd;
  const long long ld = 3LL * d;
  const int i = blockIdx.x * kAttThreads + threadIdx.x;
  const bool valid = i < n;
  const long long row = valid ? seq_row(sd, i) : 0;
  const long long stat = (static_cast<long long>(s) * H + h) * max_seq_len + i;

  float q[kHd], go[kHd], dq[kHd];
  float delta = 0.f, L = 0.f;
  if (valid) {
    load_row64(qkv + row * ld + h * kHd, q);
    load_row64(dout + row * d + h * kHd, go);
    const T* o = out + row * d + h * kHd;
#pragma unroll
    for (int e = 0; e < kHd; e += 4) {
      const float4 v = load4(o + e);
      delta += go[e] * v.x + go[e + 1] * v.y + go[e + 2] * v.z + go[e + 3] * v.w;
    }
    delta_ws[stat] = delta;
    L = lse[stat];
  }
#pragma unroll
  for (int e = 0; e < kHd; ++e) dq[e] = 0.f;

  for (int t0 = 0; t0 < n; t0 += kAttTile) {
    __syncthreads();
    load_tile(Ks, qkv + d + h * kHd, ld, sd, n, t0);
    load_tile(Vs, qkv + 2 * d + h * kHd, ld, sd, n, t0);
    if (threadIdx.x < kAttTile) {
      const int j = t0 + threadIdx.x;
      keep[threadIdx.x] = (j < n) && (key_mask == nullptr || key_mask[seq_row(sd, j)] != 0);
    }
    __syncthreads();
    if (!valid) continue;
    const int jn = min(kAttTile, n - t0);
    for (int j = 0; j < jn; ++j) {
      if (!keep[j]) continue;
      const float p = __expf(dot64(q, Ks[j]) * scale - L);
      const float ds = p * (dot64(go, Vs[j]) - delta);
      axpy64(dq, ds, Ks[j]);
    }
  }
  if (valid) store_row64(dqkv + row * ld + h * kHd, dq, scale);
}

// ------------------------------------------------------------------------------------------- backward: dk / dv
// MODE 0 writes dv, MODE 1 writes dk. One thread per key row; queries stream through shared memory.
template <typename T, int MODE>
__global__ void __launch_bounds__(kAttThreads) attn_bwd_dkv_simt_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                                        const int32_t* __restrict__ seq_desc,
                                                                        const uint8_t* __restrict__ key_mask, const float* __restrict__ lse,
                                                                        const float* __restrict__ delta_ws, T* __restrict__ dqkv, int H,
                                                                        int max_seq_len, float scale) {
  __shared__ __align__(16) float Qs[kAttTile][kHd];
  __shared__ __align__(16) float Gs[kAttTile][kHd];
  __shared__ float Ls[kAttTile], Ds[kAttTile];
  const int s = blockIdx.z, h = blockIdx.y;
  const SeqDesc sd = load_seq(seq_desc, s);
  const int n = sd.len0 + sd.len1;
  if (blockIdx.x * kAttThreads >= n) return;
  const int d = H * kHd;
  const long long ld = 3LL * d;
  const int j = blockIdx.x * kAttThreads + threadIdx.x;
  const bool valid = j < n;
  const long long row = valid ? seq_row(sd, j) : 0;
  const bool kept = valid && (key_mask == nullptr || key_mask[row] != 0);
  const long long stat0 = (static_cast<long long>(s) * H + h) * max_seq_len;

  float k[kHd], v[kHd], acc[kHd];  // v is dead (and eliminated) in MODE 0
  if (valid) {
    load_row64(qkv + row * ld + d + h * kHd, k);
    if (MODE == 1) load_row64(qkv + row * ld + 2 * d + h * kHd, v);
  }
#pragma unroll
  for (int e = 0; e < kHd; ++e) acc[e] = 0.f;

  for (int t0 = 0; t0 < n; t0 += kAttTile) {
    __syncthreads();
    load_tile(Qs, qkv + h * kHd, ld, sd, n, t0);
    load_tile(Gs, dout + h * kHd, static_cast<long long>(d), sd, n, t0);
    if (threadIdx.x < kAttTile) {
      const int i = t0 + threadIdx.x;
      Ls[threadIdx.x] = i < n ? lse[stat0 + i] : 0.f;
      Ds[threadIdx.x] = i < n ? delta_ws[stat0 + i] : 0.f;
    }
    __syncthreads();
    if (!kept) continue;
    const int in = min(kAttTile, n - t0);
    for (int i = 0; i < in; ++i) {
      const float p = __expf(dot64(k, Qs[i]) * scale - Ls[i]);
      if (MODE == 0) {
        axpy64(acc, p, Gs[i]);
      } else {
        const float dp = dot64(v, Gs[i]);
        axpy64(acc, p * (dp - Ds[i]), Qs[i]);
      }
    }
  }
  if (valid) store_row64(dqkv + row * ld + (MODE == 0 ? 2 * d : d) + h * kHd, acc, MODE == 0 ? 1.f : scale);
}

template <typename T>
static int attn_fwd_simt(const void* qkv, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse,
                         int num_seqs, int max_seq_len, int H, float scale, cudaStream_t stream) {
  dim3 grid((max_seq_len + kAttThreads - 1) / kAttThreads, H, num_seqs);
  attn_fwd_simt_kernel<T><<<grid, kAttThreads, 0, stream>>>(static_cast<const T*>(qkv), seq_desc, key_mask, static_cast<T*>(out), lse, H,
                                                          max_seq_len, scale);
  return check_launch("attn_fwd_simt");
}

template <typename T>
static int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const int32_t* seq_desc, const uint8_t* key_mask,
                         const float* lse, void* dqkv, float* delta_ws, int num_seqs, int max_seq_len, int H, float scale,
                         cudaStream_t stream) {
  dim3 grid((max_seq_len + kAttThreads - 1) / kAttThreads, H, num_seqs);
  const T* q = static_cast<const T*>(qkv);
  const T* go = static_cast<const T*>(dout);
  T* dq = static_cast<T*>(dqkv);
  attn_bwd_dq_simt_kernel<T><<<grid, kAttThreads, 0, stream>>>(q, static_cast<const T*>(out), go, seq_desc, key_mask, lse, dq, delta_ws, H,
                                                             max_seq_len, scale);
  int rc = check_launch("attn_bwd_dq_simt");
  if (rc != MOME_OK) return rc;
  attn_bwd_dkv_simt_kernel<T, 0><<<grid, kAttThreads, 0, stream>>>(q, go, seq_desc, key_mask, lse, delta_ws, dq, H, max_seq_len, scale);
  rc = check_launch("attn_bwd_dv_simt");
  if (rc != MOME_OK) return rc;
  attn_bwd_dkv_simt_kernel<T, 1><<<grid, kAttThreads, 0, stream>>>(q, go, seq_desc, key_mask, lse, delta_ws, dq, H, max_seq_len, scale);
  return check_launch("attn_bwd_dk_simt");
}

int attn_fwd_simt_dispatch(const void* qkv, int dtype, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse,
                           int num_seqs, int max_seq_len, int H, float scale, cudaStream_t stream) {
  return dtype == MOME_BF16 ? attn_fwd_simt<__nv_bfloat16>(qkv, seq_desc, key_mask, out, lse, num_seqs, max_seq_len, H, scale, stream)
                            : attn_fwd_simt<float>(qkv, seq_desc, key_mask, out, lse, num_seqs, max_seq_len, H, scale, stream);
}
int attn_bwd_simt_dispatch(const void* qkv, const void* out, const void* dout, int dtype, const int32_t* seq_desc,
                           const uint8_t* key_mask, const float* lse, void* dqkv, float* delta_ws, int num_seqs, int max_seq_len,
                           int H, float scale, cudaStream_t stream) {
  return dtype == MOME_BF16
             ? attn_bwd_simt<__nv_bfloat16>(qkv, out, dout, seq_desc, key_mask, lse, dqkv, delta_ws, num_seqs, max_seq_len, H, scale, stream)
             : attn_bwd_simt<float>(qkv, out, dout, seq_desc, key_mask, lse, dqkv, delta_ws, num_seqs, max_seq_len, H, scale, stream);
}

}  // namespace mome
