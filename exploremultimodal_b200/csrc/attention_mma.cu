// K1 (bf16 path): flash-style masked multi-head self-attention over packed segments on the tensor
// cores (mma.sync m16n8k16, bf16 in / fp32 accumulate), head_dim 64.
//
// A CTA of 4 warps owns a 64-row tile of one (sequence, head): 64 queries in the forward and dq
// kernels, 64 keys in the dk/dv kernel; every warp owns 16 of those rows, so each product in the
// kernel is a per-warp 16 x 64 x 64 GEMM whose A operand lives in registers and whose B operand is
// a 64 x 64 bf16 tile in XOR-swizzled shared memory (ldmatrix, bank-conflict free). The opposite
// side streams through a multi-stage cp.async ring. Softmax is online (running max / sum in
// the exp2 domain); probabilities never leave registers. Sequences are given as up to two row
// ranges of the packed token buffer ([text | image] after the fusion layer), keys are dropped by
// key_mask, query rows are never masked (reference vlmo.py:89-91).
//
// Backward = dq kernel (per query tile: S, dP, dQ) + dk/dv kernel (per key tile: S^T, dP^T, dV, dK),
// recomputing probabilities from the saved log-sum-exp; no atomics, deterministic.
//
// Replaces: reference vlmo.py:79-95 and its autograd backward. The tcgen05 variant is future work:
// at N <= 237 the whole (sequence, head) problem is 4 tiles and the kernel is latency bound.
#include "common.cuh"
#include "dropout.cuh"
#include "ptx.cuh"
#include "vec.cuh"

namespace mome {

namespace {

constexpr int kT = 64;                  // tile rows (queries or keys) and head_dim
constexpr int kThreads = 128;
constexpr int kTileBytes = kT * kT * 2;  // 8 KiB
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct Seq {
  int start0, len0, start1, len1;
};
__device__ __forceinline__ Seq load_seq(const int32_t* seq_desc, int s) {
  const int4 v = *reinterpret_cast<const int4*>(seq_desc + 4 * s);
  return Seq{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ long long seq_row(const Seq& sd, int i) {
  return i < sd.len0 ? static_cast<long long>(sd.start0) + i : static_cast<long long>(sd.start1) + (i - sd.len0);
}

// byte address of 16-byte chunk `chunk` of row `row` in a swizzled 64 x 64 bf16 tile
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}

// rows [t0, t0 + 64) of the sequence, 64 bf16 starting at `gbase` (column offset applied) -> swizzled tile
__device__ __forceinline__ void load_tile_async(uint8_t* tile, const __nv_bfloat16* gbase, long long ld, const Seq& sd, int n,
                                                int t0) {
#pragma unroll
  for (int c = threadIdx.x; c < kT * 8; c += kThreads) {
    const int r = c >> 3, ch = c & 7;
    const bool valid = t0 + r < n;
    const __nv_bfloat16* src = gbase + seq_row(sd, valid ? t0 + r : 0) * ld + ch * 8;
    cp_async_16(tile + r * 128 + ((ch ^ (r & 7)) << 4), src, valid);
  }
}

// Two tiles over the same 64 sequence rows in one pass (K and V of qkv; or Q of qkv and dO): a thread always
// copies 16-byte chunk (tid & 7) of rows (tid >> 3) + 16 i, so the row -> packed-buffer mapping is evaluated
// once per row and the swizzled destination offset is a per-thread constant.
__device__ __forceinline__ void load_tile_pair_async(uint8_t* tile_a, const __nv_bfloat16* ga, long long lda, uint8_t* tile_b,
                                                     const __nv_bfloat16* gb, long long ldb, const Seq& sd, int n, int t0) {
  const int r0 = threadIdx.x >> 3, ch = threadIdx.x & 7;
  const uint32_t dst = r0 * 128 + ((ch ^ (r0 & 7)) << 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + r0 + 16 * i;
    const bool valid = t < n;
    const long long row = seq_row(sd, valid ? t : 0);
    cp_async_16(tile_a + dst + i * 2048, ga + row * lda + ch * 8, valid);
    cp_async_16(tile_b + dst + i * 2048, gb + row * ldb + ch * 8, valid);
  }
}

// A fragments (4 k-steps) of the warp's 16 rows [r0, r0 + 16) of a swizzled tile
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], uint32_t tile, int r0) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(a[ks], tile_addr(tile, r0 + (lane & 15), ks * 2 + (lane >> 4)));
}

// acc[16 x 64] += A[16 x 64] * B, B(n, k) = tile[n][k] (TRANS = false) or tile[k][n] (TRANS = true)
template <bool TRANS>
__device__ __forceinline__ void warp_gemm(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile) {
  const int lane = threadIdx.x & 31;
  const int m = lane >> 3, l8 = lane & 7;
  if (!TRANS) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b[4];
        ldmatrix_x4(b, tile_addr(tile, p * 16 + (m >> 1) * 8 + l8, ks * 2 + (m & 1)));
        mma_bf16_16816(acc[2 * p], a[ks], b[0], b[1]);
        mma_bf16_16816(acc[2 * p + 1], a[ks], b[2], b[3]);
      }
    }
  } else {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, tile_addr(tile, ks * 16 + (m & 1) * 8 + l8, p * 2 + (m >> 1)));
        mma_bf16_16816(acc[2 * p], a[ks], b[0], b[1]);
        mma_bf16_16816(acc[2 * p + 1], a[ks], b[2], b[3]);
      }
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
}

// accumulator (16 x 64, C layout) -> bf16 A fragments for the next GEMM
__device__ __forceinline__ void acc_to_a(uint32_t (&a)[4][4], const float (&c)[8][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[j][0] = pack_bf16(c[2 * j][0], c[2 * j][1]);
    a[j][1] = pack_bf16(c[2 * j][2], c[2 * j][3]);
    a[j][2] = pack_bf16(c[2 * j + 1][0], c[2 * j + 1][1]);
    a[j][3] = pack_bf16(c[2 * j + 1][2], c[2 * j + 1][3]);
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Write the warp's 16 x 64 accumulator (times `mul` per row) as bf16 rows of the packed buffer:
// staged through the warp's own 16 rows of a swizzled tile so that global stores are 16-byte wide.
__device__ __forceinline__ void store_rows_bf16(const float (&acc)[8][4], float mul0, float mul1, uint8_t* tile, int r0,
                                                __nv_bfloat16* gbase, long long ld, const Seq& sd, int n, int t0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    // columns nt*8 + 2t, +1  -> chunk nt, byte offset 4t within the chunk
    *reinterpret_cast<uint32_t*>(tile + (r0 + g) * 128 + ((nt ^ ((r0 + g) & 7)) << 4) + 4 * t) = pack_bf16(acc[nt][0] * mul0, acc[nt][1] * mul0);
    *reinterpret_cast<uint32_t*>(tile + (r0 + g + 8) * 128 + ((nt ^ ((r0 + g + 8) & 7)) << 4) + 4 * t) = pack_bf16(acc[nt][2] * mul1, acc[nt][3] * mul1);
  }
  __syncwarp();
#pragma unroll
  for (int c = lane; c < 16 * 8; c += 32) {
    const int r = r0 + (c >> 3), ch = c & 7;
    if (t0 + r < n) {
      const uint4 v = *reinterpret_cast<const uint4*>(tile + r * 128 + ((ch ^ (r & 7)) << 4));
      *reinterpret_cast<uint4*>(gbase + seq_row(sd, t0 + r) * ld + ch * 8) = v;
    }
  }
}

// keep[0..63]: key j of the tile takes part; keep[64], keep[65]: every key of the first / second half does
// (lets the consumers skip the per-element selects on full, unmasked tiles). kKeepBytes per stage.
constexpr int kKeepBytes = kT + 16;
__device__ __forceinline__ void load_keep(uint8_t* keep, const uint8_t* key_mask, const Seq& sd, int n, int t0) {
  if (threadIdx.x < kT) {
    const int j = t0 + threadIdx.x;
    const bool k = (j < n) && (key_mask == nullptr || key_mask[seq_row(sd, j)] != 0);
    keep[threadIdx.x] = k;
    const bool all = __all_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0) keep[kT + (threadIdx.x >> 5)] = all;
  }
}
__device__ __forceinline__ bool tile_all_kept(const uint8_t* keep) { return keep[kT] != 0 && keep[kT + 1] != 0; }

// ---- dropout on the attention probabilities (reference vlmo.py:93). One 32-bit mask word per (sequence, head,
// query, pair of adjacent keys): low half-word -> even key, high half-word -> odd key (csrc/dropout.cuh).
struct AttnDrop {
  const uint32_t* seed;
  uint32_t salt, thr;
};
__device__ __forceinline__ uint32_t attn_drop_row(int s, int H, int h, int max_seq_len, int i) {
  return ((static_cast<uint32_t>(s) * H + h) * max_seq_len + i) * static_cast<uint32_t>((max_seq_len + 1) >> 1);
}
// a: even key, b: odd key of the pair the word belongs to
__device__ __forceinline__ void attn_drop_pair(uint32_t word, uint32_t thr, float scale, float& a, float& b) {
  a = (word & 255u) >= thr ? a * scale : 0.f;
  b = ((word >> 16) & 255u) >= thr ? b * scale : 0.f;
}

// ------------------------------------------------------------------------------------------- forward
// The streamed side lives in a ring of kFwdStages (K, V) tile pairs, one cp.async group per tile.
// Measured (profiles/): the kernel is bound by issue slots / occupancy, not by load latency, so the ring
// is kept shallow (2 stages, 41 KB) to fit 4 CTAs (16 warps) per SM, the limit the 128 registers allow.
constexpr int kFwdStages = 2;
// smem: Q | kFwdStages x (K, V) tiles, keep[kFwdStages][64]
constexpr int kFwdSmem = (1 + 2 * kFwdStages) * kTileBytes + kFwdStages * kKeepBytes;

template <bool DROP>
__global__ void __launch_bounds__(kThreads) attn_fwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ seq_desc,
                                                                const uint8_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out,
                                                                float* __restrict__ lse, int H, int max_seq_len, float scale,
                                                                const AttnDrop ad) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Qs = smem;
  uint8_t* ring = smem + kTileBytes;  // stage s: K at ring + 2 s kTileBytes, V right after
  uint8_t* keep = smem + (1 + 2 * kFwdStages) * kTileBytes;
  const int s = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kT;
  const Seq sd = load_seq(seq_desc, s);
  const int n = sd.len0 + sd.len1;
  if (q0 >= n) return;
  const int d = H * kT;
  const long long ld = 3LL * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = (n + kT - 1) / kT;

  auto load_stage = [&](int tile) {
    const int slot = tile % kFwdStages;
    load_tile_pair_async(ring + (2 * slot) * kTileBytes, qkv + d + h * kT, ld, ring + (2 * slot + 1) * kTileBytes,
                         qkv + 2 * d + h * kT, ld, sd, n, tile * kT);
    load_keep(keep + slot * kKeepBytes, key_mask, sd, n, tile * kT);
  };
  load_tile_async(Qs, qkv + h * kT, ld, sd, n, q0);
#pragma unroll
  for (int i = 0; i < kFwdStages; ++i) {
    if (i < ntiles) load_stage(i);
    cp_async_commit();
  }

  uint32_t qf[4][4];
  float o[8][4];
  zero_acc(o);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sl2 = scale * kLog2e;
  const uint32_t dkey = DROP ? drop_mix(ad.salt, __ldg(ad.seed)) : 0u;
  const float dscale = drop_scale(ad.thr);
  const uint32_t drow0 = attn_drop_row(s, H, h, max_seq_len, q0 + warp * 16 + g), drow1 = attn_drop_row(s, H, h, max_seq_len, q0 + warp * 16 + g + 8);

  for (int kt = 0; kt < ntiles; ++kt) {
    const int slot = kt % kFwdStages;
    cp_async_wait<kFwdStages - 1>();
    __syncthreads();
    if (kt == 0) load_a_frags(qf, smem_u32(Qs), warp * 16);

    float sacc[8][4];
    zero_acc(sacc);
    warp_gemm<false>(sacc, qf, smem_u32(ring + (2 * slot) * kTileBytes));
    // running max is tracked on the raw scores; exp2(s * sl2 - m * sl2) is one FFMA + EX2 per element
    const uint8_t* kp = keep + slot * kKeepBytes;
    float mx0 = -INFINITY, mx1 = -INFINITY;
    if (tile_all_kept(kp)) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const bool k0 = kp[nt * 8 + 2 * t] != 0, k1 = kp[nt * 8 + 2 * t + 1] != 0;
        sacc[nt][0] = k0 ? sacc[nt][0] : -INFINITY;
        sacc[nt][1] = k1 ? sacc[nt][1] : -INFINITY;
        sacc[nt][2] = k0 ? sacc[nt][2] : -INFINITY;
        sacc[nt][3] = k1 ? sacc[nt][3] : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
      }
    }
    mx0 = quad_max(mx0);
    mx1 = quad_max(mx1);
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float ms0 = mn0 == -INFINITY ? 0.f : mn0 * sl2, ms1 = mn1 == -INFINITY ? 0.f : mn1 * sl2;  // all keys masked so far
    const float c0 = exp2f(fmaf(m0, sl2, -ms0)), c1 = exp2f(fmaf(m1, sl2, -ms1));
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      sacc[nt][0] = exp2f(fmaf(sacc[nt][0], sl2, -ms0));
      sacc[nt][1] = exp2f(fmaf(sacc[nt][1], sl2, -ms0));
      sacc[nt][2] = exp2f(fmaf(sacc[nt][2], sl2, -ms1));
      sacc[nt][3] = exp2f(fmaf(sacc[nt][3], sl2, -ms1));
      rs0 += sacc[nt][0] + sacc[nt][1];
      rs1 += sacc[nt][2] + sacc[nt][3];
    }
    l0 = l0 * c0 + quad_sum(rs0);
    l1 = l1 * c1 + quad_sum(rs1);
    m0 = mn0;
    m1 = mn1;
    if (__any_sync(0xffffffffu, c0 != 1.f || c1 != 1.f)) {  // the running max rarely moves after the first tiles
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1;
      }
    }
    if (DROP) {  // the normaliser l keeps the undropped probabilities; only what multiplies V is dropped
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t jp = kt * 32 + nt * 4 + t;
        attn_drop_pair(drop_mix(drow0 + jp, dkey), ad.thr, dscale, sacc[nt][0], sacc[nt][1]);
        attn_drop_pair(drop_mix(drow1 + jp, dkey), ad.thr, dscale, sacc[nt][2], sacc[nt][3]);
      }
    }
    uint32_t pf[4][4];
    acc_to_a(pf, sacc);
    warp_gemm<true>(o, pf, smem_u32(ring + (2 * slot + 1) * kTileBytes));
    if (kt + kFwdStages < ntiles) {
      __syncthreads();  // everyone is done with this slot before it is refilled
      load_stage(kt + kFwdStages);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();

  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
  store_rows_bf16(o, inv0, inv1, Qs, warp * 16, out + h * kT, static_cast<long long>(d), sd, n, q0);
  if (t == 0) {
    const long long base = (static_cast<long long>(s) * H + h) * max_seq_len + q0 + warp * 16;
    if (q0 + warp * 16 + g < n) lse[base + g] = l0 > 0.f ? (m0 * sl2 + log2f(l0)) * kLn2 : -INFINITY;
    if (q0 + warp * 16 + g + 8 < n) lse[base + g + 8] = l1 > 0.f ? (m1 * sl2 + log2f(l1)) * kLn2 : -INFINITY;
  }
}

// ------------------------------------------------------------------------------------------- backward: dq (+ delta)
constexpr int kBwdStages = 3;
// smem: Q | dO | kBwdStages x (K, V) tiles, keep[kBwdStages][64], delta[64] floats
constexpr int kDqSmem = (2 + 2 * kBwdStages) * kTileBytes + kBwdStages * kKeepBytes + kT * 4;

template <bool DROP>
__global__ void __launch_bounds__(kThreads) attn_bwd_dq_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                                                                   const __nv_bfloat16* __restrict__ dout, const int32_t* __restrict__ seq_desc,
                                                                   const uint8_t* __restrict__ key_mask, const float* __restrict__ lse,
                                                                   __nv_bfloat16* __restrict__ dqkv, float* __restrict__ delta_ws, int H,
                                                                   int max_seq_len, float scale, const AttnDrop ad) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Qs = smem;
  uint8_t* Gs = smem + kTileBytes;
  uint8_t* ring = smem + 2 * kTileBytes;
  uint8_t* keep = smem + (2 + 2 * kBwdStages) * kTileBytes;
  float* delta_s = reinterpret_cast<float*>(keep + kBwdStages * kKeepBytes);
  const int s = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kT;
  const Seq sd = load_seq(seq_desc, s);
  const int n = sd.len0 + sd.len1;
  if (q0 >= n) return;
  const int d = H * kT;
  const long long ld = 3LL * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = (n + kT - 1) / kT;
  const long long stat0 = (static_cast<long long>(s) * H + h) * max_seq_len;

  auto load_stage = [&](int tile) {
    const int slot = tile % kBwdStages;
    load_tile_pair_async(ring + (2 * slot) * kTileBytes, qkv + d + h * kT, ld, ring + (2 * slot + 1) * kTileBytes,
                         qkv + 2 * d + h * kT, ld, sd, n, tile * kT);
    load_keep(keep + slot * kKeepBytes, key_mask, sd, n, tile * kT);
  };
  load_tile_pair_async(Qs, qkv + h * kT, ld, Gs, dout + h * kT, static_cast<long long>(d), sd, n, q0);
#pragma unroll
  for (int i = 0; i < kBwdStages; ++i) {
    if (i < ntiles) load_stage(i);
    cp_async_commit();
  }
  // delta = rowsum(dO * O): two threads per row, 32 columns each, straight from global memory
  {
    const int r = threadIdx.x >> 1, half = threadIdx.x & 1;
    float acc = 0.f;
    if (q0 + r < n) {
      const long long row = seq_row(sd, q0 + r);
      const uint4* po = reinterpret_cast<const uint4*>(out + row * d + h * kT + half * 32);
      const uint4* pg = reinterpret_cast<const uint4*>(dout + row * d + h * kT + half * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 a = __ldg(pg + c), b = __ldg(po + c);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[j]));
          const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[j]));
          acc += fa.x * fb.x + fa.y * fb.y;
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0) {
      delta_s[r] = acc;
      if (q0 + r < n) delta_ws[stat0 + q0 + r] = acc;
    }
  }

  uint32_t qf[4][4], gf[4][4];
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  // rows past the sequence end get lse = +inf => p = 0
  const float L0 = r0 < n ? lse[stat0 + r0] * kLog2e : INFINITY, L1 = r1 < n ? lse[stat0 + r1] * kLog2e : INFINITY;
  float D0 = 0.f, D1 = 0.f;
  const float sl2 = scale * kLog2e;
  float dq[8][4];
  zero_acc(dq);
  const uint32_t dkey = DROP ? drop_mix(ad.salt, __ldg(ad.seed)) : 0u;
  const float dscale = drop_scale(ad.thr);
  const uint32_t drow0 = attn_drop_row(s, H, h, max_seq_len, r0), drow1 = attn_drop_row(s, H, h, max_seq_len, r1);

  for (int kt = 0; kt < ntiles; ++kt) {
    const int slot = kt % kBwdStages;
    cp_async_wait<kBwdStages - 1>();
    __syncthreads();
    if (kt == 0) {
      load_a_frags(qf, smem_u32(Qs), warp * 16);
      load_a_frags(gf, smem_u32(Gs), warp * 16);
      D0 = delta_s[warp * 16 + g];
      D1 = delta_s[warp * 16 + g + 8];
    }
    float sacc[8][4], dp[8][4];
    zero_acc(sacc);
    zero_acc(dp);
    warp_gemm<false>(sacc, qf, smem_u32(ring + (2 * slot) * kTileBytes));
    warp_gemm<false>(dp, gf, smem_u32(ring + (2 * slot + 1) * kTileBytes));
    if (DROP) {  // dP of a dropped probability is zero, of a kept one it carries the 1 / (1 - p) scale
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t jp = kt * 32 + nt * 4 + t;
        attn_drop_pair(drop_mix(drow0 + jp, dkey), ad.thr, dscale, dp[nt][0], dp[nt][1]);
        attn_drop_pair(drop_mix(drow1 + jp, dkey), ad.thr, dscale, dp[nt][2], dp[nt][3]);
      }
    }
    const uint8_t* kp = keep + slot * kKeepBytes;
    if (tile_all_kept(kp)) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sacc[nt][0] = exp2f(fmaf(sacc[nt][0], sl2, -L0)) * (dp[nt][0] - D0);
        sacc[nt][1] = exp2f(fmaf(sacc[nt][1], sl2, -L0)) * (dp[nt][1] - D0);
        sacc[nt][2] = exp2f(fmaf(sacc[nt][2], sl2, -L1)) * (dp[nt][2] - D1);
        sacc[nt][3] = exp2f(fmaf(sacc[nt][3], sl2, -L1)) * (dp[nt][3] - D1);
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const bool k0 = kp[nt * 8 + 2 * t] != 0, k1 = kp[nt * 8 + 2 * t + 1] != 0;
        const float p00 = k0 ? exp2f(fmaf(sacc[nt][0], sl2, -L0)) : 0.f, p01 = k1 ? exp2f(fmaf(sacc[nt][1], sl2, -L0)) : 0.f;
        const float p10 = k0 ? exp2f(fmaf(sacc[nt][2], sl2, -L1)) : 0.f, p11 = k1 ? exp2f(fmaf(sacc[nt][3], sl2, -L1)) : 0.f;
        sacc[nt][0] = p00 * (dp[nt][0] - D0);
        sacc[nt][1] = p01 * (dp[nt][1] - D0);
        sacc[nt][2] = p10 * (dp[nt][2] - D1);
        sacc[nt][3] = p11 * (dp[nt][3] - D1);
      }
    }
    uint32_t dsf[4][4];
    acc_to_a(dsf, sacc);
    warp_gemm<true>(dq, dsf, smem_u32(ring + (2 * slot) * kTileBytes));
    if (kt + kBwdStages < ntiles) {
      __syncthreads();
      load_stage(kt + kBwdStages);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  store_rows_bf16(dq, scale, scale, Qs, warp * 16, dqkv + h * kT, ld, sd, n, q0);
}

// ------------------------------------------------------------------------------------------- backward: dk, dv
// smem: K | V | kBwdStages x (Q, dO) tiles, lse[kBwdStages][64], delta[kBwdStages][64] floats
constexpr int kDkvSmem = (2 + 2 * kBwdStages) * kTileBytes + 2 * kBwdStages * kT * 4;

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 3) attn_bwd_dkv_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                                                                    const int32_t* __restrict__ seq_desc, const uint8_t* __restrict__ key_mask,
                                                                    const float* __restrict__ lse, const float* __restrict__ delta_ws,
                                                                    __nv_bfloat16* __restrict__ dqkv, int H, int max_seq_len, float scale,
                                                                    const AttnDrop ad) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Ks = smem;
  uint8_t* Vs = smem + kTileBytes;
  uint8_t* ring = smem + 2 * kTileBytes;  // stage s: Q at ring + 2 s kTileBytes, dO right after
  float* lse_s = reinterpret_cast<float*>(smem + (2 + 2 * kBwdStages) * kTileBytes);
  float* delta_s = lse_s + kBwdStages * kT;
  const int s = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * kT;
  const Seq sd = load_seq(seq_desc, s);
  const int n = sd.len0 + sd.len1;
  if (k0 >= n) return;
  const int d = H * kT;
  const long long ld = 3LL * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = (n + kT - 1) / kT;
  const long long stat0 = (static_cast<long long>(s) * H + h) * max_seq_len;

  auto load_stage = [&](int tile) {
    const int slot = tile % kBwdStages;
    load_tile_pair_async(ring + (2 * slot) * kTileBytes, qkv + h * kT, ld, ring + (2 * slot + 1) * kTileBytes, dout + h * kT,
                         static_cast<long long>(d), sd, n, tile * kT);
    if (threadIdx.x < kT) {
      const int i = tile * kT + threadIdx.x;
      lse_s[slot * kT + threadIdx.x] = i < n ? lse[stat0 + i] * kLog2e : INFINITY;  // +inf => p = 0 for absent queries
      delta_s[slot * kT + threadIdx.x] = i < n ? delta_ws[stat0 + i] : 0.f;
    }
  };
  load_tile_pair_async(Ks, qkv + d + h * kT, ld, Vs, qkv + 2 * d + h * kT, ld, sd, n, k0);
#pragma unroll
  for (int i = 0; i < kBwdStages; ++i) {
    if (i < ntiles) load_stage(i);
    cp_async_commit();
  }

  // this thread's two key rows
  const int j0 = k0 + warp * 16 + g, j1 = j0 + 8;
  const bool keep0 = j0 < n && (key_mask == nullptr || key_mask[seq_row(sd, j0)] != 0);
  const bool keep1 = j1 < n && (key_mask == nullptr || key_mask[seq_row(sd, j1)] != 0);
  const float sl2 = scale * kLog2e;
  uint32_t kf[4][4], vf[4][4];
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  const uint32_t dkey = DROP ? drop_mix(ad.salt, __ldg(ad.seed)) : 0u;
  const float dscale = drop_scale(ad.thr);
  const uint32_t dsh0 = (j0 & 1) * 16, dsh1 = (j1 & 1) * 16;   // which half-word of the mask word belongs to this thread's keys
  const uint32_t kp2 = static_cast<uint32_t>((max_seq_len + 1) >> 1);

  for (int qt = 0; qt < ntiles; ++qt) {
    const int slot = qt % kBwdStages;
    cp_async_wait<kBwdStages - 1>();
    __syncthreads();
    if (qt == 0) {
      load_a_frags(kf, smem_u32(Ks), warp * 16);
      load_a_frags(vf, smem_u32(Vs), warp * 16);
    }
    const uint32_t q_tile = smem_u32(ring + (2 * slot) * kTileBytes), g_tile = smem_u32(ring + (2 * slot + 1) * kTileBytes);
    // S^T (rows = this warp's keys, columns = the tile's queries) -> P^T, packed to bf16 at once: it is both
    // the A operand of dV += P^T dO and, unpacked again, the factor of dS^T (saves 16 live registers)
    const float* Lq = lse_s + slot * kT;
    const float* Dq = delta_s + slot * kT;
    uint32_t pf[4][4];
    {
      float st[8][4];
      zero_acc(st);
      warp_gemm<false>(st, kf, q_tile);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * t;
        const float La = Lq[c], Lb = Lq[c + 1];
        st[nt][0] = keep0 ? exp2f(fmaf(st[nt][0], sl2, -La)) : 0.f;
        st[nt][1] = keep0 ? exp2f(fmaf(st[nt][1], sl2, -Lb)) : 0.f;
        st[nt][2] = keep1 ? exp2f(fmaf(st[nt][2], sl2, -La)) : 0.f;
        st[nt][3] = keep1 ? exp2f(fmaf(st[nt][3], sl2, -Lb)) : 0.f;
      }
      acc_to_a(pf, st);
    }
    if (DROP) {
      // P^T stays undropped in pf (it is the factor of dS^T); a dropped element is flagged in the sign bit of its
      // bf16 value (probabilities are >= 0), and a dropped + scaled copy feeds dV += drop(P)^T dO.
      uint32_t pd[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int nt = 2 * j + h2;
          const uint32_t qa = qt * kT + nt * 8 + 2 * t;  // this thread's two queries: qa, qa + 1
          const uint32_t ra = attn_drop_row(s, H, h, max_seq_len, qa), rb = ra + kp2;
#pragma unroll
          for (int r = 0; r < 2; ++r) {  // r = 0: key row j0 (pf[j][2 h2]), r = 1: key row j1 (pf[j][2 h2 + 1])
            const uint32_t jj = r == 0 ? j0 : j1, sh = r == 0 ? dsh0 : dsh1;
            const bool ka = ((drop_mix(ra + (jj >> 1), dkey) >> sh) & 255u) >= ad.thr;
            const bool kb = ((drop_mix(rb + (jj >> 1), dkey) >> sh) & 255u) >= ad.thr;
            uint32_t& w = pf[j][2 * h2 + r];
            const float2 pv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
            pd[j][2 * h2 + r] = pack_bf16(ka ? pv.x * dscale : 0.f, kb ? pv.y * dscale : 0.f);
            w |= (ka ? 0u : 0x8000u) | (kb ? 0u : 0x80000000u);
          }
        }
      }
      warp_gemm<true>(dv, pd, g_tile);
    } else {
      warp_gemm<true>(dv, pf, g_tile);
    }
    {
      float dpt[8][4];
      zero_acc(dpt);
      warp_gemm<false>(dpt, vf, g_tile);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // pf[j] = {(nt 2j: row g), (nt 2j: row g+8), (nt 2j+1: row g), (nt 2j+1: row g+8)}, each a bf16 pair of columns
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int nt = 2 * j + h2, c = nt * 8 + 2 * t;
          const float Da = Dq[c], Db = Dq[c + 1];
          const uint32_t w0 = pf[j][2 * h2], w1 = pf[j][2 * h2 + 1];
          const uint32_t a0 = DROP ? (w0 & 0x7fff7fffu) : w0, a1 = DROP ? (w1 & 0x7fff7fffu) : w1;
          const float2 p0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a0));
          const float2 p1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a1));
          if (DROP) {  // dP^T of dropped probabilities is zero, kept ones carry the scale
            dpt[nt][0] = (w0 & 0x8000u) ? 0.f : dpt[nt][0] * dscale;
            dpt[nt][1] = (w0 & 0x80000000u) ? 0.f : dpt[nt][1] * dscale;
            dpt[nt][2] = (w1 & 0x8000u) ? 0.f : dpt[nt][2] * dscale;
            dpt[nt][3] = (w1 & 0x80000000u) ? 0.f : dpt[nt][3] * dscale;
          }
          dpt[nt][0] = p0.x * (dpt[nt][0] - Da);
          dpt[nt][1] = p0.y * (dpt[nt][1] - Db);
          dpt[nt][2] = p1.x * (dpt[nt][2] - Da);
          dpt[nt][3] = p1.y * (dpt[nt][3] - Db);
        }
      }
      acc_to_a(pf, dpt);
    }
    warp_gemm<true>(dk, pf, q_tile);
    if (qt + kBwdStages < ntiles) {
      __syncthreads();
      load_stage(qt + kBwdStages);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  store_rows_bf16(dk, scale, scale, Ks, warp * 16, dqkv + d + h * kT, ld, sd, n, k0);
  store_rows_bf16(dv, 1.f, 1.f, Vs, warp * 16, dqkv + 2 * d + h * kT, ld, sd, n, k0);
}

template <typename K>
int opt_in(K kern, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%d): %s", what, bytes, cudaGetErrorString(e));
    return MOME_ERR_CUDA;
  }
  return MOME_OK;
}

}  // namespace

int attn_fwd_mma(const void* qkv, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse, int num_seqs,
                 int max_seq_len, int H, float scale, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    int rc = opt_in(attn_fwd_mma_kernel<false>, kFwdSmem, "attn_fwd_mma");
    if (rc == MOME_OK) rc = opt_in(attn_fwd_mma_kernel<true>, kFwdSmem, "attn_fwd_mma");
    if (rc != MOME_OK) return rc;
    configured = true;
  }
  dim3 grid((max_seq_len + kT - 1) / kT, H, num_seqs);
  const AttnDrop ad{drop_seed, drop_salt, drop_threshold(drop_p)};
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  if (drop_seed != nullptr && drop_p > 0.f)
    attn_fwd_mma_kernel<true><<<grid, kThreads, kFwdSmem, stream>>>(q, seq_desc, key_mask, static_cast<__nv_bfloat16*>(out), lse, H, max_seq_len, scale, ad);
  else
    attn_fwd_mma_kernel<false><<<grid, kThreads, kFwdSmem, stream>>>(q, seq_desc, key_mask, static_cast<__nv_bfloat16*>(out), lse, H, max_seq_len, scale, ad);
  return check_launch("attn_fwd_mma");
}

int attn_bwd_mma(const void* qkv, const void* out, const void* dout, const int32_t* seq_desc, const uint8_t* key_mask,
                 const float* lse, void* dqkv, float* delta_ws, int num_seqs, int max_seq_len, int H, float scale,
                 const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    int rc = opt_in(attn_bwd_dq_mma_kernel<false>, kDqSmem, "attn_bwd_dq_mma");
    if (rc == MOME_OK) rc = opt_in(attn_bwd_dq_mma_kernel<true>, kDqSmem, "attn_bwd_dq_mma");
    if (rc == MOME_OK) rc = opt_in(attn_bwd_dkv_mma_kernel<false>, kDkvSmem, "attn_bwd_dkv_mma");
    if (rc == MOME_OK) rc = opt_in(attn_bwd_dkv_mma_kernel<true>, kDkvSmem, "attn_bwd_dkv_mma");
    if (rc != MOME_OK) return rc;
    configured = true;
  }
  dim3 grid((max_seq_len + kT - 1) / kT, H, num_seqs);
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* o = static_cast<const __nv_bfloat16*>(out);
  const __nv_bfloat16* go = static_cast<const __nv_bfloat16*>(dout);
  __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(dqkv);
  const AttnDrop ad{drop_seed, drop_salt, drop_threshold(drop_p)};
  const bool drop = drop_seed != nullptr && drop_p > 0.f;
  if (drop)
    attn_bwd_dq_mma_kernel<true><<<grid, kThreads, kDqSmem, stream>>>(q, o, go, seq_desc, key_mask, lse, dq, delta_ws, H, max_seq_len, scale, ad);
  else
    attn_bwd_dq_mma_kernel<false><<<grid, kThreads, kDqSmem, stream>>>(q, o, go, seq_desc, key_mask, lse, dq, delta_ws, H, max_seq_len, scale, ad);
  int rc = check_launch("attn_bwd_dq_mma");
  if (rc != MOME_OK) return rc;
  if (drop)
    attn_bwd_dkv_mma_kernel<true><<<grid, kThreads, kDkvSmem, stream>>>(q, go, seq_desc, key_mask, lse, delta_ws, dq, H, max_seq_len, scale, ad);
  else
    attn_bwd_dkv_mma_kernel<false><<<grid, kThreads, kDkvSmem, stream>>>(q, go, seq_desc, key_mask, lse, delta_ws, dq, H, max_seq_len, scale, ad);
  return check_launch("attn_bwd_dkv_mma");
}

}  // namespace mome
