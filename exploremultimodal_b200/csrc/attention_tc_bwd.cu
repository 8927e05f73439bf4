// K1 backward on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), for layouts whose sequences fit 256 keys.
//
// Used for 64 < max_seq_len <= 256 (attention_api.cu; MOME_ATTN_TC_BWD=0 falls back to the mma.sync pair in
// attention_mma.cu, which also serves the other lengths). Same contract: mask semantics, lse format, dropout mask
// function, dqkv layout. Measured (tools/attn_bench.py --tc-bwd, 64 x [40 | 197] tokens, 12 heads): 132 us against
// 220 us for the two mma.sync kernels (150 against 269 us with dropout).
//
// One persistent CTA per SM walks over (sequence, head) items. An item is cut into 128 x 128 blocks (query tile i,
// key block j; at most 2 x 2), visited key block by key block so that dK_j / dV_j finish after two blocks while the
// two dQ_i accumulate over the whole item. Per block:
//
//   tcgen05    S  = Q_i K_j^T, dP = dO_i V_j^T                         (SS, K-major)        -> TMEM S, dP (128 columns each)
//   8 / 16 warps  thread = (query row, half / quarter of the block's keys): P = exp2(S scale log2e - lse), dropout,
//              dS = P o (dP - delta); bf16 P and dS -> two SWIZZLE_128B smem tiles [128 q][128 keys]
//   tcgen05    dV_j += P^T dO_i, dK_j += dS^T Q_i   (A = the smem tile read MN-major: the transposition is free)
//              dQ_i += dS K_j                        (A = the same dS tile read K-major)      -> TMEM accumulators
//
// TMEM (512 columns): dQ_0 | dQ_1 | dK_j | dV_j (64 each) | S | dP (128 each). Shared memory: Q, K, V, dO of the item
// (4 x 32 KB, TMA boxes out of the packed qkv / dout buffers), the P and dS tiles (2 x 32 KB). delta = rowsum(dO o O)
// comes from a small row kernel in front (attn_delta_kernel).
//
//   warp 0      producer (TMA; cp.async gather for layouts whose second range is not 8-row aligned)
//   warp 1      one elected thread issues every tcgen05.mma
//   warps 2-9 (2-17)  P / dS group (8 or 16 warps), also drains the accumulators (x scale) into dqkv
//
// Replaces: the autograd backward of reference vlmo.py:79-95.
#include <cuda.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "dropout.cuh"
#include "ptx.cuh"

namespace mome {

int tma_encode_bf16_2d(void* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer);

namespace {

constexpr int kHd = 64;
constexpr int kTile = 128;      // query tile = key block = MMA M = TMEM lanes
constexpr int kMaxKeys = 256;
constexpr float kLog2e = 1.4426950408889634f;

// shared memory map (offsets from a 1024-byte aligned base)
constexpr int kQOff = 0, kKOff = 32768, kVOff = 65536, kGOff = 98304;  // Q, K, V, dO: 256 rows x 128 B each
constexpr int kPOff = 131072, kSOff = 163840;                            // P, dS tiles: 2 chunks of 64 keys x [128 q x 128 B]
constexpr int kMetaOff = 196608;                                          // keep words [8], seq desc [4]
constexpr int kBarOff = kMetaOff + 64;
constexpr int kBwdSmem = kBarOff + 128 + 1024;
constexpr int kBwdThreads8 = 64 + 256, kBwdThreads16 = 64 + 512;  // producer + MMA warps + 8 or 16 P / dS warps
constexpr uint32_t kColDQ = 0, kColDK = 128, kColDV = 192, kColS = 256, kColDP = 384;

struct Seq {
  int start0, len0, start1, len1;
};
__device__ __forceinline__ Seq load_seq(const int32_t* seq_desc, int s) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(seq_desc + 4 * s));
  return Seq{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ long long seq_row(const Seq& sd, int i) {
  return i < sd.len0 ? static_cast<long long>(sd.start0) + i : static_cast<long long>(sd.start1) + (i - sd.len0);
}
__device__ __forceinline__ uint32_t attn_drop_row(int s, int H, int h, int max_seq_len, int i) {
  return ((static_cast<uint32_t>(s) * H + h) * max_seq_len + i) * static_cast<uint32_t>((max_seq_len + 1) >> 1);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct BwdParams {
  CUtensorMap qkv32, qkv8;  // qkv as [tokens][3 d] bf16, boxes of 64 columns x 32 / 8 rows
  CUtensorMap go32, go8;    // dout as [tokens][d]
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* dout;
  const int32_t* seq_desc;
  const uint8_t* key_mask;
  const float* lse;
  const float* delta;
  __nv_bfloat16* dqkv;
  int H, max_seq_len, num_items;
  float scale;
  const uint32_t* drop_seed;
  uint32_t drop_salt, drop_thr;
};

__device__ __forceinline__ int range_boxes(int len) { return (len >> 5) + (((len & 31) + 7) >> 3); }
__device__ __forceinline__ uint32_t range_bytes(int len) { return (len >> 5) * 4096u + (((len & 31) + 7) >> 3) * 1024u; }
__device__ __forceinline__ void issue_range_box(const CUtensorMap* m32, const CUtensorMap* m8, uint8_t* dst_rows, uint64_t* bar, int col,
                                                int grow, int len, int b) {
  const int nb = len >> 5;
  if (b < nb) {
    tma_load_2d(dst_rows + b * 4096, m32, bar, col, grow + 32 * b);
  } else {
    const int r = 32 * nb + 8 * (b - nb);
    tma_load_2d(dst_rows + r * 128, m8, bar, col, grow + r);
  }
}

// delta[(s H + h) max_seq_len + q] = sum_c dO[row(q)][h 64 + c] O[row(q)][h 64 + c]: one warp per (sequence, query), lanes over heads x 2
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                                         const int32_t* __restrict__ seq_desc, float* __restrict__ delta, int H,
                                                         int max_seq_len, int num_seqs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = static_cast<long long>(blockIdx.x) * 8 + warp;
  const int s = static_cast<int>(w / max_seq_len), q = static_cast<int>(w - static_cast<long long>(s) * max_seq_len);
  if (s >= num_seqs) return;
  const Seq sd = load_seq(seq_desc, s);
  if (q >= sd.len0 + sd.len1) return;
  const long long row = seq_row(sd, q);
  const int d = H * kHd;
  // a lane handles 32 consecutive columns (half a head) at a time; the two halves of a head sit in adjacent lanes
  for (int base = 0; base < d; base += 32 * 32) {
    const int c0 = base + lane * 32;
    const bool active = c0 < d;
    float acc = 0.f;
    if (active) {
      const uint4* po = reinterpret_cast<const uint4*>(out + row * d + c0);
      const uint4* pg = reinterpret_cast<const uint4*>(dout + row * d + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 a = __ldg(pg + i), b = __ldg(po + i);
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
    if (active && (lane & 1) == 0) delta[(static_cast<long long>(s) * H + (c0 >> 6)) * max_seq_len + q] = acc;
  }
}

// GW = warps of the P / dS group: 8 (a thread owns a query row x 64 keys of the block) or 16 (x 32 keys). The P / dS
// phase is a serial per-thread chain of exponentials between two tensor-core phases; with 16 warps it is half as long
// and each scheduler has four warps to interleave — yet it measured no faster (404 vs 405 us at 256 x [40 | 197] without
// dropout, 509 vs 478 us with): the block's chain is bound by the tensor-core / barrier latencies between its phases, not by
// the exponentials. 8 warps stay the default; MOME_ATTN_TC_BWD=3 selects 16.
// EARLY_S (experimental, MOME_ATTN_TC_BWD=2, 8 warps): S / dP of the next block of the item are issued before the
// accumulate MMAs of the current one (measured: no gain, 412 vs 405 us).
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x16(taddr, r); }

template <bool DROP, bool EARLY_S, int GW>
__global__ void __launch_bounds__(64 + 32 * GW, 1) attn_bwd_tc_kernel(const __grid_constant__ BwdParams p) {
  constexpr int kBwdThreads = 64 + 32 * GW;
  constexpr int kParts = GW / 4;        // key parts of a block = column parts of the 64-wide accumulators
  constexpr int KP = kTile / kParts;    // keys of the block per thread
  constexpr int DC = kHd / kParts;      // accumulator columns per thread in the drains
  constexpr int kGroupThreads = 32 * GW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* full = bars;          // producer -> everyone: the item's Q, K, V, dO and meta are in shared memory (2 arrivals + tx)
  uint64_t* empty = bars + 1;     // MMA -> producer: every MMA of the item has completed
  uint64_t* sdp_full = bars + 2;  // MMA -> group: S and dP of the block are in TMEM
  uint64_t* pds_full = bars + 3;  // group -> MMA: P and dS tiles written, S / dP consumed (256 arrivals)
  uint64_t* pds_free = bars + 4;  // MMA -> group: the MMAs reading the P / dS tiles have completed
  uint64_t* dkv_full = bars + 5;  // MMA -> group: dK_j, dV_j final
  uint64_t* dkv_free = bars + 6;  // group -> MMA: dK_j, dV_j drained (256 arrivals)
  uint64_t* dq_full = bars + 7;   // MMA -> group: dQ_0, dQ_1 final
  uint64_t* dq_free = bars + 8;   // group -> MMA: dQ drained (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, d = H * kHd;
  const int G = gridDim.x;

  // stale shared memory must at least be finite: rows past the sequence end multiply exact zeros
  for (int i = threadIdx.x; i < kMetaOff / 16; i += kBwdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(full, 2);
    mbar_init(empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, kGroupThreads);
    mbar_init(pds_free, 1);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_free, kGroupThreads);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, kGroupThreads);
    fence_barrier_init();
    tma_prefetch_desc(&p.qkv32);
    tma_prefetch_desc(&p.qkv8);
    tma_prefetch_desc(&p.go32);
    tma_prefetch_desc(&p.go8);
  }
  __syncwarp();
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------ producer
    int k = 0;
    Seq sd_next = blockIdx.x < p.num_items ? load_seq(p.seq_desc, blockIdx.x / H) : Seq{0, 0, 0, 0};
    for (int item = blockIdx.x; item < p.num_items; item += G, ++k) {
      const int s = item / H, h = item - s * H;
      const Seq sd = sd_next;
      if (item + G < p.num_items) sd_next = load_seq(p.seq_desc, (item + G) / H);
      const int n = sd.len0 + sd.len1;
      uint32_t* meta = reinterpret_cast<uint32_t*>(smem + kMetaOff);
      uint8_t mk[8];
#pragma unroll
      for (int j8 = 0; j8 < 8; ++j8) {
        const int j = j8 * 32 + lane;
        mk[j8] = j < n ? (p.key_mask == nullptr ? uint8_t(1) : __ldg(p.key_mask + seq_row(sd, j))) : uint8_t(0);
      }
      mbar_wait_park(empty, (k & 1) ^ 1);
      __syncwarp();
      const bool boxes_ok = sd.len1 == 0 || (sd.len0 & 7) == 0;
      if (boxes_ok) {
        const int nb0 = range_boxes(sd.len0), nb1 = range_boxes(sd.len1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full, 4u * (range_bytes(sd.len0) + range_bytes(sd.len1)));
#pragma unroll 1
          for (int op = 0; op < 4; ++op) {  // Q, K, V out of qkv; dO out of dout
            const CUtensorMap* m32 = op < 3 ? &p.qkv32 : &p.go32;
            const CUtensorMap* m8 = op < 3 ? &p.qkv8 : &p.go8;
            const int col = (op < 3 ? op * d : 0) + h * kHd;
            uint8_t* dst = smem + op * 32768;
            for (int bb = 0; bb < nb0; ++bb) issue_range_box(m32, m8, dst, full, col, sd.start0, sd.len0, bb);
            for (int bb = 0; bb < nb1; ++bb) issue_range_box(m32, m8, dst + sd.len0 * 128, full, col, sd.start1, sd.len1, bb);
          }
        }
        __syncwarp();
      }
#pragma unroll
      for (int j8 = 0; j8 < 8; ++j8) {
        const uint32_t w = __ballot_sync(0xffffffffu, mk[j8] != 0);
        if (lane == 0) meta[j8] = w;
      }
      if (lane == 0) {
        meta[8] = sd.start0; meta[9] = sd.len0; meta[10] = sd.start1; meta[11] = sd.len1;
        mbar_arrive(full);
      }
      if (!boxes_ok) {
        const int n_pad = (n + 15) & ~15;
        const int r0 = lane >> 3, ch = lane & 7;
        const long long ld = 3LL * d;
        for (int r = r0; r < n_pad; r += 4) {
          const bool valid = r < n;
          const long long grow = seq_row(sd, valid ? r : 0);
          const __nv_bfloat16* src = p.qkv + grow * ld + h * kHd + ch * 8;
          const uint32_t off = r * 128 + ((ch ^ (r & 7)) << 4);
          cp_async_16(smem + kQOff + off, src, valid);
          cp_async_16(smem + kKOff + off, src + d, valid);
          cp_async_16(smem + kVOff + off, src + 2 * d, valid);
          cp_async_16(smem + kGOff + off, p.dout + grow * d + h * kHd + ch * 8, valid);
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t sQ = smem_u32(smem + kQOff), sK = smem_u32(smem + kKOff), sV = smem_u32(smem + kVOff), sG = smem_u32(smem + kGOff);
      const uint32_t sP = smem_u32(smem + kPOff), sS = smem_u32(smem + kSOff);
      uint32_t blk = 0, jcount = 0;  // running counts of blocks / finished key blocks (barrier phases)
      int k = 0;
      for (int item = blockIdx.x; item < p.num_items; item += G, ++k) {
        const Seq sd = load_seq(p.seq_desc, item / H);
        const int n = sd.len0 + sd.len1;
        const int nt = n > kTile ? 2 : 1;  // query tiles = key blocks
        mbar_wait_park(full, k & 1);
        tcgen05_fence_after();
        for (int j = 0; j < nt; ++j) {
          const int kk = min(kTile, ((n - j * kTile) + 15) & ~15);  // keys of the block, padded to the MMA granularity
          for (int i = 0; i < nt; ++i, ++blk) {
            const int kq = min(kTile, ((n - i * kTile) + 15) & ~15);  // query rows of the tile
            // ---- S = Q_i K_j^T, dP = dO_i V_j^T (the group has consumed the previous S / dP: pds_full was waited for)
            auto issue_sdp = [&](int jj, int ii) {
              const int kkk = min(kTile, ((n - jj * kTile) + 15) & ~15);
              const uint32_t idesc = umma_idesc_bf16(kTile, kkk, false, false);
#pragma unroll
              for (int c = 0; c < kHd / 16; ++c)
                umma_bf16(tmem_base + kColS, umma_smem_desc(sQ + ii * 16384 + c * 32, 0, 1024), umma_smem_desc(sK + jj * 16384 + c * 32, 0, 1024),
                          idesc, c > 0 ? 1u : 0u);
#pragma unroll
              for (int c = 0; c < kHd / 16; ++c)
                umma_bf16(tmem_base + kColDP, umma_smem_desc(sG + ii * 16384 + c * 32, 0, 1024), umma_smem_desc(sV + jj * 16384 + c * 32, 0, 1024),
                          idesc, c > 0 ? 1u : 0u);
              umma_commit(sdp_full);
            };
            if (!EARLY_S || (j == 0 && i == 0)) issue_sdp(j, i);
            mbar_wait_park(pds_full, blk & 1);
            if (i == 0 && (j > 0 || k > 0)) mbar_wait_park(dkv_free, (jcount & 1) ^ 1);  // previous dK / dV drained
            if (i == 0 && j == 0 && k > 0) mbar_wait_park(dq_free, (k & 1) ^ 1);          // previous item's dQ drained
            tcgen05_fence_after();
            if (EARLY_S) {  // next block of this item, in visiting order (i fastest)
              const int ni = i + 1 < nt ? i + 1 : 0, nj = i + 1 < nt ? j : j + 1;
              if (nj < nt) issue_sdp(nj, ni);
            }
            {
              // dV_j += P^T dO_i, dK_j += dS^T Q_i: A = [q][keys] tile read MN-major (M = keys), K = query rows
              const uint32_t idesc_t = umma_idesc_bf16(kTile, kHd, true, true);
              const int steps_q = kq >> 4;
              for (int c = 0; c < steps_q; ++c)
                umma_bf16(tmem_base + kColDV, umma_smem_desc(sP + c * 2048, 16384, 1024), umma_smem_desc(sG + i * 16384 + c * 2048, 8192, 1024),
                          idesc_t, (i > 0 || c > 0) ? 1u : 0u);
              for (int c = 0; c < steps_q; ++c)
                umma_bf16(tmem_base + kColDK, umma_smem_desc(sS + c * 2048, 16384, 1024), umma_smem_desc(sQ + i * 16384 + c * 2048, 8192, 1024),
                          idesc_t, (i > 0 || c > 0) ? 1u : 0u);
              // dQ_i += dS K_j: A = the dS tile read K-major (M = query rows), K = keys
              const uint32_t idesc_q = umma_idesc_bf16(kTile, kHd, false, true);
              const int steps_k = kk >> 4;
              for (int c = 0; c < steps_k; ++c)
                umma_bf16(tmem_base + kColDQ + i * 64, umma_smem_desc(sS + (c >> 2) * 16384 + (c & 3) * 32, 0, 1024),
                          umma_smem_desc(sK + j * 16384 + c * 2048, 8192, 1024), idesc_q, (j > 0 || c > 0) ? 1u : 0u);
              umma_commit(pds_free);
            }
            if (i == nt - 1) {
              umma_commit(dkv_full);
              ++jcount;
            }
          }
        }
        umma_commit(dq_full);
        umma_commit(empty);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------ P / dS group
    const int wi = warp - 2, part = wi >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const float sl2 = p.scale * kLog2e;
    const uint32_t dkey = DROP ? drop_mix(p.drop_salt, __ldg(p.drop_seed)) : 0u;
    const float dscale = drop_scale(p.drop_thr);
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // this thread's 128-byte segment of the P / dS tiles (64 keys), as shared-space addresses (st.shared, not generic st)
    const uint32_t prow = smem_u32(smem + kPOff + ((part * KP) >> 6) * 16384 + row * 128);
    const uint32_t srow = smem_u32(smem + kSOff + ((part * KP) >> 6) * 16384 + row * 128);
    const int cb = ((part * KP) & 63) >> 3;  // first 16-byte chunk of this thread's keys inside the 128-byte row segment
    const int sw = row & 7;
    const long long ld3 = 3LL * d;
    uint32_t blk = 0, jcount = 0;
    int k = 0;
    for (int item = blockIdx.x; item < p.num_items; item += G, ++k) {
      const int s = item / H, h = item - s * H;
      mbar_wait_park(full, k & 1);
      const uint32_t* meta = reinterpret_cast<const uint32_t*>(smem + kMetaOff);
      const Seq sd{static_cast<int>(meta[8]), static_cast<int>(meta[9]), static_cast<int>(meta[10]), static_cast<int>(meta[11])};
      const int n = sd.len0 + sd.len1;
      const int nt = n > kTile ? 2 : 1;
      const long long stat0 = (static_cast<long long>(s) * H + h) * p.max_seq_len;
      // per query tile: log-sum-exp (log2 domain; +inf for absent rows => P = 0) and delta of this thread's row
      float Lr[2], Dr[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int q = i * kTile + row;
        Lr[i] = q < n ? __ldg(p.lse + stat0 + q) * kLog2e : INFINITY;
        Dr[i] = q < n ? __ldg(p.delta + stat0 + q) : 0.f;
      }
      for (int j = 0; j < nt; ++j) {
        // keep bits of this thread's KP keys: keys j 128 + part KP + [0, KP)
        const uint32_t kw_lo = meta[j * 4 + ((part * KP) >> 5)], kw_hi = KP > 32 ? meta[j * 4 + ((part * KP) >> 5) + 1] : 0u;
        for (int i = 0; i < nt; ++i, ++blk) {
          const int q = i * kTile + row;
          const float L = Lr[i], D = Dr[i];
          const uint32_t drow = attn_drop_row(s, H, h, p.max_seq_len, q) + ((j * kTile + part * KP) >> 1);
          mbar_wait_park(sdp_full, blk & 1);
          if (!EARLY_S && blk > 0) mbar_wait_park(pds_free, (blk & 1) ^ 1);  // the previous block's MMAs are done with the P / dS tiles
          __syncwarp();
          tcgen05_fence_after();
#pragma unroll 1
          for (int c = 0; c < KP / 16; ++c) {  // 16 keys at a time
            uint32_t sv[16], dv[16];
            tmem_ld_32x16(trow + kColS + part * KP + c * 16, sv);
            tmem_ld_32x16(trow + kColDP + part * KP + c * 16, dv);
            tmem_ld_wait();
            const uint32_t kw = ((c < 2 ? kw_lo : kw_hi) >> (16 * (c & 1))) & 0xffffu;  // KP = 32: one word, c in {0, 1}
            uint32_t pp[8], ds[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float p0 = ((kw >> (2 * e)) & 1u) ? ex2_approx(fmaf(__uint_as_float(sv[2 * e]), sl2, -L)) : 0.f;
              float p1 = ((kw >> (2 * e + 1)) & 1u) ? ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), sl2, -L)) : 0.f;
              float g0 = __uint_as_float(dv[2 * e]), g1 = __uint_as_float(dv[2 * e + 1]);
              float pd0 = p0, pd1 = p1;
              if (DROP) {  // dropped probabilities feed dV; dP of a dropped one is zero, of a kept one carries the scale
                const uint32_t word = drop_mix(drow + c * 8 + e, dkey);
                const bool k0 = (word & 255u) >= p.drop_thr, k1 = ((word >> 16) & 255u) >= p.drop_thr;
                pd0 = k0 ? p0 * dscale : 0.f;
                pd1 = k1 ? p1 * dscale : 0.f;
                g0 = k0 ? g0 * dscale : 0.f;
                g1 = k1 ? g1 * dscale : 0.f;
              }
              // keys that do not take part: exact zeros whatever the (possibly stale) S / dP columns hold
              const float s0 = ((kw >> (2 * e)) & 1u) ? p0 * (g0 - D) : 0.f;
              const float s1 = ((kw >> (2 * e + 1)) & 1u) ? p1 * (g1 - D) : 0.f;
              pp[e] = pack_bf16(pd0, pd1);
              ds[e] = pack_bf16(s0, s1);
            }
            if (EARLY_S && c == 0 && blk > 0) mbar_wait_park(pds_free, (blk & 1) ^ 1);  // first store of the block
            // keys c 16 + [0, 16) of this thread's KP: 16-byte chunks cb + 2 c and cb + 2 c + 1 of the row segment
            sts_v4_u32(prow + (((cb + 2 * c) ^ sw) << 4), pp[0], pp[1], pp[2], pp[3]);
            sts_v4_u32(prow + (((cb + 2 * c + 1) ^ sw) << 4), pp[4], pp[5], pp[6], pp[7]);
            sts_v4_u32(srow + (((cb + 2 * c) ^ sw) << 4), ds[0], ds[1], ds[2], ds[3]);
            sts_v4_u32(srow + (((cb + 2 * c + 1) ^ sw) << 4), ds[4], ds[5], ds[6], ds[7]);
          }
          fence_proxy_async();  // the tiles are read by the tensor core (async proxy)
          tcgen05_fence_before();
          mbar_arrive(pds_full);
        }
        // ---- dK_j, dV_j: lane = key row, this thread's 32 of the 64 columns of each
        mbar_wait_park(dkv_full, jcount & 1);
        ++jcount;
        __syncwarp();
        tcgen05_fence_after();
        {
          uint32_t a[DC], b[DC];
          tmem_ld_cols<DC>(trow + kColDK + part * DC, a);
          tmem_ld_cols<DC>(trow + kColDV + part * DC, b);
          tmem_ld_wait();
          tcgen05_fence_before();
          mbar_arrive(dkv_free);
          const int key = j * kTile + row;
          if (key < n) {
            __nv_bfloat16* base = p.dqkv + seq_row(sd, key) * ld3 + h * kHd + part * DC;
            uint4* dk = reinterpret_cast<uint4*>(base + d);
            uint4* dvp = reinterpret_cast<uint4*>(base + 2 * d);
#pragma unroll
            for (int e = 0; e < DC / 8; ++e) {
              dk[e] = make_uint4(pack_bf16(__uint_as_float(a[8 * e]) * p.scale, __uint_as_float(a[8 * e + 1]) * p.scale),
                                 pack_bf16(__uint_as_float(a[8 * e + 2]) * p.scale, __uint_as_float(a[8 * e + 3]) * p.scale),
                                 pack_bf16(__uint_as_float(a[8 * e + 4]) * p.scale, __uint_as_float(a[8 * e + 5]) * p.scale),
                                 pack_bf16(__uint_as_float(a[8 * e + 6]) * p.scale, __uint_as_float(a[8 * e + 7]) * p.scale));
              dvp[e] = make_uint4(pack_bf16(__uint_as_float(b[8 * e]), __uint_as_float(b[8 * e + 1])),
                                  pack_bf16(__uint_as_float(b[8 * e + 2]), __uint_as_float(b[8 * e + 3])),
                                  pack_bf16(__uint_as_float(b[8 * e + 4]), __uint_as_float(b[8 * e + 5])),
                                  pack_bf16(__uint_as_float(b[8 * e + 6]), __uint_as_float(b[8 * e + 7])));
            }
          }
        }
      }
      // ---- dQ_0, dQ_1
      mbar_wait_park(dq_full, k & 1);
      __syncwarp();
      tcgen05_fence_after();
      {
        uint32_t a[DC], b[DC];
        tmem_ld_cols<DC>(trow + kColDQ + part * DC, a);
        if (nt > 1) tmem_ld_cols<DC>(trow + kColDQ + 64 + part * DC, b);
        tmem_ld_wait();
        tcgen05_fence_before();
        mbar_arrive(dq_free);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int q = i * kTile + row;
          if (i < nt && q < n) {
            uint4* dq = reinterpret_cast<uint4*>(p.dqkv + seq_row(sd, q) * ld3 + h * kHd + part * DC);
            const uint32_t* v = i == 0 ? a : b;
#pragma unroll
            for (int e = 0; e < DC / 8; ++e)
              dq[e] = make_uint4(pack_bf16(__uint_as_float(v[8 * e]) * p.scale, __uint_as_float(v[8 * e + 1]) * p.scale),
                                 pack_bf16(__uint_as_float(v[8 * e + 2]) * p.scale, __uint_as_float(v[8 * e + 3]) * p.scale),
                                 pack_bf16(__uint_as_float(v[8 * e + 4]) * p.scale, __uint_as_float(v[8 * e + 5]) * p.scale),
                                 pack_bf16(__uint_as_float(v[8 * e + 6]) * p.scale, __uint_as_float(v[8 * e + 7]) * p.scale));
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
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

int attn_delta_launch(const void* out, const void* dout, const int32_t* seq_desc, float* delta, int H, int max_seq_len, int num_seqs,
                      cudaStream_t stream) {
  const long long rows = static_cast<long long>(num_seqs) * max_seq_len;
  attn_delta_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(out),
                                                                                static_cast<const __nv_bfloat16*>(dout), seq_desc, delta, H,
                                                                                max_seq_len, num_seqs);
  return check_launch("attn_delta");
}

int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const int32_t* seq_desc, const uint8_t* key_mask, const float* lse,
                void* dqkv, float* delta_ws, int64_t tokens, int num_seqs, int max_seq_len, int H, float scale, const uint32_t* drop_seed,
                uint32_t drop_salt, float drop_p, cudaStream_t stream) {
  MOME_REQUIRE(max_seq_len <= kMaxKeys, "attn_bwd_tc: max_seq_len %d > %d", max_seq_len, kMaxKeys);
  static bool configured = false;
  if (!configured) {
    int rc = opt_in(attn_bwd_tc_kernel<false, false, 8>, kBwdSmem, "attn_bwd_tc");
    if (rc == MOME_OK) rc = opt_in(attn_bwd_tc_kernel<true, false, 8>, kBwdSmem, "attn_bwd_tc");
    if (rc == MOME_OK) rc = opt_in(attn_bwd_tc_kernel<false, true, 8>, kBwdSmem, "attn_bwd_tc");
    if (rc == MOME_OK) rc = opt_in(attn_bwd_tc_kernel<true, true, 8>, kBwdSmem, "attn_bwd_tc");
    if (rc == MOME_OK) rc = opt_in(attn_bwd_tc_kernel<false, false, 16>, kBwdSmem, "attn_bwd_tc");
    if (rc == MOME_OK) rc = opt_in(attn_bwd_tc_kernel<true, false, 16>, kBwdSmem, "attn_bwd_tc");
    if (rc != MOME_OK) return rc;
    configured = true;
  }
  const char* variant = getenv("MOME_ATTN_TC_BWD");
  const bool early_s = variant != nullptr && variant[0] == '2';
  const bool wide = variant != nullptr && variant[0] == '3';  // 16 P / dS warps (measured: 404 vs 405 us without, 509 vs 478 us with dropout: no gain)
  BwdParams p;
  const int64_t d = static_cast<int64_t>(H) * kHd, d3 = 3 * d;
  int rc = tma_encode_bf16_2d(&p.qkv32, qkv, d3, tokens, d3, kHd, 32);
  if (rc == MOME_OK) rc = tma_encode_bf16_2d(&p.qkv8, qkv, d3, tokens, d3, kHd, 8);
  if (rc == MOME_OK) rc = tma_encode_bf16_2d(&p.go32, dout, d, tokens, d, kHd, 32);
  if (rc == MOME_OK) rc = tma_encode_bf16_2d(&p.go8, dout, d, tokens, d, kHd, 8);
  if (rc != MOME_OK) return rc;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.seq_desc = seq_desc;
  p.key_mask = key_mask;
  p.lse = lse;
  p.delta = delta_ws;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.H = H;
  p.max_seq_len = max_seq_len;
  p.num_items = num_seqs * H;
  p.scale = scale;
  p.drop_seed = drop_seed;
  p.drop_salt = drop_salt;
  p.drop_thr = drop_threshold(drop_p);
  rc = attn_delta_launch(out, dout, seq_desc, delta_ws, H, max_seq_len, num_seqs, stream);
  if (rc != MOME_OK) return rc;
  const int grid = std::min(p.num_items, sm_count());
  const bool drop = drop_seed != nullptr && drop_p > 0.f;
  if (wide && drop) attn_bwd_tc_kernel<true, false, 16><<<grid, kBwdThreads16, kBwdSmem, stream>>>(p);
  else if (wide) attn_bwd_tc_kernel<false, false, 16><<<grid, kBwdThreads16, kBwdSmem, stream>>>(p);
  else if (drop && early_s) attn_bwd_tc_kernel<true, true, 8><<<grid, kBwdThreads8, kBwdSmem, stream>>>(p);
  else if (drop) attn_bwd_tc_kernel<true, false, 8><<<grid, kBwdThreads8, kBwdSmem, stream>>>(p);
  else if (early_s) attn_bwd_tc_kernel<false, true, 8><<<grid, kBwdThreads8, kBwdSmem, stream>>>(p);
  else attn_bwd_tc_kernel<false, false, 8><<<grid, kBwdThreads8, kBwdSmem, stream>>>(p);
  return check_launch("attn_bwd_tc");
}

}  // namespace mome
