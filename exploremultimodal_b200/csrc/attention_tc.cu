// K1 on the 5th-generation tensor cores: masked multi-head self-attention over packed segments with
// tcgen05.mma, accumulators in TMEM. Sequences of the MoME path are short (40 text + 197 image tokens = 237
// at most for the pre-training configs), so a whole row of scores fits one MMA (N <= 256): no online softmax,
// no running rescale — one S = Q K^T per 128-query tile, an exact two-pass softmax straight out of TMEM, one
// O = P V with P fed back from TMEM. Same contract as attention_mma.cu (mask semantics, lse format, dropout
// mask function), so the forward kernels are interchangeable and the mma.sync backward works with either.
//
// Replaces: reference vlmo.py:79-95. Used for 64 < max_seq_len <= 256 (attention_api.cu; MOME_ATTN_TC=0 turns it
// off). Measured (tools/attn_bench.py, 256 x [40 | 197] tokens, 12 heads): 166 us against 243 us for the mma.sync
// kernel. Lessons kept in the code: issue tcgen05.mma / TMA under elect.sync, never under `lane == 0` (the
// compiler wraps every instruction in a uniform-register waterfall loop, ~220 cycles per MMA, tools/umma_probe.cu).
#include <cuda.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "dropout.cuh"
#include "ptx.cuh"

namespace mome {

namespace {

constexpr int kHd = 64;           // head_dim
constexpr int kQTile = 128;       // queries per CTA = MMA M = TMEM lanes
constexpr int kMaxKeys = 256;     // MMA N limit
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

constexpr uint32_t kTmemCols = 512;

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

struct AttnDrop {
  const uint32_t* seed;
  uint32_t salt, thr;
};
// same mask function as attention_mma.cu: one word per (sequence, head, query, pair of adjacent keys)
__device__ __forceinline__ uint32_t attn_drop_row(int s, int H, int h, int max_seq_len, int i) {
  return ((static_cast<uint32_t>(s) * H + h) * max_seq_len + i) * static_cast<uint32_t>((max_seq_len + 1) >> 1);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Forward: persistent and warp-specialised. One CTA per SM walks over (sequence, head) items;
// both 128-query tiles of an item share one load of K and V, the probabilities never touch shared memory.
//
//   warp 0      producer: TMA boxes (8 / 32 rows x 64 columns, SWIZZLE_128B) of the item's Q, K, V rows into a
//               2-stage ring (96 KB per stage), key-mask words and the sequence descriptor next to them;
//               layouts whose second range does not start on an 8-row boundary take a cp.async gather instead
//   warp 1      one thread issues every tcgen05.mma: S = Q K^T per tile (SS), O = P V per tile (A = P from TMEM)
//   warps 2-9   softmax group 0, warps 10-17 softmax group 1: a group owns one 256-column TMEM slot; tiles
//               ("units") alternate between the groups, so one group's exponentials overlap the other's
//               MMAs, epilogue and barrier latencies. Two threads per query row (TMEM lane), one per half of the
//               keys. P (bf16) is written back over S in place (tcgen05.st), O accumulates in columns [64, 128).
// Barriers: full / empty per stage (producer <-> MMA), s_full / p_full / o_full / tmem_free per slot.
constexpr int kStageBytes = 3 * 32768;       // Q (256 rows) | K | V
constexpr int kStages = 2;
constexpr int kMetaOff = kStages * kStageBytes;   // per stage 64 B: keep words [8], seq desc [4]
constexpr int kBarOff = kMetaOff + kStages * 64;  // 12 barriers, TMEM slot
constexpr int kXchgOff = kBarOff + 128;            // row max / row sum exchange between the two threads of a row: 2 x 2 KB
constexpr int kFwdSmem = kXchgOff + 4096 + 1024;
constexpr int kFwdThreads = 64 + 2 * 256;         // producer, MMA, 2 softmax groups of 8 warps
constexpr uint32_t kSlotCols = 256, kOCol = 64, kPHiCol = 128;

struct FwdParams {
  CUtensorMap map32, map8;  // qkv as [tokens][3 d] bf16, boxes of 64 columns x 32 / 8 rows
  const __nv_bfloat16* qkv;
  const int32_t* seq_desc;
  const uint8_t* key_mask;
  __nv_bfloat16* out;
  float* lse;
  int H, max_seq_len, num_items;
  float scale;
  AttnDrop ad;
  long long* dbg;  // MOME_ATTN_DBG: clock64 of pipeline events of CTA 0, [unit or item][16]
};

#define DBG(idx, ev) do { if (p.dbg != nullptr && blockIdx.x == 0 && (idx) < 64) p.dbg[(idx) * 16 + (ev)] = clock64(); } while (0)

__device__ __forceinline__ int seq_len(const int32_t* seq_desc, int s) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(seq_desc + 4 * s));
  return v.y + v.w;
}

// Boxes of one row range: 32-row boxes, then 8-row boxes for the remainder (the last may run past the range:
// it reads rows of the next sequence or zero fill, both finite, and those keys are masked).
__device__ __forceinline__ int range_boxes(int len) { return (len >> 5) + (((len & 31) + 7) >> 3); }
__device__ __forceinline__ uint32_t range_bytes(int len) { return (len >> 5) * 4096u + (((len & 31) + 7) >> 3) * 1024u; }
__device__ __forceinline__ void issue_range_box(const FwdParams& p, uint8_t* dst_rows, uint64_t* bar, int col, int grow, int len, int b) {
  const int nb = len >> 5;
  if (b < nb) {
    tma_load_2d(dst_rows + b * 4096, &p.map32, bar, col, grow + 32 * b);
  } else {
    const int r = 32 * nb + 8 * (b - nb);
    tma_load_2d(dst_rows + r * 128, &p.map8, bar, col, grow + r);
  }
}

template <bool DROP>
__global__ void __launch_bounds__(kFwdThreads, 1) attn_fwd_tc_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* full = bars;            // [2] producer -> MMA / softmax: stage loaded
  uint64_t* empty = bars + 2;       // [2] MMA -> producer: every MMA reading the stage has completed
  uint64_t* s_full = bars + 4;      // [2] MMA -> group: S of the unit is in the slot
  uint64_t* p_full = bars + 6;      // [2] group -> MMA: P written (128 arrivals)
  uint64_t* o_full = bars + 8;      // [2] MMA -> group: O complete
  uint64_t* tmem_free = bars + 10;  // [2] group -> MMA: O has been read, the slot may be overwritten (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, d = H * kHd;
  const int G = gridDim.x;

  // stale shared memory must at least be finite: rows between the last box and n_pad multiply p = 0
  for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += kFwdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&o_full[i], 1);
      mbar_init(&tmem_free[i], 256);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.map32);
    tma_prefetch_desc(&p.map8);
  }
  __syncwarp();
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------ producer
    // full[stage] takes two arrivals: the data (expect_tx + TMA bytes, or the gather) and the meta words.
    // Everything that does not need the stage (descriptor, key-mask bytes) is fetched before the wait for it.
    int k = 0;
    Seq sd_next = blockIdx.x < p.num_items ? load_seq(p.seq_desc, blockIdx.x / H) : Seq{0, 0, 0, 0};
    for (int item = blockIdx.x; item < p.num_items; item += G, ++k) {
      const int stage = k & 1;
      const int s = item / H, h = item - s * H;
      const Seq sd = sd_next;
      if (item + G < p.num_items) sd_next = load_seq(p.seq_desc, (item + G) / H);
      const int n = sd.len0 + sd.len1;
      uint8_t* st = smem + stage * kStageBytes;
      uint32_t* meta = reinterpret_cast<uint32_t*>(smem + kMetaOff + stage * 64);
      uint8_t mk[8];
#pragma unroll
      for (int j8 = 0; j8 < 8; ++j8) {  // 8 independent loads in flight
        const int j = j8 * 32 + lane;
        mk[j8] = j < n ? (p.key_mask == nullptr ? uint8_t(1) : __ldg(p.key_mask + seq_row(sd, j))) : uint8_t(0);
      }
      mbar_wait_park(&empty[stage], ((k >> 1) & 1) ^ 1);
      __syncwarp();
      if (lane == 0) DBG(k, 9);
      const bool boxes_ok = sd.len1 == 0 || (sd.len0 & 7) == 0;
      if (boxes_ok) {
        const int nb0 = range_boxes(sd.len0), nb1 = range_boxes(sd.len1);
        if (elect_one()) {  // one lane, uniform operands: per-lane boxes made the compiler serialise the warp around every TMA
          mbar_arrive_expect_tx(&full[stage], 3u * (range_bytes(sd.len0) + range_bytes(sd.len1)));
#pragma unroll 1
          for (int op = 0; op < 3; ++op) {  // operand 0 / 1 / 2 = Q / K / V
            const int col = op * d + h * kHd;
            uint8_t* dst = st + op * 32768;
            for (int bb = 0; bb < nb0; ++bb) issue_range_box(p, dst, &full[stage], col, sd.start0, sd.len0, bb);
            for (int bb = 0; bb < nb1; ++bb) issue_range_box(p, dst + sd.len0 * 128, &full[stage], col, sd.start1, sd.len1, bb);
          }
          DBG(k, 10);
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
        mbar_arrive(&full[stage]);
      }
      if (!boxes_ok) {
        // gather path: 4 rows per pass, zero fill up to n_pad
        const int n_pad = (n + 15) & ~15;
        const int r0 = lane >> 3, ch = lane & 7;
        const long long ld = 3LL * d;
        for (int r = r0; r < n_pad; r += 4) {
          const bool valid = r < n;
          const __nv_bfloat16* src = p.qkv + seq_row(sd, valid ? r : 0) * ld + h * kHd + ch * 8;
          const uint32_t off = r * 128 + ((ch ^ (r & 7)) << 4);
          cp_async_16(st + off, src, valid);
          cp_async_16(st + 32768 + off, src + d, valid);
          cp_async_16(st + 65536 + off, src + 2 * d, valid);
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------ MMA issuer
    // One thread polls both slots: per slot the units alternate S (needs the stage and the slot's previous
    // epilogue) and PV (needs the group's P). Never blocks on one slot while the other has work.
    // elect.sync (not `lane == 0`) is what lets the compiler keep descriptors in uniform registers: with a plain
    // lane test every tcgen05.mma was wrapped in an R2UR waterfall loop, ~220 cycles per instruction (tools/umma_probe.cu).
    if (elect_one()) {
      const int first = blockIdx.x;
      auto item_n = [&](int k) { return first + k * G < p.num_items ? seq_len(p.seq_desc, (first + k * G) / H) : 0; };
      struct It { int k, t, u, n; };  // local item, tile, unit number, sequence length (0: past the end)
      auto advance = [&](It& it) {
        ++it.u;
        if (++it.t >= (it.n > kQTile ? 2 : 1)) { it.t = 0; ++it.k; it.n = item_n(it.k); }
      };
      It sIt[2], pIt[2];
      sIt[0] = It{0, 0, 0, item_n(0)};
      sIt[1] = sIt[0];
      if (sIt[1].n > 0) advance(sIt[1]);
      pIt[0] = sIt[0];
      pIt[1] = sIt[1];
      int pv_count[2] = {0, 0};
      while (pIt[0].n > 0 || pIt[1].n > 0) {
        bool progress = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          It& pp = pIt[g];
          It& sp = sIt[g];
          if (pp.n > 0 && sp.u > pp.u && mbar_test_wait(&p_full[g], (pp.u >> 1) & 1)) {
            tcgen05_fence_after();
            const int stage = pp.k & 1;
            const int n_pad = (pp.n + 15) & ~15;
            const uint32_t idesc = umma_idesc_bf16(kQTile, kHd, false, true);
            const uint32_t va = smem_u32(smem + stage * kStageBytes) + 65536;
            const uint32_t pt = tmem_base + g * kSlotCols, ot = pt + kOCol;
            const int ksteps = n_pad >> 4;
            // P of keys [0, 128) sits in columns [0, 64), of keys [128, 256) in [128, 192); V advances 16 rows per step
            uint64_t vdesc = umma_smem_desc(va, 8192, 1024);
            const int lo = min(ksteps, 8);
            for (int kk = 0; kk < lo; ++kk, vdesc += 2048 >> 4) umma_bf16_ts(ot, pt + kk * 8, vdesc, idesc, kk > 0 ? 1u : 0u);
            for (int kk = 8; kk < ksteps; ++kk, vdesc += 2048 >> 4) umma_bf16_ts(ot, pt + kPHiCol + (kk - 8) * 8, vdesc, idesc, 1u);
            umma_commit(&o_full[g]);
            DBG(pp.u, 3);
            if (++pv_count[stage] == (pp.n > kQTile ? 2 : 1)) {  // every MMA reading the stage has been issued
              umma_commit(&empty[stage]);
              pv_count[stage] = 0;
            }
            advance(pp);
            if (pp.n > 0) advance(pp);
            progress = true;
          }
          if (sp.n > 0 && sp.u == pp.u && mbar_test_wait(&full[sp.k & 1], (sp.k >> 1) & 1) &&
              mbar_test_wait(&tmem_free[g], ((sp.u >> 1) & 1) ^ 1)) {
            tcgen05_fence_after();
            const int stage = sp.k & 1;
            const int n_pad = (sp.n + 15) & ~15;
            const uint32_t idesc = umma_idesc_bf16(kQTile, n_pad, false, false);
            const uint32_t qa = smem_u32(smem + stage * kStageBytes) + sp.t * 16384, ka = smem_u32(smem + stage * kStageBytes) + 32768;
            const uint32_t dt = tmem_base + g * kSlotCols;
#pragma unroll
            for (int kk = 0; kk < kHd / 16; ++kk)
              umma_bf16(dt, umma_smem_desc(qa + kk * 32, 0, 1024), umma_smem_desc(ka + kk * 32, 0, 1024), idesc, kk > 0 ? 1u : 0u);
            umma_commit(&s_full[g]);
            DBG(sp.u, 1);
            advance(sp);
            if (sp.n > 0) advance(sp);
            progress = true;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------ softmax groups
    // Group g = 8 warps on TMEM slot g. Two threads per query row: `half` 0 takes keys [0, 128), half 1 keys
    // [128, n_pad); each converts its part of S to bf16 P in place (columns [0, 64) and [128, 192) of the slot),
    // O accumulates in columns [64, 128), which half 0 has finished reading by then.
    const int wi = warp - 2, g = wi >> 3, half = (wi >> 2) & 1, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const float sl2 = p.scale * kLog2e;
    const uint32_t dkey = DROP ? drop_mix(p.ad.salt, __ldg(p.ad.seed)) : 0u;
    const float dscale = drop_scale(p.ad.thr);
    const uint32_t trow = tmem_base + g * kSlotCols + (static_cast<uint32_t>(quarter * 32) << 16);
    float* xmax = reinterpret_cast<float*>(smem + kXchgOff) + (g * 2 + half) * 128;      // this thread's slot; partner: ^ 128
    float* xsum = reinterpret_cast<float*>(smem + kXchgOff + 2048) + (g * 2 + half) * 128;
    const int pair_bar = 1 + g * 4 + quarter;  // named barrier of the two warps sharing these 32 rows
    int u = 0, k = 0;
    for (int item = blockIdx.x; item < p.num_items; item += G, ++k) {
      const int s = item / H, h = item - s * H;
      const int n = seq_len(p.seq_desc, s);
      const int tiles = n > kQTile ? 2 : 1;
      // every group takes every item's `full` phase, also when the item's only tile belongs to the other group: a group that
      // skipped phases could test a later phase of the same parity before the producer got there (parities alias two fills apart)
      mbar_wait_park(&full[k & 1], (k >> 1) & 1);
      for (int t = 0; t < tiles; ++t, ++u) {
        if ((u & 1) != g) continue;
        const int stage = k & 1;
        const uint32_t uph = (u >> 1) & 1;
        mbar_wait_park(&full[stage], (k >> 1) & 1);  // orders the producer's meta writes before the reads below
        const uint32_t* meta = reinterpret_cast<const uint32_t*>(smem + kMetaOff + stage * 64);
        const Seq sd{static_cast<int>(meta[8]), static_cast<int>(meta[9]), static_cast<int>(meta[10]), static_cast<int>(meta[11])};
        const uint32_t* keepw = meta;  // read before this thread's p_full arrive, i.e. before the stage can be recycled
        const int nsub = ((n + 15) & ~15) >> 4;                              // 16-key sub-chunks of the row
        const int c0 = half * 8, c1 = half == 0 ? min(nsub, 8) : nsub;        // this thread's sub-chunks [c0, c1)
        const int q = t * kQTile + row;
        const bool warp_live = t * kQTile + quarter * 32 < n;  // some row of this warp is a real query
        mbar_wait_park(&s_full[g], uph);
        __syncwarp();
        tcgen05_fence_after();
        const bool dbg_thread = (wi & 7) == 0 && lane == 0;
        if (dbg_thread) DBG(u, 4);
        float l = 0.f, ms = 0.f;
        if (warp_live) {
          uint32_t ra[16], rb[16];
          auto keep16 = [&](int j) { return (keepw[j >> 1] >> (16 * (j & 1))) & 0xffffu; };
          // ---- pass 1: row maximum over the keys that take part (TMEM loads run one sub-chunk ahead)
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          auto max_chunk = [&](const uint32_t (&r)[16], uint32_t kw) {
            if (kw == 0xffffu) {
#pragma unroll
              for (int c = 0; c < 16; c += 4) {
                m0 = fmaxf(m0, __uint_as_float(r[c]));
                m1 = fmaxf(m1, __uint_as_float(r[c + 1]));
                m2 = fmaxf(m2, __uint_as_float(r[c + 2]));
                m3 = fmaxf(m3, __uint_as_float(r[c + 3]));
              }
            } else {
#pragma unroll
              for (int c = 0; c < 16; c += 4) {
                m0 = fmaxf(m0, ((kw >> c) & 1u) ? __uint_as_float(r[c]) : -INFINITY);
                m1 = fmaxf(m1, ((kw >> (c + 1)) & 1u) ? __uint_as_float(r[c + 1]) : -INFINITY);
                m2 = fmaxf(m2, ((kw >> (c + 2)) & 1u) ? __uint_as_float(r[c + 2]) : -INFINITY);
                m3 = fmaxf(m3, ((kw >> (c + 3)) & 1u) ? __uint_as_float(r[c + 3]) : -INFINITY);
              }
            }
          };
          if (c0 < c1) tmem_ld_32x16(trow + c0 * 16, ra);
#pragma unroll 1
          for (int j = c0; j < c1; j += 2) {
            tmem_ld_wait();
            if (j + 1 < c1) tmem_ld_32x16(trow + (j + 1) * 16, rb);
            max_chunk(ra, keep16(j));
            if (j + 1 < c1) {
              tmem_ld_wait();
              if (j + 2 < c1) tmem_ld_32x16(trow + (j + 2) * 16, ra);
              max_chunk(rb, keep16(j + 1));
            }
          }
          const float mine = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          xmax[row] = mine;
          if (c0 < c1) tmem_ld_32x16(trow + c0 * 16, ra);  // first sub-chunk of pass 2 flies during the exchange
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          const float mx = fmaxf(mine, xmax[row + (half == 0 ? 128 : -128)]);
          if (dbg_thread) DBG(u, 5);
          ms = mx == -INFINITY ? 0.f : mx * sl2;  // every key masked: probabilities stay 0
          // ---- pass 2: probabilities, normaliser, dropout, bf16 P over this thread's part of S in place
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
          const uint32_t drow = attn_drop_row(s, H, h, p.max_seq_len, q);
          auto prob_chunk = [&](uint32_t (&r)[16], int j) {
            const uint32_t kw = keep16(j);
            if (kw == 0xffffu) {
#pragma unroll
              for (int c = 0; c < 16; ++c) r[c] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(r[c]), sl2, -ms)));
            } else if (kw == 0u) {  // padding: no exponentials
#pragma unroll
              for (int c = 0; c < 16; ++c) r[c] = 0u;
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                r[c] = ((kw >> c) & 1u) ? __float_as_uint(ex2_approx(fmaf(__uint_as_float(r[c]), sl2, -ms))) : 0u;
            }
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              l0 += __uint_as_float(r[c]); l1 += __uint_as_float(r[c + 1]); l2 += __uint_as_float(r[c + 2]); l3 += __uint_as_float(r[c + 3]);
            }
            uint32_t pk[8];
            if (DROP) {  // the normaliser keeps the undropped probabilities; only what multiplies V is dropped
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint32_t word = drop_mix(drow + 8 * j + c, dkey);
                const float a = (word & 255u) >= p.ad.thr ? __uint_as_float(r[2 * c]) * dscale : 0.f;
                const float b = ((word >> 16) & 255u) >= p.ad.thr ? __uint_as_float(r[2 * c + 1]) * dscale : 0.f;
                pk[c] = pack_bf16(a, b);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c) pk[c] = pack_bf16(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1]));
            }
            tmem_st_32x8(trow + half * 128 + (j - c0) * 8, pk);
          };
#pragma unroll 1
          for (int j = c0; j < c1; j += 2) {
            tmem_ld_wait();
            if (j + 1 < c1) tmem_ld_32x16(trow + (j + 1) * 16, rb);
            prob_chunk(ra, j);
            if (j + 1 < c1) {
              tmem_ld_wait();
              if (j + 2 < c1) tmem_ld_32x16(trow + (j + 2) * 16, ra);
              prob_chunk(rb, j + 1);
            }
          }
          l = (l0 + l1) + (l2 + l3);
          xsum[row] = l;
          tmem_st_wait();
        }
        tcgen05_fence_before();
        mbar_arrive(&p_full[g]);
        if (dbg_thread) DBG(u, 6);
        mbar_wait_park(&o_full[g], uph);
        if (dbg_thread) DBG(u, 7);
        __syncwarp();
        tcgen05_fence_after();
        if (warp_live) {
          uint32_t a[32];
          tmem_ld_32x32(trow + kOCol + half * 32, a);  // this thread's 32 of the row's 64 output columns
          l += xsum[row + (half == 0 ? 128 : -128)];   // the partner stored it before its p_full arrive
          tmem_ld_wait();
          tcgen05_fence_before();
          mbar_arrive(&tmem_free[g]);  // O is in registers: the slot can take the next S while the rows are written out
          if (q < n) {
            const float inv = l > 0.f ? 1.f / l : 0.f;
            uint4* dst = reinterpret_cast<uint4*>(p.out + seq_row(sd, q) * d + h * kHd + half * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 v;
              v.x = pack_bf16(__uint_as_float(a[8 * i]) * inv, __uint_as_float(a[8 * i + 1]) * inv);
              v.y = pack_bf16(__uint_as_float(a[8 * i + 2]) * inv, __uint_as_float(a[8 * i + 3]) * inv);
              v.z = pack_bf16(__uint_as_float(a[8 * i + 4]) * inv, __uint_as_float(a[8 * i + 5]) * inv);
              v.w = pack_bf16(__uint_as_float(a[8 * i + 6]) * inv, __uint_as_float(a[8 * i + 7]) * inv);
              dst[i] = v;
            }
            if (half == 0) p.lse[(static_cast<long long>(s) * H + h) * p.max_seq_len + q] = l > 0.f ? (ms + log2f(l)) * kLn2 : -INFINITY;
          }
        }
        if (!warp_live) {
          tcgen05_fence_before();
          mbar_arrive(&tmem_free[g]);
        }
        if (dbg_thread) DBG(u, 8);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
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

int tma_encode_bf16_2d(void* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer);

int attn_fwd_tc(const void* qkv, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse, int64_t tokens, int num_seqs,
                 int max_seq_len, int H, float scale, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, cudaStream_t stream) {
  MOME_REQUIRE(max_seq_len <= kMaxKeys, "attn_fwd_tc: max_seq_len %d > %d", max_seq_len, kMaxKeys);
  static bool configured = false;
  if (!configured) {
    int rc = opt_in(attn_fwd_tc_kernel<false>, kFwdSmem, "attn_fwd_tc");
    if (rc == MOME_OK) rc = opt_in(attn_fwd_tc_kernel<true>, kFwdSmem, "attn_fwd_tc");
    if (rc != MOME_OK) return rc;
    configured = true;
  }
  FwdParams p;
  const int64_t d3 = 3LL * H * kHd;
  int rc = tma_encode_bf16_2d(&p.map32, qkv, d3, tokens, d3, kHd, 32);
  if (rc == MOME_OK) rc = tma_encode_bf16_2d(&p.map8, qkv, d3, tokens, d3, kHd, 8);
  if (rc != MOME_OK) return rc;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.seq_desc = seq_desc;
  p.key_mask = key_mask;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.H = H;
  p.max_seq_len = max_seq_len;
  p.num_items = num_seqs * H;
  p.scale = scale;
  p.ad = AttnDrop{drop_seed, drop_salt, drop_threshold(drop_p)};
  p.dbg = nullptr;
  const bool dbg = getenv("MOME_ATTN_DBG") != nullptr;  // debugging aid: prints the event times of CTA 0, synchronises
  if (dbg) {
    cudaMalloc(&p.dbg, 64 * 16 * 8);
    cudaMemsetAsync(p.dbg, 0, 64 * 16 * 8, stream);
  }
  const int grid = std::min(p.num_items, sm_count());
  if (drop_seed != nullptr && drop_p > 0.f)
    attn_fwd_tc_kernel<true><<<grid, kFwdThreads, kFwdSmem, stream>>>(p);
  else
    attn_fwd_tc_kernel<false><<<grid, kFwdThreads, kFwdSmem, stream>>>(p);
  if (dbg) {
    static long long h[64 * 16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    long long t0 = h[0];
    fprintf(stderr, "attn_fwd_tc events of CTA 0 (cycles since first S wait): idx: S.start S.issued PV.start PV.issued | s_full pass1 p_arrive o_full epi_done | prod.start prod.issued full.seen\n");
    for (int i = 0; i < 24; ++i) {
      fprintf(stderr, "%2d:", i);
      for (int e = 0; e < 12; ++e) fprintf(stderr, " %7lld", h[i * 16 + e] ? h[i * 16 + e] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  return check_launch("attn_fwd_tc");
}

}  // namespace mome
