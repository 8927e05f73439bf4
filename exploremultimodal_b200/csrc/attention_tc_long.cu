// K1 forward on tcgen05 / TMEM / TMA for sequences of 257 .. 1024 tokens (VQA at 480 px: 40 + 901; 384 px: 40 + 577).
// Same contract as attention_tc.cu / attention_mma.cu (mask semantics, lse format, dropout mask function). A row of scores
// no longer fits one MMA, so the keys are visited in blocks of 192 with an online softmax:
//
//   item        (sequence, head, pair of 128-query tiles); softmax group g owns query tile g of the pair and one 256-column
//               TMEM slot: S_j in columns [0, 192), O in [192, 256)
//   per block   tcgen05: S_j = Q K_j^T (SS)  ->  group: block maximum; if the running maximum grows, O and the running sum
//               are rescaled in TMEM (tcgen05.ld / st of this thread's 32 columns) once PV_(j-1) has completed;
//               P_j = exp2(S_j scale log2e - max) as bf16 over S_j in place  ->  tcgen05: O += P_j V_j (A = P from TMEM)
//   the two groups alternate on the tensor core, so one group's exponentials overlap the other's MMAs
//
//   warp 0      producer: Q tiles of the pair (double buffered per item), K_j / V_j blocks into a 3-stage ring (TMA boxes;
//               cp.async gather for layouts whose second range is not 8-row aligned), key-mask words of the sequence
//   warp 1      one elected thread issues every tcgen05.mma (polls both groups' barriers with mbarrier.test_wait)
//   warps 2-9 / 10-17  softmax groups: two threads per query row, 96 keys of the block each
//
// Replaces: reference vlmo.py:79-95 at the VQA resolutions.
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
constexpr int kQTile = 128;
constexpr int kKB = 192;          // keys per block
constexpr int kMaxKeys = 1024;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

constexpr int kQBytes = 2 * kQTile * 128;        // the pair's two query tiles
constexpr int kKVBytes = kKB * 128;               // one operand of a key block
constexpr int kKVStages = 3;
constexpr int kQOff = 0;                           // 2 item parities x 32 KB
constexpr int kKVOff = 2 * kQBytes;                // 3 stages x (K | V)
constexpr int kMetaOff = kKVOff + kKVStages * 2 * kKVBytes;  // 4 slots x 192 B: keep words [32], seq desc [4]
constexpr int kMetaBytes = 192, kMetaSlots = 4;  // slot k & 3: the producer is never three items ahead of the groups (2 query buffers)
constexpr int kBarOff = kMetaOff + kMetaSlots * kMetaBytes;
constexpr int kXchgOff = kBarOff + 256;            // row max / row sum exchange between the two threads of a row: 2 x 2 KB
constexpr int kSmemBytes = kXchgOff + 4096 + 1024;
constexpr int kThreads = 64 + 2 * 256;
constexpr uint32_t kSlotCols = 256, kOCol = 192, kPHiCol = 96;

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
__device__ __forceinline__ int pad16(int x) { return (x + 15) & ~15; }
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
    if (ok) return;
  }
  __trap();
}

struct Params {
  CUtensorMap map32, map8;  // qkv as [tokens][3 d] bf16, boxes of 64 columns x 32 / 8 rows
  const __nv_bfloat16* qkv;
  const int32_t* seq_desc;
  const uint8_t* key_mask;
  __nv_bfloat16* out;
  float* lse;
  int H, max_seq_len, num_items, QP;
  float scale;
  const uint32_t* drop_seed;
  uint32_t drop_salt, drop_thr;
};

struct Geo {
  int sh, qp, n, nqt, nkv;  // (sequence, head), query pair, length, query tiles of the pair (0: skip), key blocks
};
__device__ __forceinline__ Geo item_geo(const Params& p, int item) {
  Geo g;
  g.sh = item / p.QP;
  g.qp = item - g.sh * p.QP;
  const int4 v = __ldg(reinterpret_cast<const int4*>(p.seq_desc + 4 * (g.sh / p.H)));
  g.n = v.y + v.w;
  g.nqt = max(0, min(2, (g.n + kQTile - 1) / kQTile - 2 * g.qp));
  g.nkv = (g.n + kKB - 1) / kKB;
  return g;
}

// sequence-local rows [r0, r1) of one operand -> TMA boxes (32-row boxes, then 8-row boxes; a range's last box may run past
// it into finite rows nobody reads unmasked). Returns the bytes the boxes carry.
__device__ __forceinline__ uint32_t span_boxes(const Params& p, uint8_t* dst, uint64_t* bar, int col, const Seq& sd, int r0, int r1, bool issue) {
  uint32_t bytes = 0;
#pragma unroll 1
  for (int part = 0; part < 2; ++part) {
    const int a = part == 0 ? r0 : max(r0, sd.len0);
    const int b = part == 0 ? min(r1, sd.len0) : r1;
    if (b <= a) continue;
    const int grow = part == 0 ? sd.start0 + a : sd.start1 + (a - sd.len0);
    const int len = b - a, nb32 = len >> 5, nb8 = ((len & 31) + 7) >> 3;
    uint8_t* d = dst + (a - r0) * 128;
    if (issue) {
      for (int x = 0; x < nb32; ++x) tma_load_2d(d + x * 4096, &p.map32, bar, col, grow + 32 * x);
      for (int x = 0; x < nb8; ++x) tma_load_2d(d + nb32 * 4096 + x * 1024, &p.map8, bar, col, grow + 32 * nb32 + 8 * x);
    }
    bytes += nb32 * 4096u + nb8 * 1024u;
  }
  return bytes;
}

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 1) attn_fwd_tc_long_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* q_full = bars;          // [2] producer -> MMA / groups: the item's query tiles and meta are there (2 arrivals + tx)
  uint64_t* q_empty = bars + 2;     // [2] MMA -> producer: every S MMA of the item has completed
  uint64_t* kv_full = bars + 4;     // [3] producer -> MMA
  uint64_t* kv_empty = bars + 7;    // [3] MMA -> producer: every MMA reading the stage has completed
  uint64_t* s_full = bars + 10;     // [2] MMA -> group: S_j is in the slot
  uint64_t* p_full = bars + 12;     // [2] group -> MMA: P_j written, O rescaled (256 arrivals)
  uint64_t* o_full = bars + 14;     // [2] MMA -> group: O += P_j V_j has completed
  uint64_t* o_free = bars + 16;     // [2] group -> MMA: the item's O has been read (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, d = H * kHd;
  const int G = gridDim.x;

  // stale shared memory must at least be finite: rows past the sequence end multiply p = 0
  for (int i = threadIdx.x; i < kMetaOff / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 2);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 256);
    }
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.map32);
    tma_prefetch_desc(&p.map8);
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
    int k = 0;          // non-empty items so far
    uint32_t nkv = 0;   // key blocks loaded so far: stage = nkv % 3, fill number = nkv / 3
    for (int item = blockIdx.x; item < p.num_items; item += G) {
      const Geo g = item_geo(p, item);
      if (g.nqt == 0) continue;
      const int s = g.sh / H, h = g.sh - s * H;
      const Seq sd = load_seq(p.seq_desc, s);
      const int n = g.n;
      const bool boxes_ok = sd.len1 == 0 || (sd.len0 & 7) == 0;
      const int par = k & 1;
      // gather path for one operand span: 4 rows per pass, zero fill up to the padded end
      auto gather = [&](uint8_t* dst, int op, int r0, int r1) {
        const int rows_pad = pad16(r1 - r0);
        const int rr = lane >> 3, ch = lane & 7;
        const __nv_bfloat16* base = p.qkv + op * d + h * kHd + ch * 8;
        for (int r = rr; r < rows_pad; r += 4) {
          const bool valid = r0 + r < n;
          cp_async_16(dst + r * 128 + ((ch ^ (r & 7)) << 4), base + seq_row(sd, valid ? r0 + r : 0) * (3LL * d), valid);
        }
      };
      // ---- query tiles of the pair + key words of the whole sequence
      wait_bar(&q_empty[par], ((k >> 1) & 1) ^ 1);
      __syncwarp();
      {
        uint8_t* qdst = smem + kQOff + par * kQBytes;
        const int r0 = 2 * g.qp * kQTile, r1 = min(n, r0 + 2 * kQTile);
        if (boxes_ok) {
          if (elect_one()) {
            mbar_arrive_expect_tx(&q_full[par], span_boxes(p, qdst, &q_full[par], h * kHd, sd, r0, r1, false));
            span_boxes(p, qdst, &q_full[par], h * kHd, sd, r0, r1, true);
          }
          __syncwarp();
        }
        uint32_t* meta = reinterpret_cast<uint32_t*>(smem + kMetaOff + (k & (kMetaSlots - 1)) * kMetaBytes);
#pragma unroll 1
        for (int j8 = 0; j8 < (n + 31) / 32; ++j8) {
          const int j = j8 * 32 + lane;
          const bool keep = j < n && (p.key_mask == nullptr || __ldg(p.key_mask + seq_row(sd, j)) != 0);
          const uint32_t w = __ballot_sync(0xffffffffu, keep);
          if (lane == 0) meta[j8] = w;
        }
        if (lane == 0) {
          for (int j8 = (n + 31) / 32; j8 < 32; ++j8) meta[j8] = 0u;
          meta[32] = sd.start0; meta[33] = sd.len0; meta[34] = sd.start1; meta[35] = sd.len1;
          mbar_arrive(&q_full[par]);
        }
        if (!boxes_ok) {
          gather(qdst, 0, r0, r1);
          cp_async_commit();
          cp_async_wait<0>();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&q_full[par]);
        }
      }
      ++k;
      // ---- key blocks
      for (int j = 0; j < g.nkv; ++j, ++nkv) {
        const int stage = nkv % kKVStages;
        uint8_t* kdst = smem + kKVOff + stage * 2 * kKVBytes;
        const int r0 = j * kKB, r1 = min(n, r0 + kKB);
        wait_bar(&kv_empty[stage], ((nkv / kKVStages) & 1) ^ 1);
        __syncwarp();
        if (boxes_ok) {
          if (elect_one()) {
            mbar_arrive_expect_tx(&kv_full[stage], 2u * span_boxes(p, kdst, &kv_full[stage], d + h * kHd, sd, r0, r1, false));
            span_boxes(p, kdst, &kv_full[stage], d + h * kHd, sd, r0, r1, true);
            span_boxes(p, kdst + kKVBytes, &kv_full[stage], 2 * d + h * kHd, sd, r0, r1, true);
          }
          __syncwarp();
        } else {
          gather(kdst, 1, r0, r1);
          gather(kdst + kKVBytes, 2, r0, r1);
          cp_async_commit();
          cp_async_wait<0>();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&kv_full[stage]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------ MMA issuer
    // Per group g a cursor over (item, key block, phase): S_j needs the item's Q, the block's K and — for j = 0 — the group to
    // have read the previous item's O; PV_j needs the group's P_j. The tensor pipe runs in issue order, so S_(j+1) may follow
    // PV_j at once (it overwrites the P that PV_j reads).
    if (elect_one()) {
      const uint32_t sQ = smem_u32(smem + kQOff), sKV = smem_u32(smem + kKVOff);
      struct Cur {
        int item, k, n, nqt, nkv, j, pv;  // pv: 0 = S_j is next, 1 = PV_j is next
        uint32_t kv0;                      // key blocks of the items before this one
      };
      auto seek = [&](Cur& c, int item) {
        for (; item < p.num_items; item += G) {
          const Geo g = item_geo(p, item);
          if (g.nqt > 0) {
            c.item = item; c.n = g.n; c.nqt = g.nqt; c.nkv = g.nkv; c.j = 0; c.pv = 0;
            return;
          }
        }
        c.item = item; c.n = 0; c.nqt = 0; c.nkv = 0;
      };
      auto next_item = [&](Cur& c) {
        c.kv0 += c.nkv;
        ++c.k;
        seek(c, c.item + G);
      };
      Cur cur[2];
      cur[0].k = 0; cur[0].kv0 = 0;
      seek(cur[0], blockIdx.x);
      cur[1] = cur[0];
      uint32_t nS[2] = {0, 0};        // S MMAs issued per group = phase of s_full / p_full
      uint32_t items_done[2] = {0, 0};  // items finished per group = phase of o_free
      // stage / item bookkeeping: a stage is released when both groups (or the only active one) have issued PV_j,
      // a query buffer when both have issued their last S
      auto release_after = [&](const Cur& c, int g, bool is_pv) {
        const bool both = c.nqt == 2;
        // the other group's cursor is past this point iff it is on a later item, or on this item at a later (j, phase)
        const Cur& o = cur[g ^ 1];
        auto other_passed = [&](int j, int pv) {
          if (!both) return true;
          if (o.n == 0 || o.item != c.item) return o.n == 0 || o.k > c.k;
          return o.j > j || (o.j == j && o.pv > pv) ;
        };
        if (is_pv) {
          // PV_j of this group just issued; release the stage if the other group has issued its PV_j too
          if (other_passed(c.j, 1)) umma_commit(&kv_empty[(c.kv0 + c.j) % kKVStages]);
        } else if (c.j == c.nkv - 1) {
          if (other_passed(c.j, 0)) umma_commit(&q_empty[c.k & 1]);
        }
      };
      while (cur[0].n > 0 || cur[1].n > 0) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          Cur& c = cur[g];
          if (c.n == 0) continue;
          if (g >= c.nqt) {
            // group 1 has no tile in this item: its cursor moves on once group 0's cursor has left the item, so that every
            // barrier phase it tests later has a completed predecessor (phase parities alias two fills apart)
            if (cur[0].n == 0 || cur[0].k > c.k) next_item(c);
            continue;
          }
          const uint32_t stage = (c.kv0 + c.j) % kKVStages, fill = (c.kv0 + c.j) / kKVStages;
          const int nk = min(kKB, pad16(c.n - c.j * kKB));
          const uint32_t slot = tmem_base + g * kSlotCols;
          const uint32_t st = sKV + stage * 2 * kKVBytes;
          if (c.pv == 0) {
            if (mbar_test_wait(&q_full[c.k & 1], (c.k >> 1) & 1) && mbar_test_wait(&kv_full[stage], fill & 1) &&
                (c.j > 0 || mbar_test_wait(&o_free[g], (items_done[g] & 1) ^ 1))) {
              tcgen05_fence_after();
              const uint32_t idesc = umma_idesc_bf16(kQTile, nk, false, false);
              const uint32_t qa = sQ + (c.k & 1) * kQBytes + g * (kQTile * 128);
#pragma unroll
              for (int x = 0; x < kHd / 16; ++x)
                umma_bf16(slot, umma_smem_desc(qa + x * 32, 0, 1024), umma_smem_desc(st + x * 32, 0, 1024), idesc, x > 0 ? 1u : 0u);
              umma_commit(&s_full[g]);
              release_after(c, g, false);
              c.pv = 1;
            }
          } else if (mbar_test_wait(&p_full[g], nS[g] & 1)) {
            tcgen05_fence_after();
            const uint32_t idesc = umma_idesc_bf16(kQTile, kHd, false, true);
            uint64_t vdesc = umma_smem_desc(st + kKVBytes, 8192, 1024);
            const int ksteps = nk >> 4;
            // P of keys [0, 96) sits in columns [0, 48), of keys [96, 192) in [96, 144); V advances 16 rows per step
            for (int x = 0; x < ksteps; ++x, vdesc += 2048 >> 4)
              umma_bf16_ts(slot + kOCol, slot + (x < 6 ? x * 8 : kPHiCol + (x - 6) * 8), vdesc, idesc, (c.j > 0 || x > 0) ? 1u : 0u);
            umma_commit(&o_full[g]);
            release_after(c, g, true);
            ++nS[g];
            c.pv = 0;
            if (++c.j == c.nkv) {
              ++items_done[g];
              next_item(c);
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------ softmax groups
    const int wi = warp - 2, g = wi >> 3, half = (wi >> 2) & 1, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const float sl2 = p.scale * kLog2e;
    const uint32_t dkey = DROP ? drop_mix(p.drop_salt, __ldg(p.drop_seed)) : 0u;
    const float dscale = drop_scale(p.drop_thr);
    const uint32_t trow = tmem_base + g * kSlotCols + (static_cast<uint32_t>(quarter * 32) << 16);
    float* xmax = reinterpret_cast<float*>(smem + kXchgOff) + (g * 2 + half) * 128;      // this thread's slot; partner: ^ 128
    float* xsum = reinterpret_cast<float*>(smem + kXchgOff + 2048) + (g * 2 + half) * 128;
    const int pair_bar = 1 + g * 4 + quarter;  // named barrier of the two warps sharing these 32 rows
    uint32_t nS = 0, nO = 0;  // blocks processed by this group (phase of s_full / p_full), PV completions waited for (o_full)
    int k = 0;
    for (int item = blockIdx.x; item < p.num_items; item += G) {
      const Geo geo = item_geo(p, item);
      if (geo.nqt == 0) continue;
      const int par = k & 1;
      const uint32_t kph = (k >> 1) & 1;
      const uint32_t meta = smem_u32(smem + kMetaOff + (k & (kMetaSlots - 1)) * kMetaBytes);
      ++k;
      if (g >= geo.nqt) {
        // Odd last pair (or a one-tile sequence): group 1 has no tile, but it still takes the item's barrier phase. A group
        // that ran ahead through a stretch of such items (the text sequences of a split layout) would otherwise test a later
        // phase of the same parity before the producer got there, and read a stale descriptor.
        wait_bar(&q_full[par], kph);
        continue;
      }
      const int s = geo.sh / H, h = geo.sh - s * H;
      const int n = geo.n;
      const int q = (2 * geo.qp + g) * kQTile + row;
      const bool warp_live = (2 * geo.qp + g) * kQTile + quarter * 32 < n;
      wait_bar(&q_full[par], kph);  // orders the producer's meta writes before the reads below
      const Seq sd{static_cast<int>(lds_u32(meta + 128)), static_cast<int>(lds_u32(meta + 132)), static_cast<int>(lds_u32(meta + 136)),
                   static_cast<int>(lds_u32(meta + 140))};
      const uint32_t drow = attn_drop_row(s, H, h, p.max_seq_len, q);
      float m_run = -INFINITY, l = 0.f;  // running maximum (raw score units) and this thread's share of the running sum
      for (int j = 0; j < geo.nkv; ++j, ++nS) {
        const int nk = min(kKB, pad16(n - j * kKB));
        const int nsub = nk >> 4;                                       // 16-key sub-chunks of the block
        const int c0 = half * 6, c1 = half == 0 ? min(nsub, 6) : nsub;   // this thread's sub-chunks [c0, c1)
        const int key_base = j * kKB;                                    // sequence-local key of the block's first column
        // keep bits of 16 keys starting at sequence-local key kb (multiple of 16)
        auto keep16 = [&](int kb) { return (lds_u32(meta + 4 * (kb >> 5)) >> (kb & 16)) & 0xffffu; };
        wait_bar(&s_full[g], nS & 1);
        __syncwarp();
        tcgen05_fence_after();
        if (warp_live) {
          uint32_t ra[16];
          // ---- pass 1: block maximum over the keys that take part
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
          for (int c = c0; c < c1; ++c) {
            tmem_ld_32x16(trow + c * 16, ra);
            tmem_ld_wait();
            const uint32_t kw = keep16(key_base + c * 16);
            if (kw == 0xffffu) {
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                m0 = fmaxf(m0, __uint_as_float(ra[e]));
                m1 = fmaxf(m1, __uint_as_float(ra[e + 1]));
                m2 = fmaxf(m2, __uint_as_float(ra[e + 2]));
                m3 = fmaxf(m3, __uint_as_float(ra[e + 3]));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                m0 = fmaxf(m0, ((kw >> e) & 1u) ? __uint_as_float(ra[e]) : -INFINITY);
                m1 = fmaxf(m1, ((kw >> (e + 1)) & 1u) ? __uint_as_float(ra[e + 1]) : -INFINITY);
                m2 = fmaxf(m2, ((kw >> (e + 2)) & 1u) ? __uint_as_float(ra[e + 2]) : -INFINITY);
                m3 = fmaxf(m3, ((kw >> (e + 3)) & 1u) ? __uint_as_float(ra[e + 3]) : -INFINITY);
              }
            }
          }
          const float mine = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          xmax[row] = mine;
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          const float m_new = fmaxf(m_run, fmaxf(mine, xmax[row + (half == 0 ? 128 : -128)]));
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");  // xmax is rewritten in the next block
          // ---- the running maximum grew: rescale what has been accumulated (both threads of a row see the same maxima)
          if (j > 0) {
            wait_bar(&o_full[g], nO & 1);  // PV_(j-1) has completed
            __syncwarp();
            tcgen05_fence_after();
            const float alpha = m_new > m_run ? (m_run == -INFINITY ? 0.f : ex2_approx((m_run - m_new) * sl2)) : 1.f;
            if (__any_sync(0xffffffffu, alpha != 1.f)) {
              uint32_t o[32];
              tmem_ld_32x32(trow + kOCol + half * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st_32x16(trow + kOCol + half * 32, reinterpret_cast<const uint32_t(&)[16]>(o[0]));
              tmem_st_32x16(trow + kOCol + half * 32 + 16, reinterpret_cast<const uint32_t(&)[16]>(o[16]));
            }
            l *= alpha;
          }
          m_run = m_new;
          const float ms = m_run == -INFINITY ? 0.f : m_run * sl2;  // every key masked so far: probabilities stay 0
          // ---- pass 2: probabilities, running sum, dropout, bf16 P over this thread's part of S in place
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll 1
          for (int c = c0; c < c1; ++c) {
            tmem_ld_32x16(trow + c * 16, ra);
            tmem_ld_wait();
            const uint32_t kw = keep16(key_base + c * 16);
            if (kw == 0xffffu) {
#pragma unroll
              for (int e = 0; e < 16; ++e) ra[e] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(ra[e]), sl2, -ms)));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                ra[e] = ((kw >> e) & 1u) ? __float_as_uint(ex2_approx(fmaf(__uint_as_float(ra[e]), sl2, -ms))) : 0u;
            }
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              l0 += __uint_as_float(ra[e]); l1 += __uint_as_float(ra[e + 1]); l2 += __uint_as_float(ra[e + 2]); l3 += __uint_as_float(ra[e + 3]);
            }
            uint32_t pk[8];
            if (DROP) {  // the normaliser keeps the undropped probabilities; only what multiplies V is dropped
              const uint32_t dw0 = drow + ((key_base + c * 16) >> 1);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t word = drop_mix(dw0 + e, dkey);
                const float a = (word & 255u) >= p.drop_thr ? __uint_as_float(ra[2 * e]) * dscale : 0.f;
                const float b = ((word >> 16) & 255u) >= p.drop_thr ? __uint_as_float(ra[2 * e + 1]) * dscale : 0.f;
                pk[e] = pack_bf16(a, b);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) pk[e] = pack_bf16(__uint_as_float(ra[2 * e]), __uint_as_float(ra[2 * e + 1]));
            }
            tmem_st_32x8(trow + half * kPHiCol + (c - c0) * 8, pk);
          }
          l += (l0 + l1) + (l2 + l3);
          tmem_st_wait();
        } else if (j > 0) {
          wait_bar(&o_full[g], nO & 1);
        }
        if (j > 0) ++nO;
        tcgen05_fence_before();
        mbar_arrive(&p_full[g]);
      }
      // ---- epilogue: O / l, log-sum-exp
      wait_bar(&o_full[g], nO & 1);
      ++nO;
      __syncwarp();
      tcgen05_fence_after();
      if (warp_live) {
        uint32_t a[32];
        tmem_ld_32x32(trow + kOCol + half * 32, a);
        xsum[row] = l;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        l += xsum[row + (half == 0 ? 128 : -128)];
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        tmem_ld_wait();
        tcgen05_fence_before();
        mbar_arrive(&o_free[g]);  // O is in registers: the slot can take the next item
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
          if (half == 0) {
            const float ms = m_run == -INFINITY ? 0.f : m_run * sl2;
            p.lse[static_cast<long long>(geo.sh) * p.max_seq_len + q] = l > 0.f ? (ms + log2f(l)) * kLn2 : -INFINITY;
          }
        }
      } else {
        tcgen05_fence_before();
        mbar_arrive(&o_free[g]);
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

int attn_fwd_tc_long_max_seq_len() { return kMaxKeys; }

int attn_fwd_tc_long(const void* qkv, const int32_t* seq_desc, const uint8_t* key_mask, void* out, float* lse, int64_t tokens, int num_seqs,
                     int max_seq_len, int H, float scale, const uint32_t* drop_seed, uint32_t drop_salt, float drop_p, cudaStream_t stream) {
  MOME_REQUIRE(max_seq_len <= kMaxKeys, "attn_fwd_tc_long: max_seq_len %d > %d", max_seq_len, kMaxKeys);
  static bool configured = false;
  if (!configured) {
    int rc = opt_in(attn_fwd_tc_long_kernel<false>, kSmemBytes, "attn_fwd_tc_long");
    if (rc == MOME_OK) rc = opt_in(attn_fwd_tc_long_kernel<true>, kSmemBytes, "attn_fwd_tc_long");
    if (rc != MOME_OK) return rc;
    configured = true;
  }
  Params p;
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
  p.QP = ((max_seq_len + kQTile - 1) / kQTile + 1) / 2;
  p.num_items = num_seqs * H * p.QP;
  p.scale = scale;
  p.drop_seed = drop_seed;
  p.drop_salt = drop_salt;
  p.drop_thr = drop_threshold(drop_p);
  const int grid = std::min(p.num_items, sm_count());
  if (drop_seed != nullptr && drop_p > 0.f)
    attn_fwd_tc_long_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(p);
  else
    attn_fwd_tc_long_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(p);
  return check_launch("attn_fwd_tc_long");
}

}  // namespace mome
