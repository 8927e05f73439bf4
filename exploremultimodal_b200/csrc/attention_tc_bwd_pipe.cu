// K1 backward on tcgen05 / TMEM / TMA, software-pipelined (round 2). Same contract and same arithmetic as
// attention_tc_bwd.cu (mask semantics, lse format, dropout mask function, dqkv layout); what changes is the schedule,
// and sequences may be longer than 256 tokens (VQA at 480 px: 941).
//
// ncu of the first kernel (profiles/r02_attn_bwd_ncu.md): the P / dS warps spend 22 % of their time on the P / dS
// arithmetic and the rest waiting — for S / dP of the block (the tensor core only starts them once the previous block's
// accumulate MMAs were issued), for dK / dV / dQ to become final so that they can drain them themselves, for the next
// item's TMA loads (issued after the last MMA of the item), for lse / delta loads at the start of an item. Here:
//
//   * an item is (sequence, head, PAIR of 128-query tiles); it visits every 128-key block of the sequence, key block by key
//     block, so dQ_0 / dQ_1 accumulate in TMEM over the whole item and dK_j / dV_j over the item's two query tiles;
//   * S / dP are computed per SUB-BLOCK of 128 queries x 64 keys into a two-deep ring of TMEM buffers
//     (2 x (64 + 64) columns); the P / dS warps copy a sub-block into registers and hand the buffer back at once,
//     so the tensor core is one to two sub-blocks ahead of them;
//   * the accumulate MMAs (dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j) stay per 128 x 128 block, M = 128 keys;
//   * dK / dV / dQ are drained by four extra warps, not by the P / dS warps. A sequence with one query pair (<= 256
//     tokens) stores dK_j / dV_j directly; longer ones add them (fp32 red.add) into a scratch [tokens, 2 d] that a small
//     kernel converts afterwards — the only atomics, and only on the long-sequence path;
//   * Q, dO live in a ring of four 128-row tiles, K, V in another ring of four, a full / empty barrier pair per tile. A tile
//     is released by the last MMA that reads it, so the next key block (or K_0 / V_0 of the next item) arrives while this
//     one is processed, and two 1-tile (text) items are in flight at once;
//   * a thread fetches the next item's lse / delta while it works on this one.
//
// What bounds it now (clock64 event trace of a debugging build, profiles/r02_attn_bwd_ncu.md, + tools/umma_probe.cu): with head_dim 64 every MMA has N = 64, i.e. 6 KB of
// shared-memory operands for 32 tensor-pipe cycles; operand fetch (~64 B/clk) and the issuing thread (~90 cycles per
// tcgen05.mma) both sit at ~95 cycles per MMA, 40 MMAs per 128 x 128 block = ~3800 cycles, next to ~3600 cycles of
// P / dS work per block. Three issuing threads and register-deferred tile stores were tried and measured no faster.
//
// TMEM (512 columns): dQ_0 | dQ_1 | dK_j | dV_j (64 each) | S_0 dP_0 | S_1 dP_1 (64 each).
// Shared memory: 8 operand tiles (128 KB), P and dS tiles (2 x 32 KB, two 64-key chunks each), 4 meta slots, barriers.
//
//   warp 0       producer (TMA boxes; cp.async gather for layouts whose second range is not 8-row aligned)
//   warp 1       one elected thread issues every tcgen05.mma (event loop over two cursors: S / dP and accumulate)
//   warps 2-9    P / dS group: thread = (query row, 32 keys of the sub-block)
//   warps 10-13  drain group: dK_j / dV_j after every key block, dQ_i at the end of the item (x scale) into dqkv
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
int attn_delta_launch(const void* out, const void* dout, const int32_t* seq_desc, float* delta, int H, int max_seq_len, int num_seqs,
                      cudaStream_t stream);

namespace {

constexpr int kHd = 64;
constexpr int kTile = 128;  // query tile = key block = MMA M = TMEM lanes
constexpr int kSub = 64;    // keys of a sub-block (S / dP granularity)
constexpr int kMaxKeys = 1024;
constexpr int kMaxKeyBlocks = kMaxKeys / kTile;
constexpr float kLog2e = 1.4426950408889634f;

constexpr int kTileBytes = kTile * 128;  // 128 rows x 64 bf16
constexpr int kSlots = 8;                // 0-3: Q / dO ring, 4-7: K / V ring
constexpr int kPOff = kSlots * kTileBytes, kSOff = kPOff + 2 * kTileBytes;  // P, dS: 2 chunks of 64 keys x [128 q x 128 B]
constexpr int kMetaOff = kSOff + 2 * kTileBytes;                              // 4 slots: keep words [32], seq desc [4]
constexpr int kMetaSlots = 4, kMetaBytes = 192;
constexpr int kBarOff = kMetaOff + kMetaSlots * kMetaBytes;
constexpr int kSmemBytes = kBarOff + 512 + 1024;
constexpr int kGroupWarps = 8, kDrainWarps = 4;
constexpr int kThreads = 64 + 32 * (kGroupWarps + kDrainWarps);
constexpr uint32_t kColDQ = 0, kColDK = 128, kColDV = 192, kColSdp = 256;  // S_b at kColSdp + 128 b, dP_b 64 columns further

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
// Bounded wait with a suspend-time hint, no printf (keeps the stack frame and argument registers out of the hot loops):
// a protocol bug traps (launch failure) instead of hanging the GPU.
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
__device__ __forceinline__ void red_add_v4(float* gptr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct Params {
  CUtensorMap qkv32, qkv8;  // qkv as [tokens][3 d] bf16, boxes of 64 columns x 32 / 8 rows
  CUtensorMap go32, go8;    // dout as [tokens][d]
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* dout;
  const int32_t* seq_desc;
  const uint8_t* key_mask;
  const float* lse;
  const float* delta;
  __nv_bfloat16* dqkv;
  float* dkv_acc;  // [tokens][2 d] fp32, zeroed: dK / dV of sequences with several query pairs (nullptr when QP == 1)
  int H, max_seq_len, num_items, QP;  // QP: query pairs per (sequence, head) = items per (sequence, head)
  float scale;
  const uint32_t* drop_seed;
  uint32_t drop_salt, drop_thr;
};

// An item = (sequence, head, query pair). Geometry from the sequence length: query tiles of the pair (0: the pair lies past
// the end of this sequence — every role skips the item), key blocks of the sequence.
struct Geo {
  int sh, qp, n, nqt, nkb;
};
__device__ __forceinline__ Geo item_geo(const Params& p, int item, int n) {
  Geo g;
  g.sh = item / p.QP;
  g.qp = item - g.sh * p.QP;
  g.n = n;
  g.nkb = (n + kTile - 1) / kTile;
  g.nqt = max(0, min(2, g.nkb - 2 * g.qp));
  return g;
}
__device__ __forceinline__ int item_len(const Params& p, int item) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(p.seq_desc + 4 * ((item / p.QP) / p.H)));
  return v.y + v.w;
}

struct Bars {
  uint64_t* full;       // [8] producer -> MMA: tile loaded
  uint64_t* empty;      // [8] MMA -> producer: every MMA reading the tile has completed
  uint64_t* meta_full;  // [4] producer -> groups: key words and descriptor of the item written
  uint64_t* sdp_full;   // [2] MMA -> P / dS group: S and dP of the sub-block are in TMEM buffer b
  uint64_t* sdp_free;   // [2] P / dS group -> MMA: buffer b copied to registers (256 arrivals)
  uint64_t* pds_full;   // P / dS group -> MMA: P and dS tiles of the block written (256 arrivals)
  uint64_t* pds_free;   // MMA -> P / dS group: the MMAs reading the tiles have completed
  uint64_t* dkv_full;   // MMA -> drain: dK_j, dV_j final
  uint64_t* dkv_free;   // drain -> MMA: copied to registers (128 arrivals)
  uint64_t* dq_full;    // MMA -> drain: dQ_0, dQ_1 final
  uint64_t* dq_free;    // drain -> MMA (128 arrivals)
};
__device__ __forceinline__ Bars make_bars(uint8_t* smem) {
  uint64_t* b = reinterpret_cast<uint64_t*>(smem + kBarOff);
  return Bars{b, b + 8, b + 16, b + 20, b + 22, b + 24, b + 25, b + 26, b + 27, b + 28, b + 29};
}
// ring positions: the Q / dO ring (slots 0-3) and the K / V ring (slots 4-7) count their tiles separately
__device__ __forceinline__ uint32_t q_slot(uint32_t cq) { return cq & 3; }
__device__ __forceinline__ uint32_t k_slot(uint32_t ck) { return 4 + (ck & 3); }

// one tile of an operand: sequence-local rows [r0, r1) -> TMA boxes (32-row boxes, then 8-row boxes; a range's last box may
// run past it into finite rows nobody reads unmasked). Returns the bytes the boxes carry.
__device__ __forceinline__ uint32_t tile_boxes(const CUtensorMap* m32, const CUtensorMap* m8, uint8_t* dst, uint64_t* bar, int col, const Seq& sd,
                                               int r0, int r1, bool issue) {
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
      for (int x = 0; x < nb32; ++x) tma_load_2d(d + x * 4096, m32, bar, col, grow + 32 * x);
      for (int x = 0; x < nb8; ++x) tma_load_2d(d + nb32 * 4096 + x * 1024, m8, bar, col, grow + 32 * nb32 + 8 * x);
    }
    bytes += nb32 * 4096u + nb8 * 1024u;
  }
  return bytes;
}

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_tc_pipe_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const Bars B = make_bars(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kBarOff + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, d = H * kHd;
  const int G = gridDim.x;

  // stale shared memory must at least be finite: rows past the sequence end multiply exact zeros
  for (int i = threadIdx.x; i < kMetaOff / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&B.full[i], 1);
      mbar_init(&B.empty[i], 1);
    }
    for (int i = 0; i < kMetaSlots; ++i) mbar_init(&B.meta_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&B.sdp_full[i], 1);
      mbar_init(&B.sdp_free[i], 32 * kGroupWarps);
    }
    mbar_init(B.pds_full, 32 * kGroupWarps);
    mbar_init(B.pds_free, 1);
    mbar_init(B.dkv_full, 1);
    mbar_init(B.dkv_free, 32 * kDrainWarps);
    mbar_init(B.dq_full, 1);
    mbar_init(B.dq_free, 32 * kDrainWarps);
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
    uint32_t cq = 0, ck = 0;  // tiles loaded so far into the Q / dO ring and the K / V ring (fill number = count >> 2)
    int k = 0;                // non-empty items so far
    for (int item = blockIdx.x; item < p.num_items; item += G) {
      const Geo g = item_geo(p, item, item_len(p, item));
      if (g.nqt == 0) continue;
      const int s = g.sh / H, h = g.sh - s * H;
      const Seq sd = load_seq(p.seq_desc, s);
      const int n = g.n;
      // the meta slot's previous user is item k - 4: the K / V ring slot waited for here was released by item k - 1 or
      // k - 2, and the groups finished item k - 4 long before that
      wait_bar(&B.empty[k_slot(ck)], ((ck >> 2) & 1) ^ 1);
      uint32_t* meta = reinterpret_cast<uint32_t*>(smem + kMetaOff + (k & (kMetaSlots - 1)) * kMetaBytes);
#pragma unroll 1
      for (int j8 = 0; j8 < g.nkb * 4; ++j8) {
        const int j = j8 * 32 + lane;
        const bool keep = j < n && (p.key_mask == nullptr || __ldg(p.key_mask + seq_row(sd, j)) != 0);
        const uint32_t w = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) meta[j8] = w;
      }
      if (lane == 0) {
        meta[32] = sd.start0; meta[33] = sd.len0; meta[34] = sd.start1; meta[35] = sd.len1;
        mbar_arrive(&B.meta_full[k & (kMetaSlots - 1)]);
      }
      ++k;
      const bool boxes_ok = sd.len1 == 0 || (sd.len0 & 7) == 0;
      // loading order = order of first use: K_0 V_0 | Q_0 dO_0 [Q_1 dO_1] | K_1 V_1 | K_2 V_2 ...
      const int ntiles = 2 * g.nqt + 2 * g.nkb;
#pragma unroll 1
      for (int t = 0; t < ntiles; ++t) {
        int op, rt;  // operand 0 Q, 1 K, 2 V, 3 dO; 128-row tile of the sequence
        if (t < 2) {
          op = 1 + t; rt = 0;
        } else if (t < 2 + 2 * g.nqt) {
          op = (t & 1) ? 3 : 0; rt = 2 * g.qp + ((t - 2) >> 1);
        } else {
          op = 1 + (t & 1); rt = 1 + ((t - 2 - 2 * g.nqt) >> 1);
        }
        const bool is_q = op == 0 || op == 3;
        const uint32_t cnt = is_q ? cq : ck;
        const uint32_t slot = is_q ? q_slot(cnt) : k_slot(cnt);
        if (is_q) ++cq; else ++ck;
        uint8_t* dst = smem + slot * kTileBytes;
        const int r0 = rt * kTile, r1 = min(n, r0 + kTile);
        wait_bar(&B.empty[slot], ((cnt >> 2) & 1) ^ 1);
        __syncwarp();
        if (boxes_ok) {
          if (elect_one()) {
            const CUtensorMap* m32 = op < 3 ? &p.qkv32 : &p.go32;
            const CUtensorMap* m8 = op < 3 ? &p.qkv8 : &p.go8;
            const int col = (op < 3 ? op * d : 0) + h * kHd;
            const uint32_t bytes = tile_boxes(m32, m8, dst, &B.full[slot], col, sd, r0, r1, false);
            mbar_arrive_expect_tx(&B.full[slot], bytes);
            tile_boxes(m32, m8, dst, &B.full[slot], col, sd, r0, r1, true);
          }
          __syncwarp();
        } else {
          // gather path: 4 rows per pass, zero fill up to the padded end of the tile's rows
          const int rows_pad = pad16(r1 - r0);
          const int rr = lane >> 3, ch = lane & 7;
          const long long ld = op < 3 ? 3LL * d : d;
          const __nv_bfloat16* base = (op < 3 ? p.qkv + op * d : p.dout) + h * kHd + ch * 8;
          for (int r = rr; r < rows_pad; r += 4) {
            const bool valid = r0 + r < n;
            const long long grow = seq_row(sd, valid ? r0 + r : 0);
            cp_async_16(dst + r * 128 + ((ch ^ (r & 7)) << 4), base + grow * ld, valid);
          }
          cp_async_commit();
          cp_async_wait<0>();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&B.full[slot]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t sP = smem_u32(smem + kPOff), sS = smem_u32(smem + kSOff), sR = smem_u32(smem);
      struct Cur {
        int item, k, n, nqt, nkb, j, i, c;
        uint32_t cq, ck;  // ring counts at the start of the item
      };
      // move to the next non-empty item at or after `item` (n = 0: past the end)
      auto seek = [&](Cur& c, int item) {
        for (; item < p.num_items; item += G) {
          const Geo g = item_geo(p, item, item_len(p, item));
          if (g.nqt > 0) {
            c.item = item; c.n = g.n; c.nqt = g.nqt; c.nkb = g.nkb;
            c.j = c.i = c.c = 0;
            return;
          }
        }
        c.item = item; c.n = 0; c.nqt = c.nkb = 0;
      };
      auto next_item = [&](Cur& c) {
        c.cq += 2 * c.nqt;
        c.ck += 2 * c.nkb;
        ++c.k;
        seek(c, c.item + G);
      };
      auto keys_of = [](const Cur& c, int j) { return min(kTile, pad16(c.n - j * kTile)); };
      auto q_addr = [&](uint32_t cnt) { return sR + q_slot(cnt) * kTileBytes; };
      auto k_addr = [&](uint32_t cnt) { return sR + k_slot(cnt) * kTileBytes; };
      auto q_ready = [&](uint32_t cnt) { return mbar_test_wait(&B.full[q_slot(cnt)], (cnt >> 2) & 1); };
      auto k_ready = [&](uint32_t cnt) { return mbar_test_wait(&B.full[k_slot(cnt)], (cnt >> 2) & 1); };
      Cur S;
      S.k = 0; S.cq = S.ck = 0;
      seek(S, blockIdx.x);
      Cur A = S;
      uint32_t nsb = 0, nblk = 0, jcount = 0;
      while (A.n > 0) {
        // ---- S = Q_i K_j^T, dP = dO_i V_j^T of the next sub-block, as soon as its TMEM buffer and its tiles are there
        if (S.n > 0) {
          const uint32_t buf = nsb & 1;
          const uint32_t tq = S.cq + 2 * S.i, tk = S.ck + 2 * S.j;
          if (mbar_test_wait(&B.sdp_free[buf], ((nsb >> 1) & 1) ^ 1) && q_ready(tq) && q_ready(tq + 1) && k_ready(tk) && k_ready(tk + 1)) {
            tcgen05_fence_after();
            const int kk = keys_of(S, S.j);
            const int ncols = min(kSub, kk - S.c * kSub);
            const uint32_t idesc = umma_idesc_bf16(kTile, ncols, false, false);
            const uint32_t qa = q_addr(tq), ga = q_addr(tq + 1), ka = k_addr(tk) + S.c * (kSub * 128), va = k_addr(tk + 1) + S.c * (kSub * 128);
            const uint32_t ts = tmem_base + kColSdp + buf * 128;
#pragma unroll
            for (int c = 0; c < kHd / 16; ++c)
              umma_bf16(ts, umma_smem_desc(qa + c * 32, 0, 1024), umma_smem_desc(ka + c * 32, 0, 1024), idesc, c > 0 ? 1u : 0u);
#pragma unroll
            for (int c = 0; c < kHd / 16; ++c)
              umma_bf16(ts + 64, umma_smem_desc(ga + c * 32, 0, 1024), umma_smem_desc(va + c * 32, 0, 1024), idesc, c > 0 ? 1u : 0u);
            umma_commit(&B.sdp_full[buf]);
            ++nsb;
            // advance: 64-key half, query tile, key block, item
            if (S.c == 0 && kk > kSub) {
              S.c = 1;
            } else {
              S.c = 0;
              if (++S.i == S.nqt) {
                S.i = 0;
                if (++S.j == S.nkb) next_item(S);
              }
            }
          }
        }
        // ---- accumulate MMAs of the next block whose P / dS tiles are written
        {
          bool ok = mbar_test_wait(B.pds_full, nblk & 1);
          if (ok && A.i == 0 && jcount > 0) ok = mbar_test_wait(B.dkv_free, (jcount & 1) ^ 1);       // previous dK / dV drained
          if (ok && A.i == 0 && A.j == 0 && A.k > 0) ok = mbar_test_wait(B.dq_free, (A.k & 1) ^ 1);  // previous item's dQ drained
          if (ok) {
            tcgen05_fence_after();
            const int i = A.i, j = A.j;
            const int kk = keys_of(A, j);
            const int kq = min(kTile, pad16(A.n - (2 * (A.item % p.QP) + i) * kTile));
            const uint32_t tq = A.cq + 2 * i, tk = A.ck + 2 * j;
            const uint32_t qa = q_addr(tq), ga = q_addr(tq + 1), ka = k_addr(tk);
            // dV_j += P^T dO_i, dK_j += dS^T Q_i: A = [q][keys] tile read MN-major (M = keys), K = query rows
            const uint32_t idesc_t = umma_idesc_bf16(kTile, kHd, true, true);
            const int steps_q = kq >> 4;
            for (int c = 0; c < steps_q; ++c)
              umma_bf16(tmem_base + kColDV, umma_smem_desc(sP + c * 2048, 16384, 1024), umma_smem_desc(ga + c * 2048, 8192, 1024), idesc_t,
                        (i > 0 || c > 0) ? 1u : 0u);
            for (int c = 0; c < steps_q; ++c)
              umma_bf16(tmem_base + kColDK, umma_smem_desc(sS + c * 2048, 16384, 1024), umma_smem_desc(qa + c * 2048, 8192, 1024), idesc_t,
                        (i > 0 || c > 0) ? 1u : 0u);
            // dQ_i += dS K_j: A = the dS tile read K-major (M = query rows), K = keys
            const uint32_t idesc_q = umma_idesc_bf16(kTile, kHd, false, true);
            const int steps_k = kk >> 4;
            for (int c = 0; c < steps_k; ++c)
              umma_bf16(tmem_base + kColDQ + i * 64, umma_smem_desc(sS + (c >> 2) * 16384 + (c & 3) * 32, 0, 1024),
                        umma_smem_desc(ka + c * 2048, 8192, 1024), idesc_q, (j > 0 || c > 0) ? 1u : 0u);
            umma_commit(B.pds_free);
            ++nblk;
            if (i == A.nqt - 1) {  // key block done: dK_j / dV_j final, K_j / V_j no longer needed
              umma_commit(B.dkv_full);
              ++jcount;
              umma_commit(&B.empty[k_slot(tk)]);
              umma_commit(&B.empty[k_slot(tk + 1)]);
            }
            if (j == A.nkb - 1) {  // last key block: Q_i / dO_i no longer needed
              umma_commit(&B.empty[q_slot(tq)]);
              umma_commit(&B.empty[q_slot(tq + 1)]);
            }
            if (++A.i == A.nqt) {
              A.i = 0;
              if (++A.j == A.nkb) {
                umma_commit(B.dq_full);
                next_item(A);
              }
            }
          }
        }
      }
    }
  } else if (warp < 2 + kGroupWarps) {
    // ------------------------------------------------------------------------------------ P / dS group
    const int wi = warp - 2, part = wi >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const float sl2 = p.scale * kLog2e;
    const uint32_t dkey = DROP ? drop_mix(p.drop_salt, __ldg(p.drop_seed)) : 0u;
    const float dscale = drop_scale(p.drop_thr);
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kColSdp + part * 32;
    // this thread's 64-byte piece (32 keys) of a 128-byte row segment of the P / dS tiles: 16-byte chunks cb .. cb + 3
    const uint32_t prow = smem_u32(smem + kPOff + row * 128), srow = smem_u32(smem + kSOff + row * 128);
    const int cb = part * 4, sw = row & 7;
    uint32_t nsb = 0, nblk = 0;
    int k = 0;
    // lse (log2 domain; +inf for absent rows => P = 0) and delta of this thread's row in both query tiles, fetched one item
    // ahead. Rows are clamped into the (sequence, head)'s own max_seq_len slots, so the loads do not wait for the sequence
    // length; absent rows are recognised when the values are used.
    float Ln[2], Dn[2];
    auto fetch = [&](int item, float (&L)[2], float (&D)[2]) {
      if (item < p.num_items) {
        const int sh = item / p.QP, qp = item - sh * p.QP;
        const long long stat0 = static_cast<long long>(sh) * p.max_seq_len;  // sh = s H + h
        const int q0 = min(2 * qp * kTile + row, p.max_seq_len - 1), q1 = min((2 * qp + 1) * kTile + row, p.max_seq_len - 1);
        L[0] = __ldg(p.lse + stat0 + q0);
        L[1] = __ldg(p.lse + stat0 + q1);
        D[0] = __ldg(p.delta + stat0 + q0);
        D[1] = __ldg(p.delta + stat0 + q1);
      }
    };
    fetch(blockIdx.x, Ln, Dn);
    for (int item = blockIdx.x; item < p.num_items; item += G) {
      const Geo g = item_geo(p, item, item_len(p, item));
      float Lr[2], Dr[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        Lr[i] = (2 * g.qp + i) * kTile + row < g.n ? Ln[i] * kLog2e : INFINITY;
        Dr[i] = (2 * g.qp + i) * kTile + row < g.n ? Dn[i] : 0.f;
      }
      fetch(item + G, Ln, Dn);
      if (g.nqt == 0) continue;
      const int s = g.sh / H, h = g.sh - s * H;
      const int n = g.n;
      wait_bar(&B.meta_full[k & (kMetaSlots - 1)], (k >> 2) & 1);
      const uint32_t meta = smem_u32(smem + kMetaOff + (k & (kMetaSlots - 1)) * kMetaBytes);
      ++k;
      for (int j = 0; j < g.nkb; ++j) {
        const int kk = min(kTile, pad16(n - j * kTile));
        const int nsub = kk > kSub ? 2 : 1;
        for (int i = 0; i < g.nqt; ++i, ++nblk) {
          const int q = (2 * g.qp + i) * kTile + row;
          const float L = Lr[i], D = Dr[i];
          const bool warp_live = (2 * g.qp + i) * kTile + quarter * 32 < n;  // some row of this warp is a real query
          for (int c = 0; c < nsub; ++c, ++nsb) {
            const uint32_t buf = nsb & 1;
            const int key0 = c * kSub + part * 32;  // first of this thread's 32 keys inside the block
            const uint32_t kw32 = lds_u32(meta + 4 * (j * 4 + c * 2 + part));
            wait_bar(&B.sdp_full[buf], (nsb >> 1) & 1);
            __syncwarp();
            tcgen05_fence_after();
            const bool live0 = warp_live && key0 < kk, live1 = warp_live && key0 + 16 < kk;
            const uint32_t drow = attn_drop_row(s, H, h, p.max_seq_len, q) + ((j * kTile + key0) >> 1);
            auto math16 = [&](const uint32_t (&sv)[16], const uint32_t (&dv)[16], int e, uint32_t (&po)[8], uint32_t (&so)[8]) {
              const uint32_t kw = (kw32 >> (16 * e)) & 0xffffu;
#pragma unroll
              for (int x = 0; x < 8; ++x) {
                float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * x]), sl2, -L));
                float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * x + 1]), sl2, -L));
                float g0 = __uint_as_float(dv[2 * x]), g1 = __uint_as_float(dv[2 * x + 1]);
                if (kw != 0xffffu) {  // keys that do not take part: exact zeros whatever the (possibly stale) columns hold
                  p0 = ((kw >> (2 * x)) & 1u) ? p0 : 0.f;
                  p1 = ((kw >> (2 * x + 1)) & 1u) ? p1 : 0.f;
                  g0 = ((kw >> (2 * x)) & 1u) ? g0 : 0.f;
                  g1 = ((kw >> (2 * x + 1)) & 1u) ? g1 : 0.f;
                }
                float pd0 = p0, pd1 = p1;
                if (DROP) {  // dropped probabilities feed dV; dP of a dropped one is zero, of a kept one carries the scale
                  const uint32_t word = drop_mix(drow + e * 8 + x, dkey);
                  const float t0 = (word & 255u) >= p.drop_thr ? dscale : 0.f;
                  const float t1 = ((word >> 16) & 255u) >= p.drop_thr ? dscale : 0.f;
                  pd0 = p0 * t0;
                  pd1 = p1 * t1;
                  g0 *= t0;
                  g1 *= t1;
                }
                po[x] = pack_bf16(pd0, pd1);
                so[x] = pack_bf16(p0 * (g0 - D), p1 * (g1 - D));
              }
            };
            auto store16 = [&](int e, const uint32_t (&po)[8], const uint32_t (&so)[8]) {
              const uint32_t o0 = c * kTileBytes + (((cb + 2 * e) ^ sw) << 4), o1 = c * kTileBytes + (((cb + 2 * e + 1) ^ sw) << 4);
              sts_v4_u32(prow + o0, po[0], po[1], po[2], po[3]);
              sts_v4_u32(prow + o1, po[4], po[5], po[6], po[7]);
              sts_v4_u32(srow + o0, so[0], so[1], so[2], so[3]);
              sts_v4_u32(srow + o1, so[4], so[5], so[6], so[7]);
            };
            {
              uint32_t sa[16], da[16], sb[16], db[16], po[8], so[8];
              if (live0) {
                tmem_ld_32x16(trow + buf * 128, sa);
                tmem_ld_32x16(trow + buf * 128 + 64, da);
                tmem_ld_wait();
              }
              if (live1) {  // the second half's loads fly during the first half's arithmetic
                tmem_ld_32x16(trow + buf * 128 + 16, sb);
                tmem_ld_32x16(trow + buf * 128 + 64 + 16, db);
              }
              if (live0) math16(sa, da, 0, po, so);
              // the previous block's MMAs must be done with the tiles before the first store of this block
              if (c == 0 && nblk > 0) wait_bar(B.pds_free, (nblk & 1) ^ 1);
              if (live0) store16(0, po, so);
              if (live1) tmem_ld_wait();
              tcgen05_fence_before();
              mbar_arrive(&B.sdp_free[buf]);  // the sub-block is in registers: the tensor core may overwrite the buffer
              if (live1) {
                math16(sb, db, 1, po, so);
                store16(1, po, so);
              }
            }
          }
          fence_proxy_async();  // the tiles are read by the tensor core (async proxy)
          mbar_arrive(B.pds_full);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------ drain group
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const long long ld3 = 3LL * d;
    const float scale = p.scale;
    const bool atomic_kv = p.dkv_acc != nullptr;
    uint32_t jcount = 0;
    int k = 0;
    // 32 fp32 accumulator columns of this thread's lane -> 32 bf16 (x mul) = 64 bytes of a row
    auto drain32_store = [&](uint32_t col, float mul, __nv_bfloat16* dst, bool live, bool valid) {
      if (live) {  // warp-uniform
        uint32_t a[32];
        tmem_ld_32x32(trow + col, a);
        tmem_ld_wait();
        if (valid) {
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            d4[e] = make_uint4(pack_bf16(__uint_as_float(a[8 * e]) * mul, __uint_as_float(a[8 * e + 1]) * mul),
                               pack_bf16(__uint_as_float(a[8 * e + 2]) * mul, __uint_as_float(a[8 * e + 3]) * mul),
                               pack_bf16(__uint_as_float(a[8 * e + 4]) * mul, __uint_as_float(a[8 * e + 5]) * mul),
                               pack_bf16(__uint_as_float(a[8 * e + 6]) * mul, __uint_as_float(a[8 * e + 7]) * mul));
        }
      }
    };
    auto drain32_add = [&](uint32_t col, float* dst, bool live, bool valid) {
      if (live) {
        uint32_t a[32];
        tmem_ld_32x32(trow + col, a);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            red_add_v4(dst + 4 * e, __uint_as_float(a[4 * e]), __uint_as_float(a[4 * e + 1]), __uint_as_float(a[4 * e + 2]),
                       __uint_as_float(a[4 * e + 3]));
        }
      }
    };
    for (int item = blockIdx.x; item < p.num_items; item += G) {
      const Geo g = item_geo(p, item, item_len(p, item));
      if (g.nqt == 0) continue;
      const int s = g.sh / H, h = g.sh - s * H;
      const Seq sd = load_seq(p.seq_desc, s);
      const int n = g.n;
      for (int j = 0; j < g.nkb; ++j, ++jcount) {
        const int key = j * kTile + row;
        const bool live = j * kTile + quarter * 32 < n, valid = key < n;
        const long long grow = seq_row(sd, min(key, n - 1));
        wait_bar(B.dkv_full, jcount & 1);
        __syncwarp();
        tcgen05_fence_after();
        if (!atomic_kv) {
          __nv_bfloat16* base = p.dqkv + grow * ld3 + h * kHd;
          drain32_store(kColDK, scale, base + d, live, valid);
          drain32_store(kColDK + 32, scale, base + d + 32, live, valid);
          drain32_store(kColDV, 1.f, base + 2 * d, live, valid);
          drain32_store(kColDV + 32, 1.f, base + 2 * d + 32, live, valid);
        } else {  // several query pairs add into the same rows (the scale is applied by the conversion kernel)
          float* base = p.dkv_acc + grow * (2LL * d) + h * kHd;
          drain32_add(kColDK, base, live, valid);
          drain32_add(kColDK + 32, base + 32, live, valid);
          drain32_add(kColDV, base + d, live, valid);
          drain32_add(kColDV + 32, base + d + 32, live, valid);
        }
        tcgen05_fence_before();
        mbar_arrive(B.dkv_free);
      }
      wait_bar(B.dq_full, k & 1);
      ++k;
      __syncwarp();
      tcgen05_fence_after();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int q = (2 * g.qp + i) * kTile + row;
        const bool live = i < g.nqt && (2 * g.qp + i) * kTile + quarter * 32 < n, valid = i < g.nqt && q < n;
        __nv_bfloat16* base = p.dqkv + seq_row(sd, min(q, n - 1)) * ld3 + h * kHd;
        drain32_store(kColDQ + i * 64, scale, base, live, valid);
        drain32_store(kColDQ + i * 64 + 32, scale, base + 32, live, valid);
      }
      tcgen05_fence_before();
      mbar_arrive(B.dq_free);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// dK / dV of the long-sequence path: scratch [tokens][2 d] fp32 -> the k and v thirds of dqkv (k scaled)
__global__ void __launch_bounds__(256) attn_dkv_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, long long tokens,
                                                               int d, float scale) {
  const long long per_row = 2LL * d / 8;  // 8 elements per thread
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= tokens * per_row) return;
  const long long row = idx / per_row;
  const int c = static_cast<int>(idx - row * per_row) * 8;
  const float mul = c < d ? scale : 1.f;
  const float4 a = *reinterpret_cast<const float4*>(acc + row * 2LL * d + c), b = *reinterpret_cast<const float4*>(acc + row * 2LL * d + c + 4);
  *reinterpret_cast<uint4*>(dqkv + row * 3LL * d + d + c) =
      make_uint4(pack_bf16(a.x * mul, a.y * mul), pack_bf16(a.z * mul, a.w * mul), pack_bf16(b.x * mul, b.y * mul), pack_bf16(b.z * mul, b.w * mul));
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

int attn_bwd_tc_pipe_max_seq_len() { return kMaxKeys; }

// floats of workspace behind the num_seqs * H * max_seq_len delta values: the dK / dV scratch of the long-sequence path
int64_t attn_bwd_tc_pipe_extra_ws_floats(int64_t tokens, int max_seq_len, int H) {
  return max_seq_len > 2 * kTile ? tokens * 2 * H * kHd : 0;
}

int attn_bwd_tc_pipe(const void* qkv, const void* out, const void* dout, const int32_t* seq_desc, const uint8_t* key_mask, const float* lse,
                     void* dqkv, float* delta_ws, int64_t tokens, int num_seqs, int max_seq_len, int H, float scale, const uint32_t* drop_seed,
                     uint32_t drop_salt, float drop_p, cudaStream_t stream) {
  MOME_REQUIRE(max_seq_len <= kMaxKeys, "attn_bwd_tc_pipe: max_seq_len %d > %d", max_seq_len, kMaxKeys);
  static bool configured = false;
  if (!configured) {
    int rc = opt_in(attn_bwd_tc_pipe_kernel<false>, kSmemBytes, "attn_bwd_tc_pipe");
    if (rc == MOME_OK) rc = opt_in(attn_bwd_tc_pipe_kernel<true>, kSmemBytes, "attn_bwd_tc_pipe");
    if (rc != MOME_OK) return rc;
    configured = true;
  }
  Params p;
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
  p.QP = ((max_seq_len + kTile - 1) / kTile + 1) / 2;
  p.num_items = num_seqs * H * p.QP;
  p.scale = scale;
  p.drop_seed = drop_seed;
  p.drop_salt = drop_salt;
  p.drop_thr = drop_threshold(drop_p);
  p.dkv_acc = nullptr;
  if (p.QP > 1) {  // workspace layout: delta [num_seqs H max_seq_len] | dK / dV scratch [tokens][2 d]
    const int64_t delta_floats = (static_cast<int64_t>(num_seqs) * H * max_seq_len + 3) & ~int64_t(3);
    p.dkv_acc = delta_ws + delta_floats;
    cudaError_t e = cudaMemsetAsync(p.dkv_acc, 0, static_cast<size_t>(tokens) * 2 * d * sizeof(float), stream);
    if (e != cudaSuccess) {
      set_error("attn_bwd_tc_pipe: cudaMemsetAsync: %s", cudaGetErrorString(e));
      return MOME_ERR_CUDA;
    }
  }
  rc = attn_delta_launch(out, dout, seq_desc, delta_ws, H, max_seq_len, num_seqs, stream);
  if (rc != MOME_OK) return rc;
  const int grid = std::min(p.num_items, sm_count());
  if (drop_seed != nullptr && drop_p > 0.f)
    attn_bwd_tc_pipe_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(p);
  else
    attn_bwd_tc_pipe_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(p);
  rc = check_launch("attn_bwd_tc_pipe");
  if (rc != MOME_OK || p.QP == 1) return rc;
  const long long work = tokens * (2 * d / 8);
  attn_dkv_convert_kernel<<<static_cast<unsigned>((work + 255) / 256), 256, 0, stream>>>(p.dkv_acc, p.dqkv, tokens, static_cast<int>(d), scale);
  return check_launch("attn_dkv_convert");
}

}  // namespace mome
