// K1 backward on tcgen05 / TMEM / TMA, software-pipelined (round 2). Same contract and same arithmetic as
// attention_tc_bwd.cu (mask semantics, lse format, dropout mask function, dqkv layout); what changes is the schedule.
//
// ncu of the first kernel (profiles/r02_attn_bwd_ncu.md): the P / dS warps spend 22 % of their time on the P / dS
// arithmetic and the rest waiting — for S / dP of the block (the tensor core only starts them once the previous block's
// accumulate MMAs were issued), for dK / dV / dQ to become final so that they can drain them themselves, for the next
// item's TMA loads (issued after the last MMA of the item), for lse / delta loads at the start of an item. Here every one
// of these waits has something else to overlap with:
//
//   * S / dP are computed per SUB-BLOCK of 128 queries x 64 keys into a two-deep ring of TMEM buffers
//     (2 x (64 + 64) columns); the P / dS warps copy a sub-block into registers and hand the buffer back at once,
//     so the tensor core is always one to two sub-blocks ahead of them;
//   * the accumulate MMAs (dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j) stay per 128 x 128 block, M = 128 keys;
//     the P / dS warps keep a sub-block's results in registers and only wait for the previous block's MMAs to have
//     finished reading the P / dS tiles at their first store;
//   * dK / dV / dQ are drained by four extra warps, not by the P / dS warps;
//   * Q, K, V, dO live in a ring of eight 128-row tiles with a full / empty barrier pair each. A tile is released by the
//     last MMA that reads it, so K_0 / V_0 of the next item arrive while the second key block of this item is processed,
//     and two 1-tile (text) items are in flight at once;
//   * a thread fetches the next item's lse / delta while it works on this one.
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
constexpr int kMaxKeys = 256;
constexpr float kLog2e = 1.4426950408889634f;

constexpr int kTileBytes = kTile * 128;  // 128 rows x 64 bf16
constexpr int kSlots = 8;
constexpr int kPOff = kSlots * kTileBytes, kSOff = kPOff + 2 * kTileBytes;  // P, dS: 2 chunks of 64 keys x [128 q x 128 B]
constexpr int kMetaOff = kSOff + 2 * kTileBytes;                              // 4 slots x 64 B: keep words [8], seq desc [4]
constexpr int kMetaSlots = 4;
constexpr int kBarOff = kMetaOff + kMetaSlots * 64;
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
  int H, max_seq_len, num_items;
  float scale;
  const uint32_t* drop_seed;
  uint32_t drop_salt, drop_thr;
};

// Tiles of an item in loading order = order of first use: K_0 V_0 Q_0 dO_0 [Q_1 dO_1] [K_1 V_1]
__device__ __forceinline__ int tile_k(int j, int nt) { return j == 0 ? 0 : 2 + 2 * nt + 2 * (j - 1); }
__device__ __forceinline__ int tile_q(int i) { return 2 + 2 * i; }

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
    uint32_t cnt = 0;  // tiles loaded so far: slot = cnt & 7, fill number = cnt >> 3
    int k = 0;
    Seq sd_next = blockIdx.x < p.num_items ? load_seq(p.seq_desc, blockIdx.x / H) : Seq{0, 0, 0, 0};
    for (int item = blockIdx.x; item < p.num_items; item += G, ++k) {
      const int s = item / H, h = item - s * H;
      const Seq sd = sd_next;
      if (item + G < p.num_items) sd_next = load_seq(p.seq_desc, (item + G) / H);
      const int n = sd.len0 + sd.len1;
      const int nt = n > kTile ? 2 : 1;
      // ---- meta: keep words of the item's keys, its descriptor (slot k & 3: the groups are at most two items behind)
      uint32_t* meta = reinterpret_cast<uint32_t*>(smem + kMetaOff + (k & (kMetaSlots - 1)) * 64);
      uint8_t mk[8];
#pragma unroll
      for (int j8 = 0; j8 < 8; ++j8) {
        const int j = j8 * 32 + lane;
        mk[j8] = j < n ? (p.key_mask == nullptr ? uint8_t(1) : __ldg(p.key_mask + seq_row(sd, j))) : uint8_t(0);
      }
      // the slot's previous user is item k - 4; item k - 2 has released the first ring slot this item needs before the
      // loads below can start, and the groups finished item k - 4 long before that
      wait_bar(&B.empty[cnt & 7], ((cnt >> 3) & 1) ^ 1);
#pragma unroll
      for (int j8 = 0; j8 < 8; ++j8) {
        const uint32_t w = __ballot_sync(0xffffffffu, mk[j8] != 0);
        if (lane == 0) meta[j8] = w;
      }
      if (lane == 0) {
        meta[8] = sd.start0; meta[9] = sd.len0; meta[10] = sd.start1; meta[11] = sd.len1;
        mbar_arrive(&B.meta_full[k & (kMetaSlots - 1)]);
      }
      const bool boxes_ok = sd.len1 == 0 || (sd.len0 & 7) == 0;
      const int ntiles = 4 * nt;
#pragma unroll 1
      for (int t = 0; t < ntiles; ++t, ++cnt) {
        // t -> (operand, row tile): 0 K_0, 1 V_0, 2 Q_0, 3 dO_0, 4 Q_1, 5 dO_1, 6 K_1, 7 V_1
        const int op = (t == 0 || t == 6) ? 1 : (t == 1 || t == 7) ? 2 : (t & 1) ? 3 : 0;  // 0 Q, 1 K, 2 V, 3 dO
        const int rt = (t < 4) ? 0 : 1;
        const int slot = cnt & 7;
        uint8_t* dst = smem + slot * kTileBytes;
        const int r0 = rt * kTile, r1 = min(n, r0 + kTile);
        wait_bar(&B.empty[slot], ((cnt >> 3) & 1) ^ 1);
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
      const int first = blockIdx.x;
      struct Cur {
        int k, n, nt, j, i, c;
        uint32_t tb;  // tiles of the items before this one
      };
      auto item_n = [&](int k) {
        const int item = first + k * G;
        if (item >= p.num_items) return 0;
        const int4 v = __ldg(reinterpret_cast<const int4*>(p.seq_desc + 4 * (item / H)));
        return v.y + v.w;
      };
      auto next_item = [&](Cur& c) {
        c.tb += 4 * c.nt;
        ++c.k;
        c.n = item_n(c.k);
        c.nt = c.n > kTile ? 2 : 1;
        c.j = c.i = c.c = 0;
      };
      auto keys_of = [](const Cur& c, int j) { return min(kTile, pad16(c.n - j * kTile)); };
      auto slot_addr = [&](uint32_t cnt) { return sR + (cnt & 7) * kTileBytes; };
      auto tile_ready = [&](uint32_t cnt) { return mbar_test_wait(&B.full[cnt & 7], (cnt >> 3) & 1); };
      Cur S{0, item_n(0), 0, 0, 0, 0, 0};
      S.nt = S.n > kTile ? 2 : 1;
      Cur A = S;
      uint32_t nsb = 0, nblk = 0, jcount = 0;
      while (A.n > 0) {
        bool progress = false;
        // ---- S = Q_i K_j^T, dP = dO_i V_j^T of the next sub-block, as soon as its TMEM buffer and its tiles are there
        if (S.n > 0) {
          const uint32_t buf = nsb & 1;
          const uint32_t tq = S.tb + tile_q(S.i), tk = S.tb + tile_k(S.j, S.nt);
          if (mbar_test_wait(&B.sdp_free[buf], ((nsb >> 1) & 1) ^ 1) && tile_ready(tq) && tile_ready(tq + 1) && tile_ready(tk) &&
              tile_ready(tk + 1)) {
            tcgen05_fence_after();
            const int kk = keys_of(S, S.j);
            const int ncols = min(kSub, kk - S.c * kSub);
            const uint32_t idesc = umma_idesc_bf16(kTile, ncols, false, false);
            const uint32_t qa = slot_addr(tq), ga = slot_addr(tq + 1), ka = slot_addr(tk) + S.c * (kSub * 128), va = slot_addr(tk + 1) + S.c * (kSub * 128);
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
              if (++S.i == S.nt) {
                S.i = 0;
                if (++S.j == S.nt) next_item(S);
              }
            }
            progress = true;
          }
        }
        // ---- accumulate MMAs of the next block whose P / dS tiles are written
        {
          bool ok = mbar_test_wait(B.pds_full, nblk & 1);
          if (ok && A.i == 0 && jcount > 0) ok = mbar_test_wait(B.dkv_free, (jcount & 1) ^ 1);       // previous dK / dV drained
          if (ok && A.i == 0 && A.j == 0 && A.k > 0) ok = mbar_test_wait(B.dq_free, (A.k & 1) ^ 1);  // previous item's dQ drained
          if (ok) {
            tcgen05_fence_after();
            const int i = A.i, j = A.j, nt = A.nt;
            const int kk = keys_of(A, j);
            const int kq = min(kTile, pad16(A.n - i * kTile));
            const uint32_t tq = A.tb + tile_q(i), tk = A.tb + tile_k(j, nt);
            const uint32_t qa = slot_addr(tq), ga = slot_addr(tq + 1), ka = slot_addr(tk);
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
            if (i == nt - 1) {  // key block done: dK_j / dV_j final, K_j / V_j no longer needed
              umma_commit(B.dkv_full);
              ++jcount;
              umma_commit(&B.empty[tk & 7]);
              umma_commit(&B.empty[(tk + 1) & 7]);
            }
            if (j == nt - 1) {  // last key block: Q_i / dO_i no longer needed
              umma_commit(&B.empty[tq & 7]);
              umma_commit(&B.empty[(tq + 1) & 7]);
            }
            if (++A.i == nt) {
              A.i = 0;
              if (++A.j == nt) {
                umma_commit(B.dq_full);
                next_item(A);
              }
            }
            progress = true;
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
    // lse (log2 domain; +inf for absent rows => P = 0) and delta of this thread's row in both query tiles, fetched one item ahead
    float Ln[2], Dn[2];
    // rows are clamped into the (sequence, head)'s own max_seq_len slots, so the loads do not wait for the sequence length;
    // absent rows are recognised when the values are used
    const int qc0 = min(row, p.max_seq_len - 1), qc1 = min(kTile + row, p.max_seq_len - 1);
    auto fetch = [&](int item, float (&L)[2], float (&D)[2]) {
      if (item < p.num_items) {
        const long long stat0 = static_cast<long long>(item) * p.max_seq_len;  // item = s H + h
        L[0] = __ldg(p.lse + stat0 + qc0);
        L[1] = __ldg(p.lse + stat0 + qc1);
        D[0] = __ldg(p.delta + stat0 + qc0);
        D[1] = __ldg(p.delta + stat0 + qc1);
      }
    };
    fetch(blockIdx.x, Ln, Dn);
    for (int item = blockIdx.x; item < p.num_items; item += G, ++k) {
      const int s = item / H, h = item - s * H;
      wait_bar(&B.meta_full[k & (kMetaSlots - 1)], (k >> 2) & 1);
      const uint32_t meta = smem_u32(smem + kMetaOff + (k & (kMetaSlots - 1)) * 64);
      const int n = static_cast<int>(lds_u32(meta + 36) + lds_u32(meta + 44));
      const int nt = n > kTile ? 2 : 1;
      float Lr[2], Dr[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        Lr[i] = i * kTile + row < n ? Ln[i] * kLog2e : INFINITY;
        Dr[i] = i * kTile + row < n ? Dn[i] : 0.f;
      }
      fetch(item + G, Ln, Dn);
      for (int j = 0; j < nt; ++j) {
        const int kk = min(kTile, pad16(n - j * kTile));
        const int nsub = kk > kSub ? 2 : 1;
        for (int i = 0; i < nt; ++i, ++nblk) {
          const int q = i * kTile + row;
          const float L = Lr[i], D = Dr[i];
          const bool warp_live = i * kTile + quarter * 32 < n;  // some row of this warp is a real query
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
    uint32_t jcount = 0;
    int k = 0;
    // 64 fp32 accumulator columns of this thread's lane -> 64 bf16 (x mul) = one 128-byte row segment
    auto load64 = [&](uint32_t col, float mul, uint4 (&o)[8]) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t a[32];
        tmem_ld_32x32(trow + col + hf * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 4; ++e)
          o[hf * 4 + e] = make_uint4(pack_bf16(__uint_as_float(a[8 * e]) * mul, __uint_as_float(a[8 * e + 1]) * mul),
                                     pack_bf16(__uint_as_float(a[8 * e + 2]) * mul, __uint_as_float(a[8 * e + 3]) * mul),
                                     pack_bf16(__uint_as_float(a[8 * e + 4]) * mul, __uint_as_float(a[8 * e + 5]) * mul),
                                     pack_bf16(__uint_as_float(a[8 * e + 6]) * mul, __uint_as_float(a[8 * e + 7]) * mul));
      }
    };
    auto store64 = [&](__nv_bfloat16* dst, const uint4 (&o)[8]) {
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int e = 0; e < 8; ++e) d4[e] = o[e];
    };
    for (int item = blockIdx.x; item < p.num_items; item += G, ++k) {
      const int s = item / H, h = item - s * H;
      const Seq sd = load_seq(p.seq_desc, s);
      const int n = sd.len0 + sd.len1;
      const int nt = n > kTile ? 2 : 1;
      for (int j = 0; j < nt; ++j, ++jcount) {
        const int key = j * kTile + row;
        const bool warp_live = j * kTile + quarter * 32 < n;
        wait_bar(B.dkv_full, jcount & 1);
        __syncwarp();
        tcgen05_fence_after();
        __nv_bfloat16* base = p.dqkv + seq_row(sd, min(key, n - 1)) * ld3 + h * kHd;
        {
          uint4 o[8];
          if (warp_live) load64(kColDK, scale, o);
          if (key < n) store64(base + d, o);
          if (warp_live) load64(kColDV, 1.f, o);
          tcgen05_fence_before();
          mbar_arrive(B.dkv_free);
          if (key < n) store64(base + 2 * d, o);
        }
      }
      wait_bar(B.dq_full, k & 1);
      __syncwarp();
      tcgen05_fence_after();
      {
        uint4 o[8];
        const bool live0 = quarter * 32 < n, live1 = kTile + quarter * 32 < n;
        if (live0) load64(kColDQ, scale, o);
        if (row < n) store64(p.dqkv + seq_row(sd, row) * ld3 + h * kHd, o);
        if (live1) load64(kColDQ + 64, scale, o);
        tcgen05_fence_before();
        mbar_arrive(B.dq_free);
        if (kTile + row < n) store64(p.dqkv + seq_row(sd, kTile + row) * ld3 + h * kHd, o);
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
  p.num_items = num_seqs * H;
  p.scale = scale;
  p.drop_seed = drop_seed;
  p.drop_salt = drop_salt;
  p.drop_thr = drop_threshold(drop_p);
  rc = attn_delta_launch(out, dout, seq_desc, delta_ws, H, max_seq_len, num_seqs, stream);
  if (rc != MOME_OK) return rc;
  const int grid = std::min(p.num_items, sm_count());
  if (drop_seed != nullptr && drop_p > 0.f)
    attn_bwd_tc_pipe_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(p);
  else
    attn_bwd_tc_pipe_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(p);
  return check_launch("attn_bwd_tc_pipe");
}

}  // namespace mome
