// K2: persistent, warp-specialised grouped GEMM on the 5th-gen tensor cores, CTA pairs.
//
//   out[m, n] = epilogue( sum_k A(m, k) * B(n, k) )   per group (expert segment of the packed tokens)
//
// Launch: 148 CTAs as 74 two-CTA clusters (one pair per TPC), 384 threads per CTA:
//   warp 0      TMA producer (one elected lane): its CTA's 128 rows of A and its half of B per k-block
//   warp 1      tcgen05.mma issuer (one elected lane of the LEADER CTA): cta_group::2, M = 256, N = BLOCK_N
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: TMEM -> registers -> per-warp padded smem tile (transpose) -> fused epilogue with
//               128-bit row-contiguous global loads / stores
// A work item is a 256 x BLOCK_N output tile (x one k-split): each CTA of the pair owns 128 accumulator
// lanes (rows) x BLOCK_N columns in TMEM, double-buffered so the epilogue of item i overlaps the MMAs
// of item i+1. Pairing halves the shared-memory / L2 traffic per FLOP for B: at 128 x 256 single-CTA
// tiles the kernel was L2-bandwidth bound (profiles/r01_launches_v1_summary.md).
// Operands are staged by TMA into a ring of SWIZZLE_128B tiles; both may be K-major or MN-major
// (transpose bits of the instruction descriptor), so forward, dgrad and wgrad all run on this kernel
// without transposed copies (include/mome.h). wgrad uses split-K with fp32 red.add.
//
// Replaces: F.linear / timm Mlp / residual adds of the reference Block (vlmo.py:76-78, 96, 190-196).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <stdlib.h>
#include <algorithm>
#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "dropout.cuh"
#include "ptx.cuh"

namespace mome {

namespace {

constexpr int PAIR_M = 256;   // rows of a work item (two CTAs x 128)
constexpr int CTA_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
// Epilogue warps per CTA. What bounds the K = 768 GEMMs is the 128 B/clk of the SM's shared-memory data path: the main
// loop alone moves 2 x 32 KB per k-block through it (TMA fill + tcgen05 operand reads = 96 B/clk at the measured MMA
// rate), and the epilogue's transpose tile adds 256 KB per 128 x 256 fp32 tile, the whole remaining budget of a tile's
// ~8k MMA cycles (r02 measurements: TMEM drain only 1560-1600 TFLOP/s, + transpose tile 1560 / 1600 / 1100, + math and
// stores 1216 / 979 / 921 for STORE / GELU / DGELU). More epilogue warps therefore buy little: 16 (with a 4-stage ring
// instead of 5 to pay for their transpose tiles) measured +3 % on GELU, whose fp16 math is latency bound as well, and
// -3 .. -7 % on the STORE and DGELU variants, which keep 8.
constexpr int epi_warps(int epi) { return epi == MOME_EPI_GELU ? 16 : 8; }
constexpr int gemm_threads(int epi) { return 128 + 32 * epi_warps(epi); }
constexpr int kAtomBytes = BLOCK_K * 128;  // one 64x64 MN-major box / 64 rows of a K-major tile
constexpr int kStagePitch = 36;            // floats per row of the per-warp transpose tile (32 + pad, 16 B aligned)
constexpr int kStageBytesPerWarp = 32 * kStagePitch * 4;
constexpr int kStage16BytesPerWarp = 32 * 64;  // bf16 transpose tile: 32 rows x 32 columns, 16-byte pieces XOR-swizzled
// Variants whose epilogue consumes bf16(acc + bias) — GELU (the reference's autocast Linear hands bf16 to the activation)
// and RESIDUAL (b = bf16(acc + bias) is the branch value) — add the bias in the accumulator's row layout and stage bf16:
// half the staging bytes, and the 16-warp GELU variant keeps a 5-stage ring. Measured: GELU 1037 -> 1050 TFLOP/s; STORE
// gained nothing (1219 -> 1199) and its fp32-output form must not round, so it keeps the fp32 tile.
constexpr bool stage16(int epi) { return epi == MOME_EPI_GELU || epi == MOME_EPI_RESIDUAL; }

struct GemmGroupDev {
  void* out;
  void* out2;
  const float* bias;
  const float* res;
  const void* aux;
  float* colsum;
  long long row0;
  int M;
  int k_blocks;
  int item_start;
  int pad;
};

struct alignas(64) GemmParams {
  CUtensorMap tma_a[MOME_MAX_GROUPS];
  CUtensorMap tma_b[MOME_MAX_GROUPS];
  GemmGroupDev g[MOME_MAX_GROUPS];
  const float* gamma;
  const uint32_t* drop_seed;
  const float* row_scale;
  uint32_t drop_salt, drop_thr;
  long long ldo, ldo2, ldres, ldaux;
  int num_groups, total_items, n_tiles, splits;
  int N, epilogue, out_bf16, debug;
};

struct WorkItem {
  int g, m_tile, n_tile, kb0, kb1;
};

__device__ __forceinline__ WorkItem decode_item(const GemmParams& p, int item) {
  WorkItem w;
  w.g = 0;
#pragma unroll
  for (int i = 1; i < MOME_MAX_GROUPS; ++i)
    if (i < p.num_groups && item >= p.g[i].item_start) w.g = i;
  int local = item - p.g[w.g].item_start;
  const int split = local % p.splits;
  local /= p.splits;
  w.n_tile = local % p.n_tiles;
  w.m_tile = local / p.n_tiles;
  const int kb = p.g[w.g].k_blocks;
  const int per = (kb + p.splits - 1) / p.splits;
  w.kb0 = split * per;
  w.kb1 = min(w.kb0 + per, kb);
  return w;
}

__device__ __forceinline__ uint2 pack4_bf16(float4 v) { return make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w)); }
__device__ __forceinline__ float4 unpack4_bf16(uint2 u) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// bf16x2 <-> fp32 helpers of the epilogue
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u)); }

// GELU (erf form, reference vlmo.py:114 nn.GELU) and its derivative for two elements at once in packed fp16: the
// epilogue of the K = 768 fc1 GEMM has ~0.2 cycles per element per SM before it, not the MMAs, bounds the kernel,
// and fp32 erff costs ~25 issue slots per element. Phi(z) = 0.5 (1 + erf(z / sqrt 2)) is evaluated through the
// odd approximant
//     Phi(z) ~= 0.5 (1 + tanh(z (a + b z^2 + c z^4))),   a = 0.797627599, b = 0.0369255429, c = -3.41174313e-4
// (minimax fit of gelu over [-8, 8]: |gelu - gelu_erf| < 3.1e-5, |gelu' - gelu_erf'| < 1.2e-4 in exact arithmetic,
// 15x tighter than the usual two-term "tanh GELU" and below half an ulp of the bf16 results for |value| > 0.016);
// z^2 is clamped at 64, beyond which tanh saturates. One MUFU (tanh.approx.f16x2) per pair: the derivative is the
// approximant's own, gelu'(z) = Phi + 0.5 z (1 - t^2) (a + 3 b z^2 + 5 c z^4), so forward and backward are consistent.
// fp16 arithmetic (11-bit significand) adds ~5e-4 relative, the fp32 validation path uses exact erff.
__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
template <bool DROP>
__device__ __forceinline__ void gelu_fast2(float z0, float z1, float m0, float m1, uint32_t& g_bf16x2, uint32_t& dg_bf16x2) {
  const __half2 z = __floats2half2_rn(z0, z1);
  const __half2 z2 = __hmin2(__hmul2(z, z), __float2half2_rn(64.f));
  const __half2 q = __hfma2(__hfma2(z2, __float2half2_rn(-3.41174313e-4f), __float2half2_rn(0.0369255429f)), z2,
                            __float2half2_rn(0.797627599f));
  uint32_t th;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(th) : "r"(h2_as_u32(__hmul2(z, q))));
  const __half2 t = u32_as_h2(th);
  const __half2 cdf = __hfma2(t, __float2half2_rn(0.5f), __float2half2_rn(0.5f));
  // 0.5 * d/dz [z (a + b z^2 + c z^4)] = 0.5 a + 1.5 b z^2 + 2.5 c z^4
  const __half2 hdu = __hfma2(__hfma2(z2, __float2half2_rn(2.5f * -3.41174313e-4f), __float2half2_rn(1.5f * 0.0369255429f)), z2,
                              __float2half2_rn(0.5f * 0.797627599f));
  const __half2 sech2 = __hfma2(__hneg2(t), t, __float2half2_rn(1.f));
  __half2 g = __hmul2(z, cdf);
  __half2 dg = __hfma2(__hmul2(z, hdu), sech2, cdf);
  if (DROP) {  // dropout after the activation (timm Mlp): both the value and the derivative carry mask * scale
    const __half2 m = __floats2half2_rn(m0, m1);
    g = __hmul2(g, m);
    dg = __hmul2(dg, m);
  }
  const float2 gf = __half22float2(g), dgf = __half22float2(dg);
  g_bf16x2 = pack_bf16(gf.x, gf.y);
  dg_bf16x2 = pack_bf16(dgf.x, dgf.y);
}

template <int BLOCK_N, int EW, bool S16>
struct GemmCfg {
  static constexpr int A_BYTES = CTA_M * 128;            // this CTA's 128 rows x 64 k
  static constexpr int B_BYTES = (BLOCK_N / 2) * 128;    // this CTA's half of the N tile x 64 k
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // the ring shares the 227 KB with one transpose tile per epilogue warp
  static constexpr int STAGES = (BLOCK_N == 256) ? ((EW == 16 && !S16) ? 4 : 5) : ((EW == 16 && !S16) ? 6 : 7);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int STAGE_TILE_BYTES = S16 ? kStage16BytesPerWarp : kStageBytesPerWarp;
  static constexpr int EPI_BYTES = EW * STAGE_TILE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 + 1024;  // + barriers + alignment slack
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, bool DROP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm_threads(EPI), 1) gemm_pair_kernel(const __grid_constant__ GemmParams p) {
  constexpr int kEpiWarps = epi_warps(EPI);
  constexpr int kParts = kEpiWarps / 4;            // column parts of a tile (one per group of 4 warps = 128 TMEM lanes)
  constexpr int kPartCols = BLOCK_N / kParts;
  constexpr bool S16 = stage16(EPI);
  using Cfg = GemmCfg<BLOCK_N, kEpiWarps, S16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.num_groups; ++g) {
      tma_prefetch_desc(&p.tma_a[g]);
      tma_prefetch_desc(&p.tma_b[g]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's copy is the one in use: one arrive.expect_tx + both CTAs' TMA bytes
      mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);                // multicast tcgen05.commit
      mbar_init(&tempty_bar[s], 2 * kEpiWarps);   // leader's copy: every epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  cluster_sync_all();  // barrier inits and TMEM allocations of both CTAs are visible before any remote arrive / MMA
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = pair; item < p.total_items; item += num_pairs) {
        const WorkItem w = decode_item(p, item);
        const int m0 = w.m_tile * PAIR_M + cta_rank * CTA_M;
        const int n0 = w.n_tile * BLOCK_N + cta_rank * (BLOCK_N / 2);
        // Epilogue operand (gelu' / residual) of this CTA's slab: pulled into L2 by bulk prefetches spread over
        // the item's k-blocks, one main loop ahead of the epilogue warps that read it with plain loads (they
        // then see L2 instead of HBM latency).
        constexpr bool kPrefetch = EPI == MOME_EPI_DGELU || EPI == MOME_EPI_RESIDUAL;
        const GemmGroupDev& gp = p.g[w.g];
        const int pf_esz = EPI == MOME_EPI_DGELU ? 2 : 4;
        const long long pf_ld = EPI == MOME_EPI_DGELU ? p.ldaux : p.ldres;
        const char* pf_base = (EPI == MOME_EPI_DGELU ? static_cast<const char*>(gp.aux) : reinterpret_cast<const char*>(gp.res)) +
                              (static_cast<long long>(m0) * pf_ld + static_cast<long long>(w.n_tile) * BLOCK_N) * pf_esz;
        const int pf_bytes = min(BLOCK_N, p.N - w.n_tile * BLOCK_N) * pf_esz;
        const int pf_rows = kPrefetch ? min(CTA_M, gp.M - m0) : 0;
        const int pf_per_kb = (pf_rows + (w.kb1 - w.kb0) - 1) / max(1, w.kb1 - w.kb0);
        int pf_next = 0;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + Cfg::A_BYTES;
          if (!A_MN) {
            tma_load_2d_pair(a_dst, &p.tma_a[w.g], &full_bar[stage], kb * BLOCK_K, m0);
          } else {
#pragma unroll
            for (int j = 0; j < CTA_M / 64; ++j)
              tma_load_2d_pair(a_dst + j * kAtomBytes, &p.tma_a[w.g], &full_bar[stage], m0 + j * 64, kb * BLOCK_K);
          }
          if (!B_MN) {
            tma_load_2d_pair(b_dst, &p.tma_b[w.g], &full_bar[stage], kb * BLOCK_K, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 128; ++j)
              tma_load_2d_pair(b_dst + j * kAtomBytes, &p.tma_b[w.g], &full_bar[stage], n0 + j * 64, kb * BLOCK_K);
          }
          if (kPrefetch) {
            const int pf_end = min(pf_rows, pf_next + pf_per_kb);
            for (; pf_next < pf_end; ++pf_next) l2_prefetch(pf_base + pf_next * pf_ld * pf_esz, pf_bytes);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (cta_rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR_M, BLOCK_N, A_MN, B_MN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = pair; item < p.total_items; item += num_pairs) {
        const WorkItem w = decode_item(p, item);
        if (w.kb0 >= w.kb1) continue;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_base = a_base + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t a_desc = A_MN ? umma_smem_desc(a_base + k * (UMMA_K * 128), kAtomBytes, 1024)
                                         : umma_smem_desc(a_base + k * (UMMA_K * 2), 0, 1024);
            const uint64_t b_desc = B_MN ? umma_smem_desc(b_base + k * (UMMA_K * 128), kAtomBytes, 1024)
                                         : umma_smem_desc(b_base + k * (UMMA_K * 2), 0, 1024);
            umma_bf16_pair(d_tmem, a_desc, b_desc, idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage]);  // frees the slot in both CTAs once these MMAs have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull_bar[acc]);  // accumulator complete -> epilogue warps of both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int ew = warp - 4;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int part = ew >> 2;              // which column part of the tile
    const uint32_t stage_u32 = smem_u32(epi_stage + ew * (Cfg::STAGE_TILE_BYTES / 4));  // this warp's transpose tile
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = pair; item < p.total_items; item += num_pairs) {
      const WorkItem w = decode_item(p, item);
      if (w.kb0 >= w.kb1) continue;
      // ---- per-item setup: group fields into registers, running byte pointers at (row0 + rsub, col_base),
      // number of this lane's 8 row-iterations that fall inside the group
      const GemmGroupDev g = p.g[w.g];
      const long long row0 = static_cast<long long>(w.m_tile) * PAIR_M + cta_rank * CTA_M + quarter * 32;
      constexpr int kChunks = kPartCols / 32;  // 32-column chunks per warp
      const int col_base = w.n_tile * BLOCK_N + part * kPartCols + c4;
      const long long rfirst = row0 + rsub;
      const int nvalid = static_cast<int>(max(0LL, min(8LL, (static_cast<long long>(g.M) - rfirst + 3) >> 2)));
      const int nchunks = max(0, min(kChunks, (p.N - col_base + 31) >> 5));  // chunks inside N (N % 32 == 0)
      const int osize = (EPI == MOME_EPI_ATOMIC || EPI == MOME_EPI_RESIDUAL || !p.out_bf16) ? 4 : 2;
      char* out_p = static_cast<char*>(g.out) + (rfirst * p.ldo + col_base) * osize;
      const long long out_step = 4 * p.ldo * osize;
      char* out2_p = static_cast<char*>(g.out2) + (rfirst * p.ldo2 + col_base) * 2;
      const long long out2_step = 4 * p.ldo2 * 2;
      const bool has_out2 = g.out2 != nullptr;
      const bool do_cs = (EPI == MOME_EPI_STORE || EPI == MOME_EPI_DGELU) && g.colsum != nullptr;
      // dropout: one 32-bit mask word per lane-iteration (its 4 columns); group index advances by N / 4 * 4 rows = N per iteration
      const uint32_t dkey = DROP ? drop_mix(p.drop_salt, __ldg(p.drop_seed)) : 0u;
      const float dscale = drop_scale(p.drop_thr);
      const uint32_t didx0 = drop_group(g.row0 + rfirst, p.N, col_base);
      const float* rs_p = (EPI == MOME_EPI_RESIDUAL && p.row_scale != nullptr) ? p.row_scale + g.row0 + rfirst : nullptr;

      // The whole per-item epilogue is instantiated twice: FULL = every row / column this warp touches lies inside
      // the group (all tiles but the last row tile of a group), with no per-iteration bounds predicates, and the
      // general version for ragged edges.
      auto epilogue_item = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        // ---- operands the epilogue reads besides the accumulator are requested BEFORE the wait for the MMAs:
        // DGELU: the stashed bf16 gelu'(z) (8 B per lane-iteration), RESIDUAL: the fp32 residual (16 B per
        // lane-iteration): the first two 32-column chunks up front, then two chunks ahead of their use (three deep was measured:
        // 908 against 921 TFLOP/s, the extra registers cost more than the loads' exposed latency).
        constexpr int kAuxDepth = EPI == MOME_EPI_DGELU ? (kChunks < 2 ? kChunks : 2) : 1;
        constexpr int kResDepth = EPI == MOME_EPI_RESIDUAL ? 2 : 1;
        uint2 aux[kAuxDepth][8];
        float4 res[kResDepth][8];
        if (EPI == MOME_EPI_DGELU && !(p.debug & 16)) {
          const char* ap = static_cast<const char*>(g.aux) + (rfirst * p.ldaux + col_base) * 2;
          const long long astep = 4 * p.ldaux * 2;
  #pragma unroll
          for (int c = 0; c < kAuxDepth; ++c) {
  #pragma unroll
            for (int it = 0; it < 8; ++it)
              if (FULL || (it < nvalid && c < nchunks)) aux[c][it] = *reinterpret_cast<const uint2*>(ap + it * astep + c * 64);
          }
        }
        const char* res_p = static_cast<const char*>(static_cast<const void*>(g.res)) + (rfirst * p.ldres + col_base) * 4;
        const long long res_step = 4 * p.ldres * 4;
        if (EPI == MOME_EPI_RESIDUAL) {
  #pragma unroll
          for (int c = 0; c < kResDepth && c < kChunks; ++c) {
  #pragma unroll
            for (int it = 0; it < 8; ++it)
              if (FULL || (it < nvalid && c < nchunks)) res[c][it] = *reinterpret_cast<const float4*>(res_p + it * res_step + c * 128);
          }
        }

        mbar_wait(&tfull_bar[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N + part * kPartCols;
  #pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tacc + c * 32, r);
          tmem_ld_wait();
          if (c == kChunks - 1) {
            // all of this warp's accumulator reads are done: hand the TMEM stage back to the MMA issuer
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
          }
          if (p.debug & 2) continue;  // measurement knob: TMEM drain only
          if (S16) {
            // lane = row: add the bias of the chunk's 32 columns (the same for every lane: broadcast loads), round to
            // bf16 and store the row's 64 bytes as four 16-byte pieces, piece p at p ^ ((row >> 1) & 3)
            const bool has_bias = g.bias != nullptr && (FULL || c < nchunks);
            const float* bp = g.bias + (col_base - c4 + c * 32);
            uint32_t w[16];
  #pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 b4r = make_float4(0.f, 0.f, 0.f, 0.f);
              if (has_bias) b4r = __ldg(reinterpret_cast<const float4*>(bp) + i);
              w[2 * i] = pack_bf16(__uint_as_float(r[4 * i]) + b4r.x, __uint_as_float(r[4 * i + 1]) + b4r.y);
              w[2 * i + 1] = pack_bf16(__uint_as_float(r[4 * i + 2]) + b4r.z, __uint_as_float(r[4 * i + 3]) + b4r.w);
            }
            const uint32_t mine16 = stage_u32 + lane * 64;
            const int sw = (lane >> 1) & 3;
  #pragma unroll
            for (int i = 0; i < 4; ++i) sts_v4_u32(mine16 + ((i ^ sw) << 4), w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
          } else {
            const uint32_t mine = stage_u32 + lane * (kStagePitch * 4);
  #pragma unroll
            for (int i = 0; i < 8; ++i)
              sts_v4_u32(mine + 16 * i, r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          }
          __syncwarp();
          if ((p.debug & 1) || (!FULL && c >= nchunks)) { __syncwarp(); continue; }  // knob: no epilogue math / global IO
          const int col = col_base + c * 32;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), gm4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (!S16 && EPI != MOME_EPI_ATOMIC && EPI != MOME_EPI_DGELU && g.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col));
          if (EPI == MOME_EPI_RESIDUAL && p.gamma != nullptr) gm4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
          float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
          const uint32_t lds_a = stage_u32 + (rsub * kStagePitch + c4) * 4;
          char* op = out_p + c * 32 * osize;
          char* o2p = out2_p + c * 64;
  #pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (FULL || it < nvalid) {
              float4 v;
              if (S16) {  // row rsub + 4 it, this lane's 4 columns = 8 bytes: piece (c4 >> 3) of the row, swizzled as written
                const int row = rsub + 4 * it;
                v = unpack4_bf16(lds_v2_u32(stage_u32 + row * 64 + ((((c4 >> 3) ^ ((row >> 1) & 3)) << 4) | ((c4 & 4) << 1))));
              } else {
                v = lds_v4(lds_a + it * (16 * kStagePitch));
              }
              if (EPI == MOME_EPI_ATOMIC) {
                atomicAdd(reinterpret_cast<float4*>(op), v);
              } else {
                if (!S16 && EPI != MOME_EPI_DGELU) { v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w; }
                if (EPI == MOME_EPI_GELU) {
                  // out = gelu(z), out2 = gelu'(z), both from the fp32 accumulator + bias rounded once to fp16 (the
                  // reference's autocast rounds z to bf16 first; fp16 keeps 3 more bits of it)
                  uint2 u, du;
                  float4 dm = make_float4(1.f, 1.f, 1.f, 1.f);
                  if (DROP) dm = drop_mul4(didx0 + c * 8 + it * p.N, dkey, p.drop_thr, dscale);
                  gelu_fast2<DROP>(v.x, v.y, dm.x, dm.y, u.x, du.x);
                  gelu_fast2<DROP>(v.z, v.w, dm.z, dm.w, u.y, du.y);
                  if (!(p.debug & 8)) *reinterpret_cast<uint2*>(op) = u;
                  if (!(p.debug & 8)) *reinterpret_cast<uint2*>(o2p) = du;
                } else if (EPI == MOME_EPI_RESIDUAL) {
                  // b = bf16(acc + bias) is what the reference's autocast Linear returns; the residual stream stays fp32
                  uint2 bb = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
                  if (DROP) {  // proj_drop / Mlp output dropout acts on the bf16 Linear output
                    const float4 dm = drop_mul4(didx0 + c * 8 + it * p.N, dkey, p.drop_thr, dscale);
                    const float2 t0 = unpack_bf16x2(bb.x), t1 = unpack_bf16x2(bb.y);
                    bb = make_uint2(pack_bf16(t0.x * dm.x, t0.y * dm.y), pack_bf16(t1.x * dm.z, t1.y * dm.w));
                  }
                  if (has_out2) *reinterpret_cast<uint2*>(o2p) = bb;
                  float2 ba = unpack_bf16x2(bb.x), bc = unpack_bf16x2(bb.y);
                  if (rs_p != nullptr) {  // stochastic depth: the whole branch of a dropped sample vanishes
                    const float rsc = __ldg(rs_p + it * 4);
                    ba.x *= rsc; ba.y *= rsc; bc.x *= rsc; bc.y *= rsc;
                  }
                  float4 rr = res[c % kResDepth][it];
                  rr.x = fmaf(gm4.x, ba.x, rr.x); rr.y = fmaf(gm4.y, ba.y, rr.y);
                  rr.z = fmaf(gm4.z, bc.x, rr.z); rr.w = fmaf(gm4.w, bc.y, rr.w);
                  *reinterpret_cast<float4*>(op) = rr;
                } else {
                  if (EPI == MOME_EPI_DGELU) {
                    const float2 a0 = unpack_bf16x2(aux[c % kAuxDepth][it].x), a1 = unpack_bf16x2(aux[c % kAuxDepth][it].y);
                    v.x *= a0.x; v.y *= a0.y; v.z *= a1.x; v.w *= a1.y;
                  }
                  if (osize == 2) {
                    const uint2 ob = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
                    if (!(p.debug & 8)) *reinterpret_cast<uint2*>(op) = ob;
                    if (do_cs) {
                      const float2 oa = unpack_bf16x2(ob.x), oc = unpack_bf16x2(ob.y);
                      cs.x += oa.x; cs.y += oa.y; cs.z += oc.x; cs.w += oc.y;
                    }
                  } else {
                    *reinterpret_cast<float4*>(op) = v;
                    if (do_cs) { cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w; }
                  }
                }
              }
            }
            op += out_step;
            o2p += out2_step;
          }
          if (EPI == MOME_EPI_DGELU && c + kAuxDepth < kChunks && !(p.debug & 16)) {
            const char* ap = static_cast<const char*>(g.aux) + (rfirst * p.ldaux + col_base) * 2;
            const long long astep = 4 * p.ldaux * 2;
  #pragma unroll
            for (int it = 0; it < 8; ++it)
              if (FULL || (it < nvalid && c + kAuxDepth < nchunks))
                aux[c % kAuxDepth][it] = *reinterpret_cast<const uint2*>(ap + it * astep + (c + kAuxDepth) * 64);
          }
          if (EPI == MOME_EPI_RESIDUAL && c + kResDepth < kChunks) {
            // the slot just consumed is refilled with the residual of chunk c + 2
  #pragma unroll
            for (int it = 0; it < 8; ++it)
              if (FULL || (it < nvalid && c + kResDepth < nchunks))
                res[c % kResDepth][it] = *reinterpret_cast<const float4*>(res_p + it * res_step + (c + kResDepth) * 128);
          }
          if (do_cs) {
            // fused bias gradient, stage 1: the stored values summed over this warp's 32 rows go to row
            // (row0 / 32) of the partials buffer (no atomics; mome_colreduce adds the parts)
  #pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
              cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
            }
            if (rsub == 0 && (FULL || row0 < g.M)) *reinterpret_cast<float4*>(g.colsum + (row0 >> 5) * p.N + col) = cs;
          }
          __syncwarp();
        }
      };
      if (row0 + 32 <= static_cast<long long>(g.M) && nchunks == kChunks) epilogue_item(std::true_type{});
      else epilogue_item(std::false_type{});
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer can still touch its barriers / TMEM
  if (warp == 2) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D bf16 tensor [outer][inner] with row pitch ld (elements), 128B-swizzled box {box_inner, box_outer}.
int encode_bf16_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MOME_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple pitch (base %p, ld %lld)", base, (long long)ld);
    return MOME_ERR_ARG;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): inner %lld outer %lld ld %lld box %dx%d", (int)r, (long long)inner,
              (long long)outer, (long long)ld, box_inner, box_outer);
    return MOME_ERR_CUDA;
  }
  return MOME_OK;
}

// ---- event-based profiling of every launch (bench.py roofline) ----
struct ProfRec {
  cudaEvent_t a, b;
  double flops;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, bool DROP = false>
int launch_one(const GemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, epi_warps(EPI), stage16(EPI)>;
  static bool configured = false;
  auto kern = gemm_pair_kernel<BLOCK_N, A_MN, B_MN, EPI, DROP>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem %d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return MOME_ERR_CUDA;
    }
    configured = true;
  }
  kern<<<grid, gemm_threads(EPI), Cfg::SMEM_BYTES, stream>>>(p);
  return check_launch("gemm_pair");
}

// The epilogue is a template parameter; only the (major, epilogue) pairs the block uses are instantiated.
template <int BLOCK_N>
int launch_gemm(const GemmParams& p, bool a_mn, bool b_mn, int grid, cudaStream_t stream) {
  if (!a_mn && !b_mn) {
    switch (p.epilogue) {
      case MOME_EPI_STORE: return launch_one<BLOCK_N, false, false, MOME_EPI_STORE>(p, grid, stream);
      case MOME_EPI_GELU:
        return p.drop_seed ? launch_one<BLOCK_N, false, false, MOME_EPI_GELU, true>(p, grid, stream)
                           : launch_one<BLOCK_N, false, false, MOME_EPI_GELU>(p, grid, stream);
      case MOME_EPI_RESIDUAL:
        return p.drop_seed ? launch_one<BLOCK_N, false, false, MOME_EPI_RESIDUAL, true>(p, grid, stream)
                           : launch_one<BLOCK_N, false, false, MOME_EPI_RESIDUAL>(p, grid, stream);
      case MOME_EPI_ATOMIC: return launch_one<BLOCK_N, false, false, MOME_EPI_ATOMIC>(p, grid, stream);
    }
  } else if (!a_mn && b_mn) {
    switch (p.epilogue) {
      case MOME_EPI_STORE: return launch_one<BLOCK_N, false, true, MOME_EPI_STORE>(p, grid, stream);
      case MOME_EPI_DGELU: return launch_one<BLOCK_N, false, true, MOME_EPI_DGELU>(p, grid, stream);
      case MOME_EPI_ATOMIC: return launch_one<BLOCK_N, false, true, MOME_EPI_ATOMIC>(p, grid, stream);
    }
  } else if (a_mn && b_mn) {
    switch (p.epilogue) {
      case MOME_EPI_STORE: return launch_one<BLOCK_N, true, true, MOME_EPI_STORE>(p, grid, stream);
      case MOME_EPI_ATOMIC: return launch_one<BLOCK_N, true, true, MOME_EPI_ATOMIC>(p, grid, stream);
    }
  }
  set_error("gemm(bf16): operand majors (%d, %d) with epilogue %d are not instantiated", (int)a_mn, (int)b_mn, p.epilogue);
  return MOME_ERR_UNSUPPORTED;
}

}  // namespace

// the same encoder for other translation units (attention_tc.cu); `map` is a CUtensorMap*
int tma_encode_bf16_2d(void* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer) {
  return encode_bf16_2d(static_cast<CUtensorMap*>(map), base, inner, outer, ld, box_inner, box_outer);
}

int gemm_bf16(const MomeGemmArgs* a, cudaStream_t stream) {
  MOME_REQUIRE(a->num_groups >= 1 && a->num_groups <= MOME_MAX_GROUPS, "gemm: num_groups %d", a->num_groups);
  MOME_REQUIRE(a->N > 0 && a->N % 32 == 0, "gemm(bf16): N=%lld must be a positive multiple of 32", (long long)a->N);
  const bool a_mn = a->a_major == 1, b_mn = a->b_major == 1;
  MOME_REQUIRE(!(a_mn && !b_mn), "gemm(bf16): A MN-major with B K-major is not instantiated");
  MOME_REQUIRE(a->epilogue != MOME_EPI_ATOMIC || a->out_dtype == MOME_F32, "gemm: ATOMIC epilogue needs fp32 out");
  MOME_REQUIRE(a->epilogue != MOME_EPI_RESIDUAL || a->out_dtype == MOME_F32, "gemm: RESIDUAL epilogue needs fp32 out");
  MOME_REQUIRE((a->epilogue != MOME_EPI_GELU && a->epilogue != MOME_EPI_DGELU) || a->out_dtype == MOME_BF16,
               "gemm(bf16): GELU/DGELU epilogues write bf16");
  MOME_REQUIRE(a->drop_seed == nullptr || a->drop_p <= 0.f || a->epilogue == MOME_EPI_GELU || a->epilogue == MOME_EPI_RESIDUAL,
               "gemm: dropout is defined for the GELU and RESIDUAL epilogues only");
  MOME_REQUIRE(a->row_scale == nullptr || a->epilogue == MOME_EPI_RESIDUAL, "gemm: row_scale needs the RESIDUAL epilogue");
  MOME_REQUIRE(a->ldo % 8 == 0 && a->ldo2 % 8 == 0 && a->ldres % 4 == 0 && a->ldaux % 8 == 0, "gemm: leading dims must keep rows 16-byte aligned");

  double flops = 0;
  for (int g = 0; g < a->num_groups; ++g) {
    const MomeGemmGroup& s = a->group[g];
    MOME_REQUIRE(s.M > 0 && s.K > 0, "gemm: group %d has M=%lld K=%lld (drop empty groups on the host)", g, (long long)s.M, (long long)s.K);
    flops += 2.0 * double(s.M) * double(a->N) * double(s.K);
  }
  ProfRec rec{};
  bool prof = false;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof = g_prof_on;
  }
  if (prof) {
    cudaEventCreate(&rec.a);
    cudaEventCreate(&rec.b);
    rec.flops = flops;
    cudaEventRecord(rec.a, stream);
  }
  int rc;
  {
    const int block_n = (a->N % 256 == 0) ? 256 : 128;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.num_groups = a->num_groups;
    p.N = static_cast<int>(a->N);
    p.n_tiles = static_cast<int>((a->N + block_n - 1) / block_n);
    p.epilogue = a->epilogue;
    p.out_bf16 = a->out_dtype == MOME_BF16;
    p.ldo = a->ldo; p.ldo2 = a->ldo2; p.ldres = a->ldres; p.ldaux = a->ldaux;
    p.gamma = a->gamma;
    p.row_scale = a->row_scale;
    p.drop_seed = (a->drop_seed != nullptr && a->drop_p > 0.f) ? a->drop_seed : nullptr;
    p.drop_salt = a->drop_salt;
    p.drop_thr = drop_threshold(a->drop_p);
    {
      static const int dbg = [] { const char* e = getenv("MOME_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
      p.debug = dbg;
    }
    const int pairs = std::max(1, sm_count() / 2);
    long long tiles = 0, max_kb = 0;
    for (int g = 0; g < a->num_groups; ++g) {
      const MomeGemmGroup& s = a->group[g];
      tiles += ((s.M + PAIR_M - 1) / PAIR_M) * p.n_tiles;
      max_kb = std::max<long long>(max_kb, (s.K + BLOCK_K - 1) / BLOCK_K);
    }
    int splits = 1;
    if (a->epilogue == MOME_EPI_ATOMIC) {
      if (a->split_k > 0) {
        splits = a->split_k;
      } else {
        // fewest k-splits that fill the machine for >= 2 waves with <= 10 % tail loss (else the best seen)
        const int max_splits = static_cast<int>(std::max<long long>(1, max_kb / 4));
        double best_eff = -1.0;
        for (int s = 1; s <= std::min(max_splits, 64); ++s) {
          const long long items = tiles * s;
          const long long waves = (items + pairs - 1) / pairs;
          const double eff = double(items) / double(waves * pairs);
          if (eff > best_eff + 1e-9) { best_eff = eff; splits = s; }
          if (waves >= 2 && eff >= 0.9) { splits = s; break; }
        }
      }
      splits = std::max(1, std::min<int>(splits, static_cast<int>(std::max<long long>(1, max_kb))));
    }
    p.splits = splits;
    int item = 0;
    rc = MOME_OK;
    for (int g = 0; g < a->num_groups && rc == MOME_OK; ++g) {
      const MomeGemmGroup& s = a->group[g];
      GemmGroupDev& d = p.g[g];
      d.out = s.out; d.out2 = s.out2; d.bias = s.bias; d.res = s.res; d.aux = s.aux; d.colsum = s.colsum; d.row0 = s.row0;
      d.M = static_cast<int>(s.M);
      d.k_blocks = static_cast<int>((s.K + BLOCK_K - 1) / BLOCK_K);
      d.item_start = item;
      item += static_cast<int>((s.M + PAIR_M - 1) / PAIR_M) * p.n_tiles * splits;
      // K-major operand: tensor [rows][K], box {64 (K), rows per CTA}; MN-major: tensor [K][rows], box {64 (MN), 64 (K)}
      rc = a_mn ? encode_bf16_2d(&p.tma_a[g], s.a, s.M, s.K, a->lda, 64, BLOCK_K)
                : encode_bf16_2d(&p.tma_a[g], s.a, s.K, s.M, a->lda, BLOCK_K, CTA_M);
      if (rc != MOME_OK) break;
      rc = b_mn ? encode_bf16_2d(&p.tma_b[g], s.b, a->N, s.K, a->ldb, 64, BLOCK_K)
                : encode_bf16_2d(&p.tma_b[g], s.b, s.K, a->N, a->ldb, BLOCK_K, block_n / 2);
    }
    if (rc == MOME_OK) {
      p.total_items = item;
      const int grid = 2 * std::min(item, pairs);
      rc = block_n == 256 ? launch_gemm<256>(p, a_mn, b_mn, grid, stream) : launch_gemm<128>(p, a_mn, b_mn, grid, stream);
    }
  }
  if (prof) {
    cudaEventRecord(rec.b, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(rec);
  }
  return rc;
}

}  // namespace mome

namespace mome {
// block.cu: while per-launch profiling is on, every GEMM stays on the caller's stream (a launch timed next to concurrent kernels
// would measure the sharing, not the kernel)
bool gemm_prof_active() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  return g_prof_on;
}
}  // namespace mome

extern "C" int mome_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(mome::g_prof_mu);
  mome::g_prof_on = on != 0;
  return MOME_OK;
}

extern "C" int mome_prof_read(int64_t* launches, double* ms, double* flops, int reset) {
  std::lock_guard<std::mutex> lk(mome::g_prof_mu);
  int64_t n = 0;
  double t = 0, f = 0;
  for (auto& r : mome::g_prof) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    float dt = 0;
    if (cudaEventElapsedTime(&dt, r.a, r.b) == cudaSuccess) {
      t += dt;
      f += r.flops;
      ++n;
    }
  }
  if (reset) {
    for (auto& r : mome::g_prof) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    mome::g_prof.clear();
  }
  if (launches) *launches = n;
  if (ms) *ms = t;
  if (flops) *flops = f;
  return MOME_OK;
}
