// K2 (first revision, kept for A/B runs with MOME_GEMM_V1=1): persistent, warp-specialised grouped GEMM
// on the 5th-gen tensor cores, one CTA per tile (cta_group::1), 4 epilogue warps, row-per-thread stores.
//
//   out[m, n] = epilogue( sum_k A(m, k) * B(n, k) )   per group (expert segment of the packed tokens)
//
// One CTA per SM, 256 threads: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected
// lane), warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers -> fused epilogue ->
// global). Operands are staged by TMA into a 4/6-deep ring of SWIZZLE_128B shared-memory tiles;
// accumulators (128 x BLOCK_N fp32) live in TMEM, double-buffered so that the epilogue of work item
// i overlaps the MMAs of item i+1. Both operands may be K-major or MN-major (transpose bits of the
// instruction descriptor), which is how forward, dgrad and wgrad all run on this one kernel without
// materialising transposed copies (see include/mome.h). wgrad uses split-K with fp32 red.add.
//
// Replaces: F.linear / timm Mlp / residual adds of the reference Block (vlmo.py:76-78, 96, 190-196).
#include <cuda.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace mome {
namespace v1 {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 256;
constexpr int kAtomBytes = BLOCK_K * 128;  // one 64x64 MN-major box / 64 rows of a K-major tile

struct GemmGroupDev {
  void* out;
  void* out2;
  const float* bias;
  const float* res;
  const void* aux;
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
  long long ldo, ldo2, ldres, ldaux;
  int num_groups, total_items, n_tiles, splits;
  int N, epilogue, out_bf16, pad;
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

// Fused epilogue for 32 consecutive columns of one output row.
__device__ __forceinline__ void epilogue_row32(const GemmParams& p, const GemmGroupDev& g, float (&v)[32], long long row,
                                               int col0) {
  if (p.epilogue == MOME_EPI_ATOMIC) {
    float* o = reinterpret_cast<float*>(g.out) + row * p.ldo + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 4) atomicAdd(reinterpret_cast<float4*>(o + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    return;
  }
  if (g.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(g.bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(b4 + i);
      v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if (p.epilogue == MOME_EPI_GELU) {
    // z is rounded to bf16 first (what an autocast Linear hands to GELU); out = gelu(z), out2 = gelu'(z)
    uint4* z4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out2) + row * p.ldo2 + col0);
    uint4* u4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + row * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t zp[4], up[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float z0 = __bfloat162float(__float2bfloat16_rn(v[8 * i + 2 * j]));
        const float z1 = __bfloat162float(__float2bfloat16_rn(v[8 * i + 2 * j + 1]));
        zp[j] = pack_bf16(gelu_erf_grad(z0), gelu_erf_grad(z1));
        up[j] = pack_bf16(gelu_erf(z0), gelu_erf(z1));
      }
      z4[i] = make_uint4(zp[0], zp[1], zp[2], zp[3]);
      u4[i] = make_uint4(up[0], up[1], up[2], up[3]);
    }
    return;
  }
  if (p.epilogue == MOME_EPI_RESIDUAL) {
    // b = bf16(acc + bias) is what the reference's autocast Linear returns; residual stream stays fp32
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
    if (g.out2 != nullptr) {
      uint4* b4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out2) + row * p.ldo2 + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        b4[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
    }
    const float4* r4 = reinterpret_cast<const float4*>(g.res + row * p.ldres + col0);
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + row * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 r = r4[i];
      float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p.gamma != nullptr) gm = __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + i);
      r.x += gm.x * v[4 * i]; r.y += gm.y * v[4 * i + 1]; r.z += gm.z * v[4 * i + 2]; r.w += gm.w * v[4 * i + 3];
      o4[i] = r;
    }
    return;
  }
  if (p.epilogue == MOME_EPI_DGELU) {
    const uint4* a4 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(g.aux) + row * p.ldaux + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 a = a4[i];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 z = *reinterpret_cast<const __nv_bfloat162*>(&aw[j]);
        v[8 * i + 2 * j] *= __low2float(z);
        v[8 * i + 2 * j + 1] *= __high2float(z);
      }
    }
  }
  // STORE (and the tail of DGELU)
  if (p.out_bf16) {
    uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + row * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o4[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                         pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
  } else {
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + row * p.ldo + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;  // + barriers + alignment slack
};

template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.num_groups; ++g) {
      tma_prefetch_desc(&p.tma_a[g]);
      tma_prefetch_desc(&p.tma_b[g]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const WorkItem w = decode_item(p, item);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + Cfg::A_BYTES;
          if (!A_MN) {
            tma_load_2d(a_dst, &p.tma_a[w.g], &full_bar[stage], kb * BLOCK_K, w.m_tile * BLOCK_M);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d(a_dst + j * kAtomBytes, &p.tma_a[w.g], &full_bar[stage], w.m_tile * BLOCK_M + j * 64, kb * BLOCK_K);
          }
          if (!B_MN) {
            tma_load_2d(b_dst, &p.tma_b[w.g], &full_bar[stage], kb * BLOCK_K, w.n_tile * BLOCK_N);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(b_dst + j * kAtomBytes, &p.tma_b[w.g], &full_bar[stage], w.n_tile * BLOCK_N + j * 64, kb * BLOCK_K);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, A_MN, B_MN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
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
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const WorkItem w = decode_item(p, item);
      if (w.kb0 >= w.kb1) continue;
      const GemmGroupDev& g = p.g[w.g];
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const long long row = static_cast<long long>(w.m_tile) * BLOCK_M + ew * 32 + lane;
      const bool row_ok = row < g.M;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N + c * 32, r);
        tmem_ld_wait();
        const int col0 = w.n_tile * BLOCK_N + c * 32;
        if (row_ok && col0 < p.N) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          epilogue_row32(p, g, v, row, col0);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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
static int encode_bf16_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                          int box_outer) {
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

template <int BLOCK_N, bool A_MN, bool B_MN>
static int launch_gemm(const GemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  static bool configured = false;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem %d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return MOME_ERR_CUDA;
    }
    configured = true;
  }
  kern<<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(p);
  return check_launch("gemm_tcgen05");
}

int gemm_bf16(const MomeGemmArgs* a, cudaStream_t stream) {
  MOME_REQUIRE(a->num_groups >= 1 && a->num_groups <= MOME_MAX_GROUPS, "gemm: num_groups %d", a->num_groups);
  MOME_REQUIRE(a->N > 0 && a->N % 32 == 0, "gemm(bf16): N=%lld must be a positive multiple of 32", (long long)a->N);
  const bool a_mn = a->a_major == 1, b_mn = a->b_major == 1;
  MOME_REQUIRE(!(a_mn && !b_mn), "gemm(bf16): A MN-major with B K-major is not instantiated");
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
  MOME_REQUIRE(a->epilogue != MOME_EPI_ATOMIC || a->out_dtype == MOME_F32, "gemm: ATOMIC epilogue needs fp32 out");
  MOME_REQUIRE(a->epilogue != MOME_EPI_RESIDUAL || a->out_dtype == MOME_F32, "gemm: RESIDUAL epilogue needs fp32 out");
  MOME_REQUIRE((a->epilogue != MOME_EPI_GELU && a->epilogue != MOME_EPI_DGELU) || a->out_dtype == MOME_BF16,
               "gemm(bf16): GELU/DGELU epilogues write bf16");
  MOME_REQUIRE(a->ldo % 8 == 0 && a->ldo2 % 8 == 0 && a->ldres % 4 == 0 && a->ldaux % 8 == 0, "gemm: leading dims must keep rows 16-byte aligned");

  long long tiles = 0, max_kb = 0;
  for (int g = 0; g < a->num_groups; ++g) {
    const MomeGemmGroup& s = a->group[g];
    MOME_REQUIRE(s.M > 0 && s.K > 0, "gemm: group %d has M=%lld K=%lld (drop empty groups on the host)", g, (long long)s.M, (long long)s.K);
    tiles += ((s.M + BLOCK_M - 1) / BLOCK_M) * p.n_tiles;
    max_kb = std::max<long long>(max_kb, (s.K + BLOCK_K - 1) / BLOCK_K);
  }
  int splits = 1;
  if (a->epilogue == MOME_EPI_ATOMIC) {
    splits = a->split_k > 0 ? a->split_k : static_cast<int>((2 * sm_count() + tiles - 1) / tiles);
    splits = std::max(1, std::min<int>(splits, static_cast<int>(std::max<long long>(1, max_kb / 4))));
  }
  p.splits = splits;
  int item = 0;
  for (int g = 0; g < a->num_groups; ++g) {
    const MomeGemmGroup& s = a->group[g];
    GemmGroupDev& d = p.g[g];
    d.out = s.out; d.out2 = s.out2; d.bias = s.bias; d.res = s.res; d.aux = s.aux;
    d.M = static_cast<int>(s.M);
    d.k_blocks = static_cast<int>((s.K + BLOCK_K - 1) / BLOCK_K);
    d.item_start = item;
    item += static_cast<int>((s.M + BLOCK_M - 1) / BLOCK_M) * p.n_tiles * splits;
    int rc;
    // K-major operand: tensor [rows][K], box {64 (K), tile rows}; MN-major: tensor [K][rows], box {64 (MN), 64 (K)}
    rc = a_mn ? encode_bf16_2d(&p.tma_a[g], s.a, s.M, s.K, a->lda, 64, BLOCK_K)
              : encode_bf16_2d(&p.tma_a[g], s.a, s.K, s.M, a->lda, BLOCK_K, BLOCK_M);
    if (rc != MOME_OK) return rc;
    rc = b_mn ? encode_bf16_2d(&p.tma_b[g], s.b, a->N, s.K, a->ldb, 64, BLOCK_K)
              : encode_bf16_2d(&p.tma_b[g], s.b, s.K, a->N, a->ldb, BLOCK_K, block_n);
    if (rc != MOME_OK) return rc;
  }
  p.total_items = item;
  const int grid = std::min(item, sm_count());

  int rc;
  if (block_n == 256) {
    rc = a_mn ? launch_gemm<256, true, true>(p, grid, stream)
              : (b_mn ? launch_gemm<256, false, true>(p, grid, stream) : launch_gemm<256, false, false>(p, grid, stream));
  } else {
    rc = a_mn ? launch_gemm<128, true, true>(p, grid, stream)
              : (b_mn ? launch_gemm<128, false, true>(p, grid, stream) : launch_gemm<128, false, false>(p, grid, stream));
  }
  return rc;
}

}  // namespace v1
}  // namespace mome
