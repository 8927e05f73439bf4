// fp32 validation path of mome_gemm: a plain CUDA-core tiled GEMM with the same operand-major and
// epilogue semantics as the tcgen05 kernel. It exists so that the whole block can be checked
// against the fp32 oracle at 1e-4 (north_star's fp32 tolerance); it is not a performance path.
#include "common.cuh"
#include "ptx.cuh"

namespace mome {

int gemm_bf16(const MomeGemmArgs* a, cudaStream_t stream);  // gemm_tcgen05.cu

struct SimtGroup {
  const float* a;
  const float* b;
  float* out;
  float* out2;
  const float* bias;
  const float* res;
  const float* aux;
  float* colsum;
  int M, K;
};
struct SimtParams {
  SimtGroup g[MOME_MAX_GROUPS];
  const float* gamma;
  long long sam, sak, sbn, sbk, ldo, ldo2, ldres, ldaux;
  int N, epilogue;
};

constexpr int TS = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(const SimtParams p) {
  const SimtGroup& g = p.g[blockIdx.z];
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  if (m0 >= g.M) return;
  __shared__ float As[TK][TS + 1];
  __shared__ float Bs[TK][TS + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < g.K; k0 += TK) {
    for (int i = threadIdx.x; i < TS * TK; i += 256) {
      // keep the fastest-varying index on the contiguous dimension of each operand
      int mm, kk;
      if (p.sak == 1) { kk = i % TK; mm = i / TK; } else { mm = i % TS; kk = i / TS; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < g.K) ? g.a[m * p.sam + k * p.sak] : 0.f;
      int nn;
      if (p.sbk == 1) { kk = i % TK; nn = i / TK; } else { nn = i % TS; kk = i / TS; }
      const int n = n0 + nn;
      const int k2 = k0 + kk;
      Bs[kk][nn] = (n < p.N && k2 < g.K) ? g.b[n * p.sbn + k2 * p.sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; bv[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.epilogue == MOME_EPI_ATOMIC) { atomicAdd(g.out + m * p.ldo + n, v); continue; }
      if (g.bias) v += g.bias[n];
      switch (p.epilogue) {
        case MOME_EPI_GELU:
          g.out2[m * p.ldo2 + n] = gelu_erf_grad(v);
          g.out[m * p.ldo + n] = gelu_erf(v);
          break;
        case MOME_EPI_RESIDUAL:
          if (g.out2) g.out2[m * p.ldo2 + n] = v;
          g.out[m * p.ldo + n] = g.res[m * p.ldres + n] + (p.gamma ? p.gamma[n] : 1.f) * v;
          break;
        case MOME_EPI_DGELU:
          v *= g.aux[m * p.ldaux + n];
          g.out[m * p.ldo + n] = v;
          if (g.colsum) atomicAdd(g.colsum + (m >> 5) * p.N + n, v);
          break;
        default:
          g.out[m * p.ldo + n] = v;
          if (g.colsum) atomicAdd(g.colsum + (m >> 5) * p.N + n, v);
      }
    }
  }
}

static int gemm_f32(const MomeGemmArgs* a, cudaStream_t stream) {
  MOME_REQUIRE(a->out_dtype == MOME_F32, "gemm(fp32): out must be fp32");
  if ((a->drop_seed != nullptr && a->drop_p > 0.f) || a->row_scale != nullptr) {
    set_error("gemm(fp32): dropout / stochastic depth exist on the bf16 path only");
    return MOME_ERR_UNSUPPORTED;
  }
  MOME_REQUIRE(a->num_groups >= 1 && a->num_groups <= MOME_MAX_GROUPS, "gemm: num_groups %d", a->num_groups);
  SimtParams p;
  memset(&p, 0, sizeof(p));
  long long max_m = 0;
  for (int g = 0; g < a->num_groups; ++g) {
    const MomeGemmGroup& s = a->group[g];
    MOME_REQUIRE(s.M > 0 && s.K > 0, "gemm: group %d has M=%lld K=%lld", g, (long long)s.M, (long long)s.K);
    p.g[g] = SimtGroup{static_cast<const float*>(s.a), static_cast<const float*>(s.b), static_cast<float*>(s.out),
                       static_cast<float*>(s.out2), s.bias, s.res, static_cast<const float*>(s.aux), s.colsum, (int)s.M, (int)s.K};
    max_m = std::max<long long>(max_m, s.M);
  }
  p.gamma = a->gamma;
  p.sam = a->a_major == 0 ? a->lda : 1;
  p.sak = a->a_major == 0 ? 1 : a->lda;
  p.sbn = a->b_major == 0 ? a->ldb : 1;
  p.sbk = a->b_major == 0 ? 1 : a->ldb;
  p.ldo = a->ldo; p.ldo2 = a->ldo2; p.ldres = a->ldres; p.ldaux = a->ldaux;
  p.N = (int)a->N;
  p.epilogue = a->epilogue;
  dim3 grid((unsigned)((a->N + TS - 1) / TS), (unsigned)((max_m + TS - 1) / TS), (unsigned)a->num_groups);
  gemm_f32_kernel<<<grid, 256, 0, stream>>>(p);
  return check_launch("gemm_f32");
}

}  // namespace mome

extern "C" int mome_gemm(const MomeGemmArgs* args, void* stream) {
  MOME_REQUIRE(args != nullptr, "gemm: null args");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (args->dtype == MOME_BF16) return mome::gemm_bf16(args, s);
  if (args->dtype == MOME_F32) return mome::gemm_f32(args, s);
  mome::set_error("gemm: unknown dtype %d", args->dtype);
  return MOME_ERR_ARG;
}
