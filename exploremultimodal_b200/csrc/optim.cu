// Flat fused AdamW: one launch updates a whole flat fp32 parameter buffer (or this rank's shard of
// it) from the matching flat gradient buffer. Per-element hyper-parameters come from a group id per element and two
// small device tables (learning rate, weight decay per group), which is how the reference's three-tier parameter
// groups (utils/optim_factory.py:22-90: bottom / fusion / head layers x decay / no_decay) share one kernel.
// The gradient-clipping factor and the step count are device scalars, so a step is capturable in a CUDA graph.
//
// Replaces: apex FusedAdam / DeepSpeed Adam(adam_w_mode) (reference utils/optim_factory.py:186, conf/config.yaml:86-94)
// and clip_grad_norm_'s scaling pass (train/pretrain/multimodal.py:311-330). HBM-bound: 16 B read + 12 B written per element.
#include <algorithm>

#include "common.cuh"

namespace mome {

__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, const uint8_t* __restrict__ gid,
                                                          const float* __restrict__ lr_tab, const float* __restrict__ wd_tab,
                                                          const float* __restrict__ step, const float* __restrict__ gscale, float beta1,
                                                          float beta2, float eps, long long n) {
  const float t = __ldg(step);
  const float c1 = 1.f / (1.f - powf(beta1, t)), c2 = 1.f / (1.f - powf(beta2, t));
  const float gs = gscale != nullptr ? __ldg(gscale) : 1.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 pp = *reinterpret_cast<float4*>(p + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      const uchar4 id = *reinterpret_cast<const uchar4*>(gid + i);
      const float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x * gs, gg.y * gs, gg.z * gs, gg.w * gs};
      float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w}, out[4];
      const unsigned char ids[4] = {id.x, id.y, id.z, id.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float lr = __ldg(lr_tab + ids[k]), wd = __ldg(wd_tab + ids[k]);
        ma[k] = beta1 * ma[k] + (1.f - beta1) * ga[k];
        va[k] = beta2 * va[k] + (1.f - beta2) * ga[k] * ga[k];
        const float upd = (ma[k] * c1) / (sqrtf(va[k] * c2) + eps);
        out[k] = pa[k] * (1.f - lr * wd) - lr * upd;  // decoupled weight decay (AdamW), as torch.optim.AdamW
      }
      *reinterpret_cast<float4*>(p + i) = make_float4(out[0], out[1], out[2], out[3]);
      *reinterpret_cast<float4*>(m + i) = make_float4(ma[0], ma[1], ma[2], ma[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(va[0], va[1], va[2], va[3]);
    } else {
      for (long long j = i; j < n; ++j) {
        const float lr = __ldg(lr_tab + gid[j]), wd = __ldg(wd_tab + gid[j]);
        const float gj = g[j] * gs;
        const float mj = beta1 * m[j] + (1.f - beta1) * gj, vj = beta2 * v[j] + (1.f - beta2) * gj * gj;
        m[j] = mj;
        v[j] = vj;
        p[j] = p[j] * (1.f - lr * wd) - lr * (mj * c1) / (sqrtf(vj * c2) + eps);
      }
    }
  }
}

// sum of squares of a flat buffer into out[0] (+=), for the global gradient norm
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  float s = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 a = *reinterpret_cast<const float4*>(x + i);
      s += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
    } else {
      for (long long j = i; j < n; ++j) s += x[j] * x[j];
    }
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

}  // namespace mome

using namespace mome;

extern "C" int mome_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const uint8_t* group_id,
                               const float* lr_table, const float* wd_table, const float* step, const float* grad_scale, float beta1,
                               float beta2, float eps, int64_t n, void* stream) {
  MOME_REQUIRE(param && grad && exp_avg && exp_avg_sq && group_id && lr_table && wd_table && step, "adamw_flat: null argument");
  MOME_REQUIRE((reinterpret_cast<uintptr_t>(param) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(exp_avg) & 15) == 0 && (reinterpret_cast<uintptr_t>(exp_avg_sq) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(group_id) & 3) == 0,
               "adamw_flat: buffers must be 16-byte aligned (group ids 4-byte)");
  if (n == 0) return MOME_OK;
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, sm_count() * 8LL)));
  adamw_flat_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, group_id, lr_table, wd_table, step,
                                                                          grad_scale, beta1, beta2, eps, n);
  return check_launch("adamw_flat");
}

extern "C" int mome_sumsq(const float* x, int64_t n, float* out, void* stream) {
  MOME_REQUIRE(x != nullptr && out != nullptr && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "sumsq: null or unaligned argument");
  if (n == 0) return MOME_OK;
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, sm_count() * 4LL)));
  sumsq_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out);
  return check_launch("sumsq");
}
