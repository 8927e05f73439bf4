// Counter-based dropout masks shared by the GEMM epilogues, the LayerScale backward kernels and the
// attention kernels. A mask bit depends only on (seed, salt, element index), so the backward regenerates
// the forward's mask instead of storing it, whatever the thread mapping of the kernel that asks.
//
//   key  = mix(salt, *seed)                     seed: device scalar (bumped once per step, also under a CUDA
//                                               graph); salt: host constant naming the call site
//   word = mix(group_index, key)                one 32-bit word serves 4 consecutive elements of a row (matrix
//                                               dropouts) or 2 adjacent keys of one query (attention)
//   keep = byte >= thr,  thr = round(256 p)     i.e. the drop probability is quantised to 1/256 (p = 0.1 ->
//                                               26/256 = 0.1016) and kept values are scaled by 256 / (256 - thr)
// mix is the murmur3 finaliser; tests/helpers.py restates it in numpy for the oracle-with-masks tests.
#pragma once
#include <stdint.h>

namespace mome {

__host__ __device__ __forceinline__ uint32_t drop_mix(uint32_t idx, uint32_t key) {
  uint32_t h = idx * 0x9E3779B1u + key;
  h ^= h >> 16; h *= 0x85ebca6bu;
  h ^= h >> 13; h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  const int t = static_cast<int>(p * 256.f + 0.5f);
  return static_cast<uint32_t>(t < 0 ? 0 : (t > 255 ? 255 : t));
}
__host__ __device__ __forceinline__ float drop_scale(uint32_t thr) { return 256.f / static_cast<float>(256u - thr); }

// multipliers (0 or scale) of 4 consecutive elements whose group index is `idx`
__device__ __forceinline__ float4 drop_mul4(uint32_t idx, uint32_t key, uint32_t thr, float scale) {
  const uint32_t h = drop_mix(idx, key);
  return make_float4((h & 255u) >= thr ? scale : 0.f, ((h >> 8) & 255u) >= thr ? scale : 0.f, ((h >> 16) & 255u) >= thr ? scale : 0.f,
                     (h >> 24) >= thr ? scale : 0.f);
}
// group index of the 4 elements starting at (row, col) of a row-major matrix with `cols` columns (col % 4 == 0)
__device__ __forceinline__ uint32_t drop_group(long long row, long long cols, long long col) {
  return static_cast<uint32_t>((row * cols + col) >> 2);
}

}  // namespace mome
