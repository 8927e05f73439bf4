// Micro-probe: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N, of where A
// comes from (shared memory descriptor vs TMEM) and of how many independent accumulators the stream alternates
// between. One CTA, one issuing thread; operands are zeros (only the timing matters).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu && tools/umma_probe
#include <cstdio>
#include <cuda_runtime.h>

#include "../exploremultimodal_b200/csrc/ptx.cuh"

using namespace mome;

struct Cfg {
  int n, ts, accs, b_mn, reps;
};

__global__ void __launch_bounds__(128) probe(const Cfg* cfgs, int ncfg, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32 && elect_one()) {  // elect.sync lets the compiler keep the MMA operands in uniform registers
    uint32_t phase = 0;
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    for (int c = 0; c < ncfg; ++c) {
      const Cfg g = cfgs[c];
      const uint32_t idesc = umma_idesc_bf16(128, g.n, false, g.b_mn != 0);
      for (int rep = 0; rep < 2; ++rep) {  // second repetition is the one reported (warm)
        const long long t0 = clock64();
        for (int i = 0; i < g.reps; ++i) {
          // accumulator i % accs; A in TMEM lives in columns [384, 512), D in [0, 256) (+ 256 for the second one when it fits)
          const uint32_t d = tm + (i % g.accs) * (g.n <= 128 ? 128 : 0) + ((i % g.accs) && g.n > 128 ? 256 : 0);
          const uint64_t bd = g.b_mn ? umma_smem_desc(b + (i & 7) * 2048, 8192, 1024) : umma_smem_desc(b + (i & 3) * 32, 0, 1024);
          if (g.ts) umma_bf16_ts(d, tm + 384 + (i & 7) * 8, bd, idesc, 1u);
          else umma_bf16(d, umma_smem_desc(a + (i & 3) * 32, 0, 1024), bd, idesc, 1u);
        }
        umma_commit(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, phase);
        phase ^= 1;
        const long long t2 = clock64();
        out[c * 2 + 0] = t1 - t0;
        out[c * 2 + 1] = t2 - t0;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

// TMEM read throughput: `nw` warps each issue `reps` tcgen05.ld (32 lanes x 32 columns = 4 KB, or x 16 = 2 KB) back to back.
template <int COLS>
__global__ void __launch_bounds__(512) ld_probe(int nw, int reps, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nw) {
    for (int i = 0; i < reps; ++i) {
      if (COLS == 32) {
        uint32_t r[32];
        tmem_ld_32x32(tm + ((i * 32) & 255), r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) acc ^= r[c];
      } else {
        uint32_t r[16];
        tmem_ld_32x16(tm + ((i * 16) & 255), r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) acc ^= r[c];
      }
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[63] = acc;
  if (threadIdx.x == 0) out[0] = t1 - t0;
  __syncthreads();
  if (threadIdx.x == 32 * (nw - 1)) out[1] = t1 - t0;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  const Cfg h[] = {
      {64, 0, 1, 1, 16},  {64, 0, 1, 1, 64},  {64, 0, 2, 1, 64},  {64, 1, 1, 1, 16},  {64, 1, 1, 1, 64},  {64, 1, 2, 1, 64},
      {64, 0, 1, 0, 64},  {64, 1, 1, 0, 64},  {128, 0, 1, 1, 64}, {128, 1, 1, 1, 64}, {128, 0, 1, 0, 64}, {240, 0, 1, 0, 16},
      {240, 0, 1, 0, 64}, {256, 0, 1, 0, 64}, {256, 1, 1, 0, 64}, {240, 0, 2, 0, 64}, {32, 0, 1, 1, 64},  {16, 0, 1, 0, 64},
  };
  const int n = sizeof(h) / sizeof(h[0]);
  Cfg* d;
  long long* o;
  cudaMalloc(&d, sizeof(h));
  cudaMalloc(&o, n * 16);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  probe<<<1, 128, 100 * 1024>>>(d, n, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("probe failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  long long r[64];
  cudaMemcpy(r, o, n * 16, cudaMemcpyDeviceToHost);
  printf("%5s %3s %4s %5s %5s %12s %12s %10s\n", "N", "A", "accs", "B", "reps", "issue cyc", "total cyc", "cyc/mma");
  for (int i = 0; i < n; ++i)
    printf("%5d %3s %4d %5s %5d %12lld %12lld %10.1f\n", h[i].n, h[i].ts ? "TS" : "SS", h[i].accs, h[i].b_mn ? "MN" : "K", h[i].reps, r[2 * i],
           r[2 * i + 1], double(r[2 * i + 1]) / h[i].reps);
  printf("\nTMEM read: warps x reps of tcgen05.ld.32x32b (each followed by wait::ld)\n%6s %5s %5s %12s %14s\n", "cols", "warps", "reps", "cycles", "B/clk (SM)");
  for (int cols : {32, 16}) {
    for (int nw : {1, 4, 8, 16}) {
      const int reps = 256;
      if (cols == 32) ld_probe<32><<<1, 512>>>(nw, reps, o);
      else ld_probe<16><<<1, 512>>>(nw, reps, o);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("ld_probe failed\n"); return 1; }
      cudaMemcpy(r, o, 16, cudaMemcpyDeviceToHost);
      printf("%6d %5d %5d %12lld %14.1f\n", cols, nw, reps, r[1], double(nw) * reps * 32 * cols * 4 / double(r[1]));
    }
  }
  return 0;
}
