// Micro-probe: tcgen05.ld (TMEM -> registers) throughput per SM, and tcgen05.mma SS cost at N = 32 / 64 / 96 (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu && tools/tmem_probe
// Question behind it (profiles/r2_chain.md): can the epilogue afford to read 96 accumulator columns per 128-row block
// (horizontal taps stacked along N) instead of 32?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld_cols(uint32_t taddr, uint32_t &sink) {
  if constexpr (X == 32) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
    for (int k = 0; k < 32; ++k) sink ^= v[k];
  } else {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
    for (int k = 0; k < 8; ++k) sink ^= v[k];
  }
}

// nwarps warps (multiples of 4: warp w reads lane quarter w % 4) each read `cols` columns per iteration
template <int X>
__global__ void __launch_bounds__(512) ld_probe(int nwarps, int cols, int iters, long long *cycles, uint32_t *out) {
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
      for (int c = 0; c < cols; c += X) ld_cols<X>(lane_base + (uint32_t)((c + (warp >> 2) * 128) & 511), sink);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  out[tid] = sink;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long *cyc;
  uint32_t *out;
  cudaMalloc(&cyc, sizeof(long long));
  cudaMalloc(&out, 512 * sizeof(uint32_t));
  for (int x : {32, 8})
    for (int nw : {4, 8, 16}) {
      const int cols = 96, iters = 2000;
      if (x == 32) ld_probe<32><<<1, 512>>>(nw, cols, iters, cyc, out);
      else ld_probe<8><<<1, 512>>>(nw, cols, iters, cyc, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      const double bytes = (double)nw * 32 * cols * 4 * iters;
      printf("tcgen05.ld 32x32b.x%-2d  %2d warps x %d columns: %.1f cycles per warp-pass, %.1f B/cycle per SM\n", x, nw, cols,
             (double)c / iters, bytes / (double)c);
    }
  return 0;
}
