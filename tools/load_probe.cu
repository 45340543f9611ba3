// Probe: global -> shared streaming throughput on sm_100a for the three candidate load paths of the patch loader.
//   mode 0: cp.async.cg 16 B per thread (LDGSTS), U copies in flight per thread per round
//   mode 1: ld.global.nc.v4 -> st.shared.v4 (registers), U loads in flight per thread per round
//   mode 2: cp.async.bulk (1-D TMA) of CH-byte chunks issued by one thread, NS chunks in flight
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/load_probe tools/load_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int U>
__global__ void k_cpasync(const uint4 *__restrict__ src, size_t n16_per_cta, int rounds_smem16, unsigned long long *sink) {
  extern __shared__ uint4 sm[];
  const uint4 *p = src + (size_t)blockIdx.x * n16_per_cta;
  const int T = blockDim.x;
  for (size_t base = 0; base + (size_t)U * T <= n16_per_cta; base += (size_t)U * T) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + (size_t)u * T + threadIdx.x;
      const uint32_t dst = smem_u32(sm + ((u * T + threadIdx.x) % rounds_smem16));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");      // keep two rounds in flight
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = sm[0].x;
}

template <int U>
__global__ void k_ldg(const uint4 *__restrict__ src, size_t n16_per_cta, int rounds_smem16, unsigned long long *sink) {
  extern __shared__ uint4 sm[];
  const uint4 *p = src + (size_t)blockIdx.x * n16_per_cta;
  const int T = blockDim.x;
  for (size_t base = 0; base + (size_t)U * T <= n16_per_cta; base += (size_t)U * T) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(p + base + (size_t)u * T + threadIdx.x);
#pragma unroll
    for (int u = 0; u < U; ++u) sm[(u * T + threadIdx.x) % rounds_smem16] = v[u];
  }
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = sm[0].x;
}

// the conv loader's pattern: thread -> (pixel = t >> logS, slab = t & (S-1)); src = pixel-major rows of S*16 bytes,
// dst = [slab][pixel][16 B] with an odd pixel count per slab (as in cs_tc.cu) or a 128B-multiple one (PADODD = 0)
template <int U, int LOGS, int PADODD>
__global__ void k_patch(const uint4 *__restrict__ src, size_t n16_per_cta, unsigned long long *sink) {
  extern __shared__ uint4 sm[];
  constexpr int S = 1 << LOGS;
  const uint4 *p = src + (size_t)blockIdx.x * n16_per_cta;
  const int T = blockDim.x, pstep = T >> LOGS, slab = threadIdx.x & (S - 1), p0 = threadIdx.x >> LOGS;
  const int npix = U * pstep, slabpix = PADODD ? (npix | 1) : npix;
  for (size_t base = 0; base + (size_t)U * T <= n16_per_cta; base += (size_t)U * T) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int px = p0 + u * pstep;
      const uint32_t dst = smem_u32(sm + slab * slabpix + px);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p + base + (size_t)px * S + slab) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = sm[0].x;
}

__global__ void k_bulk(const uint8_t *__restrict__ src, size_t bytes_per_cta, int CH, int NS, unsigned long long *sink) {
  extern __shared__ __align__(128) uint8_t smb[];
  __shared__ uint64_t bar[16];
  const uint8_t *p = src + (size_t)blockIdx.x * bytes_per_cta;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    const size_t nch = bytes_per_cta / CH;
    for (size_t c = 0; c < nch + NS; ++c) {
      const int s = (int)(c % NS);
      if (c >= (size_t)NS) {      // wait for the chunk issued NS iterations ago
        const uint32_t par = (uint32_t)(((c / NS) - 1) & 1);
        asm volatile(
            "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(
                smem_u32(&bar[s])),
            "r"(par)
            : "memory");
      }
      if (c < nch) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(CH) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smb + (size_t)s * CH)),
                     "l"(p + c * CH), "r"(CH), "r"(smem_u32(&bar[s]))
                     : "memory");
      }
    }
    sink[blockIdx.x] = smb[0];
  }
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  const size_t total = 2ull << 30;      // 2 GiB source, far larger than L2
  uint8_t *src;
  unsigned long long *sink;
  cudaMalloc(&src, total);
  {
    // pseudo-random fill so that memory compression cannot inflate the numbers
    std::vector<uint32_t> h(64 << 20);
    uint32_t x = 12345;
    for (auto &v : h) { x = x * 1664525u + 1013904223u; v = x; }
    for (size_t off = 0; off < total; off += h.size() * 4) cudaMemcpy(src + off, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  }
  cudaMalloc(&sink, 4096 * 8);
  const int ctas = 148;
  const size_t per_cta = (total / ctas) & ~(size_t)((1 << 20) - 1);
  const int smem = 96 * 1024;
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
#define RUN_CP(U, T)                                                                                                   \
  {                                                                                                                    \
    cudaFuncSetAttribute(k_cpasync<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                             \
    float ms = time_ms([&] { k_cpasync<U><<<ctas, T, smem>>>((const uint4 *)src, per_cta / 16, smem / 16, sink); });   \
    printf("cp.async.cg16  threads=%4d U=%2d (%5.1f KB in flight/SM): %7.1f GB/s\n", T, U, 2.0 * U * T * 16 / 1024.0,   \
           ctas * (double)per_cta / ms / 1e6);                                                                         \
  }
#define RUN_LD(U, T)                                                                                                   \
  {                                                                                                                    \
    cudaFuncSetAttribute(k_ldg<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                                 \
    float ms = time_ms([&] { k_ldg<U><<<ctas, T, smem>>>((const uint4 *)src, per_cta / 16, smem / 16, sink); });       \
    printf("ldg.v4->sts    threads=%4d U=%2d (%5.1f KB in flight/SM): %7.1f GB/s\n", T, U, 1.0 * U * T * 16 / 1024.0,   \
           ctas * (double)per_cta / ms / 1e6);                                                                         \
  }
#define RUN_PT(U, T, LOGS, ODD)                                                                                          \
  {                                                                                                                    \
    cudaFuncSetAttribute(k_patch<U, LOGS, ODD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                    \
    float ms = time_ms([&] { k_patch<U, LOGS, ODD><<<ctas, T, smem>>>((const uint4 *)src, per_cta / 16, sink); });     \
    printf("patch pattern  threads=%4d U=%2d S=%2d odd=%d: %7.1f GB/s\n", T, U, 1 << LOGS, ODD,                          \
           ctas * (double)per_cta / ms / 1e6);                                                                         \
  }
  RUN_PT(8, 256, 2, 1) RUN_PT(8, 256, 2, 0) RUN_PT(8, 256, 3, 1) RUN_PT(8, 256, 3, 0) RUN_PT(8, 256, 4, 1) RUN_PT(8, 256, 4, 0)
  RUN_CP(4, 256) RUN_CP(8, 256) RUN_CP(16, 256) RUN_CP(4, 512) RUN_CP(8, 512) RUN_CP(16, 512) RUN_CP(8, 1024)
  RUN_LD(4, 256) RUN_LD(8, 256) RUN_LD(16, 256) RUN_LD(8, 512) RUN_LD(8, 1024)
  for (int CH : {2048, 8192, 16384, 32768})
    for (int NS : {2, 4, 6}) {
      if ((size_t)CH * NS > 190 * 1024) continue;
      float ms = time_ms([&] { k_bulk<<<ctas, 32, (size_t)CH * NS>>>(src, per_cta, CH, NS, sink); });
      printf("cp.async.bulk  chunk=%5d B x %d in flight (%5.1f KB/SM): %7.1f GB/s\n", CH, NS, CH * NS / 1024.0,
             ctas * (double)per_cta / ms / 1e6);
    }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
