// Issue-rate probe for the wgrad MMA sequence: MN-major A (shift-stacked, LBO = one pixel row) and B, M = 128, K = 16,
// nacc accumulators visited round-robin with A row offsets of `urows`, ksteps K-steps per "tile", repeated `tiles` times.
// Reports cycles per MMA for several (RBx, RBy, N, nacc, urows) and for K-major descriptors of the same footprint
// (timing only; the K-major results are not checked).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_mn_rate tools/umma_mn_rate.cu && ./umma_mn_rate
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args {
  int RBx, RBy, N, nacc, urows, ksteps, tiles, kmajor, mstack;   // mstack: LBO of A = RBx (1) or a far block (0)
  long long *cycles;
};

__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xFFFFFFFF;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128) probe(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t x_base = base, y_base = base + 120 * 1024;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem + (base - smem_u32(smem)))[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t major = a.kmajor ? 0u : ((1u << 15) | (1u << 16));
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | major | ((uint32_t)(a.N >> 3) << 17) | (8u << 24);
    auto lt = [](int RB) { return RB == 128 ? 2u : (RB == 64 ? 4u : 6u); };
    const uint64_t a_hi = (uint64_t)(((uint32_t)(8 * a.RBx) >> 4) | (1u << 14) | (lt(a.RBx) << 29)) << 32;
    const uint64_t b_hi = (uint64_t)(((uint32_t)(8 * a.RBy) >> 4) | (1u << 14) | (lt(a.RBy) << 29)) << 32;
    const uint32_t a_lbo = (a.kmajor ? 1u : (a.mstack ? ((uint32_t)a.RBx >> 4) : (uint32_t)(40 * 1024 >> 4))) << 16;
    const uint32_t b_lbo = 1u << 16;
    const uint32_t a_kstep = a.kmajor ? 2u : ((uint32_t)(16 * a.RBx) >> 4);     // K-major: 32 bytes along the row
    const uint32_t b_kstep = a.kmajor ? 2u : ((uint32_t)(16 * a.RBy) >> 4);
    const uint32_t a_ustep = ((uint32_t)a.urows * a.RBx) >> 4;
    const long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < a.tiles; ++t) {
#pragma unroll 1
      for (int ks = 0; ks < a.ksteps; ++ks) {
        const int kk = a.kmajor ? (ks & 1) : ks;
        const uint32_t a_ks = a_lbo | ((x_base >> 4) + (uint32_t)kk * a_kstep), b_ks = b_lbo | ((y_base >> 4) + (uint32_t)kk * b_kstep);
        uint32_t a_u = a_ks, dcol = tmem;
#pragma unroll 1
        for (int u = 0; u < a.nacc; ++u, a_u += a_ustep, dcol += (uint32_t)a.N)
          mma(dcol, a_hi | (uint64_t)a_u, b_hi | (uint64_t)b_ks, idesc, 1u);
      }
    }
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xFFFFFFFF;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(
            smem_u32(&bar))
        : "memory");
    const long long t1 = clock64();
    if (tid == 0) a.cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long *cyc;
  cudaMalloc(&cyc, sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  struct Cfg { int RBx, RBy, N, nacc, urows, kmajor, mstack; };
  const Cfg cfgs[] = {
      {64, 64, 32, 1, 0, 0, 1},  {64, 64, 32, 3, 50, 0, 1}, {64, 64, 32, 3, 48, 0, 1}, {64, 64, 32, 3, 50, 0, 0},
      {64, 64, 32, 3, 50, 1, 1}, {128, 128, 64, 3, 26, 0, 1}, {128, 128, 64, 6, 26, 0, 1}, {128, 128, 64, 3, 26, 1, 1},
      {128, 64, 32, 6, 50, 0, 1}, {64, 64, 32, 1, 0, 1, 1},  {128, 128, 64, 1, 0, 0, 1},  {128, 128, 64, 1, 0, 0, 0},
      {32, 32, 16, 3, 50, 0, 1},  {128, 128, 64, 3, 24, 0, 1}};
  for (const Cfg &c : cfgs) {
    Args a{c.RBx, c.RBy, c.N, c.nacc, c.urows, 32, 8, c.kmajor, c.mstack, cyc};
    probe<<<1, 128, 210 * 1024>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    long long cy;
    cudaMemcpy(&cy, cyc, sizeof(cy), cudaMemcpyDeviceToHost);
    printf("RBx=%3d RBy=%3d N=%3d nacc=%d urows=%2d %s %s : %.1f cycles/MMA\n", c.RBx, c.RBy, c.N, c.nacc, c.urows,
           c.kmajor ? "K-major " : "MN-major", c.mstack ? "stacked" : "blocked", (double)cy / (32.0 * 8 * c.nacc));
  }
  return 0;
}
