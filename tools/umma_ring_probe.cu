// Probe for the row-streamed convolution kernel (csrc/cs_tc_rs.cu): do back-to-back tcgen05.mma instructions whose
// accumulator COLUMN RANGES OVERLAP PARTIALLY execute in issue order?  The kernel stacks the three kernel rows along N and
// lets the tensor core do the shift-add: the MMA of input row y writes tensor-memory columns [slot(y-2), slot(y-1), slot(y)],
// the MMA of row y+1 writes [slot(y-1), slot(y), slot(y+1)], ... -- every slot is accumulated by three different
// instructions with three different accumulator base addresses.  Each slot is zeroed by an MMA against a zero B operand
// (accumulate = 0, N = Cp) issued right before the first real MMA that touches it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ring_probe tools/umma_ring_probe.cu && ./umma_ring_probe
// Checks Y[y'] = sum_j X[y'+2-j] * B_j exactly (integer data) for Cp = 32 / 64, with and without ring wrap-around, and
// reports cycles per input row.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args {
  int Cp;        // output channels per part (slot width in columns)
  int R;         // input rows streamed
  int ksteps;    // K = 16 steps per row
  int zero_by_mma;   // 1: zero each slot with an MMA against zeros; 0: first real MMA split into acc=0 / acc=1 pieces
  float *D;      // [128][512] dump of the whole tensor memory
  long long *cycles;
};

__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xFFFFFFFF;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

__host__ __device__ inline int xval(int p, int k) { return ((p * 7 + k * 3) % 13) - 6; }
__host__ __device__ inline int bval(int n, int k) { return ((n * 5 + k) % 7) - 3; }

__global__ void __launch_bounds__(128) probe(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t *gen = smem + (base - smem_u32(smem));
  const int rows = 128 + a.R, Nt = 3 * a.Cp, K = 16 * a.ksteps;
  const uint32_t a_base = base, b_base = base + 64 * 1024, z_base = base + 128 * 1024;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // A: [k8][rows][8], B: [k8][Nt][8] (no swizzle), zeros: 8 KB
  for (int i = tid; i < rows * K; i += 128) {
    const int p = i / K, k = i % K;
    reinterpret_cast<__nv_bfloat16 *>(gen)[(size_t)(k / 8) * rows * 8 + p * 8 + (k % 8)] = __float2bfloat16((float)xval(p, k));
  }
  for (int i = tid; i < Nt * K; i += 128) {
    const int n = i / K, k = i % K;
    reinterpret_cast<__nv_bfloat16 *>(gen + 64 * 1024)[(size_t)(k / 8) * Nt * 8 + n * 8 + (k % 8)] = __float2bfloat16((float)bval(n, k));
  }
  for (int i = tid; i < 8 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(gen + 128 * 1024)[i] = 0u;
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
  // fill the tensor memory with a sentinel so that a missing zeroing shows up
  for (int n0 = 0; n0 < 512; n0 += 16) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
            tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0),
        "r"(__float_as_uint(1000.f)));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) {
    const int NS = 512 / a.Cp;
    const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint64_t a_fix = hi | ((uint64_t)(((uint32_t)rows * 16) >> 4) << 16);
    const uint64_t b_fix = hi | ((uint64_t)(((uint32_t)Nt * 16) >> 4) << 16);
    const uint64_t z_fix = hi | ((uint64_t)(((uint32_t)a.Cp * 16) >> 4) << 16);
    const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 24);
    const long long t0 = clock64();
#pragma unroll 1
    for (int y = 0; y < a.R; ++y) {
      const uint64_t a_row = a_fix | ((a_base + (uint32_t)y * 16) >> 4);
      // parts j = 0..2 -> output row y - 2 + j (>= 0)
      const int jlo = y >= 2 ? 0 : 2 - y;
      const uint32_t colz = (uint32_t)(y % NS) * a.Cp;
      if (a.zero_by_mma)
        mma(tmem + colz, a_row, z_fix | (z_base >> 4), idesc0 | ((uint32_t)(a.Cp >> 3) << 17), 0u);
      // segments of contiguous slots
      int j = jlo;
      while (j <= 2) {
        const int s0 = (y - 2 + j) % NS;
        int np = 1;
        while (j + np <= 2 && s0 + np < NS) ++np;
        const uint32_t idesc = idesc0 | ((uint32_t)((np * a.Cp) >> 3) << 17);
        const uint32_t col = (uint32_t)s0 * a.Cp;
#pragma unroll 1
        for (int ks = 0; ks < a.ksteps; ++ks) {
          const uint64_t ad = a_row + (uint64_t)(((uint32_t)(2 * ks) * rows * 16) >> 4);
          const uint64_t bd = (b_fix | ((b_base + (uint32_t)j * a.Cp * 16) >> 4)) + (uint64_t)(((uint32_t)(2 * ks) * Nt * 16) >> 4);
          if (!a.zero_by_mma && ks == 0 && j + np == 3) {
            // part 2 is the first touch of its slot: overwrite there, accumulate on the others
            if (np > 1) mma(tmem + col, ad, bd, idesc0 | ((uint32_t)(((np - 1) * a.Cp) >> 3) << 17), 1u);
            mma(tmem + col + (uint32_t)(np - 1) * a.Cp, ad, bd + (uint64_t)(((uint32_t)(np - 1) * a.Cp * 16) >> 4),
                idesc0 | ((uint32_t)(a.Cp >> 3) << 17), 0u);
          } else {
            mma(tmem + col, ad, bd, idesc, 1u);
          }
        }
        j += np;
      }
    }
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xFFFFFFFF;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar))
        : "memory");
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(
            smem_u32(&bar))
        : "memory");
    const long long t1 = clock64();
    if (tid == 0) a.cycles[0] = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int n0 = 0; n0 < 512; n0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int k = 0; k < 16; ++k) a.D[(size_t)tid * 512 + n0 + k] = __uint_as_float(v[k]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  float *D;
  long long *cyc;
  cudaMalloc(&D, 128 * 512 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int bad = 0;
  for (int zb = 1; zb >= 0; --zb)
    for (int Cp : {32, 64})
      for (int ksteps : {2, 4})
        for (int R : {10, 40, 200}) {
          if ((128 + R) * 16 * ksteps * 2 > 64 * 1024) continue;
          Args a{Cp, R, ksteps, zb, D, cyc};
          probe<<<1, 128, 200 * 1024>>>(a);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("Cp=%d R=%d: CUDA error %s\n", Cp, R, cudaGetErrorString(e));
            return 1;
          }
          std::vector<float> h(128 * 512);
          long long c;
          cudaMemcpy(h.data(), D, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
          cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
          const int NS = 512 / Cp, K = 16 * ksteps;
          double maxerr = 0;
          int checked = 0;
          for (int yo = (R - NS > 0 ? R - NS : 0); yo <= R - 3; ++yo) {      // complete output rows still in the ring
            const int col = (yo % NS) * Cp;
            for (int m = 0; m < 128; ++m)
              for (int o = 0; o < Cp; ++o) {
                double ref = 0;
                for (int j = 0; j < 3; ++j) {
                  const int y = yo + 2 - j;
                  for (int k = 0; k < K; ++k) ref += (double)xval(y + m, k) * (double)bval(j * Cp + o, k);
                }
                const double d = fabs(ref - h[(size_t)m * 512 + col + o]);
                if (d > maxerr) maxerr = d;
                ++checked;
              }
          }
          const int mmas = R * (ksteps + zb);
          printf("zero_by_mma=%d Cp=%2d K=%2d R=%3d : max|err| = %-8g %s (%d values) | %.1f cycles per row, %.1f per MMA\n", zb, Cp,
                 K, R, maxerr, maxerr == 0 ? "OK " : "BAD", checked, (double)c / R, (double)c / mmas);
          bad += maxerr != 0;
        }
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
