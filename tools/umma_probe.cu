// Micro-probe of tcgen05.mma shared-memory descriptor semantics and issue cost on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe tools/umma_probe.cu && ./umma_probe
// For each layout (no swizzle / 128B / 64B / 32B swizzle), row offset r and base_offset policy it checks
// D[m][n] = sum_k A[r+m][k] * B[n][k] (K = 64) against the host and reports cycles per MMA.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args {
  int mode;        // 0 none, 1 = 128B, 2 = 64B, 3 = 32B swizzle
  int N;           // MMA N
  int r;           // row offset of the A start address
  int use_base_offset;
  int iters;       // timing repetitions of the 4-MMA K loop
  int rows;        // rows of A resident in smem
  float *D;        // [128][N]
  long long *cycles;
};

__device__ __forceinline__ uint32_t a_addr(int mode, uint32_t base, int rows, int p, int k8) {
  // byte address of the 16-byte chunk (pixel p, channels 8*k8..8*k8+7); K = 64 -> k8 in 0..7
  if (mode == 0) return base + (uint32_t)k8 * rows * 16 + (uint32_t)p * 16;
  const int rowB = mode == 1 ? 128 : (mode == 2 ? 64 : 32);          // bytes per swizzle row
  const int cpr = rowB / 16;                                         // chunks per row
  const int kb = k8 / cpr, c = k8 % cpr;                             // K block, chunk within the row
  const uint32_t blk = base + (uint32_t)kb * rows * rowB;
  const uint32_t lin = blk + (uint32_t)p * rowB + (uint32_t)c * 16;
  // Swizzle<B,4,3>: XOR address bits [4,4+B) with bits [7,7+B)
  const int B = mode == 1 ? 3 : (mode == 2 ? 2 : 1);
  return lin ^ (((lin >> 7) & ((1u << B) - 1)) << 4);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(128) probe(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 96 * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  // fill A (rows x 64) and B (N x 64) in the chosen layout
  for (int i = tid; i < a.rows * 8; i += 128) {
    const int p = i / 8, k8 = i % 8;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((float)(((p * 7 + (k8 * 8 + e) * 3) % 13) - 6));
    const uint32_t dst = a_addr(a.mode, a_base, a.rows, p, k8);
    const uint4 o = *reinterpret_cast<uint4 *>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w));
  }
  for (int i = tid; i < a.N * 8; i += 128) {
    const int n = i / 8, k8 = i % 8;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((float)(((n * 5 + (k8 * 8 + e)) % 7) - 3));
    const uint32_t dst = a_addr(a.mode, b_base, a.N, n, k8);
    const uint4 o = *reinterpret_cast<uint4 *>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w));
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.N >> 3) << 17) | (8u << 24);
    const int rowB = a.mode == 0 ? 16 : (a.mode == 1 ? 128 : (a.mode == 2 ? 64 : 32));
    uint32_t a_lo, b_lo, a_hi, b_hi;
    if (a.mode == 0) {
      a_lo = (((uint32_t)a.rows * 16) >> 4) << 16;
      b_lo = (((uint32_t)a.N * 16) >> 4) << 16;
      a_hi = b_hi = (128u >> 4) | (1u << 14);
    } else {
      const uint32_t lt = a.mode == 1 ? 2u : (a.mode == 2 ? 4u : 6u);
      a_lo = b_lo = 1u << 16;
      a_hi = b_hi = ((uint32_t)(8 * rowB) >> 4) | (1u << 14) | (lt << 29);
    }
    const uint32_t a_start = a_base + (uint32_t)a.r * rowB;
    long long t0 = 0, t1 = 0;
    uint32_t alo[4], blo[4];
    for (int j = 0; j < 4; ++j) {          // K = 64 as four K = 16 steps
      uint32_t ao, bo;
      if (a.mode == 0) {
        ao = (uint32_t)(2 * j) * a.rows * 16;
        bo = (uint32_t)(2 * j) * a.N * 16;
      } else {
        const int kpr = rowB / 32;           // K=16 steps per swizzle row
        ao = (uint32_t)(j / kpr) * a.rows * rowB + (uint32_t)(j % kpr) * 32;
        bo = (uint32_t)(j / kpr) * a.N * rowB + (uint32_t)(j % kpr) * 32;
      }
      alo[j] = a_lo | (((a_start + ao) & 0x3FFFFu) >> 4);
      blo[j] = b_lo | (((b_base + bo) & 0x3FFFFu) >> 4);
    }
#define MMA(J, ACC) asm volatile( \
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t" \
            "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem + (uint32_t)(a.use_base_offset ? (J) * a.N : 0)), \
            "r"(alo[J]), "r"(a_hi), "r"(blo[J]), "r"(b_hi), "r"(idesc), "r"(ACC) : "memory")
    if (elect_one()) { MMA(0, 0u); MMA(1, 1u); MMA(2, 1u); MMA(3, 1u); }
    __syncwarp();
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < a.iters; ++it) {
      if (elect_one()) { MMA(0, 1u); MMA(1, 1u); MMA(2, 1u); MMA(3, 1u); }
      __syncwarp();
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(
            smem_u32(&bar))
        : "memory");
    t1 = clock64();
    if (tid == 0) a.cycles[0] = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // D accumulated (iters+1) times the same product: read back and divide on the host
  for (int n0 = 0; n0 < a.N; n0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int k = 0; k < 16; ++k) a.D[(size_t)tid * a.N + n0 + k] = __uint_as_float(v[k]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main() {
  float *D;
  long long *cyc;
  cudaMalloc(&D, 128 * 256 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char *names[] = {"none", "sw128", "sw64", "sw32"};
  for (int mode = 0; mode < 1; ++mode)
    for (int N : {32, 64})
      for (int r : {0, 3})
        for (int rows : {320, 617, 361, 313}) for (int ubo = 0; ubo < 1; ++ubo) {
          if (ubo && N > 64) continue;
          for (int iters : {1, 64}) {
            Args a{mode, N, r, ubo, iters, rows, D, cyc};
            probe<<<1, 128, 200 * 1024>>>(a);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("mode=%s N=%d r=%d ubo=%d: CUDA error %s\n", names[mode], N, r, ubo, cudaGetErrorString(e));
              return 1;
            }
            std::vector<float> h(128 * N);
            long long c;
            cudaMemcpy(h.data(), D, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
            cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
              for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (double)((((r + m) * 7 + k * 3) % 13) - 6) * (double)(((n * 5 + k) % 7) - 3);
                ref *= (iters + 1);
                const double d = fabs(ref - h[m * N + n]);
                if (d > maxerr) maxerr = d;
              }
            if (iters == 1)
              printf("mode=%-5s N=%3d r=%2d rows=%d : max|err| = %-8g %s", names[mode], N, r, rows, maxerr, maxerr == 0 ? "OK " : "BAD");
            else
              printf("  | %d MMAs: %.1f cycles/MMA (err %g)\n", 4 * iters, (double)c / (4.0 * iters), maxerr);
          }
        }
  return 0;
}
