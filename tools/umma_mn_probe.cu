// Micro-probe of tcgen05.mma with MN-major (pixel-major) operands on sm_100a -- the operand form of the wgrad GEMM
//   dW[(shift j, ci), co] = sum_pixels X[p + r + j][ci] * dY[p][co]
// where both tensors sit in shared memory as one row of RB bytes per pixel, 16-byte chunks XOR-swizzled with the row's
// address bits (the forward kernel's patch layout).  Checks (1) the MN-major canonical layouts of CUTLASS
// (cute/atom/mma_traits_sm100.hpp: ((T,s,m),(8,k)):((1,T,LBO),(sT,SBO))), (2) that the start address may sit at any row,
// (3) that LBO = RB stacks pixel-shifted copies of the patch along M ("tap stacking"), and reports cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_mn_probe tools/umma_mn_probe.cu && ./umma_mn_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args {
  int RBx, RBy;    // bytes per pixel row of X (A operand) and dY (B operand): 32 / 64 / 128
  int N;           // MMA N (<= RBy/2: one atom; or a multiple: atoms at LBO = yblock)
  int r;           // row offset of the A start address
  int ksteps;      // K = 16 * ksteps pixels
  int iters;       // timing repetitions
  int rows;        // pixel rows resident
  int M64;         // 1: M = 64 instead of 128
  float *D;        // [128][N]
  long long *cycles;
};

__host__ __device__ inline int xval(int p, int c) { return ((p * 7 + c * 3) % 13) - 6; }
__host__ __device__ inline int yval(int p, int c) { return ((p * 5 + c * 11) % 7) - 3; }

__device__ __forceinline__ uint32_t row_chunk_addr(uint32_t base, int RB, int p, int chunk) {
  const uint32_t row = base + (uint32_t)p * RB;
  const uint32_t mask = RB / 16 - 1;   // 1 / 3 / 7
  return row + ((((uint32_t)chunk) ^ ((row >> 7) & mask)) << 4);
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
  const uint32_t x_base = base, y_base = base + 96 * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int cx = a.RBx / 16, cy = a.RBy / 16;
  for (int i = tid; i < a.rows * cx; i += 128) {
    const int p = i / cx, ch = i % cx;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((float)xval(p, ch * 8 + e));
    const uint4 o = *reinterpret_cast<uint4 *>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row_chunk_addr(x_base, a.RBx, p, ch)), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w));
  }
  // dY: N channels as N / (RBy/2) blocks of [rows][RBy]
  const int nyb = (a.N * 2 + a.RBy - 1) / a.RBy;
  for (int i = tid; i < a.rows * cy * nyb; i += 128) {
    const int blk = i / (a.rows * cy), rem = i % (a.rows * cy);
    const int p = rem / cy, ch = rem % cy;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((float)yval(p, blk * (a.RBy / 2) + ch * 8 + e));
    const uint4 o = *reinterpret_cast<uint4 *>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row_chunk_addr(y_base + blk * a.rows * a.RBy, a.RBy, p, ch)), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w));
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
    // bf16 x bf16 -> f32, A and B MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(a.N >> 3) << 17) |
                           ((a.M64 ? 4u : 8u) << 24);
    auto lt = [](int RB) { return RB == 128 ? 2u : (RB == 64 ? 4u : 6u); };
    // A: atoms of RBx/2 channels x 8 pixels; next atom along M = the patch shifted by one pixel (LBO = RBx); next 8
    // pixels along K at SBO = 8 * RBx
    const uint32_t a_hi = ((uint32_t)(8 * a.RBx) >> 4) | (1u << 14) | (lt(a.RBx) << 29);
    const uint32_t a_lo_fix = ((uint32_t)a.RBx >> 4) << 16;
    // B: next atom along N = next channel block (LBO = rows * RBy)
    const uint32_t b_hi = ((uint32_t)(8 * a.RBy) >> 4) | (1u << 14) | (lt(a.RBy) << 29);
    const uint32_t b_lo_fix = ((((uint32_t)a.rows * a.RBy) >> 4) & 0x3FFFu) << 16;
    long long t0 = 0, t1 = 0;
#define MMA(KS, ACC) do { \
      const uint32_t alo = a_lo_fix | (((x_base + (uint32_t)(a.r + 16 * (KS)) * a.RBx) & 0x3FFFFu) >> 4); \
      const uint32_t blo = b_lo_fix | (((y_base + (uint32_t)(16 * (KS)) * a.RBy) & 0x3FFFFu) >> 4); \
      asm volatile( \
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t" \
            "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem), \
            "r"(alo), "r"(a_hi), "r"(blo), "r"(b_hi), "r"(idesc), "r"(ACC) : "memory"); } while (0)
    if (elect_one()) {
      for (int ks = 0; ks < a.ksteps; ++ks) MMA(ks, ks > 0 ? 1u : 0u);
    }
    __syncwarp();
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < a.iters; ++it) {
      if (elect_one()) {
        for (int ks = 0; ks < a.ksteps; ++ks) MMA(ks, 1u);
      }
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
  struct Cfg { int RBx, RBy, N; };
  const Cfg cfgs[] = {{64, 64, 32}, {128, 128, 64}, {128, 64, 32}, {64, 128, 64}, {32, 32, 16}, {128, 128, 128}, {64, 32, 16}};
  int bad = 0;
  for (const Cfg &c : cfgs)
    for (int M64 = 0; M64 < 2; ++M64)
      for (int r : {0, 3, 50}) {
        const int rows = 360, ksteps = 4;
        for (int iters : {1, 64}) {
          Args a{c.RBx, c.RBy, c.N, r, ksteps, iters, rows, M64, D, cyc};
          cudaMemset(D, 0, 128 * 256 * sizeof(float));
          probe<<<1, 128, 200 * 1024>>>(a);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("RBx=%d RBy=%d N=%d r=%d: CUDA error %s\n", c.RBx, c.RBy, c.N, r, cudaGetErrorString(e));
            return 1;
          }
          std::vector<float> h(128 * c.N);
          long long cy;
          cudaMemcpy(h.data(), D, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
          cudaMemcpy(&cy, cyc, sizeof(cy), cudaMemcpyDeviceToHost);
          const int CA = c.RBx / 2, M = M64 ? 64 : 128;
          double maxerr = 0;
          // M = 64: report which TMEM lanes hold the result (tries lane = m and lane = 32*(m/16) + m%16)
          double maxerr_alt = 0;
          for (int m = 0; m < M; ++m)
            for (int n = 0; n < c.N; ++n) {
              const int j = m / CA, ci = m % CA;
              double ref = 0;
              for (int k = 0; k < 16 * ksteps; ++k) ref += (double)xval(r + k + j, ci) * (double)yval(k, n);
              ref *= (iters + 1);
              const double d = fabs(ref - h[m * c.N + n]);
              if (d > maxerr) maxerr = d;
              if (M64) {
                const int alt = 32 * (m / 16) + (m % 16);
                const double d2 = fabs(ref - h[alt * c.N + n]);
                if (d2 > maxerr_alt) maxerr_alt = d2;
              }
            }
          if (iters == 1) {
            const bool ok = maxerr == 0 || (M64 && maxerr_alt == 0);
            if (!ok) ++bad;
            printf("RBx=%3d RBy=%3d N=%3d M=%3d r=%2d : max|err| = %-8g (alt lane map %-8g) %s", c.RBx, c.RBy, c.N, M, r, maxerr,
                   maxerr_alt, ok ? "OK " : "BAD");
          } else {
            printf("  | %d MMAs: %.1f cycles/MMA\n", ksteps * iters, (double)cy / (ksteps * iters));
          }
        }
      }
  printf(bad ? "%d configurations BAD\n" : "all configurations OK\n", bad);
  return 0;
}
