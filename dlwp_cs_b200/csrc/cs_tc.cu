// tcgen05 (sm_100a) bf16 implicit-GEMM kernel of the cubed-sphere convolution.
//
// Reference semantics: DLWP/custom.py:921-1002 (CubeSphereConv2D.call) with the preceding CubeSpherePadding2D
// (custom.py:1198-1308) and the U-Net's pool / upsample / concatenate (Azure/train_cs.py:197-199, 282-299) folded
// into the load stage, bias + capped leaky ReLU in the epilogue.
//
// GEMM view of one face:  D[q, o] = sum_{tap,c} A[q + toff(tap), c] * W[tap, c, o]
//   q     linear position r*Wv + c over the *virtual* (zero/halo padded) face of width Wv; positions with c >= Wout
//         are computed and discarded (Wout/Wv = 48/50 useful at C48), which makes every tap a pure row offset of the
//         same shared-memory patch: the A operand descriptor of tap (u,v) is the patch base + (u*dh*Wv + v*dw) rows.
//   A     the patch, K-major without swizzle: [channel slab of 8][patch pixel][8 bf16] -- core matrices of 8 pixels x
//         16 bytes are contiguous (SBO = 128 B), slabs are LBO apart, and a row offset is a plain +16 B per pixel.
//   W     packed once per layer in exactly the shared-memory image, streamed through an mbarrier ring by TMA bulk copy.
//   D     fp32 in tensor memory: MB accumulators of 128 lanes x CoutP columns.
// Warp roles (192 threads): warp 0 = TMA weight producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = patch loaders (halo gather through the index table, cp.async 16 B per pixel-slab; 2x2 mean or scalar
// gathers go through registers) and then the epilogue (tcgen05.ld -> bias/activation -> global stores).
#include <stdlib.h>
#include <string.h>
#include "cs_common.cuh"

namespace dlwpcs {

namespace {

constexpr int TC_THREADS = 192;
constexpr int TC_LOADERS = 128;
constexpr int MAX_CHUNKS = 8;
constexpr int MAX_STAGES = 8;
constexpr int MAX_MB = 8;

struct TcPlan {
  int CinP, CoutP, KC, nch, SPC;     // padded channels, channels per K chunk, chunks, 8-channel slabs per chunk
  int taps, NU, UPS, NST, nstages;   // weight units (chunk,tap), units per ring stage, ring depth, total stage loads
  int unitBytes, stageBytes;
  int Wv, Hv, Q, nmb, haloExt;       // virtual face, linear outputs per face, 128-row blocks per face, patch overhang
  int MB, tiles, NPIXp, slabBytes;   // m-blocks per CTA, CTAs per face, patch rows (padded), bytes per slab
  int tmemCols;
  int smemBytes;
  int64_t groupBytes;                // packed weights per face group
  int vec;                           // 16-byte gather path usable
};

struct TcP {
  const __nv_bfloat16 *x0, *x1;
  const int32_t *lut;
  const uint8_t *wpack;
  const float *bias;        // [3][CoutP] fp32 (zeros when the layer has no bias)
  void *y;
  const void *mask_y;       // dgrad: forward output whose activation derivative scales the gathered dy (nullptr: none)
  int y_f32, mask_f32;
  int n, Hin, Win, Hout, Wout;
  int cin, cout, c0, c1, mode0, mode1;
  int kw, dh, dw;
  int pt[3], pl;
  int act;
  float slope, maxv;
  int mask_act;
  float mask_slope, mask_max;
  TcPlan pl_;
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start address, leading byte offset
// (between the two 8-channel core matrices of one K=16 step), stride byte offset (between 8-row groups), version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ float act_apply(float v, int act, float slope, float maxv) {
  if (act == DLWPCS_ACT_CAPPED_LEAKY_RELU) v = v < 0.f ? slope * v : fminf(v, maxv);
  return v;
}
__device__ __forceinline__ float act_grad_from_y(float y, int act, float slope, float maxv) {
  if (act == DLWPCS_ACT_CAPPED_LEAKY_RELU) return y < 0.f ? slope : (y < maxv ? 1.f : 0.f);
  return 1.f;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4 &v, float *f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
  }
}

// physical pixel index (into a (B,6,e,e,C) tensor) of logical source pixel s = f*n*n + i*n + j for one input source
__device__ __forceinline__ int phys_pixel(int s, int n, int mode, int b) {
  if (mode == DLWPCS_SRC_SAME) return b * 6 * n * n + s;
  const int nn = n * n;
  const int f = s / nn, rem = s - f * nn, i = rem / n, j = rem - i * n;
  if (mode == DLWPCS_SRC_UP2) {
    const int h = n >> 1;
    return ((b * 6 + f) * h + (i >> 1)) * h + (j >> 1);
  }
  const int h = n * 2;  // POOL2: top-left pixel of the 2x2 block
  return ((b * 6 + f) * h + 2 * i) * h + 2 * j;
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ TcP P) {
  extern __shared__ uint8_t smem_raw[];
  const TcPlan &L = P.pl_;
  // carve: [barriers 256 B][patch][weight ring][pixel tables]
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_wfull = base, bar_wempty = base + 8 * MAX_STAGES, bar_pfull = base + 16 * MAX_STAGES,
                 bar_acc = base + 16 * MAX_STAGES + 8 * MAX_CHUNKS, tmem_slot = bar_acc + 8;
  const uint32_t patch = base + 256;
  const uint32_t wring = patch + (uint32_t)(L.CinP / 8) * L.slabBytes;
  int *s_pix0 = reinterpret_cast<int *>(gen + 256 + (size_t)(L.CinP / 8) * L.slabBytes + (size_t)L.NST * L.stageBytes);
  int *s_pix1 = s_pix0 + L.NPIXp;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, bf = blockIdx.y, f = bf % 6, b = bf / 6;
  const int grp = f < 4 ? 0 : f - 3;
  const int mb0 = tile * L.MB;
  const int MBc = min(L.MB, L.nmb - mb0);
  const int q0 = mb0 * 128;
  const int npix = MBc * 128 + L.haloExt;

  if (tid == 0) {
    for (int i = 0; i < L.NST; ++i) {
      mbar_init(bar_wfull + 8 * i, 1);
      mbar_init(bar_wempty + 8 * i, 1);
    }
    for (int i = 0; i < L.nch; ++i) mbar_init(bar_pfull + 8 * i, TC_LOADERS);
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)L.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(gen + (tmem_slot - base));

  if (warp == 0) {
    // ===== weight producer: TMA bulk copies of whole ring stages =====
    if (lane == 0) {
      const uint8_t *wg = P.wpack + (size_t)grp * L.groupBytes;
      for (int s = 0; s < L.nstages; ++s) {
        const int st = s % L.NST, ph = (s / L.NST) & 1;
        mbar_wait(bar_wempty + 8 * st, ph ^ 1);
        const int units = min(L.UPS, L.NU - s * L.UPS);
        const uint32_t bytes = (uint32_t)units * L.unitBytes;
        mbar_expect_tx(bar_wfull + 8 * st, bytes);
        tma_bulk_g2s(wring + st * L.stageBytes, wg + (size_t)s * L.stageBytes, bytes, bar_wfull + 8 * st);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(L.CoutP >> 3) << 17) | (8u << 24);
      int unit = 0, st = 0, ph = 0;
      for (int kc = 0; kc < L.nch; ++kc) {
        mbar_wait(bar_pfull + 8 * kc, 0);
        tc_fence_after();
        for (int tap = 0; tap < L.taps; ++tap) {
          const int uis = unit % L.UPS;
          if (uis == 0) {
            mbar_wait(bar_wfull + 8 * st, ph);
            tc_fence_after();
          }
          const int u = tap / P.kw, v = tap - u * P.kw;
          const uint32_t a0 = patch + (uint32_t)(kc * L.SPC) * L.slabBytes + (uint32_t)(u * P.dh * L.Wv + v * P.dw) * 16u;
          const uint32_t b0 = wring + st * L.stageBytes + uis * L.unitBytes;
          for (int mb = 0; mb < MBc; ++mb) {
            for (int j = 0; j < L.KC / 16; ++j) {
              const uint64_t ad = make_desc(a0 + (uint32_t)(2 * j) * L.slabBytes + (uint32_t)mb * 2048u, L.slabBytes, 128);
              const uint64_t bd = make_desc(b0 + (uint32_t)(2 * j) * L.CoutP * 16u, (uint32_t)L.CoutP * 16u, 128);
              umma_bf16(tmem_base + (uint32_t)(mb * L.CoutP), ad, bd, idesc, (unit > 0 || j > 0) ? 1u : 0u);
            }
          }
          ++unit;
          if (unit % L.UPS == 0 || unit == L.NU) {
            umma_commit(bar_wempty + 8 * st);
            if (++st == L.NST) { st = 0; ph ^= 1; }
          }
        }
      }
      umma_commit(bar_acc);
    }
  } else {
    // ===== patch loaders, then epilogue =====
    const int lt = tid - 64;
    // phase 1: logical source pixel of every patch row, mapped to the physical pixel of each input source
    for (int i = lt; i < npix; i += TC_LOADERS) {
      const int g = q0 + i;
      const int rv = g / L.Wv, cv = g - rv * L.Wv;
      const int r = rv - P.pt[grp], c = cv - P.pl;
      int p0 = -1, p1 = -1;
      if (r >= 0 && r < P.Hin && c >= 0 && c < P.Win) {
        const int s = P.lut ? __ldg(P.lut + (f * P.Hin + r) * P.Win + c) : (f * P.Hin + r) * P.Win + c;
        p0 = phys_pixel(s, P.n, P.mode0, b);
        if (P.c1 > 0) p1 = phys_pixel(s, P.n, P.mode1, b);
      }
      s_pix0[i] = p0;
      s_pix1[i] = p1;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(TC_LOADERS) : "memory");   // the tables are read by other loader threads
    // phase 2: gather, one cp.async group per K chunk
    const int items = npix * L.SPC;
    for (int kc = 0; kc < L.nch; ++kc) {
      for (int it = lt; it < items; it += TC_LOADERS) {
        const int i = it / L.SPC, sl = it - i * L.SPC;
        const int slab = kc * L.SPC + sl;
        const int c = slab * 8;
        const uint32_t dst = patch + (uint32_t)slab * L.slabBytes + (uint32_t)i * 16u;
        if (L.vec) {
          const bool first = c < P.c0;
          const __nv_bfloat16 *src = first ? P.x0 : P.x1;
          const int C = first ? P.c0 : P.c1, cc = first ? c : c - P.c0, mode = first ? P.mode0 : P.mode1;
          const int px = first ? s_pix0[i] : s_pix1[i];
          const bool valid = px >= 0 && c < P.cin;
          if (mode != DLWPCS_SRC_POOL2 && !P.mask_y) {
            const __nv_bfloat16 *g = valid ? src + (size_t)px * C + cc : P.x0;
            cp_async16(dst, g, valid ? 16u : 0u);
          } else {
            uint4 o = make_uint4(0, 0, 0, 0);
            if (valid) {
              float a[8];
              if (mode == DLWPCS_SRC_POOL2) {
                const int w2 = P.n * 2;
                float t[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] = 0.f;
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                  for (int dx = 0; dx < 2; ++dx) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)(px + dy * w2 + dx) * C + cc));
                    unpack_bf16x8(v, t);
#pragma unroll
                    for (int k = 0; k < 8; ++k) a[k] += t[k];
                  }
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] *= 0.25f;
              } else {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)px * C + cc));
                unpack_bf16x8(v, a);
              }
              if (P.mask_y) {  // dgrad: scale dy by the activation derivative taken from the forward output
                float m[8];
                if (P.mask_f32) {
                  const float *mp = reinterpret_cast<const float *>(P.mask_y) + (size_t)px * C + cc;
                  const float4 m0 = __ldg(reinterpret_cast<const float4 *>(mp));
                  const float4 m1 = __ldg(reinterpret_cast<const float4 *>(mp) + 1);
                  m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w; m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
                } else {
                  const uint4 mv = __ldg(reinterpret_cast<const uint4 *>(
                      reinterpret_cast<const __nv_bfloat16 *>(P.mask_y) + (size_t)px * C + cc));
                  unpack_bf16x8(mv, m);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] *= act_grad_from_y(m[k], P.mask_act, P.mask_slope, P.mask_max);
              }
              o = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]),
                             pack_bf16x2(a[6], a[7]));
            }
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                         : "memory");
          }
        } else {
          // scalar gather: channel counts that are not multiples of 8 (the 18-channel network input)
          float a[8];
          const int px0 = s_pix0[i], px1 = s_pix1[i];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int ch = c + k;
            float v = 0.f;
            if (ch < P.cin) {
              const bool first = ch < P.c0;
              const __nv_bfloat16 *src = first ? P.x0 : P.x1;
              const int C = first ? P.c0 : P.c1, cc = first ? ch : ch - P.c0, mode = first ? P.mode0 : P.mode1;
              const int px = first ? px0 : px1;
              if (px >= 0) {
                if (mode == DLWPCS_SRC_POOL2) {
                  const int w2 = P.n * 2;
                  v = 0.25f * (__bfloat162float(src[(size_t)px * C + cc]) + __bfloat162float(src[(size_t)(px + 1) * C + cc]) +
                               __bfloat162float(src[(size_t)(px + w2) * C + cc]) +
                               __bfloat162float(src[(size_t)(px + w2 + 1) * C + cc]));
                } else {
                  v = __bfloat162float(src[(size_t)px * C + cc]);
                }
                if (P.mask_y) {
                  const float m = P.mask_f32 ? reinterpret_cast<const float *>(P.mask_y)[(size_t)px * C + cc]
                                             : __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(P.mask_y)[(size_t)px * C + cc]);
                  v *= act_grad_from_y(m, P.mask_act, P.mask_slope, P.mask_max);
                }
              }
            }
            a[k] = v;
          }
          const uint4 o = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]),
                                     pack_bf16x2(a[6], a[7]));
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                       : "memory");
        }
      }
      cp_async_commit();
    }
    for (int kc = 0; kc < L.nch; ++kc) {
      cp_async_wait_dyn(L.nch - 1 - kc);
      fence_proxy_async();
      mbar_arrive(bar_pfull + 8 * kc);
    }

    // ===== epilogue: TMEM -> registers -> bias / activation -> global =====
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int quarter = warp & 3;                  // TMEM lanes this warp may read: 32*quarter .. +31
    const float *bias = P.bias + grp * L.CoutP;
    const size_t face_px = (size_t)(b * 6 + f) * P.Hout * P.Wout;
    const bool vec_out = P.y_f32 ? (P.cout % 4 == 0) : (P.cout % 8 == 0);
    for (int mb = 0; mb < MBc; ++mb) {
      const int q = q0 + mb * 128 + quarter * 32 + lane;
      const int r = q / L.Wv, c = q - r * L.Wv;
      const bool ok = q < L.Q && c < P.Wout;
      const size_t opix = face_px + (size_t)r * P.Wout + c;
      for (int n0 = 0; n0 < L.CoutP; n0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mb * L.CoutP + n0), v);
        tmem_ld_wait();
        if (!ok) continue;
        float o[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
          o[k] = act_apply(__uint_as_float(v[k]) + __ldg(bias + n0 + k), P.act, P.slope, P.maxv);
        if (P.y_f32) {
          float *yp = reinterpret_cast<float *>(P.y) + opix * P.cout + n0;
          if (vec_out && n0 + 16 <= P.cout) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              reinterpret_cast<float4 *>(yp)[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (n0 + k < P.cout) yp[k] = o[k];
          }
        } else {
          __nv_bfloat16 *yp = reinterpret_cast<__nv_bfloat16 *>(P.y) + opix * P.cout + n0;
          if (vec_out && n0 + 16 <= P.cout) {
            reinterpret_cast<uint4 *>(yp)[0] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                          pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
            reinterpret_cast<uint4 *>(yp)[1] = make_uint4(pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]),
                                                          pack_bf16x2(o[12], o[13]), pack_bf16x2(o[14], o[15]));
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (n0 + k < P.cout) yp[k] = __float2bfloat16_rn(o[k]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)L.tmemCols);
}

// ---- weight packing -----------------------------------------------------------------------------------------------
// packed[g][kc][tap][k8][n][8] bf16  (+ fp32 bias[3][CoutP] after the three groups).  transposed (dgrad): the GEMM's
// K runs over the forward cout, N over the forward cin, taps rotated 180 degrees.
__global__ void pack_tc_kernel(const float *__restrict__ w_eq, const float *__restrict__ w_pol,
                               const float *__restrict__ w_np, const float *__restrict__ b_eq,
                               const float *__restrict__ b_pol, const float *__restrict__ b_np, uint8_t *__restrict__ out,
                               int kh, int kw, int cin, int cout, int flip, int transposed, int CinP, int CoutP, int KC,
                               long long groupElems) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int taps = kh * kw;
  if (i < 3 * groupElems) {
    const int g = (int)(i / groupElems);
    long long r = i % groupElems;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % CoutP); r /= CoutP;
    const int k8 = (int)(r % (KC / 8)); r /= (KC / 8);
    const int tap = (int)(r % taps);
    const int kc = (int)(r / taps);
    const int k = kc * KC + k8 * 8 + e;            // GEMM K index (input channel of this GEMM)
    int u = tap / kw, v = tap % kw;
    const int gemm_cin = transposed ? cout : cin, gemm_cout = transposed ? cin : cout;
    float val = 0.f;
    if (k < gemm_cin && n < gemm_cout) {
      const float *src = g == 0 ? w_eq : (g == 1 ? w_pol : (w_np ? w_np : w_pol));
      if (transposed) { u = kh - 1 - u; v = kw - 1 - v; }
      const int us = (g == 2 && flip) ? kh - 1 - u : u;
      const int ci = transposed ? n : k, co = transposed ? k : n;
      val = src[(((long long)us * kw + v) * cin + ci) * cout + co];
    }
    reinterpret_cast<__nv_bfloat16 *>(out)[i] = __float2bfloat16_rn(val);
  } else if (i < 3 * groupElems + 3LL * CoutP) {
    const int j = (int)(i - 3 * groupElems), g = j / CoutP, co = j % CoutP;
    const float *src = g == 0 ? b_eq : (g == 1 ? b_pol : (b_np ? b_np : b_pol));
    float *bo = reinterpret_cast<float *>(out + 3 * groupElems * 2);
    bo[j] = (src && !transposed && co < cout) ? src[co] : 0.f;
  }
}

int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

// Tiling / staging plan for one layer; returns a reason string when the configuration is outside the kernel's reach.
const char *make_plan(const dlwpcs_conv_desc *d, const Geometry &g, int gemm_cin, int gemm_cout, TcPlan *L) {
  if (d->stride_h != 1 || d->stride_w != 1) return "strides must be 1";
  L->CinP = (gemm_cin + 15) / 16 * 16;
  L->CoutP = (gemm_cout + 15) / 16 * 16;
  if (L->CoutP > 256) return "more than 256 output channels";
  L->KC = 16;
  for (int kc : {64, 48, 32, 16})
    if (L->CinP % kc == 0) { L->KC = kc; break; }
  L->nch = L->CinP / L->KC;
  if (L->nch > MAX_CHUNKS) return "too many input channels";
  L->SPC = L->KC / 8;
  L->taps = d->kh * d->kw;
  L->NU = L->nch * L->taps;
  L->unitBytes = L->KC * L->CoutP * 2;
  L->UPS = 16384 / L->unitBytes;
  if (L->UPS < 1) L->UPS = 1;
  if (L->UPS > L->NU) L->UPS = L->NU;
  L->stageBytes = L->UPS * L->unitBytes;
  L->nstages = (L->NU + L->UPS - 1) / L->UPS;
  L->NST = L->nstages < 4 ? L->nstages : 4;
  L->groupBytes = (int64_t)L->NU * L->unitBytes;
  L->Wv = g.Wout + (d->kw - 1) * d->dil_w;
  L->Hv = g.Hout + (d->kh - 1) * d->dil_h;
  L->Q = (g.Hout - 1) * L->Wv + g.Wout;
  L->nmb = (L->Q + 127) / 128;
  L->haloExt = (d->kh - 1) * d->dil_h * L->Wv + (d->kw - 1) * d->dil_w;
  const int smem_cap = 225 * 1024, smem_pair = 110 * 1024;
  auto smem_for = [&](int MB, int *npixp, int *slab) {
    const int np = ((MB * 128 + L->haloExt + 7) / 8) * 8 + 1;     // odd multiple of 16 B: slab-strided stores spread
    *npixp = np;
    *slab = np * 16;
    return 128 + 256 + (L->CinP / 8) * np * 16 + L->NST * L->stageBytes + 2 * np * 4 + 64;
  };
  int mbmax = 512 / L->CoutP;
  if (mbmax > MAX_MB) mbmax = MAX_MB;
  if (mbmax > L->nmb) mbmax = L->nmb;
  int best = 0, np = 0, sb = 0;
  const int forced = env_int("DLWPCS_TC_MB", 0);
  if (forced > 0) {
    best = forced < mbmax ? forced : mbmax;
    if (smem_for(best, &np, &sb) > smem_cap) return "forced DLWPCS_TC_MB does not fit shared memory";
  } else {
    for (int MB = mbmax; MB >= 1; --MB)          // largest tile that still lets two CTAs share an SM ...
      if (smem_for(MB, &np, &sb) <= smem_pair && MB * L->CoutP <= 256) { best = MB; break; }
    if (!best)
      for (int MB = mbmax; MB >= 1; --MB)        // ... else the largest that fits at all
        if (smem_for(MB, &np, &sb) <= smem_cap) { best = MB; break; }
    if (!best) return "input patch does not fit shared memory (too many input channels for this face width)";
    const int tiles = (L->nmb + best - 1) / best;
    best = (L->nmb + tiles - 1) / tiles;           // even out the tiles of a face
  }
  L->MB = best;
  L->tiles = (L->nmb + best - 1) / best;
  L->smemBytes = smem_for(best, &L->NPIXp, &L->slabBytes);
  int cols = 32;
  while (cols < best * L->CoutP) cols *= 2;
  L->tmemCols = cols;
  if (L->slabBytes >= (1 << 18)) return "patch too large";
  return nullptr;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool tc_supported(const dlwpcs_conv_desc *d, const Geometry &g, const char **why) {
  TcPlan L;
  const char *r = make_plan(d, g, d->cin, d->cout, &L);
  if (r) {
    *why = r;
    return false;
  }
  if (d->x_dtype != DLWPCS_BF16) {
    *why = "activations must be bfloat16";
    return false;
  }
  return true;
}

int64_t tc_packed_weight_bytes(const dlwpcs_conv_desc *d, const Geometry &g, int transposed) {
  TcPlan L;
  const char *r = make_plan(d, g, transposed ? d->cout : d->cin, transposed ? d->cin : d->cout, &L);
  if (r) {
    set_error("bf16 tensor-core path does not support this configuration: %s", r);
    return -1;
  }
  return 3 * L.groupBytes + 3LL * L.CoutP * 4;
}

int tc_pack_weights(const dlwpcs_conv_desc *d, const Geometry &g, const dlwpcs_conv_weights *w, int transposed,
                    void *packed, cudaStream_t st) {
  TcPlan L;
  const char *r = make_plan(d, g, transposed ? d->cout : d->cin, transposed ? d->cin : d->cout, &L);
  CS_CHECK(r == nullptr, "bf16 tensor-core path does not support this configuration: %s", r);
  const long long groupElems = L.groupBytes / 2;
  const long long total = 3 * groupElems + 3LL * L.CoutP;
  pack_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      w->w_eq, w->w_pol, d->independent_north_pole ? w->w_np : nullptr, d->use_bias ? w->b_eq : nullptr,
      d->use_bias ? w->b_pol : nullptr, (d->use_bias && d->independent_north_pole) ? w->b_np : nullptr,
      (uint8_t *)packed, d->kh, d->kw, d->cin, d->cout, d->flip_north_pole, transposed, L.CinP, L.CoutP, L.KC,
      groupElems);
  CS_CUDA(cudaGetLastError());
  return 0;
}

static int launch_tc(TcP &P, int batch, cudaStream_t st) {
  const TcPlan &L = P.pl_;
  static int cur_max = 48 * 1024;
  if (L.smemBytes > cur_max) {
    CS_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cur_max = 227 * 1024;
  }
  CS_CHECK((long long)batch * 6 <= 65535, "batch too large for one launch (B*6 = %lld > 65535)", (long long)batch * 6);
  dim3 grid(L.tiles, batch * 6);
  conv_tc_kernel<<<grid, TC_THREADS, L.smemBytes, st>>>(P);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int tc_conv_fwd(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed,
                void *y, cudaStream_t st) {
  TcP P;
  memset(&P, 0, sizeof(P));
  const char *r = make_plan(d, g, d->cin, d->cout, &P.pl_);
  CS_CHECK(r == nullptr, "bf16 tensor-core path does not support this configuration: %s", r);
  TcPlan &L = P.pl_;
  P.x0 = (const __nv_bfloat16 *)x0;
  P.x1 = (const __nv_bfloat16 *)x1;
  P.lut = nullptr;
  if (d->halo > 0) {
    const HaloTables *t = get_halo_tables(d->n, d->halo);
    if (!t) return 3;
    P.lut = t->lut;
  }
  P.wpack = (const uint8_t *)packed;
  P.bias = reinterpret_cast<const float *>(P.wpack + 3 * L.groupBytes);
  P.y = y;
  P.y_f32 = d->y_dtype == DLWPCS_F32;
  P.mask_y = nullptr;
  P.n = d->n; P.Hin = g.Hin; P.Win = g.Win; P.Hout = g.Hout; P.Wout = g.Wout;
  P.cin = d->cin; P.cout = d->cout; P.c0 = d->c0; P.c1 = d->c1; P.mode0 = d->mode0; P.mode1 = d->mode1;
  P.kw = d->kw; P.dh = d->dil_h; P.dw = d->dil_w;
  P.pt[0] = g.pt[0]; P.pt[1] = g.pt[1]; P.pt[2] = g.pt[2]; P.pl = g.pl;
  P.act = d->act; P.slope = d->act_slope; P.maxv = d->act_max;
  L.vec = (d->c0 % 8 == 0) && (d->c1 % 8 == 0) && aligned16(x0) && (d->c1 == 0 || aligned16(x1));
  return launch_tc(P, d->batch, st);
}

}  // namespace dlwpcs
