// tcgen05 (sm_100a) bf16 implicit-GEMM kernel of the cubed-sphere convolution for NARROW layers (<= 32 output channels,
// 3 kernel columns): the three horizontal taps are STACKED ALONG N.
//
// Why (profiles/r2_chain.md): the classic kernel (cs_tc.cu) issues one M128 x N32 MMA per tap, and an SS-mode MMA re-reads
// its whole A tile from shared memory each time -- 4 KB of A for 1 KB of B.  Nine taps x K steps saturate the SM's
// shared-memory port (128 B / cycle); everything else on the SM queues behind it.  Here one MMA computes, for a kernel row
// u and 16 input channels, all three kernel columns at once:
//
//     E[q, v*32 + o] = sum_{u,c} A[q + u*dh*Wv, c] * W[u, v, c, o]            (N = 96: 56.8 cycles instead of 3 x 41.8)
//     out[q, o]      = E[q, o] + E[q + 1, 32 + o] + E[q + 2, 64 + o]          (shift-add across rows in the epilogue)
//
// q = linear position over the virtual (halo-padded) face of width Wv as in cs_tc.cu; a third of the A reads.
// The shift crosses tensor-memory LANES (rows), so the epilogue adds in registers: shuffles inside a warp's 32-row
// quarter, a 640-byte shared-memory exchange across the three quarter boundaries of a 128-row block, and a block stride
// of 126 rows (the last two rows of every block only feed their predecessors) so that nothing crosses a block.
// Accumulators: a ring of five 96-column slots in tensor memory, filled and drained block by block.
//
// Reference semantics: DLWP/custom.py:921-1002 (CubeSphereConv2D.call) with the preceding CubeSpherePadding2D
// (custom.py:1198-1308) and the U-Net's up-sampling / concatenation (Azure/train_cs.py:198, 293, 299) folded into the load
// stage, bias + capped leaky ReLU (train_cs.py:199) in the epilogue -- identical results to cs_tc.cu (same products, the
// float32 sums associate differently).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cs_common.cuh"
#include "cs_ptx.cuh"

namespace dlwpcs {

namespace {

constexpr int NS_THREADS = 608, NS_LOADERS = 256, NS_EPI = 256;
constexpr int NS_EPI_WARP0 = 8, NS_TMA_WARP = 16, NS_MMA_WARP = 17, NS_MMA_WARPS = 2;
constexpr int NS_STRIDE = 126;     // output rows per 128-row block
constexpr int NS_SLOTS = 5;        // accumulator ring: 5 x 96 of the 512 tensor-memory columns
constexpr int NS_N = 96;           // 3 kernel columns x 32 output channels
constexpr int NS_SMEM_CAP = 227 * 1024;
constexpr int NS_MAX_PS = 3;

struct NsPlan {
  int CinP, KC, nch, S, logS;                  // padded input channels, channels per K block, K blocks, 16-byte chunks per pixel
  int RB, cprLog, swzMask, layoutType, blockBytes;
  int kh, Wv, Q, nmb, MB, tpf, haloV, NPIXp, G;
  int NU, unitBytes, groupBytes;               // weight units (K block, kernel row) of KC x 96 bf16
  int patchBytes, PS, tabBytes, twoTabs;
  int off_w, off_misc, off_bias, off_tab, off_stg, off_xch, smemBytes;
};

struct NsP {
  const __nv_bfloat16 *x0, *x1;
  const int32_t *tab0, *tab1;
  const uint8_t *wpack;
  const float *bias;
  __nv_bfloat16 *y;
  int batch, Hout, Wout, cin, cout, c0, c1, ppb0, ppb1, dh, act;
  float slope, maxv;
  NsPlan L;
};

struct NsTile {
  int b, f, grp, tf;
};
__device__ __forceinline__ NsTile ns_decode(int id, int batch, int tpf) {
  NsTile t;
  const int eq = 4 * batch * tpf, pol = batch * tpf;
  if (id < eq) {
    const int bf = id / tpf;
    t.tf = id - bf * tpf;
    t.b = bf >> 2;
    t.f = bf & 3;
    t.grp = 0;
  } else {
    int r = id - eq;
    t.grp = 1;
    if (r >= pol) { r -= pol; t.grp = 2; }
    t.b = r / tpf;
    t.tf = r - t.b * tpf;
    t.f = 3 + t.grp;
  }
  return t;
}

__device__ __forceinline__ void st_shared_f32x4(uint32_t dst, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f32x4(uint32_t src) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src));
  return v;
}

// bias + activation + bf16 pack of eight consecutive output channels -> one 16-byte chunk
__device__ __forceinline__ uint4 ns_finish8(const float *s, const float *bias, int act, float slope, float maxv) {
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    o[e] = s[e] + bias[e];
    o[e] = act_apply(o[e], act, slope, maxv);
  }
  return make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
// KC16T = K = 16 steps per K block (KC / 16)
template <int KC16T>
__global__ void __launch_bounds__(NS_THREADS, 1) conv_tc_ns_kernel(const __grid_constant__ NsP P) {
  extern __shared__ uint8_t smem_raw[];
  const NsPlan &L = P.L;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t patch0 = base, wbase = base + L.off_w, misc = base + L.off_misc;
  // barriers (8 bytes each): wfull, wempty, pfull[3], pempty[3], afull[5], aempty[5], tfull[2], tempty[2]; tmem slot
  const uint32_t bar_wfull = misc, bar_wempty = misc + 8, bar_pfull = misc + 16, bar_pempty = misc + 48,
                 bar_afull = misc + 80, bar_aempty = misc + 128, bar_tfull = misc + 176, bar_tempty = misc + 192,
                 tmem_slot = misc + 208;
  float *s_bias = reinterpret_cast<float *>(gen + L.off_bias);
  const uint32_t tabring = base + L.off_tab;
  const int32_t *s_tab = reinterpret_cast<const int32_t *>(gen + L.off_tab);
  const uint32_t stg0 = base + L.off_stg, xch0 = base + L.off_xch;

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int ntiles = 6 * P.batch * L.tpf;

  if (tid == 0) {
    mbar_init(bar_wfull, 1);
    mbar_init(bar_wempty, NS_MMA_WARPS);
    for (int i = 0; i < L.PS; ++i) {
      mbar_init(bar_pfull + 8 * i, NS_LOADERS);
      mbar_init(bar_pempty + 8 * i, NS_MMA_WARPS);
    }
    for (int i = 0; i < NS_SLOTS; ++i) {
      mbar_init(bar_afull + 8 * i, 1);
      mbar_init(bar_aempty + 8 * i, 128);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, NS_LOADERS);
    }
    fence_mbar_init();
  }
  if (warp == NS_MMA_WARP) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(gen + L.off_misc + 208);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == NS_TMA_WARP) {
    if (lane == 0) {
      // ===== weights: one face group resident at a time; reloaded when the walk enters the next group =====
      int ph = 0, prev_grp = -1, loads = 0;
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int grp = ns_decode(tile, P.batch, L.tpf).grp;
        if (grp == prev_grp) continue;
        prev_grp = grp;
        if (loads > 0) {                         // the issuers have released the previous group's image
          mbar_wait(bar_wempty, ph);
          ph ^= 1;
        }
        ++loads;
        const uint8_t *wg = P.wpack + (size_t)grp * L.groupBytes;
        mbar_expect_tx(bar_wfull, (uint32_t)L.groupBytes);
        for (int off = 0; off < L.groupBytes; off += 16384) {
          const int n = min(16384, L.groupBytes - off);
          tma_bulk_g2s(wbase + off, wg + off, (uint32_t)n, bar_wfull);
        }
      }
    } else if (lane == 1) {
      // ===== patch-table slices (16-byte aligned source: the slice starts at the entry below, `rem` entries early) =====
      int ts = 0, tp = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const NsTile T = ns_decode(tile, P.batch, L.tpf);
        const int MBc = min(L.MB, L.nmb - T.tf * L.MB);
        const int npix = NS_STRIDE * (MBc - 1) + 128 + L.haloV;
        const long long first = (long long)T.f * L.G + (long long)T.tf * L.MB * NS_STRIDE;
        const int rem = (int)(first & 3);
        const uint32_t bytes = (uint32_t)((npix + rem + 3) / 4 * 4) * 4u;
        mbar_wait(bar_tempty + 8 * ts, tp ^ 1);
        mbar_expect_tx(bar_tfull + 8 * ts, bytes * (1 + L.twoTabs));
        tma_bulk_g2s(tabring + ts * L.tabBytes, P.tab0 + (first - rem), bytes, bar_tfull + 8 * ts);
        if (L.twoTabs)
          tma_bulk_g2s(tabring + ts * L.tabBytes + (L.NPIXp + 8) * 4, P.tab1 + (first - rem), bytes, bar_tfull + 8 * ts);
        if (++ts == 2) { ts = 0; tp ^= 1; }
      }
    }
  } else if (warp >= NS_MMA_WARP && warp < NS_MMA_WARP + NS_MMA_WARPS) {
    // ===== MMA issuers: block-major -- every 128-row block gets its (K block, kernel row, K step) MMAs back to back into
    // its own ring slot and is committed on its own; the two warps take alternate blocks =====
    const int mw = warp - NS_MMA_WARP;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NS_N >> 3) << 17) | (8u << 24);
    const uint64_t a_fix = ((uint64_t)(((uint32_t)(8 * L.RB) >> 4) | (1u << 14) | ((uint32_t)L.layoutType << 29)) << 32) |
                           (1ull << 16);
    const uint64_t b_fix = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((uint64_t)((uint32_t)NS_N & 0x3FFFu) << 16);
    const uint32_t a_jstep = 32u >> 4, b_jstep = (uint32_t)(2 * NS_N * 16) >> 4;
    const uint32_t a_mbstep = (uint32_t)(NS_STRIDE * L.RB) >> 4, blk16 = (uint32_t)L.blockBytes >> 4,
                   urow16 = ((uint32_t)(P.dh * L.Wv) * L.RB) >> 4, unit16 = (uint32_t)L.unitBytes >> 4;
    const uint64_t b_base = b_fix | (wbase >> 4);
    int sp = 0, pp = 0, wph = 0, prev_grp = -1, gcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const NsTile T = ns_decode(tile, P.batch, L.tpf);
      const int MBc = min(L.MB, L.nmb - T.tf * L.MB);
      if (T.grp != prev_grp) {                   // a new weight image lands
        mbar_wait(bar_wfull, wph);
        wph ^= 1;
        tc_fence_after();
        prev_grp = T.grp;
      }
      const int next = tile + gridDim.x;
      const bool release_w = next >= ntiles || ns_decode(next, P.batch, L.tpf).grp != T.grp;
      mbar_wait(bar_pfull + 8 * sp, pp);
      tc_fence_after();
      const uint64_t a_stage = a_fix | ((patch0 + (uint32_t)sp * L.patchBytes) >> 4);
#pragma unroll 1
      for (int mb = 0; mb < MBc; ++mb) {
        const int g = gcount + mb;
        if ((g & 1) != mw) continue;
        const int slot = g % NS_SLOTS, sphase = (g / NS_SLOTS) & 1;
        mbar_wait(bar_aempty + 8 * slot, sphase ^ 1);
        tc_fence_after();
        const uint32_t d_slot = tmem_base + (uint32_t)(slot * NS_N);
        const uint64_t a_mb = a_stage + (uint32_t)mb * a_mbstep;
        uint64_t b_unit = b_base;
        uint32_t a_kc = 0;
#pragma unroll 1
        for (int kc = 0; kc < L.nch; ++kc, a_kc += blk16) {
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            if (u < L.kh) {
#pragma unroll
              for (int j = 0; j < KC16T; ++j)
                umma_bf16_elect(d_slot, a_mb + a_kc + (uint32_t)u * urow16 + (uint32_t)j * a_jstep,
                                b_unit + (uint32_t)j * b_jstep, idesc, (kc > 0 || u > 0 || j > 0) ? 1u : 0u);
              b_unit += unit16;
            }
          }
        }
        umma_commit_elect(bar_afull + 8 * slot);
      }
      gcount += MBc;
      umma_commit_elect(bar_pempty + 8 * sp);          // this warp's MMAs on the patch stage have been issued
      if (release_w) umma_commit_elect(bar_wempty);
      if (++sp == L.PS) { sp = 0; pp ^= 1; }
    }
  } else if (warp >= NS_EPI_WARP0 && warp < NS_EPI_WARP0 + 8) {
    // ===== epilogue: two groups of four warps (one per tensor-memory lane quarter) take alternate 128-row blocks =====
    const int quarter = warp & 3, grp_h = (warp - NS_EPI_WARP0) >> 2;
    const uint32_t stg = stg0 + (uint32_t)(warp - NS_EPI_WARP0) * (32u * 64u);
    const uint32_t rowB = (uint32_t)P.cout * 2u, cpr = rowB >> 4;
    const uint32_t smask = (cpr & (cpr - 1)) == 0 ? min(cpr, 8u) - 1u : 0u;
    // exchange area of this group: [parity][quarter 1..3][d1 of lane 0 | d2 of lane 0 | d2 of lane 1][32] float, then the
    // partial sums of lanes 30 / 31: [parity][quarter 0..2][2][32] float
    const uint32_t xgrp = xch0 + (uint32_t)grp_h * (2u * 3u * 5u * 128u);
    int gcount = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = tid - NS_EPI_WARP0 * 32; i < 3 * 32; i += NS_EPI) s_bias[i] = P.bias[i];
    asm volatile("bar.sync 2, %0;" ::"n"(NS_EPI) : "memory");
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const NsTile T = ns_decode(tile, P.batch, L.tpf);
      const int MBc = min(L.MB, L.nmb - T.tf * L.MB);
      const float *bias = s_bias + T.grp * 32;
      const size_t face_px = (size_t)(T.b * 6 + T.f) * P.Hout * P.Wout;
      for (int mb = 0; mb < MBc; ++mb) {
        const int g = gcount + mb;
        if ((g & 1) != grp_h) continue;
        const int slot = g % NS_SLOTS, sphase = (g / NS_SLOTS) & 1, par = (g >> 1) & 1;
        const int i = quarter * 32 + lane;
        const int q = (T.tf * L.MB + mb) * NS_STRIDE + i;
        const int r = q / L.Wv, c = q - r * L.Wv;
        const bool ok = i < NS_STRIDE && q < L.Q && c < P.Wout;
        const size_t opix = face_px + (size_t)r * P.Wout + c;
        const unsigned okmask = __ballot_sync(0xffffffffu, ok);
        const int nvalid = __popc(okmask);
        const size_t opix0 = __shfl_sync(0xffffffffu, opix, okmask ? __ffs(okmask) - 1 : 0);
        const uint32_t prow = (uint32_t)(opix - opix0);
        const uint32_t srow = stg + prow * rowB;
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * NS_N);
        const uint32_t xq_out = xgrp + (uint32_t)((par * 3 + (quarter - 1)) * 5) * 128u;       // what this quarter publishes
        const uint32_t xq_in = xgrp + (uint32_t)((par * 3 + quarter) * 5) * 128u;              // what the next one published
        mbar_wait(bar_afull + 8 * slot, sphase);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t d0[16], d1[16], d2[16];
          tmem_ld16(trow + (uint32_t)(16 * h), d0);
          tmem_ld16(trow + (uint32_t)(32 + 16 * h), d1);
          tmem_ld16(trow + (uint32_t)(64 + 16 * h), d2);
          tmem_ld_wait();
          if (h == 1) {                      // the slot has been read: hand it back to the issuers
            tc_fence_before();
            mbar_arrive(bar_aempty + 8 * slot);
          }
          float s[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float a = __shfl_down_sync(0xffffffffu, __uint_as_float(d1[e]), 1);
            const float b = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[e]), 2);
            s[e] = __uint_as_float(d0[e]) + (lane < 31 ? a : 0.f) + (lane < 30 ? b : 0.f);
          }
          if (quarter > 0 && lane < 2) {     // rows 0 / 1 of this quarter complete rows 30 / 31 of the previous one
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              if (lane == 0)
                st_shared_f32x4(xq_out + (uint32_t)(16 * h + 4 * e4) * 4u, __uint_as_float(d1[4 * e4]), __uint_as_float(d1[4 * e4 + 1]),
                                __uint_as_float(d1[4 * e4 + 2]), __uint_as_float(d1[4 * e4 + 3]));
              st_shared_f32x4(xq_out + (uint32_t)(lane == 0 ? 128 : 256) + (uint32_t)(16 * h + 4 * e4) * 4u, __uint_as_float(d2[4 * e4]),
                              __uint_as_float(d2[4 * e4 + 1]), __uint_as_float(d2[4 * e4 + 2]), __uint_as_float(d2[4 * e4 + 3]));
            }
          }
          if (quarter < 3 && lane >= 30) {   // partial sums of the two rows that need the next quarter
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4)
              st_shared_f32x4(xq_in + (uint32_t)(lane == 30 ? 384 : 512) + (uint32_t)(16 * h + 4 * e4) * 4u, s[4 * e4], s[4 * e4 + 1],
                              s[4 * e4 + 2], s[4 * e4 + 3]);
          } else if (ok) {
#pragma unroll
            for (int c8 = 0; c8 < 2; ++c8) {
              const int nb = 16 * h + 8 * c8;
              if (nb + 8 <= P.cout)
                st_shared16(srow + ((((uint32_t)nb >> 3) ^ (prow & smask)) << 4),
                            ns_finish8(s + 8 * c8, bias + nb, P.act, P.slope, P.maxv));
            }
          }
        }
        // the four warps of the block meet: rows 30 / 31 of quarters 0..2 are finished by eight lanes each
        asm volatile("bar.sync %0, 128;" ::"r"(3 + grp_h) : "memory");
        if (quarter < 3 && lane < 8) {
          const int rowsel = lane >> 2, nb = (lane & 3) * 8;
          const bool rok = (okmask >> (30 + rowsel)) & 1u;
          if (rok && nb + 8 <= P.cout) {
            const uint32_t pr = (uint32_t)__popc(okmask & ((1u << (30 + rowsel)) - 1u));
            float s[8];
#pragma unroll
            for (int e4 = 0; e4 < 2; ++e4) {
              const uint32_t o4 = (uint32_t)(nb + 4 * e4) * 4u;
              const float4 p = ld_shared_f32x4(xq_in + (uint32_t)(rowsel == 0 ? 384 : 512) + o4);
              float4 t;
              if (rowsel == 0) {
                t = ld_shared_f32x4(xq_in + 128u + o4);                    // + d2 of the next quarter's lane 0
              } else {
                const float4 t1 = ld_shared_f32x4(xq_in + o4);             // + d1 of lane 0 + d2 of lane 1
                const float4 t2 = ld_shared_f32x4(xq_in + 256u + o4);
                t = make_float4(t1.x + t2.x, t1.y + t2.y, t1.z + t2.z, t1.w + t2.w);
              }
              s[4 * e4] = p.x + t.x; s[4 * e4 + 1] = p.y + t.y; s[4 * e4 + 2] = p.z + t.z; s[4 * e4 + 3] = p.w + t.w;
            }
            st_shared16(stg + pr * rowB + ((((uint32_t)nb >> 3) ^ (pr & smask)) << 4),
                        ns_finish8(s, bias + nb, P.act, P.slope, P.maxv));
          }
        }
        __syncwarp();
        {   // the warp's valid rows are one contiguous run in HBM: coalesced 512-byte stores
          uint8_t *gdst = reinterpret_cast<uint8_t *>(P.y) + opix0 * rowB;
          const uint32_t total = (uint32_t)nvalid * rowB;
          for (uint32_t off = (uint32_t)lane * 16u; off < total; off += 512u) {
            const uint32_t row = off / rowB, ch = (off - row * rowB) >> 4;
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stg + row * rowB + ((ch ^ (row & smask)) << 4)));
            *reinterpret_cast<uint4 *>(gdst + off) = v;
          }
        }
        __syncwarp();
      }
      gcount += MBc;
    }
  } else if (warp < 8) {
    // ===== patch loaders: one table lookup per patch row, 16-byte cp.async gathers (SAME / UP2 sources) =====
    const int lt = tid;
    int si = 0, pi = 0, ts = 0, tp = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int chunk = lt & (L.S - 1), c = chunk * 8;
    const bool first = c < P.c0;
    const __nv_bfloat16 *src = first ? P.x0 : P.x1;
    const int C = first ? P.c0 : P.c1, cc = first ? c : c - P.c0;
    const bool chan_ok = c < P.cin;
    const int pstep = NS_LOADERS >> L.logS;
    const uint32_t cw = (uint32_t)chunk & ((1u << L.cprLog) - 1u);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const NsTile T = ns_decode(tile, P.batch, L.tpf);
      const int MBc = min(L.MB, L.nmb - T.tf * L.MB);
      const int npix = NS_STRIDE * (MBc - 1) + 128 + L.haloV;
      const int rem = (int)(((long long)T.f * L.G + (long long)T.tf * L.MB * NS_STRIDE) & 3);
      mbar_wait(bar_pempty + 8 * si, pi ^ 1);
      const uint32_t stage = patch0 + (uint32_t)si * L.patchBytes;
      mbar_wait(bar_tfull + 8 * ts, tp);
      const int32_t *t0 = s_tab + (size_t)ts * (L.tabBytes / 4) + rem;
      const int32_t *tab = (first || !L.twoTabs) ? t0 : t0 + (L.NPIXp + 8);
      const size_t boff = (size_t)T.b * (first ? P.ppb0 : P.ppb1);
      const uint32_t dst0 = stage + (uint32_t)(chunk >> L.cprLog) * L.blockBytes;
      for (int i0 = lt >> L.logS; i0 < npix; i0 += 8 * pstep) {
        int px[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int i = i0 + e * pstep;
          px[e] = (i < npix && chan_ok) ? tab[i] : -1;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int i = i0 + e * pstep;
          if (i < npix) {
            const uint32_t row = dst0 + (uint32_t)i * L.RB;
            const __nv_bfloat16 *gp = px[e] >= 0 ? src + (boff + px[e]) * C + cc : P.x0;
            cp_async16(row + ((cw ^ ((row >> 7) & (uint32_t)L.swzMask)) << 4), gp, px[e] >= 0 ? 16u : 0u);
          }
        }
      }
      mbar_arrive(bar_tempty + 8 * ts);
      if (++ts == 2) { ts = 0; tp ^= 1; }
      fence_proxy_async();
      cp_async_mbar_arrive(bar_pfull + 8 * si);
      if (++si == L.PS) { si = 0; pi ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NS_MMA_WARP) tmem_dealloc(tmem_base, 512u);
}

// ---- weight packing: packed[g][kc][u][k8][n = v*32 + o][8] bf16, then float bias[3][32] --------------------------------
__global__ void pack_ns_kernel(const float *__restrict__ w_eq, const float *__restrict__ w_pol, const float *__restrict__ w_np,
                               const float *__restrict__ b_eq, const float *__restrict__ b_pol, const float *__restrict__ b_np,
                               uint8_t *__restrict__ out, int kh, int cin, int cout, int scin, int scout, int flip, int KC,
                               long long groupElems) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < 3 * groupElems) {
    const int g = (int)(i / groupElems);
    long long r = i % groupElems;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % NS_N); r /= NS_N;
    const int k8 = (int)(r % (KC / 8)); r /= (KC / 8);
    const int u = (int)(r % kh);
    const int kc = (int)(r / kh);
    const int k = kc * KC + k8 * 8 + e, v = n >> 5, o = n & 31;
    float val = 0.f;
    if (k < cin && o < cout && k < scin && o < scout) {
      const float *src = g == 0 ? w_eq : (g == 1 ? w_pol : (w_np ? w_np : w_pol));
      const int us = (g == 2 && flip) ? kh - 1 - u : u;
      val = src[(((long long)us * 3 + v) * scin + k) * scout + o];
    }
    reinterpret_cast<__nv_bfloat16 *>(out)[i] = __float2bfloat16_rn(val);
  } else if (i < 3 * groupElems + 96) {
    const int j = (int)(i - 3 * groupElems), g = j >> 5, co = j & 31;
    const float *src = g == 0 ? b_eq : (g == 1 ? b_pol : (b_np ? b_np : b_pol));
    reinterpret_cast<float *>(out + 3 * groupElems * 2)[j] = (src && co < scout) ? src[co] : 0.f;
  }
}

int ns_env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

// returns nullptr when the narrow-layer kernel applies, else the reason
const char *make_ns_plan(const dlwpcs_conv_desc *d, const Geometry &g, NsPlan *L) {
  static const int enabled = ns_env_int("DLWPCS_TC_NS", 1);
  if (!enabled) return "disabled (DLWPCS_TC_NS=0)";
  if (d->x_dtype != DLWPCS_BF16 || d->y_dtype != DLWPCS_BF16) return "bf16 in and out only";
  if (d->stride_h != 1 || d->stride_w != 1 || d->kw != 3 || d->dil_w != 1 || d->kh > 3) return "3 kernel columns, stride 1";
  if (d->cout > 32 || d->cout <= 16 || d->cout % 8) return "17..32 output channels in multiples of 8";
  if (d->c0 % 8 || d->c1 % 8) return "input channels in multiples of 8";
  if (d->mode0 == DLWPCS_SRC_POOL2 || (d->c1 > 0 && d->mode1 == DLWPCS_SRC_POOL2)) return "no pooled source";
  L->CinP = (d->cin + 15) / 16 * 16;
  if (L->CinP > 32) L->CinP = (d->cin + 63) / 64 * 64;
  L->KC = L->CinP < 64 ? L->CinP : 64;
  L->nch = L->CinP / L->KC;
  L->S = L->CinP / 8;
  L->logS = -1;
  for (int l = 0; l <= 5; ++l)
    if ((1 << l) == L->S) L->logS = l;
  if (L->logS < 0) return "input channels do not give a power-of-two chunk count";
  L->RB = L->KC * 2;
  L->cprLog = L->RB == 128 ? 3 : (L->RB == 64 ? 2 : 1);
  L->swzMask = (1 << L->cprLog) - 1;
  L->layoutType = L->RB == 128 ? 2 : (L->RB == 64 ? 4 : 6);
  L->kh = d->kh;
  L->Wv = g.Wout + 2;
  L->haloV = (d->kh - 1) * d->dil_h * L->Wv;
  L->Q = (g.Hout - 1) * L->Wv + g.Wout;
  L->nmb = (L->Q + NS_STRIDE - 1) / NS_STRIDE;
  L->NU = L->nch * d->kh;
  L->unitBytes = L->KC * NS_N * 2;
  L->groupBytes = L->NU * L->unitBytes;
  if (L->groupBytes > 96 * 1024) return "weights of one face group do not fit";
  L->twoTabs = (d->c1 > 0 && d->mode1 != d->mode0) ? 1 : 0;
  L->G = ((L->nmb - 1) * NS_STRIDE + 128 + L->haloV + 8 + 3) / 4 * 4;
  const int fixed = 1024 /*align*/ + ((L->groupBytes + 1023) / 1024 * 1024) + 1024 /*misc + bias*/ + 8 * 2048 /*staging*/ +
                    2 * 2 * 3 * 5 * 128 /*exchange*/ + 256;
  int best = 0;
  for (int MB = L->nmb < 4 ? L->nmb : 4; MB >= 1 && !best; --MB) {
    const int np = (NS_STRIDE * (MB - 1) + 128 + L->haloV + 7) / 8 * 8;
    const int patch = (L->nch * np * L->RB + 1023) / 1024 * 1024;
    const int tabs = 2 * (1 + L->twoTabs) * (np + 8) * 4;
    if (fixed + tabs + 2 * patch <= NS_SMEM_CAP) best = MB;
  }
  if (!best) return "input patch does not fit shared memory";
  const int tiles = (L->nmb + best - 1) / best;
  best = (L->nmb + tiles - 1) / tiles;
  L->MB = best;
  L->tpf = (L->nmb + best - 1) / best;
  L->NPIXp = (NS_STRIDE * (best - 1) + 128 + L->haloV + 7) / 8 * 8;
  L->blockBytes = L->NPIXp * L->RB;
  L->patchBytes = (L->nch * L->NPIXp * L->RB + 1023) / 1024 * 1024;
  L->tabBytes = (1 + L->twoTabs) * (L->NPIXp + 8) * 4;
  L->PS = (NS_SMEM_CAP - fixed - 2 * L->tabBytes) / L->patchBytes;
  if (L->PS > NS_MAX_PS) L->PS = NS_MAX_PS;
  if (L->PS < 2) return "input patch does not fit shared memory twice";
  L->off_w = L->PS * L->patchBytes;
  L->off_misc = L->off_w + (L->groupBytes + 1023) / 1024 * 1024;
  L->off_bias = L->off_misc + 512;
  L->off_tab = L->off_misc + 1024;
  L->off_stg = (L->off_tab + 2 * L->tabBytes + 127) / 128 * 128;
  L->off_xch = L->off_stg + 8 * 2048;
  L->smemBytes = 1024 + L->off_xch + 2 * 2 * 3 * 5 * 128;
  if (L->smemBytes > NS_SMEM_CAP) return "shared memory plan exceeds the SM";
  return nullptr;
}

bool ns_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool ns_eligible(const dlwpcs_conv_desc *d, const Geometry &g) {
  NsPlan L;
  return make_ns_plan(d, g, &L) == nullptr;
}

int64_t ns_packed_weight_bytes(const dlwpcs_conv_desc *d, const Geometry &g) {
  NsPlan L;
  if (make_ns_plan(d, g, &L)) return -1;
  return 3LL * L.groupBytes + 96 * 4;
}

int ns_pack_weights(const dlwpcs_conv_desc *d, const Geometry &g, const dlwpcs_conv_weights *w, int src_cin, int src_cout,
                    void *packed, cudaStream_t st) {
  NsPlan L;
  const char *r = make_ns_plan(d, g, &L);
  CS_CHECK(r == nullptr, "narrow-layer kernel does not apply: %s", r);
  const long long groupElems = L.groupBytes / 2, total = 3 * groupElems + 96;
  pack_ns_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      w->w_eq, w->w_pol, d->independent_north_pole ? w->w_np : nullptr, d->use_bias ? w->b_eq : nullptr,
      d->use_bias ? w->b_pol : nullptr, (d->use_bias && d->independent_north_pole) ? w->b_np : nullptr, (uint8_t *)packed,
      d->kh, d->cin, d->cout, src_cin, src_cout, d->flip_north_pole, L.KC, groupElems);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int ns_conv_fwd(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed, void *y,
                cudaStream_t st) {
  NsP P;
  memset(&P, 0, sizeof(P));
  const char *r = make_ns_plan(d, g, &P.L);
  CS_CHECK(r == nullptr, "narrow-layer kernel does not apply: %s", r);
  const NsPlan &L = P.L;
  CS_CHECK(ns_aligned16(x0) && (d->c1 == 0 || ns_aligned16(x1)) && ns_aligned16(y), "tensors must be 16-byte aligned");
  P.x0 = (const __nv_bfloat16 *)x0;
  P.x1 = (const __nv_bfloat16 *)x1;
  P.tab0 = get_patch_table(g, L.Wv, L.G, d->n, d->halo, d->mode0);
  if (!P.tab0) return 3;
  P.tab1 = P.tab0;
  if (d->c1 > 0) {
    P.tab1 = get_patch_table(g, L.Wv, L.G, d->n, d->halo, d->mode1);
    if (!P.tab1) return 3;
  }
  auto ppb = [&](int mode) {
    const int e = mode == DLWPCS_SRC_SAME ? d->n : (mode == DLWPCS_SRC_UP2 ? d->n / 2 : d->n * 2);
    return 6 * e * e;
  };
  P.ppb0 = ppb(d->mode0);
  P.ppb1 = ppb(d->mode1);
  P.wpack = (const uint8_t *)packed;
  P.bias = reinterpret_cast<const float *>(P.wpack + 3LL * L.groupBytes);
  P.y = (__nv_bfloat16 *)y;
  P.batch = d->batch; P.Hout = g.Hout; P.Wout = g.Wout;
  P.cin = d->cin; P.cout = d->cout; P.c0 = d->c0; P.c1 = d->c1;
  P.dh = d->dil_h;
  P.act = d->act; P.slope = d->act_slope; P.maxv = d->act_max;
  typedef void (*kern_t)(const NsP);
  static const kern_t kerns[4] = {conv_tc_ns_kernel<1>, conv_tc_ns_kernel<2>, conv_tc_ns_kernel<3>, conv_tc_ns_kernel<4>};
  static bool attr_set[kMaxDevices][4] = {};
  const int kj = L.KC / 16 - 1, dev_i = current_device_index();
  CS_CHECK(kj >= 0 && kj < 4, "internal: bad K block");
  if (!attr_set[dev_i][kj]) {
    CS_CUDA(cudaFuncSetAttribute(kerns[kj], cudaFuncAttributeMaxDynamicSharedMemorySize, NS_SMEM_CAP));
    attr_set[dev_i][kj] = true;
  }
  const long long ntiles = 6LL * d->batch * L.tpf;
  CS_CHECK(ntiles < (1LL << 30), "batch too large");
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_i) != cudaSuccess || sms <= 0) sms = 148;
  int grid = sms < ntiles ? sms : (int)ntiles;
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  while (grid > 1 && L.nmb % L.MB != 0 && gcd(grid, L.tpf) != 1) --grid;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NS_THREADS);
  cfg.dynamicSmemBytes = (size_t)L.smemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CS_CUDA(cudaLaunchKernelEx(&cfg, kerns[kj], P));
  return 0;
}

}  // namespace dlwpcs
