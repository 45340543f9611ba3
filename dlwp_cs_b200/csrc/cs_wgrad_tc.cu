// tcgen05 (sm_100a) bf16 weight-gradient kernel of the cubed-sphere convolution.
//
// Reference semantics: what TensorFlow autodiff derives from DLWP/custom.py:921-1002 (CubeSphereConv2D.call) after
// CubeSpherePadding2D (custom.py:1198-1308) -- SURVEY.md section 8 row a7:
//   dW_g[u, v, ci, co] = sum_{faces f of group g} sum_{b, r, c} x_pad[b, f, r + u*dh, c + v, ci] * dyM[b, f, r, c, co]
//   db_g[co]           = sum dyM,          dyM = dy * act'(y)        (groups: equatorial 0-3, south pole, north pole)
//
// GEMM view.  K runs over the linear positions q = r*Wv + c of the virtual (halo-padded) face -- the same linearisation
// as the forward kernel (cs_tc.cu), so a tap is a row offset of one shared-memory patch and the halo exchange is one
// table lookup per patch row.  Both operands sit in shared memory exactly as they sit in HBM, one row of channels per
// pixel (16-byte chunks XOR-swizzled with the row's address bits), which is the *MN-major* canonical UMMA layout
// (cute/atom/mma_traits_sm100.hpp; verified on B200 with tools/umma_mn_probe.cu):
//   A[m = (shift j, ci), k = q] = patch[q + u*dh*Wv + j][ci]   the leading-dimension byte offset of the descriptor is ONE
//                                                              PIXEL ROW, so the 128 / CinBlk atoms along M are the
//                                                              patch shifted by 0, 1, 2, ... pixels: the horizontal
//                                                              taps v = j of one kernel row u share a single MMA
//   B[n = co, k = q]            = dyM[q][co]                   rows with c >= Wout or q >= Q are zero
//   D[u][(j, ci), co]           fp32 in tensor memory, accumulated over every tile the CTA owns, read once at the end.
// A CTA belongs to one job (block of 64 input channels x group of output channels whose accumulators fit the 512 TMEM
// columns) and one face group; it walks its tiles of TP = 128*TPB positions, then dumps its partial sums to the
// workspace; a second kernel adds the partials in a fixed order (deterministic), un-flips the north-pole rows and
// merges the two polar faces (custom.py:965-996).
// Warp roles (288 threads): warps 0-7 loaders (x patch: 16-byte cp.async gathers through the halo table; dy: registers,
// activation-derivative mask from the forward output, bias partial sums) and, after the loop, the accumulator dump;
// warp 8 allocates tensor memory and issues the MMAs.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <vector>
#include "cs_common.cuh"
#include "cs_ptx.cuh"

namespace dlwpcs {

namespace {

constexpr int WG_THREADS = 320;      // warps 0-3 x loaders, 4-7 dy loaders, 8 MMA issuer, 9 patch-table TMA producer
constexpr int WG_XL = 128, WG_YL = 128;
constexpr int WG_LOADERS = WG_XL + WG_YL;
constexpr int WG_MMA_WARP = 8, WG_TAB_WARP = 9;
constexpr int WG_MAX_UNITS = 32;
constexpr int WG_MAX_PS = 4;
constexpr int WG_TS = 3;             // patch-table ring slots
constexpr int WG_SMEM_CAP = 227 * 1024;
constexpr int WG_MISC = 256 + WG_YL * 8 * 4;     // barriers + tensor-memory slot, bias partials

struct WgPlan {
  int CinP, CoutP;
  int CinBlk, NCB;            // input channels per A atom row (16 / 32 / 64), 64-channel blocks (jobs along cin)
  int RBx, RBy;               // bytes per pixel row in shared memory
  int SPB, MBu;               // pixel shifts per 128-row M block, M blocks per kernel row
  int NBlk, NB, NBJ, NNG, NJ; // N per MMA, dy channel blocks: total / per job, job groups along cout, channels per job
  int J, nc, ncta, ng[3];     // jobs, CTAs per job, CTAs, CTAs per face group inside a job
  int kh, kw, dh;
  int Wv, Q, TPB, TP, tpf;    // virtual width, linear outputs per face, tile size in 128-blocks / positions, tiles per face
  int NPX, G;                 // patch rows per stage, patch-table entries per face
  int xBytes, yBlockBytes, stageBytes, PS;
  int nacc, NU, tmemCols;
  int smemBytes;
  int vecx, vecy;
  int logSx, logSy;
};

struct WgP {
  const __nv_bfloat16 *x, *dy, *mask_y;
  const int32_t *tabx;        // [6][G] patch position -> source pixel of one batch element or -1
  float *ws, *ws_b;           // [ncta][nacc][128][NJ], [ncta][NJ]
  int batch, ppbx, ppfy;      // x pixels per batch element, dy pixels per face
  int Ho, Wo;
  int knock;                  // bottleneck analysis (DLWPCS_WG_KNOCK): 1 no x gathers, 2 no dy loads, 4 no MMAs, 8 no dump
  int cin, cout;
  int act;
  float slope, maxv;
  WgPlan pl;
};

__device__ __forceinline__ uint32_t layout_code(int RB) { return RB == 128 ? 2u : (RB == 64 ? 4u : 6u); }

// eight consecutive channels [c0, c0+8) of one pixel as floats (zero beyond C); vec: one 16-byte load
__device__ __forceinline__ void load8(const __nv_bfloat16 *src, size_t pix, int C, int c0, bool vec, float *f) {
  if (vec) {
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4 *>(src + pix * C + c0)), f);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = c0 + e < C ? __bfloat162float(src[pix * C + c0 + e]) : 0.f;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// KH / MBU / NBJ: compile-time copies of the plan's kh / MBu / NBJ (0 = read them at run time) so that the MMA issue
// loops unroll -- an un-unrolled iteration costs ~110 cycles of uniform-datapath latency (tools/umma_mn_rate.cu)
template <int KH, int MBU, int NBJT>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgP P) {
  extern __shared__ uint8_t smem_raw[];
  const WgPlan &L = P.pl;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage0 = base;
  const uint32_t misc = stage0 + (uint32_t)L.PS * L.stageBytes;
  const uint32_t bar_full = misc, bar_empty = misc + 32, bar_done = misc + 64, tmem_slot = misc + 72,
                 bar_tfull = misc + 96, bar_tempty = misc + 128;
  float *s_bsum = reinterpret_cast<float *>(gen + (misc - base) + 256);
  const uint32_t tabring = misc + WG_MISC;                                  // WG_TS slots of NPX int32
  const int32_t *s_tab = reinterpret_cast<const int32_t *>(gen + (tabring - base));

  // warp index broadcast from lane 0: tells the compiler the role branches are warp-uniform (uniform datapath usable)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

  // ---- which job / face group / tiles this CTA owns
  const int job = blockIdx.x / L.nc, ic = blockIdx.x - job * L.nc;
  const int cb = job % L.NCB, ngi = job / L.NCB;
  int grp = 0, idx = ic, nin = L.ng[0];
  if (ic >= L.ng[0] + L.ng[1]) { grp = 2; idx = ic - L.ng[0] - L.ng[1]; nin = L.ng[2]; }
  else if (ic >= L.ng[0]) { grp = 1; idx = ic - L.ng[0]; nin = L.ng[1]; }
  const int T = (grp == 0 ? 4 : 1) * P.batch * L.tpf;
  const int my_tiles = idx < T ? (T - idx + nin - 1) / nin : 0;
  // tile k of this CTA -> (batch element, face, tile of the face)
  auto decode = [&](int k, int &b, int &f, int &tf) {
    const int t = idx + k * nin;
    if (grp == 0) { const int bf = t / L.tpf; tf = t - bf * L.tpf; b = bf >> 2; f = bf & 3; }
    else { b = t / L.tpf; tf = t - b * L.tpf; f = 3 + grp; }
  };

  if (tid == 0) {
    for (int i = 0; i < L.PS; ++i) {
      mbar_init(bar_full + 8 * i, WG_LOADERS);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < WG_TS; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, WG_XL);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, (uint32_t)L.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(gen + (tmem_slot - base));

  if (warp == WG_TAB_WARP) {
    // ===== patch-table producer: the slice of every tile's halo table goes to shared memory ahead of the x loaders, so
    // that their lookups do not queue behind their own gathers =====
    if (lane == 0) {
      int ts = 0, tp = 0;
      for (int k = 0; k < my_tiles; ++k) {
        int b, f, tf;
        decode(k, b, f, tf);
        mbar_wait(bar_tempty + 8 * ts, tp ^ 1);
        mbar_expect_tx(bar_tfull + 8 * ts, (uint32_t)L.NPX * 4u);
        tma_bulk_g2s(tabring + (uint32_t)ts * L.NPX * 4u, P.tabx + (size_t)f * L.G + (size_t)tf * L.TP, (uint32_t)L.NPX * 4u,
                     bar_tfull + 8 * ts);
        if (++ts == WG_TS) { ts = 0; tp ^= 1; }
      }
    }
  } else if (warp == WG_MMA_WARP) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(L.NBlk >> 3) << 17) |
                           (8u << 24);
    const uint64_t a_hi = (uint64_t)((((uint32_t)(8 * L.RBx)) >> 4) | (1u << 14) | (layout_code(L.RBx) << 29)) << 32;
    const uint64_t b_hi = (uint64_t)((((uint32_t)(8 * L.RBy)) >> 4) | (1u << 14) | (layout_code(L.RBy) << 29)) << 32;
    const uint32_t a_lbo = ((uint32_t)L.RBx >> 4) << 16, b_lbo = 1u << 16;
    const uint32_t a_kstep = ((uint32_t)(16 * L.RBx)) >> 4, b_kstep = ((uint32_t)(16 * L.RBy)) >> 4;
    const int ksteps = L.TP / 16;
    const uint32_t a_ustep = ((uint32_t)(L.dh * L.Wv) * L.RBx) >> 4, a_mstep = ((uint32_t)L.SPB * L.RBx) >> 4;
    const uint32_t b_nstep = (uint32_t)L.yBlockBytes >> 4;
    int s = 0, ph = 0;
    for (int k = 0; k < my_tiles; ++k) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      const uint32_t st_addr = stage0 + (uint32_t)s * L.stageBytes;
      const uint32_t a_stage = a_lbo | (st_addr >> 4), b_stage = b_lbo | ((st_addr + (uint32_t)L.xBytes) >> 4);
      // whole-warp loops, one elected lane per instruction: every descriptor stays in uniform registers
      const int khc = KH ? KH : L.kh, mbuc = MBU ? MBU : L.MBu, nbjc = NBJT ? NBJT : L.NBJ;
#pragma unroll 2
      for (int ks = 0; ks < ((P.knock & 4) ? 0 : ksteps); ++ks) {
        const uint32_t a_ks = a_stage + (uint32_t)ks * a_kstep, b_ks = b_stage + (uint32_t)ks * b_kstep;
        const uint32_t acc = (k > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
        for (int u = 0; u < khc; ++u) {
#pragma unroll
          for (int mbu = 0; mbu < mbuc; ++mbu) {
#pragma unroll
            for (int nbj = 0; nbj < nbjc; ++nbj)
              umma_bf16_elect(tmem_base + (uint32_t)((u * mbuc + mbu) * nbjc + nbj) * (uint32_t)L.NBlk,
                              a_hi | (uint64_t)(a_ks + (uint32_t)u * a_ustep + (uint32_t)mbu * a_mstep),
                              b_hi | (uint64_t)(b_ks + (uint32_t)nbj * b_nstep), idesc, acc);
          }
        }
      }
      umma_commit_elect(bar_empty + 8 * s);        // the stage may be refilled once these MMAs have read it
      if (k == my_tiles - 1) umma_commit_elect(bar_done);
      if (++s == L.PS) { s = 0; ph ^= 1; }
    }
  } else if (warp < 4) {
    // ===== x loaders: rows q0 .. q0 + NPX of the virtual face through the halo table, 16-byte cp.async gathers =====
    const int lt = tid;
    const int Sx = L.CinBlk / 8;
    const int chx = lt & (Sx - 1);
    const int cx = cb * L.CinBlk + chx * 8;                       // first input channel of this thread's chunk
    const bool okx = cx < P.cin;
    const int pstx = WG_XL >> L.logSx;
    const uint32_t swx = (uint32_t)(L.RBx / 16 - 1);
    int s = 0, ph = 0, ts = 0, tp = 0;
    for (int k = 0; k < my_tiles; ++k) {
      int b, f, tf;
      decode(k, b, f, tf);
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      mbar_wait(bar_tfull + 8 * ts, tp);
      const uint32_t st_addr = stage0 + (uint32_t)s * L.stageBytes;
      const int32_t *tab = s_tab + (size_t)ts * L.NPX;
      const size_t xb = (size_t)b * P.ppbx;
      if (P.knock & 1) {
      } else if (L.vecx) {
        for (int i0 = lt >> L.logSx; i0 < L.NPX; i0 += 8 * pstx) {
          int px[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = i0 + e * pstx;
            px[e] = (i < L.NPX && okx) ? tab[i] : -1;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = i0 + e * pstx;
            if (i < L.NPX) {
              const uint32_t row = st_addr + (uint32_t)i * L.RBx;
              const __nv_bfloat16 *g = px[e] >= 0 ? P.x + (xb + px[e]) * P.cin + cx : P.x;
              cp_async16(row + ((((uint32_t)chx) ^ ((row >> 7) & swx)) << 4), g, px[e] >= 0 ? 16u : 0u);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int i = lt >> L.logSx; i < L.NPX; i += pstx) {
          const int px = okx ? tab[i] : -1;
          float a[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] = 0.f;
          if (px >= 0) load8(P.x, xb + px, P.cin, cx, false, a);
          const uint32_t row = st_addr + (uint32_t)i * L.RBx;
          st_shared16(row + ((((uint32_t)chx) ^ ((row >> 7) & swx)) << 4),
                      make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7])));
        }
      }
      mbar_arrive(bar_tempty + 8 * ts);
      if (++ts == WG_TS) { ts = 0; tp ^= 1; }
      fence_proxy_async();
      cp_async_mbar_arrive(bar_full + 8 * s);
      if (++s == L.PS) { s = 0; ph ^= 1; }
    }
  } else {
    // ===== dy loaders: through registers -- activation-derivative mask, bias partial sums, bf16 rounding of the product.
    // Rows with a discarded column (c >= Wout) or beyond the face are zero. =====
    const int lt = tid - WG_XL;
    const int Sy = L.NJ / 8;
    const int chy = lt & (Sy - 1);
    const int cy = ngi * L.NJ + chy * 8;                          // first output channel of this thread's chunk
    const bool oky = cy < P.cout;
    const int psty = WG_YL >> L.logSy;
    const uint32_t swy = (uint32_t)(L.RBy / 16 - 1);
    const uint32_t yblk = (uint32_t)(chy / (L.RBy / 16)), ycw = (uint32_t)(chy % (L.RBy / 16));
    const bool direct = L.vecy && !P.mask_y;
    const bool want_bias = cb == 0 && P.ws_b != nullptr;
    float bsum[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bsum[e] = 0.f;
    uint32_t prev_ybase = 0;
    // add this thread's own chunks of one dy tile (zero rows included: they add nothing) to its bias partial sums
    auto readback = [&](uint32_t yb_s) {
      for (int i = lt >> L.logSy; i < L.TP; i += psty) {
        const uint32_t row = yb_s + (uint32_t)i * L.RBy;
        uint4 v4;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v4.x), "=r"(v4.y), "=r"(v4.z), "=r"(v4.w)
                     : "r"(row + ((ycw ^ ((row >> 7) & swy)) << 4)));
        float f[8];
        unpack_bf16x8(v4, f);
#pragma unroll
        for (int c = 0; c < 8; ++c) bsum[c] += f[c];
      }
    };
    int s = 0, ph = 0;
    for (int k = 0; k < my_tiles; ++k) {
      int b, f, tf;
      decode(k, b, f, tf);
      const int q0 = tf * L.TP;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      const uint32_t st_addr = stage0 + (uint32_t)s * L.stageBytes;
      const size_t yb = (size_t)(b * 6 + f) * P.ppfy;
      const uint32_t ybase = st_addr + (uint32_t)L.xBytes + yblk * (uint32_t)L.yBlockBytes;
      if (P.knock & 2) {
      } else if (direct) {
        // no mask: 16-byte cp.async straight into the tile.  The bias partial sums are taken from shared memory one tile
        // later: every thread reads back the chunks it copied itself once its previous copy group has landed.
        for (int i0 = lt >> L.logSy; i0 < L.TP; i0 += 8 * psty) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = i0 + e * psty, q = q0 + i;
            if (i < L.TP) {
              const int r = q / L.Wv, c = q - r * L.Wv;
              const bool ok = oky && r < P.Ho && c < P.Wo;
              const uint32_t row = ybase + (uint32_t)i * L.RBy;
              const __nv_bfloat16 *g = ok ? P.dy + (yb + (size_t)(r * P.Wo + c)) * P.cout + cy : P.dy;
              cp_async16(row + ((ycw ^ ((row >> 7) & swy)) << 4), g, ok ? 16u : 0u);
            }
          }
        }
        if (want_bias) {
          cp_async_commit();
          if (k > 0) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            readback(prev_ybase);
          }
          prev_ybase = ybase;
        }
      } else
      for (int i0 = lt >> L.logSy; i0 < L.TP; i0 += 8 * psty) {
        int px[8];
        uint4 dv[8], mv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int i = i0 + e * psty, q = q0 + i;
          const int r = q / L.Wv, c = q - r * L.Wv;
          px[e] = (i < L.TP && oky && r < P.Ho && c < P.Wo) ? r * P.Wo + c : -1;
        }
        if (L.vecy) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (px[e] >= 0) {
              dv[e] = __ldg(reinterpret_cast<const uint4 *>(P.dy + (yb + px[e]) * P.cout + cy));
              if (P.mask_y) mv[e] = __ldg(reinterpret_cast<const uint4 *>(P.mask_y + (yb + px[e]) * P.cout + cy));
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int i = i0 + e * psty;
          if (i < L.TP) {
            uint4 o = make_uint4(0, 0, 0, 0);
            if (px[e] >= 0) {
              float v[8], m[8];
              if (L.vecy) {
                unpack_bf16x8(dv[e], v);
                if (P.mask_y) unpack_bf16x8(mv[e], m);
              } else {
                load8(P.dy, yb + px[e], P.cout, cy, false, v);
                if (P.mask_y) load8(P.mask_y, yb + px[e], P.cout, cy, false, m);
              }
              if (P.mask_y) {
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] *= act_grad_from_y(m[c], P.act, P.slope, P.maxv);
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) bsum[c] += v[c];
              o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
            const uint32_t row = ybase + (uint32_t)i * L.RBy;
            st_shared16(row + ((ycw ^ ((row >> 7) & swy)) << 4), o);
          }
        }
      }
      fence_proxy_async();
      cp_async_mbar_arrive(bar_full + 8 * s);
      if (++s == L.PS) { s = 0; ph ^= 1; }
    }
    if (direct && want_bias && my_tiles > 0 && !(P.knock & 2)) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      readback(prev_ybase);
    }
    // ---- bias partials of this CTA (only the cin-block-0 jobs report them): fixed-order sum over the dy loaders
    if (want_bias) {
#pragma unroll
      for (int e = 0; e < 8; ++e) s_bsum[lt * 8 + e] = bsum[e];
      named_bar_sync(1, WG_YL);
      if (lt < L.NJ) {
        const int ch = lt >> 3, e = lt & 7;
        float acc = 0.f;
        for (int th = ch; th < WG_YL; th += Sy) acc += s_bsum[th * 8 + e];
        P.ws_b[(size_t)blockIdx.x * L.NJ + lt] = acc;
      }
    }
  }
  if (warp < 8) {
    // ---- accumulator dump by the eight loader warps: TMEM -> workspace [acc][lane][NJ]
    if (my_tiles > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
    }
    const int quarter = warp & 3, half = warp >> 2;
    float *wsc = P.ws + (size_t)blockIdx.x * L.nacc * 128 * L.NJ;
    const int nchunks = L.nacc * L.NJ / 16;
    for (int cc = half; cc < ((P.knock & 8) ? 0 : nchunks); cc += 2) {
      uint32_t v[16];
      if (my_tiles > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cc * 16), v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0u;
      }
      const int col0 = cc * 16, a = col0 / L.NJ, col = col0 - a * L.NJ;
      float4 *dst = reinterpret_cast<float4 *>(wsc + ((size_t)a * 128 + quarter * 32 + lane) * L.NJ + col);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                             __uint_as_float(v[4 * e + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) tmem_dealloc(tmem_base, (uint32_t)L.tmemCols);
}

// second pass: fixed-order sum of the CTA partials of each (job, face group); un-flip / merge the north-pole share.
// A block handles 32 consecutive outputs (lanes; consecutive output channels -> coalesced) with 8 warps that each add
// every 8th CTA partial; the 8 slice sums are then added in slice order (deterministic).
// V = 4: every thread sums four consecutive output channels with 16-byte loads (cout and the column block are multiples
// of 4: one quarter of the load instructions for the same bytes -- the pass is a latency-bound sweep over L2); V = 1 for
// odd channel counts.
template <int V>
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float *__restrict__ ws, const float *__restrict__ ws_b,
                                                              float *dw_eq, float *dw_pol, float *dw_np, float *db_eq,
                                                              float *db_pol, float *db_np, int kh, int kw, int cin,
                                                              int cout, int flip, const WgPlan L) {
  __shared__ float sh[3][8][32][V];
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const long long per = (long long)kh * kw * cin * cout;
  const long long i = (blockIdx.x * 32LL + o) * V;           // first of the V outputs of this thread
  const size_t cta_stride = (size_t)L.nacc * 128 * L.NJ;
  const int gfirst[3] = {0, L.ng[0], L.ng[0] + L.ng[1]};
  float part[3][V];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int e = 0; e < V; ++e) part[g][e] = 0.f;
  auto accumulate = [&](const float *src, int n, size_t stride, float *acc) {
    for (int c = sl; c < n; c += 8) {
      if constexpr (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(src + (size_t)c * stride));
        acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
      } else {
        acc[0] += src[(size_t)c * stride];
      }
    }
  };
  if (i < per) {
    const int co = (int)(i % cout);
    long long r = i / cout;
    const int ci = (int)(r % cin); r /= cin;
    const int v = (int)(r % kw);
    const int u = (int)(r / kw);
    const int cbk = ci / L.CinBlk, cil = ci - cbk * L.CinBlk;
    const int ngi = co / L.NJ, col = co - ngi * L.NJ;
    const int job = ngi * L.NCB + cbk;
    const int mbu = v / L.SPB, js = v - mbu * L.SPB, lane = js * L.CinBlk + cil;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int uu = (g == 2 && flip) ? kh - 1 - u : u;              // packed kernel row that reads source row u
      const size_t off = ((size_t)(uu * L.MBu + mbu) * 128 + lane) * L.NJ + col;
      accumulate(ws + (size_t)(job * L.nc + gfirst[g]) * cta_stride + off, L.ng[g], cta_stride, part[g]);
    }
  } else if (db_eq && i < per + cout) {
    const int co = (int)(i - per);
    const int ngi = co / L.NJ, col = co - ngi * L.NJ;
    const int job = ngi * L.NCB;
#pragma unroll
    for (int g = 0; g < 3; ++g)
      accumulate(ws_b + (size_t)(job * L.nc + gfirst[g]) * L.NJ + col, L.ng[g], (size_t)L.NJ, part[g]);
  }
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int e = 0; e < V; ++e) sh[g][sl][o][e] = part[g][e];
  __syncthreads();
  if (sl != 0) return;
#pragma unroll
  for (int e = 0; e < V; ++e) {
    float tot[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += sh[g][k][o][e];
      tot[g] = acc;
    }
    const long long ie = i + e;
    if (ie < per) {
      dw_eq[ie] = tot[0];
      if (dw_np) { dw_pol[ie] = tot[1]; dw_np[ie] = tot[2]; }
      else dw_pol[ie] = tot[1] + tot[2];
    } else if (db_eq && ie < per + cout) {
      const int co = (int)(ie - per);
      db_eq[co] = tot[0];
      if (db_np) { db_pol[co] = tot[1]; db_np[co] = tot[2]; }
      else db_pol[co] = tot[1] + tot[2];
    }
  }
}

int env_int_wg(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

int pad_channels(int c) {
  int p = (c + 15) / 16 * 16;
  if (p > 32) p = (c + 63) / 64 * 64;
  return p;
}

int ilog2_exact(int v) {
  for (int l = 0; l < 16; ++l)
    if ((1 << l) == v) return l;
  return -1;
}

int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

const char *make_wg_plan(const dlwpcs_conv_desc *d, const Geometry &g, WgPlan *L) {
  memset(L, 0, sizeof(*L));
  if (d->stride_h != 1 || d->stride_w != 1) return "strides must be 1";
  if (d->dil_w != 1) return "horizontal dilation must be 1 (the horizontal taps are stacked along M)";
  if (d->x_dtype != DLWPCS_BF16 || d->y_dtype != DLWPCS_BF16) return "activations must be bfloat16";
  if (d->c1 != 0 || d->mode0 != DLWPCS_SRC_SAME) return "fused input sampling is forward-only";
  L->kh = d->kh; L->kw = d->kw; L->dh = d->dil_h;
  L->CinP = pad_channels(d->cin);
  L->CoutP = pad_channels(d->cout);
  if (L->CoutP > 256) return "more than 256 output channels";
  L->CinBlk = L->CinP < 64 ? L->CinP : 64;
  L->NCB = L->CinP / L->CinBlk;
  L->RBx = 2 * L->CinBlk;
  L->SPB = 128 / L->CinBlk;
  L->MBu = (d->kw + L->SPB - 1) / L->SPB;
  L->NBlk = L->CoutP < 64 ? L->CoutP : 64;
  L->RBy = 2 * L->NBlk;
  L->NB = L->CoutP / L->NBlk;
  L->nacc = d->kh * L->MBu;
  L->NBJ = 512 / (L->nacc * L->NBlk);
  if (L->NBJ < 1) return "accumulators of one output-channel block exceed tensor memory (kernel window too large)";
  if (L->NBJ > L->NB) L->NBJ = L->NB;
  while (L->NB % L->NBJ) --L->NBJ;
  if ((L->NBJ & (L->NBJ - 1)) != 0) return "unsupported output-channel split";
  L->NNG = L->NB / L->NBJ;
  L->NJ = L->NBJ * L->NBlk;
  L->NU = L->nacc * L->NBJ;
  if (L->NU > WG_MAX_UNITS) return "kernel window too large";
  L->logSx = ilog2_exact(L->CinBlk / 8);
  L->logSy = ilog2_exact(L->NJ / 8);
  if (L->logSx < 0 || L->logSy < 0) return "internal: chunk counts must be powers of two";
  int cols = 32;
  while (cols < L->nacc * L->NJ) cols *= 2;
  L->tmemCols = cols;
  L->Wv = g.Wout + (d->kw - 1);
  L->Q = (g.Hout - 1) * L->Wv + g.Wout;
  const int nmb = (L->Q + 127) / 128;
  const int extX = (d->kh - 1) * d->dil_h * L->Wv + L->MBu * L->SPB - 1;
  int best = 0;
  const int forced = env_int_wg("DLWPCS_WG_TPB", 0);
  for (int tpb = (forced > 0 ? forced : 4); tpb >= 1 && !best; --tpb) {
    const int np = (tpb * 128 + extX + 7) / 8 * 8;
    const int xb = (np * L->RBx + 1023) / 1024 * 1024;
    const int stage = xb + L->NBJ * tpb * 128 * L->RBy;
    if (2 * stage + WG_MISC + WG_TS * np * 4 + 1024 <= WG_SMEM_CAP) best = tpb;
  }
  if (!best) return "tile does not fit shared memory";
  if (best > nmb) best = nmb;
  const int tiles = (nmb + best - 1) / best;
  best = (nmb + tiles - 1) / tiles;                 // even out the tiles of a face
  if (forced <= 0) {
    // tiles are uniform (positions beyond the face are zero rows that still cost MMAs): take a smaller tile when it
    // removes at least 10 % of the padded 128-blocks (5 blocks at 24x24: 5 tiles of 1 instead of 2 tiles of 3)
    const int padded = (nmb + best - 1) / best * best;
    for (int tpb = best - 1; tpb >= 1; --tpb) {
      const int p2 = (nmb + tpb - 1) / tpb * tpb;
      if (10 * p2 <= 9 * padded) { best = tpb; break; }
    }
  }
  L->TPB = best;
  L->TP = best * 128;
  L->tpf = (nmb + best - 1) / best;
  L->NPX = (L->TP + extX + 7) / 8 * 8;
  L->G = ((L->tpf - 1) * L->TP + L->NPX + 3) / 4 * 4;
  L->xBytes = (L->NPX * L->RBx + 1023) / 1024 * 1024;
  L->yBlockBytes = L->TP * L->RBy;
  L->stageBytes = L->xBytes + L->NBJ * L->yBlockBytes;
  L->PS = (WG_SMEM_CAP - WG_MISC - WG_TS * L->NPX * 4 - 1024) / L->stageBytes;
  int ps_max = env_int_wg("DLWPCS_WG_PS", 3);
  if (ps_max > WG_MAX_PS) ps_max = WG_MAX_PS;
  if (L->PS > ps_max) L->PS = ps_max;
  if (L->PS < 1) return "tile does not fit shared memory";
  L->smemBytes = 1024 + L->PS * L->stageBytes + WG_MISC + WG_TS * L->NPX * 4;
  // CTAs: equal share per job, split 4 : 1 : 1 over the face groups (equatorial faces hold 4/6 of the tiles)
  L->J = L->NCB * L->NNG;
  const long long tiles_all = 6LL * d->batch * L->tpf;
  int nc = sm_count() / L->J;
  if (nc < 3) return "too many jobs for the grid";
  if ((long long)nc > tiles_all) nc = tiles_all < 3 ? 3 : (int)tiles_all;
  L->nc = nc;
  int n0 = (int)((4LL * nc + 3) / 6);
  if (n0 > nc - 2) n0 = nc - 2;
  if (n0 < 1) n0 = 1;
  const int n1 = (nc - n0) / 2;
  L->ng[0] = n0; L->ng[1] = n1; L->ng[2] = nc - n0 - n1;
  L->ncta = L->J * nc;
  return nullptr;
}

bool aligned16p(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool tc_wgrad_supported(const dlwpcs_conv_desc *d, const Geometry &g) {
  static const int force_fp32 = env_int_wg("DLWPCS_WGRAD_FP32", 0);
  if (force_fp32) return false;
  WgPlan L;
  return make_wg_plan(d, g, &L) == nullptr;
}

int64_t tc_wgrad_workspace_bytes(const dlwpcs_conv_desc *d, const Geometry &g) {
  WgPlan L;
  const char *r = make_wg_plan(d, g, &L);
  if (r) {
    set_error("bf16 tensor-core wgrad does not support this configuration: %s", r);
    return -1;
  }
  return (int64_t)L.ncta * ((int64_t)L.nacc * 128 * L.NJ + L.NJ) * 4;
}

int tc_conv_wgrad(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *dy, const void *y,
                  const dlwpcs_conv_wgrads *out, void *workspace, cudaStream_t st) {
  WgP P;
  memset(&P, 0, sizeof(P));
  const char *r = make_wg_plan(d, g, &P.pl);
  CS_CHECK(r == nullptr, "bf16 tensor-core wgrad does not support this configuration: %s", r);
  WgPlan &L = P.pl;
  P.x = (const __nv_bfloat16 *)x0;
  P.dy = (const __nv_bfloat16 *)dy;
  P.mask_y = d->act != DLWPCS_ACT_NONE ? (const __nv_bfloat16 *)y : nullptr;
  P.act = d->act; P.slope = d->act_slope; P.maxv = d->act_max;
  P.tabx = get_patch_table(g, L.Wv, L.G, d->n, d->halo, DLWPCS_SRC_SAME);
  if (!P.tabx) return 3;
  P.ws = (float *)workspace;
  P.ws_b = P.ws + (size_t)L.ncta * L.nacc * 128 * L.NJ;
  P.batch = d->batch; P.ppbx = 6 * d->n * d->n; P.ppfy = g.Hout * g.Wout;
  P.cin = d->cin; P.cout = d->cout; P.Ho = g.Hout; P.Wo = g.Wout;
  L.vecx = (d->cin % 8 == 0) && aligned16p(x0);
  L.vecy = (d->cout % 8 == 0) && aligned16p(dy) && (!P.mask_y || aligned16p(y));
  static const int knock = env_int_wg("DLWPCS_WG_KNOCK", 0);
  P.knock = knock;
  if (env_int_wg("DLWPCS_WG_VERBOSE", 0))
    fprintf(stderr, "[wgrad plan] cin=%d cout=%d n=%d CinBlk=%d NCB=%d MBu=%d NBlk=%d NBJ=%d NNG=%d TPB=%d tpf=%d PS=%d NPX=%d smem=%d ncta=%d ng=%d/%d/%d tmem=%d\n",
            d->cin, d->cout, d->n, L.CinBlk, L.NCB, L.MBu, L.NBlk, L.NBJ, L.NNG, L.TPB, L.tpf, L.PS, L.NPX, L.smemBytes, L.ncta, L.ng[0], L.ng[1], L.ng[2], L.tmemCols);
  typedef void (*kern_t)(const WgP);
  static const kern_t kerns[5] = {wgrad_tc_kernel<0, 0, 0>, wgrad_tc_kernel<3, 1, 1>, wgrad_tc_kernel<3, 2, 1>,
                                  wgrad_tc_kernel<3, 1, 2>, wgrad_tc_kernel<1, 1, 1>};
  static bool attr_set[kMaxDevices][5] = {};      // the shared-memory opt-in is per device
  const int dev_i = current_device_index();
  int ki = 0;
  if (L.kh == 3 && L.MBu == 1 && L.NBJ == 1) ki = 1;
  else if (L.kh == 3 && L.MBu == 2 && L.NBJ == 1) ki = 2;
  else if (L.kh == 3 && L.MBu == 1 && L.NBJ == 2) ki = 3;
  else if (L.kh == 1 && L.MBu == 1 && L.NBJ == 1) ki = 4;
  if (!attr_set[dev_i][ki]) {
    CS_CUDA(cudaFuncSetAttribute(kerns[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_CAP));
    attr_set[dev_i][ki] = true;
  }
  kerns[ki]<<<L.ncta, WG_THREADS, L.smemBytes, st>>>(P);
  CS_CUDA(cudaGetLastError());
  CS_CUDA(cudaGetLastError());
  const long long total = (long long)g.taps * d->cin * d->cout + d->cout;
  // four outputs per thread when every group of four consecutive outputs stays inside one column block and one row of
  // the partials (cout, the column block and the bias offset `per` are multiples of 4) and the partials are 16-byte aligned
  const bool vec4 = d->cout % 4 == 0 && L.NJ % 4 == 0 && (reinterpret_cast<uintptr_t>(P.ws) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(P.ws_b) & 15) == 0;
  if (vec4)
    wgrad_tc_reduce_kernel<4><<<(unsigned)((total + 127) / 128), 256, 0, st>>>(
        P.ws, P.ws_b, out->dw_eq, out->dw_pol, d->independent_north_pole ? out->dw_np : nullptr,
        d->use_bias ? out->db_eq : nullptr, d->use_bias ? out->db_pol : nullptr,
        (d->use_bias && d->independent_north_pole) ? out->db_np : nullptr, d->kh, d->kw, d->cin, d->cout,
        d->flip_north_pole, L);
  else
    wgrad_tc_reduce_kernel<1><<<(unsigned)((total + 31) / 32), 256, 0, st>>>(
        P.ws, P.ws_b, out->dw_eq, out->dw_pol, d->independent_north_pole ? out->dw_np : nullptr,
        d->use_bias ? out->db_eq : nullptr, d->use_bias ? out->db_pol : nullptr,
        (d->use_bias && d->independent_north_pole) ? out->db_np : nullptr, d->kh, d->kw, d->cin, d->cout,
        d->flip_north_pole, L);
  CS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dlwpcs
