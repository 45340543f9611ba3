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

constexpr int WG_THREADS = 288;
constexpr int WG_LOADERS = 256;
constexpr int WG_MMA_WARP = 8;
constexpr int WG_MAX_UNITS = 32;
constexpr int WG_MAX_PS = 4;
constexpr int WG_SMEM_CAP = 227 * 1024;
constexpr int WG_MISC = 128 + WG_MAX_UNITS * 16 + WG_LOADERS * 8 * 4;     // barriers, unit table, bias partials

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
  const int32_t *ytab;        // [tpf*TP] linear position -> output pixel r*Wo + c of one face or -1
  float *ws, *ws_b;           // [ncta][nacc][128][NJ], [ncta][NJ]
  int batch, ppbx, ppfy;      // x pixels per batch element, dy pixels per face
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

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgP P) {
  extern __shared__ uint8_t smem_raw[];
  const WgPlan &L = P.pl;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage0 = base;
  const uint32_t misc = stage0 + (uint32_t)L.PS * L.stageBytes;
  const uint32_t bar_full = misc, bar_empty = misc + 32, bar_done = misc + 64, tmem_slot = misc + 72;
  float *s_bsum = reinterpret_cast<float *>(gen + (misc - base) + 128 + WG_MAX_UNITS * 16);

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

  if (tid == 0) {
    for (int i = 0; i < L.PS; ++i) {
      mbar_init(bar_full + 8 * i, WG_LOADERS);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, (uint32_t)L.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(gen + (tmem_slot - base));

  if (warp == WG_MMA_WARP) {
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
#pragma unroll 1
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t a_ks = a_stage + (uint32_t)ks * a_kstep, b_ks = b_stage + (uint32_t)ks * b_kstep;
        const uint32_t acc = (k > 0 || ks > 0) ? 1u : 0u;
        uint32_t a_u = a_ks, dcol = tmem_base;
#pragma unroll 1
        for (int u = 0; u < L.kh; ++u, a_u += a_ustep) {
          uint32_t a_m = a_u;
#pragma unroll 1
          for (int mbu = 0; mbu < L.MBu; ++mbu, a_m += a_mstep) {
            uint32_t b_n = b_ks;
#pragma unroll 1
            for (int nbj = 0; nbj < L.NBJ; ++nbj, b_n += b_nstep, dcol += (uint32_t)L.NBlk)
              umma_bf16_elect(dcol, a_hi | (uint64_t)a_m, b_hi | (uint64_t)b_n, idesc, acc);
          }
        }
      }
      umma_commit_elect(bar_empty + 8 * s);        // the stage may be refilled once these MMAs have read it
      if (k == my_tiles - 1) umma_commit_elect(bar_done);
      if (++s == L.PS) { s = 0; ph ^= 1; }
    }
  } else {
    // ===== loaders =====
    const int lt = tid;
    const int Sx = L.CinBlk / 8, Sy = L.NJ / 8;
    const int chx = lt & (Sx - 1), chy = lt & (Sy - 1);
    const int cx = cb * L.CinBlk + chx * 8;                       // first input channel of this thread's x chunk
    const int cy = ngi * L.NJ + chy * 8;                          // first output channel of this thread's dy chunk
    const bool okx = cx < P.cin, oky = cy < P.cout;
    const int pstx = WG_LOADERS >> L.logSx, psty = WG_LOADERS >> L.logSy;
    const uint32_t swx = (uint32_t)(L.RBx / 16 - 1), swy = (uint32_t)(L.RBy / 16 - 1);
    const uint32_t yblk = (uint32_t)(chy / (L.RBy / 16)), ycw = (uint32_t)(chy % (L.RBy / 16));
    float bsum[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bsum[e] = 0.f;
    int s = 0, ph = 0;
    for (int k = 0; k < my_tiles; ++k) {
      const int t = idx + k * nin;
      int b, f, tf;
      if (grp == 0) { const int bf = t / L.tpf; tf = t - bf * L.tpf; b = bf >> 2; f = bf & 3; }
      else { b = t / L.tpf; tf = t - b * L.tpf; f = 3 + grp; }
      const int q0 = tf * L.TP;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      const uint32_t st_addr = stage0 + (uint32_t)s * L.stageBytes;
      // -- x patch: rows q0 .. q0 + NPX of the virtual face through the halo table
      const int32_t *tab = P.tabx + (size_t)f * L.G + q0;
      const size_t xb = (size_t)b * P.ppbx;
      if (L.vecx) {
        for (int i0 = lt >> L.logSx; i0 < L.NPX; i0 += 8 * pstx) {
          int px[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = i0 + e * pstx;
            px[e] = (i < L.NPX && okx) ? __ldg(tab + i) : -1;
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
          const int px = okx ? __ldg(tab + i) : -1;
          float a[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] = 0.f;
          if (px >= 0) load8(P.x, xb + px, P.cin, cx, false, a);
          const uint32_t row = st_addr + (uint32_t)i * L.RBx;
          st_shared16(row + ((((uint32_t)chx) ^ ((row >> 7) & swx)) << 4),
                      make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7])));
        }
      }
      // -- dy tile through registers: activation-derivative mask, bias partial sums, bf16 rounding of the product
      const int32_t *yt = P.ytab + q0;
      const size_t yb = (size_t)(b * 6 + f) * P.ppfy;
      const uint32_t ybase = st_addr + (uint32_t)L.xBytes + yblk * (uint32_t)L.yBlockBytes;
      for (int i0 = lt >> L.logSy; i0 < L.TP; i0 += 4 * psty) {
        int px[4];
        float v[4][8], m[4][8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = i0 + e * psty;
          px[e] = (i < L.TP && oky) ? __ldg(yt + i) : -1;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (px[e] >= 0) {
            load8(P.dy, yb + px[e], P.cout, cy, L.vecy, v[e]);
            if (P.mask_y) load8(P.mask_y, yb + px[e], P.cout, cy, L.vecy, m[e]);
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = i0 + e * psty;
          if (i < L.TP) {
            uint4 o = make_uint4(0, 0, 0, 0);
            if (px[e] >= 0) {
              if (P.mask_y) {
#pragma unroll
                for (int c = 0; c < 8; ++c) v[e][c] *= act_grad_from_y(m[e][c], P.act, P.slope, P.maxv);
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) bsum[c] += v[e][c];
              o = make_uint4(pack_bf16x2(v[e][0], v[e][1]), pack_bf16x2(v[e][2], v[e][3]), pack_bf16x2(v[e][4], v[e][5]),
                             pack_bf16x2(v[e][6], v[e][7]));
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
    // ---- bias partials of this CTA (only the cin-block-0 jobs report them): fixed-order sum over the loader threads
    if (cb == 0 && P.ws_b) {
#pragma unroll
      for (int e = 0; e < 8; ++e) s_bsum[lt * 8 + e] = bsum[e];
      named_bar_sync(1, WG_LOADERS);
      if (lt < L.NJ) {
        const int ch = lt >> 3, e = lt & 7;
        float acc = 0.f;
        for (int th = ch; th < WG_LOADERS; th += Sy) acc += s_bsum[th * 8 + e];
        P.ws_b[(size_t)blockIdx.x * L.NJ + lt] = acc;
      }
    }
    // ---- accumulator dump: TMEM -> workspace [acc][lane][NJ]
    if (my_tiles > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
    }
    const int quarter = warp & 3, half = warp >> 2;
    float *wsc = P.ws + (size_t)blockIdx.x * L.nacc * 128 * L.NJ;
    const int nchunks = L.nacc * L.NJ / 16;
    for (int cc = half; cc < nchunks; cc += 2) {
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

// second pass: fixed-order sum of the CTA partials of each (job, face group); un-flip / merge the north-pole share
__global__ void wgrad_tc_reduce_kernel(const float *__restrict__ ws, const float *__restrict__ ws_b, float *dw_eq,
                                       float *dw_pol, float *dw_np, float *db_eq, float *db_pol, float *db_np, int kh,
                                       int kw, int cin, int cout, int flip, const WgPlan L) {
  const long long per = (long long)kh * kw * cin * cout;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const size_t cta_stride = (size_t)L.nacc * 128 * L.NJ;
  if (i < per) {
    const int co = (int)(i % cout);
    long long r = i / cout;
    const int ci = (int)(r % cin); r /= cin;
    const int v = (int)(r % kw), u = (int)(r / kw);
    const int cbk = ci / L.CinBlk, cil = ci - cbk * L.CinBlk;
    const int ngi = co / L.NJ, col = co - ngi * L.NJ;
    const int job = ngi * L.NCB + cbk;
    const int mbu = v / L.SPB, js = v - mbu * L.SPB, lane = js * L.CinBlk + cil;
    auto group_sum = [&](int g, int uu) {
      const int first = job * L.nc + (g == 0 ? 0 : (g == 1 ? L.ng[0] : L.ng[0] + L.ng[1]));
      const size_t off = ((size_t)(uu * L.MBu + mbu) * 128 + lane) * L.NJ + col;
      float s = 0.f;
      for (int c = 0; c < L.ng[g]; ++c) s += ws[(size_t)(first + c) * cta_stride + off];
      return s;
    };
    dw_eq[i] = group_sum(0, u);
    const float south = group_sum(1, u);
    const float north = group_sum(2, flip ? kh - 1 - u : u);       // packed kernel row that reads source row u
    if (dw_np) {
      dw_pol[i] = south;
      dw_np[i] = north;
    } else {
      dw_pol[i] = south + north;
    }
  } else if (db_eq && i < per + cout) {
    const int co = (int)(i - per);
    const int ngi = co / L.NJ, col = co - ngi * L.NJ;
    const int job = ngi * L.NCB;
    auto group_sum = [&](int g) {
      const int first = job * L.nc + (g == 0 ? 0 : (g == 1 ? L.ng[0] : L.ng[0] + L.ng[1]));
      float s = 0.f;
      for (int c = 0; c < L.ng[g]; ++c) s += ws_b[(size_t)(first + c) * L.NJ + col];
      return s;
    };
    db_eq[co] = group_sum(0);
    const float south = group_sum(1), north = group_sum(2);
    if (db_np) {
      db_pol[co] = south;
      db_np[co] = north;
    } else {
      db_pol[co] = south + north;
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
    if (2 * stage + WG_MISC + 1024 <= WG_SMEM_CAP) best = tpb;
  }
  if (!best) return "tile does not fit shared memory";
  if (best > nmb) best = nmb;
  const int tiles = (nmb + best - 1) / best;
  best = (nmb + tiles - 1) / tiles;                 // even out the tiles of a face
  L->TPB = best;
  L->TP = best * 128;
  L->tpf = (nmb + best - 1) / best;
  L->NPX = (L->TP + extX + 7) / 8 * 8;
  L->G = ((L->tpf - 1) * L->TP + L->NPX + 3) / 4 * 4;
  L->xBytes = (L->NPX * L->RBx + 1023) / 1024 * 1024;
  L->yBlockBytes = L->TP * L->RBy;
  L->stageBytes = L->xBytes + L->NBJ * L->yBlockBytes;
  L->PS = (WG_SMEM_CAP - WG_MISC - 1024) / L->stageBytes;
  int ps_max = env_int_wg("DLWPCS_WG_PS", 3);
  if (ps_max > WG_MAX_PS) ps_max = WG_MAX_PS;
  if (L->PS > ps_max) L->PS = ps_max;
  if (L->PS < 1) return "tile does not fit shared memory";
  L->smemBytes = 1024 + L->PS * L->stageBytes + WG_MISC;
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

// linear position of the virtual face -> output pixel of one face (or -1: discarded column / beyond the face)
struct YKey {
  int dev, Ho, Wo, Wv, len;
  bool operator<(const YKey &o) const { return memcmp(this, &o, sizeof(YKey)) < 0; }
};
std::mutex g_ytab_mu;
std::map<YKey, int32_t *> g_ytabs;

const int32_t *get_ytab(int Ho, int Wo, int Wv, int len) {
  YKey key;
  memset(&key, 0, sizeof(key));
  if (cudaGetDevice(&key.dev) != cudaSuccess) {
    set_error("cudaGetDevice failed");
    return nullptr;
  }
  key.Ho = Ho; key.Wo = Wo; key.Wv = Wv; key.len = len;
  std::lock_guard<std::mutex> lk(g_ytab_mu);
  auto it = g_ytabs.find(key);
  if (it != g_ytabs.end()) return it->second;
  std::vector<int32_t> tab((size_t)len);
  for (int q = 0; q < len; ++q) {
    const int r = q / Wv, c = q % Wv;
    tab[q] = (r < Ho && c < Wo) ? r * Wo + c : -1;
  }
  int32_t *dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, tab.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemcpy(dev, tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("wgrad table upload failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  g_ytabs[key] = dev;
  return dev;
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
  P.ytab = get_ytab(g.Hout, g.Wout, L.Wv, L.tpf * L.TP);
  if (!P.ytab) return 3;
  P.ws = (float *)workspace;
  P.ws_b = P.ws + (size_t)L.ncta * L.nacc * 128 * L.NJ;
  P.batch = d->batch; P.ppbx = 6 * d->n * d->n; P.ppfy = g.Hout * g.Wout;
  P.cin = d->cin; P.cout = d->cout;
  L.vecx = (d->cin % 8 == 0) && aligned16p(x0);
  L.vecy = (d->cout % 8 == 0) && aligned16p(dy) && (!P.mask_y || aligned16p(y));
  static bool attr_set = false;
  if (!attr_set) {
    CS_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_CAP));
    attr_set = true;
  }
  wgrad_tc_kernel<<<L.ncta, WG_THREADS, L.smemBytes, st>>>(P);
  CS_CUDA(cudaGetLastError());
  const long long total = (long long)g.taps * d->cin * d->cout + d->cout;
  wgrad_tc_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      P.ws, P.ws_b, out->dw_eq, out->dw_pol, d->independent_north_pole ? out->dw_np : nullptr,
      d->use_bias ? out->db_eq : nullptr, d->use_bias ? out->db_pol : nullptr,
      (d->use_bias && d->independent_north_pole) ? out->db_np : nullptr, d->kh, d->kw, d->cin, d->cout,
      d->flip_north_pole, L);
  CS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dlwpcs
