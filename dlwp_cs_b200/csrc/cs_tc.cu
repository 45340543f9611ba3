// tcgen05 (sm_100a) bf16 implicit-GEMM kernel of the cubed-sphere convolution: persistent, warp-specialised, pipelined.
//
// Reference semantics: DLWP/custom.py:921-1002 (CubeSphereConv2D.call) with the preceding CubeSpherePadding2D
// (custom.py:1198-1308) and the U-Net's pool / upsample / concatenate (Azure/train_cs.py:197-199, 282-299) folded
// into the load stage, bias + capped leaky ReLU in the epilogue.
//
// GEMM view of one face:  D[q, o] = sum_{tap,c} A[q + toff(tap), c] * W[tap, c, o]
//   q     linear position r*Wv + c over the *virtual* (zero/halo padded) face of width Wv; positions with c >= Wout
//         are computed and discarded (Wout/Wv = 48/50 useful at C48), which makes every tap a pure row offset of the
//         same shared-memory patch: the A operand descriptor of tap (u,v) is the patch base + (u*dh*Wv + v*dw) rows.
//   A     the patch, pixel-major like the tensor in HBM: one row of RB = min(2*Cin, 128) bytes per patch pixel (64
//         channels per K block), 16-byte chunks XOR-swizzled with the row's address bits (UMMA SWIZZLE_32/64/128B,
//         K-major).  The swizzle is a function of the absolute shared-memory address (checked with
//         tools/umma_probe.cu), so a row offset is a plain +RB per pixel and a warp's gathers land in contiguous
//         512-byte runs (3x the cp.async throughput of a channel-slab-major patch, tools/load_probe.cu).
//   W     packed once per layer in exactly the shared-memory image; TMA bulk copies through an mbarrier ring, kept
//         resident across tiles of the same face group when the whole set fits.
//   D     fp32 in tensor memory, double buffered: AS x MB accumulators of 128 lanes x CoutP columns.
// A tile is MB consecutive 128-row blocks of one face; a CTA (one per SM) walks tiles blockIdx.x, +gridDim.x, ...
// Warp roles (576 threads): warp 16 = TMA weight producer, warp 17 = TMEM allocator + MMA issuer, warps 8..15 = epilogue
// (tcgen05.ld -> bias / activation -> global; two warps per TMEM lane quarter), warps 0..7 = patch loaders: the halo
// exchange, pooling, upsampling and concatenation are one table lookup per patch row (physical source pixel or -1 =
// zero) followed by 16-byte cp.async gathers; they run up to PS-1 tiles ahead of the MMA warp.
// tcgen05.mma in SS mode is shared-memory-bandwidth paced (measured, tools/umma_probe.cu: M128 K16 costs
// max(N/2, (4 KB + N*32 B) / 128 B) cycles: 41.7 / 48.8 / 65 / 130 for N = 32 / 64 / 128 / 256), so the issue loop is
// kept free of divisions and table-driven: one LDS per (chunk, tap) unit, then MB x KC/16 back-to-back MMAs.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <vector>
#include "cs_common.cuh"
#include "cs_ptx.cuh"
#include "cs_pack.cuh"

namespace dlwpcs {

namespace {

// warps 0-7 loaders, 8-15 epilogue, 16 TMA weight producer, 17-18 MMA issuers (alternate m-blocks of a tile: one thread
// sustains ~58 cycles per tcgen05.mma, the tensor core ~42 at N = 32, so two issuers keep it fed).  The scheduler favours
// the highest warp id of a sub-partition (B300_MICROARCH.md), so the latency-critical issuers get the last warps.
constexpr int TC_THREADS = 608;
constexpr int NUM_MMA_WARPS = 2;
constexpr int TC_LOADERS = 256;
constexpr int TC_EPI = 256;
constexpr int LOAD_WARP0 = 0, EPI_WARP0 = 8, TMA_WARP = 16, MMA_WARP = 17;   // MMA warps: MMA_WARP .. +NUM_MMA_WARPS-1
constexpr int MAX_UNITS = 256;     // (K chunk, tap) pairs of one layer
constexpr int MAX_STAGES = 8;     // weight ring
constexpr int MAX_PS = 4;         // patch stages
constexpr int SMEM_CAP = 227 * 1024;

struct TcPlan {
  int CinP, CoutP, KC, nch, S;       // padded channels, channels per K block / weight chunk, K blocks, 16-byte chunks per pixel
  int taps, NU, UPS, NST, nstages;   // weight units (chunk,tap), units per ring stage, ring depth, stage loads per set
  int unitBytes, stageBytes, resident;
  int Wv, Hv, Q, nmb, haloExt;       // virtual face, linear outputs per face, 128-row blocks per face, patch overhang
  int MB, tpf, NPIXp, RB, blockBytes;  // m-blocks per tile, tiles per face, patch rows, bytes per row, bytes per K block
  int cprLog, swzMask, layoutType;   // log2(chunks per row), swizzle mask over address bits 7.., UMMA layout code
  int patchBytes, PS, AS, tmemCols;  // bytes per patch stage, patch stages, accumulator stages
  int G;                             // patch-table entries per face (multiple of 4)
  int TS, tabBytes, twoTabs;         // table ring slots, bytes per slot (one or two tables of NPIXp entries)
  int stgBytes, stgBufs;             // output staging per epilogue warp (32 rows x cout bf16; 0: direct stores)
  int smemBytes;
  int64_t groupBytes;                // packed weights per face group
  int vec, logS;                     // 16-byte gather path usable; log2(S) or -1
};

struct TcP {
  const __nv_bfloat16 *x0, *x1;
  const int32_t *tab0, *tab1;   // [6][G] physical source pixel within one batch element, -1 = zero
  const uint8_t *wpack;
  const float *bias;            // [3][CoutP] fp32 (zeros when the layer has no bias)
  void *y;
  const void *mask_y;           // dgrad: forward output whose activation derivative scales the gathered dy
  int y_f32, mask_f32;
  int batch, n, Hout, Wout;
  int cin, cout, c0, c1, mode0, mode1;
  int ppb0, ppb1;               // pixels per batch element of each source tensor
  int kw, dh, dw;
  int act;
  float slope, maxv;
  int mask_act;
  float mask_slope, mask_max;
  long long *dbg;               // optional phase timestamps (DLWPCS_TC_TIMING=1)
  unsigned long long *trace;    // optional per-CTA %globaltimer stamps of this launch (DLWPCS_TC_TRACE=1), [grid][8]
  // chained launch (dlwpcs_conv2d_fwd_chained): instead of waiting for the whole previous grid, a tile waits until the
  // launch that produced its input has completed that SAMPLE (per-sample completion counters), and tiles are handed out
  // by an atomic counter so that CTAs that start early (on SMs the previous layer has already left) take more of them
  const unsigned *dep;          // [batch] completion counters of the producing launch; nullptr = stream order (classic)
  unsigned dep_target;          // dep[b] >= dep_target  <=>  sample b of the input is complete
  unsigned *done;               // [batch] counters this launch raises (once per completed tile); may be nullptr
  unsigned *tile_ctr;           // dynamic tile scheduler; nullptr = static round-robin over the grid
  unsigned *chain_err;          // set to 1 if a dependency wait timed out (diagnostics; never hang the GPU)
  int knock;                    // bottleneck analysis (DLWPCS_TC_KNOCK): 1 no gathers, 2 no MMAs, 4 no epilogue math/stores, 8 no global stores
  TcPlan pl_;
};

// debug timeline (DLWPCS_TC_TIMING=1): timestamps of the CTA in the middle of the grid at its DBG_K-th tile
constexpr int DBG_K = 6;
__device__ __forceinline__ void stamp(const TcP &P, int slot, int k, bool who) {
  if (P.dbg && who && blockIdx.x == gridDim.x / 2 && (k == DBG_K || (slot == 11 && k == DBG_K + 1) || slot == 0))
    P.dbg[slot] = clock64();
}

// cross-kernel timeline (DLWPCS_TC_TRACE=1): nanosecond stamps of every CTA of every launch, read back with
// dlwpcs_trace_read -- slot 0 kernel entry, 1 prologue done, 2 loaders past the grid dependency, 3 first patch landed,
// 4 first accumulators complete, 5 last epilogue done, 6 kernel exit, 7 tiles walked by the CTA
constexpr int TRACE_LAUNCHES = 256, TRACE_CTAS = 160;
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace(const TcP &P, int slot, bool who) {
  if (P.trace && who) P.trace[blockIdx.x * 8 + slot] = gtime_ns();
}

// tile id -> (batch, face, tile in face).  Tiles are enumerated face-group-major (equatorial faces of every sample, then
// the south-pole faces, then the north-pole faces) so that a CTA changes weight set at most twice.
struct TileInfo {
  int b, f, grp, tf;
};
__device__ __forceinline__ TileInfo decode_tile(int id, int batch, int tpf) {
  TileInfo t;
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

// k-th tile of this CTA.  Static schedule: blockIdx.x + k * gridDim.x.  Dynamic schedule (chained launches): lane 1 of the
// producer warp draws tile ids from a global counter and publishes them in a 16-entry shared-memory ring tagged with
// k + 1; every role polls its entry (the producer runs at most TS + PS + AS = 7 tiles ahead of the slowest role).
constexpr int TILE_RING = 16;
// The poll lives inside one opaque asm statement (like mbar_wait): a spin loop written in C++ makes the compiler treat
// everything after it as divergent, and the MMA issuers lose their uniform-register descriptors (2x the R2UR count).
__device__ __forceinline__ int tile_poll(uint32_t ring, int k) {
  uint32_t lo;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 hi;\n\t"
      "POLL_TILE:\n\t"
      "ld.volatile.shared.v2.u32 {%0, hi}, [%1];\n\t"
      "setp.ne.u32 p, hi, %2;\n\t"
      "@p bra POLL_TILE;\n\t"
      "}"
      : "=r"(lo)
      : "r"(ring + 8u * (uint32_t)(k & (TILE_RING - 1))), "r"((uint32_t)(k + 1))
      : "memory");
  return (int)lo;
}
__device__ __forceinline__ void tile_publish(uint32_t ring, int k, int tile) {
  asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(ring + 8u * (uint32_t)(k & (TILE_RING - 1))),
               "r"((uint32_t)tile), "r"((uint32_t)(k + 1))
               : "memory");
}
// Loop header over the tiles of a CTA: the static variant is literally the round-robin for-loop of the classic kernel
// (the compiler's code for it must not change), the dynamic one polls the ring.  K is the running tile index.
#define TC_TILE_LOOP(K)                                                                                    \
  for (int tile = DYN ? tile_poll(tile_ring, 0) : (int)blockIdx.x; DYN ? tile >= 0 : tile < ntiles;        \
       tile = DYN ? tile_poll(tile_ring, K + 1) : tile + (int)gridDim.x, ++K)
#define TC_TILE_LOOP_UNIFORM(K)                                                                                   \
  for (int tile = DYN ? __shfl_sync(0xffffffffu, tile_poll(tile_ring, 0), 0) : (int)blockIdx.x;                   \
       DYN ? tile >= 0 : tile < ntiles;                                                                           \
       tile = DYN ? __shfl_sync(0xffffffffu, tile_poll(tile_ring, K + 1), 0) : tile + (int)gridDim.x, ++K)

// One 8-channel slab of one patch row through registers: 2x2 mean, activation-derivative mask, or plain copy.
__device__ __noinline__ uint4 gather_regs(const TcP &P, const __nv_bfloat16 *src, int C, int mode, size_t pix, int cc,
                                             int w2) {
  float a[8];
  if (mode == DLWPCS_SRC_POOL2) {
    float t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (pix + dy * w2 + dx) * C + cc));
        unpack_bf16x8(v, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += t[k];
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= 0.25f;
  } else {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + pix * C + cc));
    unpack_bf16x8(v, a);
  }
  if (P.mask_y) {
    float m[8];
    if (P.mask_f32) {
      const float4 *mp = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(P.mask_y) + pix * C + cc);
      const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
      m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w; m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
    } else {
      unpack_bf16x8(
          __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(P.mask_y) + pix * C + cc)), m);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= act_grad_from_y(m[k], P.mask_act, P.mask_slope, P.mask_max);
  }
  return make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
}

// Generic gather (any chunk count; channel counts that are not multiples of 8 go element by element).  Kept out of
// line: only odd layer shapes come here.
__device__ __noinline__ void generic_gather(const TcP &P, uint32_t stage, int npix, const int32_t *t0, const int32_t *t1,
                                            size_t b0, size_t b1, int lt) {
  const TcPlan &L = P.pl_;
  const int w2_0 = P.n * 2;
  const int items = npix * L.S;
#pragma unroll 1
  for (int it = lt; it < items; it += TC_LOADERS) {
    const int i = it / L.S, chunk = it - i * L.S, c = chunk * 8;
    const uint32_t row = stage + (uint32_t)(chunk >> L.cprLog) * L.blockBytes + (uint32_t)i * L.RB;
    const uint32_t dst = row + ((((uint32_t)chunk & ((1u << L.cprLog) - 1u)) ^ ((row >> 7) & (uint32_t)L.swzMask)) << 4);
    const int p0 = t0[i], p1 = P.c1 > 0 ? t1[i] : -1;
    if (L.vec) {
      const bool first = c < P.c0;
      const __nv_bfloat16 *src = first ? P.x0 : P.x1;
      const int C = first ? P.c0 : P.c1, cc = first ? c : c - P.c0, mode = first ? P.mode0 : P.mode1;
      const int px = first ? p0 : p1;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (px >= 0 && c < P.cin) o = gather_regs(P, src, C, mode, (first ? b0 : b1) + px, cc, w2_0);
      st_shared16(dst, o);
    } else {
      float a[8];
#pragma unroll 1
      for (int e = 0; e < 8; ++e) {
        const int ch = c + e;
        float val = 0.f;
        if (ch < P.cin) {
          const bool first = ch < P.c0;
          const __nv_bfloat16 *src = first ? P.x0 : P.x1;
          const int C = first ? P.c0 : P.c1, cc = first ? ch : ch - P.c0, mode = first ? P.mode0 : P.mode1;
          const int px = first ? p0 : p1;
          if (px >= 0) {
            const size_t pix = (first ? b0 : b1) + px;
            if (mode == DLWPCS_SRC_POOL2) {
              val = 0.25f * (__bfloat162float(src[pix * C + cc]) + __bfloat162float(src[(pix + 1) * C + cc]) +
                             __bfloat162float(src[(pix + w2_0) * C + cc]) + __bfloat162float(src[(pix + w2_0 + 1) * C + cc]));
            } else {
              val = __bfloat162float(src[pix * C + cc]);
            }
            if (P.mask_y) {
              const float m = P.mask_f32 ? reinterpret_cast<const float *>(P.mask_y)[pix * C + cc]
                                         : __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(P.mask_y)[pix * C + cc]);
              val *= act_grad_from_y(m, P.mask_act, P.mask_slope, P.mask_max);
            }
          }
        }
        a[e] = val;
      }
      st_shared16(dst, make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]),
                                  pack_bf16x2(a[6], a[7])));
    }
  }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
template <int MBT, int KC16T, bool DYN>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcP P) {
  extern __shared__ uint8_t smem_raw[];
  const TcPlan &L = P.pl_;
  trace(P, 0, threadIdx.x == 0);
  // carve (1024-byte aligned): [patch stages][weight ring][barriers + tmem slot: 512 B][unit table 2 KB][bias 3*CoutP]
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t patch0 = base;
  const uint32_t wring = patch0 + (uint32_t)L.PS * L.patchBytes;
  const uint32_t misc = wring + (uint32_t)L.NST * L.stageBytes;
  const uint32_t bar_wfull = misc, bar_wempty = misc + 64, bar_pfull = misc + 128, bar_pempty = misc + 160,
                 bar_afull = misc + 192, bar_aempty = misc + 208, tmem_slot = misc + 224;
  float *s_bias = reinterpret_cast<float *>(gen + (misc - base) + 512 + MAX_UNITS * 8);
  const uint32_t tabring = misc + 512 + MAX_UNITS * 8 + (uint32_t)(3 * L.CoutP * 4);          // TS slots of tabBytes
  const int32_t *s_tab = reinterpret_cast<const int32_t *>(gen + (tabring - base));
  const uint32_t bar_tfull = misc + 256, bar_tempty = misc + 288;
  const uint32_t tile_ring = misc + 320;                // 16 x {tile id, tag}
  const uint32_t epi_ctr = misc + 448;                  // epilogue warps that have finished their share of a tile (monotonic)
  const bool chained = DYN && P.dep != nullptr;
  const uint32_t stg0 = (tabring + (uint32_t)(L.TS * L.tabBytes) + 127u) & ~127u;      // 8 warps x stgBufs x stgBytes

  // warp index broadcast from lane 0 (cutlass::canonical_warp_idx_sync): the role branches become provably warp-uniform,
  // which lets the compiler keep the MMA descriptors in uniform registers
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int ntiles = 6 * P.batch * L.tpf;

  if (tid == 0) {
    for (int i = 0; i < L.NST; ++i) {
      mbar_init(bar_wfull + 8 * i, 1);
      mbar_init(bar_wempty + 8 * i, NUM_MMA_WARPS);
    }
    for (int i = 0; i < L.PS; ++i) {
      mbar_init(bar_pfull + 8 * i, TC_LOADERS);
      mbar_init(bar_pempty + 8 * i, NUM_MMA_WARPS);
    }
    for (int i = 0; i < L.AS; ++i) {
      mbar_init(bar_afull + 8 * i, NUM_MMA_WARPS);
      mbar_init(bar_aempty + 8 * i, TC_EPI);
    }
    for (int i = 0; i < L.TS; ++i) {
      mbar_init(bar_tfull + 8 * i, chained ? 2 : 1);     // chained: the table copy and the resolved dependency
      mbar_init(bar_tempty + 8 * i, TC_LOADERS);
    }
    for (int i = 0; i < TILE_RING; ++i) *reinterpret_cast<volatile unsigned long long *>(gen + (tile_ring - base) + 8 * i) = 0ull;
    *reinterpret_cast<volatile uint32_t *>(gen + (epi_ctr - base)) = 0u;
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, (uint32_t)L.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(gen + (tmem_slot - base));
  stamp(P, 0, 0, tid == 0);
  trace(P, 1, tid == 0);
  // programmatic dependent launch: the next kernel of the stream may start its prologue (barriers, tensor-memory
  // allocation, patch-table TMA) on every SM this grid has left.  Whatever may have been written by the previous kernel
  // -- activations (loaders), packed weights (TMA lane 0), bias (epilogue) -- is read behind griddepcontrol.wait; the
  // patch tables are library-owned and static; every global store is data-dependent on such a read
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == TMA_WARP) {
    // ===== producers (one thread each): lane 0 streams weights through the ring with TMA bulk copies, skipped while the
    // resident set stays valid; lane 1 prefetches each tile's slice of the patch table(s) into shared memory, so that
    // the loaders' lookups do not queue behind their own DRAM gathers in the L1 pipeline =====
    if (lane == 0) {
      int st = 0, ph = 0, prev_grp = -1;
      // the packed weights may come from the previous kernel (a chained launch's weights are older than its whole chain)
      if (!chained) asm volatile("griddepcontrol.wait;" ::: "memory");
      int k = 0;
      TC_TILE_LOOP(k) {
        const int grp = decode_tile(tile, P.batch, L.tpf).grp;
        if (L.resident && grp == prev_grp) continue;
        prev_grp = grp;
        const uint8_t *wg = P.wpack + (size_t)grp * L.groupBytes;
        for (int s = 0; s < L.nstages; ++s) {
          mbar_wait(bar_wempty + 8 * st, ph ^ 1);
          const int units = min(L.UPS, L.NU - s * L.UPS);
          const uint32_t bytes = (uint32_t)units * L.unitBytes;
          mbar_expect_tx(bar_wfull + 8 * st, bytes);
          tma_bulk_g2s(wring + st * L.stageBytes, wg + (size_t)s * L.stageBytes, bytes, bar_wfull + 8 * st);
          if (++st == L.NST) { st = 0; ph ^= 1; }
        }
      }
    } else if (DYN && lane == 2 && P.done) {
      // completion signaller: tile k is complete when all epilogue warps have counted it; publish it device-wide
      int k = 0;
      TC_TILE_LOOP(k) {
        const unsigned want = (unsigned)(k + 1) * (unsigned)(TC_EPI / 32);
        unsigned v;
        for (;;) {
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(epi_ctr) : "memory");
          if (v >= want) break;
          __nanosleep(256);
        }
        __threadfence();           // cumulative: the epilogue warps' stores observed through the CTA-scope acquire above
        atomicAdd(P.done + decode_tile(tile, P.batch, L.tpf).b, 1u);
      }
    } else if (lane == 1) {
      int ts = 0, tp = 0;
      // dynamic schedule: the draw for tile k+1 is in flight while tile k is processed (the atomic's round trip to L2
      // would otherwise sit in front of every table copy); a CTA over-draws the counter by one at the end -- harmless
      // The first tile of a CTA is its own index (148 simultaneous draws on one address would serialise in the L2 slice
      // for ~2 us at the start of every layer); the counter hands out the tiles from gridDim.x on.
      unsigned drawn = blockIdx.x;
      for (int k = 0;; ++k) {
        int tile;
        if constexpr (DYN) {
          tile = drawn < (unsigned)ntiles ? (int)drawn : -1;
          tile_publish(tile_ring, k, tile);          // to the other roles before anything else
          if (tile >= 0) drawn = gridDim.x + atomicAdd(P.tile_ctr, 1u);
        } else {
          const long long t = (long long)blockIdx.x + (long long)k * gridDim.x;
          tile = t < ntiles ? (int)t : -1;
        }
        if (tile < 0) break;
        const TileInfo T = decode_tile(tile, P.batch, L.tpf);
        const int MBc = min(L.MB, L.nmb - T.tf * L.MB);
        const uint32_t bytes = (uint32_t)((MBc * 128 + L.haloExt + 3) / 4 * 4) * 4u;
        const size_t off = (size_t)T.f * L.G + (size_t)T.tf * L.MB * 128;
        mbar_wait(bar_tempty + 8 * ts, tp ^ 1);
        mbar_expect_tx(bar_tfull + 8 * ts, bytes * (1 + L.twoTabs));
        tma_bulk_g2s(tabring + ts * L.tabBytes, P.tab0 + off, bytes, bar_tfull + 8 * ts);
        if (L.twoTabs) tma_bulk_g2s(tabring + ts * L.tabBytes + L.NPIXp * 4, P.tab1 + off, bytes, bar_tfull + 8 * ts);
        if (++ts == L.TS) { ts = 0; tp ^= 1; }
      }
    } else if (DYN && lane == 3 && chained) {
      // dependency resolver: the table slot of a tile is published (second arrival on its barrier) only when the producing
      // launch has finished sample T.b -- the loaders start behind that barrier, so every gather of the tile is ordered
      // after this acquire load.  The table copy itself (static data) does not wait for it.
      int ts = 0, tp = 0, k = 0, last_b = -1;
      TC_TILE_LOOP(k) {
        const int b = decode_tile(tile, P.batch, L.tpf).b;
        if (b != last_b) {
          const unsigned *c = P.dep + b;
          unsigned v;
          const long long t0 = clock64();
          for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            if (v >= P.dep_target) break;
            if (clock64() - t0 > (1ll << 32)) {        // ~2 s: mis-wired chain -- flag it and go on instead of hanging
              if (P.chain_err) atomicExch(P.chain_err, 1u);
              break;
            }
            __nanosleep(64);
          }
          last_b = b;
        }
        mbar_wait(bar_tempty + 8 * ts, tp ^ 1);      // the slot's previous use is over: this arrival counts for tile k
        mbar_arrive(bar_tfull + 8 * ts);
        if (++ts == L.TS) { ts = 0; tp ^= 1; }
      }
    }
  } else if (warp >= MMA_WARP && warp < MMA_WARP + NUM_MMA_WARPS) {
    // ===== MMA issuers: each warp walks the loop (uniform control flow) for its m-blocks, one elected lane issues =====
    const int mw = warp - MMA_WARP;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(L.CoutP >> 3) << 17) | (8u << 24);
    // A: swizzled K-major rows of RB bytes: SBO = 8 rows, LBO unused (1), layout code in bits 61..63, version 1
    const uint64_t a_fix = ((uint64_t)(((uint32_t)(8 * L.RB) >> 4) | (1u << 14) | ((uint32_t)L.layoutType << 29)) << 32) |
                           (1ull << 16);
    // B: un-swizzled K-major core matrices [k8][n][8]: SBO = 128 B (next 8 output channels), LBO = CoutP*16 B (next k8)
    const uint64_t b_fix = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((uint64_t)((uint32_t)L.CoutP & 0x3FFFu) << 16);
    const uint32_t a_jstep = 32u >> 4, b_jstep = (uint32_t)(2 * L.CoutP * 16) >> 4;     // one K = 16 step
    const uint32_t a_mbstep = (uint32_t)(128 * L.RB) >> 4;
    const uint32_t stage16 = (uint32_t)L.stageBytes >> 4, coutp = (uint32_t)L.CoutP, unit16 = (uint32_t)L.unitBytes >> 4;
    const int kh = L.taps / P.kw;
    const uint32_t blk16 = (uint32_t)L.blockBytes >> 4, urow16 = ((uint32_t)(P.dh * L.Wv) * L.RB) >> 4,
                   vcol16 = ((uint32_t)P.dw * L.RB) >> 4;
    int st = 0, ph = 0, sp = 0, pp = 0, sa = 0, pa = 0, prev_grp = -1;
    int k = 0;
    // dynamic schedule: the tile id is broadcast from lane 0, which keeps everything derived from it provably
    // warp-uniform (uniform registers for the MMA descriptors) although it is read from shared memory
    TC_TILE_LOOP_UNIFORM(k) {
      const TileInfo T = decode_tile(tile, P.batch, L.tpf);
      const int MBc = min(L.MB, L.nmb - T.tf * L.MB);
      const bool load_ev = !(L.resident && T.grp == prev_grp);
      prev_grp = T.grp;
      int next;
      if constexpr (DYN) {
        next = L.resident ? __shfl_sync(0xffffffffu, tile_poll(tile_ring, k + 1), 0) : -1;
      } else {
        next = tile + (int)gridDim.x;
        if (next >= ntiles) next = -1;
      }
      const bool release_w = !L.resident || next < 0 || decode_tile(next, P.batch, L.tpf).grp != T.grp;
      stamp(P, 5, k, lane == 0 && mw == 0);
      mbar_wait(bar_aempty + 8 * sa, pa ^ 1);
      mbar_wait(bar_pfull + 8 * sp, pp);
      tc_fence_after();
      trace(P, 3, k == 0 && lane == 0 && mw == 0);
      stamp(P, 6, k, lane == 0 && mw == 0);
      stamp(P, 11, k, lane == 0 && mw == 0);
      const uint64_t a_stage = a_fix | ((patch0 + (uint32_t)sp * L.patchBytes) >> 4);
      const uint64_t b_ring = b_fix | (wring >> 4);
      const uint32_t d_stage = tmem_base + (uint32_t)(sa * L.MB) * coutp;
      // whole-warp loops over (K chunk, kernel row, kernel column); one elected lane per instruction (umma_*_elect).
      // The loop bodies are unrolled over the window (the uniform-datapath address chains of one iteration are long:
      // tools/umma_mn_rate.cu measures ~110 cycles of loop overhead per un-unrolled iteration next to ~40 per MMA).
#define TC_ISSUE_UNIT(A_OFF, B_UNIT, FIRST)                                                                              \
  do {                                                                                                                   \
    const uint64_t a_unit_ = a_stage + (A_OFF);                                                                          \
    _Pragma("unroll") for (int mi = 0; mi < (MBT + NUM_MMA_WARPS - 1) / NUM_MMA_WARPS; ++mi) {                            \
      const int mb = mw + NUM_MMA_WARPS * mi;                                                                            \
      if (mb < MBc) {                                                                                                    \
        _Pragma("unroll") for (int j = 0; j < KC16T; ++j)                                                                \
            umma_bf16_elect(d_stage + (uint32_t)mb * coutp, a_unit_ + (uint32_t)mb * a_mbstep + (uint32_t)j * a_jstep,   \
                            (B_UNIT) + (uint32_t)j * b_jstep, idesc, (!(FIRST) || j > 0) ? 1u : 0u);                     \
      }                                                                                                                  \
    }                                                                                                                    \
  } while (0)
      const bool win3 = kh == 3 && P.kw == 3;
      if (P.knock & 2) {
        // bottleneck analysis: no MMAs; the weight ring is still cycled below through the generic path's bookkeeping
        if (load_ev || release_w) {
          int uis = 0, sidx = 0;
          for (int unit = 0; unit < L.NU; ++unit) {
            const int stage = load_ev ? st : sidx;
            if (load_ev && uis == 0) { mbar_wait(bar_wfull + 8 * st, ph); tc_fence_after(); }
            if (uis == L.UPS - 1 || unit == L.NU - 1) {
              if (release_w) umma_commit_elect(bar_wempty + 8 * stage);
              if (load_ev && ++st == L.NST) { st = 0; ph ^= 1; }
            }
            if (++uis == L.UPS) { uis = 0; ++sidx; }
          }
        }
      } else if (L.resident && !load_ev && !release_w) {
        // steady state of a resident weight set: units are contiguous in the ring, nothing to wait for or release
        uint64_t b_unit = b_ring;
        uint32_t a_kc = 0;
        if (win3) {
#pragma unroll 1
          for (int kc = 0; kc < L.nch; ++kc, a_kc += blk16) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              TC_ISSUE_UNIT(a_kc + (uint32_t)(t / 3) * urow16 + (uint32_t)(t % 3) * vcol16, b_unit + (uint32_t)t * unit16,
                            kc == 0 && t == 0);
            }
            b_unit += 9u * unit16;
          }
        } else {
          int unit = 0;
#pragma unroll 1
          for (int kc = 0; kc < L.nch; ++kc, a_kc += blk16) {
            uint32_t a_u = a_kc;
#pragma unroll 1
            for (int u = 0; u < kh; ++u, a_u += urow16) {
              uint32_t a_off = a_u;
#pragma unroll 1
              for (int v = 0; v < P.kw; ++v, a_off += vcol16, ++unit, b_unit += unit16) TC_ISSUE_UNIT(a_off, b_unit, unit == 0);
            }
          }
        }
      } else {
        int uis = 0, sidx = 0, unit = 0;
        uint32_t a_kc = 0;
#pragma unroll 1
        for (int kc = 0; kc < L.nch; ++kc, a_kc += blk16) {
          uint32_t a_u = a_kc;
#pragma unroll 1
          for (int u = 0; u < kh; ++u, a_u += urow16) {
            uint32_t a_off = a_u;
#pragma unroll 1
            for (int v = 0; v < P.kw; ++v, a_off += vcol16) {
              const int stage = load_ev ? st : sidx;
              if (load_ev && uis == 0) {
                mbar_wait(bar_wfull + 8 * st, ph);
                tc_fence_after();
              }
              const uint64_t b_unit = b_ring + ((uint32_t)stage * stage16 + (uint32_t)uis * unit16);
              TC_ISSUE_UNIT(a_off, b_unit, unit == 0);
              if (uis == L.UPS - 1 || unit == L.NU - 1) {
                if (release_w) umma_commit_elect(bar_wempty + 8 * stage);
                if (load_ev && ++st == L.NST) { st = 0; ph ^= 1; }
              }
              ++unit;
              if (++uis == L.UPS) { uis = 0; ++sidx; }
            }
          }
        }
      }
#undef TC_ISSUE_UNIT
      umma_commit_elect(bar_pempty + 8 * sp);     // patch stage may be refilled once these MMAs have read it
      umma_commit_elect(bar_afull + 8 * sa);      // accumulators complete
      stamp(P, 7, k, lane == 0 && mw == 0);
      if (++sp == L.PS) { sp = 0; pp ^= 1; }
      if (++sa == L.AS) { sa = 0; pa ^= 1; }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 8) {
    // ===== epilogue: TMEM -> registers -> bias / activation -> global.  Warp w reads TMEM lanes 32*(w%4)..+31; the two
    // warps of a lane quarter take alternate m-blocks. =====
    const int quarter = warp & 3, half = (warp - EPI_WARP0) >> 2;
    const bool fast_out = !P.y_f32 && P.cout % 8 == 0;
    // bf16 rows of 16-byte multiples: every warp compacts its 32 rows of an m-block in shared memory -- they are one
    // contiguous run in HBM once the discarded columns are dropped -- and writes the run back with fully coalesced
    // 512-byte stores (the per-thread 32-byte stores of the direct path cost 8x the L1 sector operations and starve
    // the loaders' cp.async of the same pipeline)
    const bool staged = L.stgBytes > 0;
    const uint32_t stg = stg0 + (uint32_t)(warp - EPI_WARP0) * L.stgBytes;
    const uint32_t rowB = (uint32_t)P.cout * 2u;
    const bool act_fast = P.act == DLWPCS_ACT_CAPPED_LEAKY_RELU && P.slope >= 0.f && P.slope <= 1.f;
    const uint32_t cpr = rowB >> 4;                                             // 16-byte chunks per output row
    const int cprLog = (cpr & (cpr - 1)) == 0 ? 31 - __clz(cpr) : -1;          // rows of 2^k chunks: shifts instead of divisions
    const uint32_t smask = (cpr & (cpr - 1)) == 0 ? min(cpr, 8u) - 1u : 0u;    // XOR swizzle of the chunk index by row
    // rows shorter than 128 bytes share a bank line: the key advances once per line (keyed by the row itself, the
    // 16-byte stores of 64-byte rows collide two-way -- ncu: 8 instead of 4 wavefronts per STS.128)
    const uint32_t kshift = rowB >= 128u ? 0u : (rowB == 64u ? 1u : (rowB == 32u ? 2u : 3u));
    int sa = 0, pa = 0, k = 0;
    if (!chained) asm volatile("griddepcontrol.wait;" ::: "memory");        // the bias sits behind the packed weights
    for (int i = tid - EPI_WARP0 * 32; i < 3 * L.CoutP; i += TC_EPI) s_bias[i] = P.bias[i];
    asm volatile("bar.sync 2, %0;" ::"n"(TC_EPI) : "memory");
    TC_TILE_LOOP(k) {
      const TileInfo T = decode_tile(tile, P.batch, L.tpf);
      const int MBc = min(L.MB, L.nmb - T.tf * L.MB);
      const float *bias = s_bias + T.grp * L.CoutP;
      const size_t face_px = (size_t)(T.b * 6 + T.f) * P.Hout * P.Wout;
      stamp(P, 8, k, tid == EPI_WARP0 * 32);
      mbar_wait(bar_afull + 8 * sa, pa);
      tc_fence_after();
      if (DYN && P.done && k > 0) {
        // count the PREVIOUS tile as stored by this warp (release at CTA scope, after the warp has converged).  Done here,
        // one tile late and before this tile's first global store, the release finds the old stores already drained and
        // costs nothing on the epilogue's critical path; lane 2 of the producer warp turns complete tiles into the
        // device-wide per-sample counter (one device-scope fence per tile per CTA)
        __syncwarp();
        if (lane == 0) asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(epi_ctr) : "memory");
      }
      trace(P, 4, k == 0 && tid == EPI_WARP0 * 32);
      stamp(P, 9, k, tid == EPI_WARP0 * 32);
      for (int mb = half; mb < MBc; mb += 2) {
        const int q = (T.tf * L.MB + mb) * 128 + quarter * 32 + lane;
        const int r = q / L.Wv, c = q - r * L.Wv;
        const bool ok = q < L.Q && c < P.Wout;
        const size_t opix = face_px + (size_t)r * P.Wout + c;
        const unsigned okmask = __ballot_sync(0xffffffffu, ok);
        const int nvalid = __popc(okmask);
        const size_t opix0 = __shfl_sync(0xffffffffu, opix, okmask ? __ffs(okmask) - 1 : 0);   // first valid row of the warp
        const uint32_t prow = (uint32_t)(opix - opix0);                          // compacted row of this lane
        const uint32_t srow = stg + prow * rowB;
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((sa * L.MB + mb) * L.CoutP);
        for (int n0 = 0; n0 < L.CoutP; n0 += 32) {
          uint32_t v[32];
          const bool two = n0 + 16 < L.CoutP;
          tmem_ld16(trow + (uint32_t)n0, v);
          if (two) tmem_ld16(trow + (uint32_t)n0 + 16u, v + 16);
          tmem_ld_wait();
          if (!ok || (P.knock & 4)) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const int nb = n0 + 16 * h;
            float o[16];
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const float4 bv = *reinterpret_cast<const float4 *>(bias + nb + 4 * k4);
              o[4 * k4 + 0] = __uint_as_float(v[16 * h + 4 * k4 + 0]) + bv.x;
              o[4 * k4 + 1] = __uint_as_float(v[16 * h + 4 * k4 + 1]) + bv.y;
              o[4 * k4 + 2] = __uint_as_float(v[16 * h + 4 * k4 + 2]) + bv.z;
              o[4 * k4 + 3] = __uint_as_float(v[16 * h + 4 * k4 + 3]) + bv.w;
            }
            if (act_fast) {            // capped leaky ReLU with 0 <= slope <= 1: min(max(v, slope v), cap)
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = fminf(fmaxf(o[e], P.slope * o[e]), P.maxv);
            } else if (P.act != DLWPCS_ACT_NONE) {
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = act_apply(o[e], P.act, P.slope, P.maxv);
            }
            if (staged) {
              if (nb + 8 <= P.cout)
                st_shared16(srow + ((((uint32_t)nb >> 3) ^ ((prow >> kshift) & smask)) << 4), make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                                  pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
              if (nb + 16 <= P.cout)
                st_shared16(srow + (((((uint32_t)nb >> 3) + 1u) ^ ((prow >> kshift) & smask)) << 4), make_uint4(pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]),
                                                                        pack_bf16x2(o[12], o[13]), pack_bf16x2(o[14], o[15])));
            } else if (fast_out && nb + 16 <= P.cout) {
              __nv_bfloat16 *yp = reinterpret_cast<__nv_bfloat16 *>(P.y) + opix * P.cout + nb;
              reinterpret_cast<uint4 *>(yp)[0] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                            pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
              reinterpret_cast<uint4 *>(yp)[1] = make_uint4(pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]),
                                                            pack_bf16x2(o[12], o[13]), pack_bf16x2(o[14], o[15]));
            } else if (P.y_f32) {
              float *yp = reinterpret_cast<float *>(P.y) + opix * P.cout + nb;
              if (P.cout % 4 == 0 && nb + 16 <= P.cout) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  reinterpret_cast<float4 *>(yp)[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
              } else {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                  if (nb + k < P.cout) yp[k] = o[k];
              }
            } else {
              __nv_bfloat16 *yp = reinterpret_cast<__nv_bfloat16 *>(P.y) + opix * P.cout + nb;
#pragma unroll
              for (int k = 0; k < 16; ++k)
                if (nb + k < P.cout) yp[k] = __float2bfloat16_rn(o[k]);
            }
          }
        }
        if (staged && !(P.knock & 12)) {
          __syncwarp();
          uint8_t *gdst = reinterpret_cast<uint8_t *>(P.y) + opix0 * rowB;
          const uint32_t total = (uint32_t)nvalid * rowB;
          for (uint32_t off = (uint32_t)lane * 16u; off < total; off += 512u) {
            uint32_t row, ch;
            if (cprLog >= 0) { row = off >> (4 + cprLog); ch = (off >> 4) & (cpr - 1u); }
            else { row = off / rowB; ch = (off - row * rowB) >> 4; }
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stg + row * rowB + ((ch ^ ((row >> kshift) & smask)) << 4)));
            *reinterpret_cast<uint4 *>(gdst + off) = v;
          }
          __syncwarp();                                 // the buffer is rewritten by the next m-block
        }
      }
      tc_fence_before();
      mbar_arrive(bar_aempty + 8 * sa);
      stamp(P, 10, k, tid == EPI_WARP0 * 32);
      if (++sa == L.AS) { sa = 0; pa ^= 1; }
    }
    if (DYN && P.done && k > 0) {                 // the last tile of the CTA
      __syncwarp();
      if (lane == 0) asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(epi_ctr) : "memory");
    }
    trace(P, 5, tid == EPI_WARP0 * 32);
    if (P.trace && tid == EPI_WARP0 * 32) P.trace[blockIdx.x * 8 + 7] = (unsigned long long)k;
  } else if (warp < LOAD_WARP0 + 8) {
    // ===== patch loaders =====
    const int lt = tid - LOAD_WARP0 * 32;
    int si = 0, pi = 0, k = 0, ts = 0, tp = 0;
    const int w2_0 = P.n * 2;
    if (!chained) asm volatile("griddepcontrol.wait;" ::: "memory");
    trace(P, 2, lt == 0);
    TC_TILE_LOOP(k) {
      const TileInfo T = decode_tile(tile, P.batch, L.tpf);
      const int MBc = min(L.MB, L.nmb - T.tf * L.MB);
      const int npix = MBc * 128 + L.haloExt;
      stamp(P, 1, k, lt == 0);
      mbar_wait(bar_pempty + 8 * si, pi ^ 1);
      stamp(P, 2, k, lt == 0);
      const uint32_t stage = patch0 + (uint32_t)si * L.patchBytes;
      mbar_wait(bar_tfull + 8 * ts, tp);
      const int32_t *t0 = s_tab + (size_t)ts * (L.tabBytes / 4);
      const int32_t *t1 = L.twoTabs ? t0 + L.NPIXp : t0;
      const size_t b0 = (size_t)T.b * P.ppb0, b1 = (size_t)T.b * P.ppb1;
      if (P.knock & 1) {
      } else if (L.vec && L.logS >= 0) {
        // each thread owns one 16-byte chunk (fixed source / channel offset) of every (256 >> logS)-th patch row
        const int chunk = lt & (L.S - 1), c = chunk * 8;
        const bool first = c < P.c0;
        const __nv_bfloat16 *src = first ? P.x0 : P.x1;
        const int C = first ? P.c0 : P.c1, cc = first ? c : c - P.c0, mode = first ? P.mode0 : P.mode1;
        const int32_t *tab = first ? t0 : t1;
        const size_t boff = first ? b0 : b1;
        const bool chan_ok = c < P.cin;
        const int pstep = TC_LOADERS >> L.logS;
        const uint32_t cw = (uint32_t)chunk & ((1u << L.cprLog) - 1u);            // chunk within its row
        const uint32_t dst0 = stage + (uint32_t)(chunk >> L.cprLog) * L.blockBytes;
        const bool direct = mode != DLWPCS_SRC_POOL2 && !P.mask_y;
        if (direct) {
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
                const __nv_bfloat16 *g = px[e] >= 0 ? src + (boff + px[e]) * C + cc : P.x0;
                cp_async16(row + ((cw ^ ((row >> 7) & (uint32_t)L.swzMask)) << 4), g, px[e] >= 0 ? 16u : 0u);
              }
            }
          }
        } else if (mode == DLWPCS_SRC_POOL2 && !P.mask_y) {
          // (ld.global.nc stays legal under a chained launch: a sample's pixels are only ever read after the producing
          // launch has completed that sample, and L1 is invalidated at the start of every kernel, so no line of it can
          // be cached from before it was final.  ld.global.cg here costs spills across the whole kernel.)
          // 2x2 mean (AveragePooling3D((1,2,2)), train_cs.py:197) through registers: two patch rows = eight 16-byte
          // loads in flight per thread; the mean is rounded to bf16 once, like a stored pooled tensor would be
          for (int i0 = lt >> L.logS; i0 < npix; i0 += 2 * pstep) {
            uint4 v[2][4];
            int px[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = i0 + e * pstep;
              px[e] = (i < npix && chan_ok) ? tab[i] : -1;
              if (px[e] >= 0) {
                const __nv_bfloat16 *g = src + (boff + px[e]) * C + cc;
                v[e][0] = __ldg(reinterpret_cast<const uint4 *>(g));
                v[e][1] = __ldg(reinterpret_cast<const uint4 *>(g + C));
                v[e][2] = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)w2_0 * C));
                v[e][3] = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)(w2_0 + 1) * C));
              }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = i0 + e * pstep;
              if (i < npix) {
                uint4 o = make_uint4(0, 0, 0, 0);
                if (px[e] >= 0) {
                  float a[8], t[8];
                  unpack_bf16x8(v[e][0], a);
#pragma unroll
                  for (int q = 1; q < 4; ++q) {
                    unpack_bf16x8(v[e][q], t);
#pragma unroll
                    for (int k2 = 0; k2 < 8; ++k2) a[k2] += t[k2];
                  }
                  o = make_uint4(pack_bf16x2(0.25f * a[0], 0.25f * a[1]), pack_bf16x2(0.25f * a[2], 0.25f * a[3]),
                                 pack_bf16x2(0.25f * a[4], 0.25f * a[5]), pack_bf16x2(0.25f * a[6], 0.25f * a[7]));
                }
                const uint32_t row = dst0 + (uint32_t)i * L.RB;
                st_shared16(row + ((cw ^ ((row >> 7) & (uint32_t)L.swzMask)) << 4), o);
              }
            }
          }
        } else {
#pragma unroll 1
          for (int i = lt >> L.logS; i < npix; i += pstep) {
            const int px = chan_ok ? tab[i] : -1;
            const uint32_t row = dst0 + (uint32_t)i * L.RB;
            uint4 o = make_uint4(0, 0, 0, 0);
            if (px >= 0) o = gather_regs(P, src, C, mode, boff + px, cc, w2_0);
            st_shared16(row + ((cw ^ ((row >> 7) & (uint32_t)L.swzMask)) << 4), o);
          }
        }
      } else {
        generic_gather(P, stage, npix, t0, t1, b0, b1, lt);
      }
      // the stage is published when this thread's copies have landed (register-path stores are already visible; the
      // proxy fence orders them before the tensor core's reads) -- no blocking, so the loaders run ahead of the MMA warp
      mbar_arrive(bar_tempty + 8 * ts);          // this thread is done reading the table slot
      if (++ts == L.TS) { ts = 0; tp ^= 1; }
      fence_proxy_async();
      cp_async_mbar_arrive(bar_pfull + 8 * si);
      stamp(P, 3, k, lt == 0);
      if (++si == L.PS) { si = 0; pi ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  trace(P, 6, tid == 0);
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, (uint32_t)L.tmemCols);
}

// ---- weight packing -----------------------------------------------------------------------------------------------
// packed[g][kc][tap][k8][n][8] bf16  (+ fp32 bias[3][CoutP] after the three groups).  transposed (dgrad): the GEMM's
// K runs over the forward cout, N over the forward cin, taps rotated 180 degrees.
// One launch packs the forward image (blockIdx.y = 0) and / or the transposed one (blockIdx.y = 1).  The source kernels
// have the logical shape (kh, kw, scin, scout) and are zero-extended to the descriptor's (cin, cout): the bf16 path's
// channel padding never materialises padded copies of the parameters.
// One launch packs any number of images (forward / transposed / row-streamed forward) of any number of layers:
// blockIdx.y = image, grid-stride over its elements (cs_pack.cuh).
__global__ void pack_batch_kernel(const __grid_constant__ PackBatch B) {
  const PackImage &S = B.im[blockIdx.y];
  const long long total = 3 * S.groupElems + 3LL * S.CoutP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    pack_element(S, i);
}

int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

// Tiling / staging plan for one layer; returns a reason string when the configuration is outside the kernel's reach.
const char *make_plan(const dlwpcs_conv_desc *d, const Geometry &g, int gemm_cin, int gemm_cout, TcPlan *L) {
  if (d->stride_h != 1 || d->stride_w != 1) return "strides must be 1";
  L->CinP = (gemm_cin + 15) / 16 * 16;
  if (L->CinP > 32) L->CinP = (gemm_cin + 63) / 64 * 64;     // rows of 32 / 64 / 128 bytes; 64 channels per K block
  L->CoutP = (gemm_cout + 15) / 16 * 16;
  if (L->CoutP > 256) return "more than 256 output channels";
  L->KC = L->CinP < 64 ? L->CinP : 64;
  L->nch = L->CinP / L->KC;
  L->RB = L->KC * 2;
  L->cprLog = L->RB == 128 ? 3 : (L->RB == 64 ? 2 : 1);
  L->swzMask = (1 << L->cprLog) - 1;
  L->layoutType = L->RB == 128 ? 2 : (L->RB == 64 ? 4 : 6);   // cute::UMMA::LayoutType SWIZZLE_128B / 64B / 32B
  L->S = L->CinP / 8;
  L->logS = -1;
  for (int l = 0; l <= 5; ++l)
    if ((1 << l) == L->S) L->logS = l;
  L->taps = d->kh * d->kw;
  L->NU = L->nch * L->taps;
  L->unitBytes = L->KC * L->CoutP * 2;
  L->UPS = 16384 / L->unitBytes;
  if (L->UPS < 1) L->UPS = 1;
  if (L->UPS > L->NU) L->UPS = L->NU;
  L->stageBytes = L->UPS * L->unitBytes;
  L->nstages = (L->NU + L->UPS - 1) / L->UPS;
  L->groupBytes = (int64_t)L->NU * L->unitBytes;
  L->resident = (L->nstages <= MAX_STAGES && (int64_t)L->nstages * L->stageBytes <= 96 * 1024) ? 1 : 0;
  L->NST = L->resident ? L->nstages : 3;
  if (!L->resident) {
    const int want = env_int("DLWPCS_TC_NST", 3);
    if (want >= 2 && want <= MAX_STAGES) L->NST = want;
  }
  L->Wv = g.Wout + (d->kw - 1) * d->dil_w;
  L->Hv = g.Hout + (d->kh - 1) * d->dil_h;
  L->Q = (g.Hout - 1) * L->Wv + g.Wout;
  L->nmb = (L->Q + 127) / 128;
  L->haloExt = (d->kh - 1) * d->dil_h * L->Wv + (d->kw - 1) * d->dil_w;
  L->G = (L->nmb * 128 + L->haloExt + 3) / 4 * 4;
  L->TS = 2;
  L->twoTabs = (d->c1 > 0 && d->mode1 != d->mode0) ? 1 : 0;
  if (L->NU > MAX_UNITS) return "kernel window x input channels too large";
  // bf16 outputs with 16-byte rows are compacted in shared memory and written with TMA bulk stores
  L->stgBytes = (d->y_dtype == DLWPCS_BF16 && gemm_cout % 8 == 0 && !env_int("DLWPCS_TC_NOSTAGE", 0)) ? 32 * gemm_cout * 2 : 0;
  L->stgBufs = L->stgBytes ? 1 : 0;
  // Streamed weights and >= 128 output channels (256-byte rows: a lane's direct 16-byte stores fill whole sectors anyway): the
  // staging buffers' 64 KB go to a deeper weight ring instead (measured, 64 -> 128 at 12 x 12, batch 64: 18.4 instead of 21.8 us;
  // with 128-byte rows the direct stores cost more than the ring gains)
  if (!L->resident && L->stgBytes && gemm_cout * 2 >= 256 && env_int("DLWPCS_TC_WIDE_DIRECT", 1)) {
    L->stgBytes = 0;
    L->stgBufs = 0;
    if (L->NST == 3) L->NST = 5;
  }
  int fixed0 = 1024 + 512 + MAX_UNITS * 8 + 3 * L->CoutP * 4 + L->NST * L->stageBytes + 8 * L->stgBufs * L->stgBytes + 128;
  auto patch_for = [&](int MB, int *npixp) {
    const int np = ((MB * 128 + L->haloExt + 7) / 8) * 8;
    *npixp = np;
    return (L->nch * np * L->RB + 1023) / 1024 * 1024;            // stages stay 1024-byte aligned (swizzle period)
  };
  auto tabs_for = [&](int MB) { return L->TS * (1 + L->twoTabs) * (((MB * 128 + L->haloExt + 7) / 8) * 8) * 4; };
  int mbmax = 512 / L->CoutP;
  if (mbmax > 4) mbmax = 4;
  if (mbmax > L->nmb) mbmax = L->nmb;
  int best = 0, np = 0;
  const int forced = env_int("DLWPCS_TC_MB", 0);
  const int stg_unit = 8 * L->stgBytes;           // one staging buffer for each of the 8 epilogue warps
  fixed0 -= L->stgBufs * stg_unit;                // staging is decided together with the tile size below
  const int bufs_max = L->stgBufs;
  for (int MB = (forced > 0 && forced < mbmax) ? forced : mbmax; MB >= 1 && !best; --MB)
    for (int bufs = bufs_max; bufs >= 0; --bufs)
      if (fixed0 + bufs * stg_unit + tabs_for(MB) + 2 * patch_for(MB, &np) <= SMEM_CAP) {
        best = MB;
        L->stgBufs = bufs;
        break;
      }
  if (best) {
    if (L->stgBufs == 0) L->stgBytes = 0;
    fixed0 += L->stgBufs * stg_unit;
  } else {
    L->stgBufs = 0;
    L->stgBytes = 0;
  }
  if (!best) {
    best = 1;                                      // single-buffered patch as the last resort
    if (fixed0 + tabs_for(1) + patch_for(1, &np) > SMEM_CAP) return "input patch does not fit shared memory (too many input channels)";
  }
  const int tiles = (L->nmb + best - 1) / best;
  best = (L->nmb + tiles - 1) / tiles;             // even out the tiles of a face
  L->MB = best;
  L->tpf = (L->nmb + best - 1) / best;
  L->patchBytes = patch_for(best, &L->NPIXp);
  const int fixed = fixed0 + tabs_for(best);
  L->blockBytes = L->NPIXp * L->RB;
  L->tabBytes = (1 + L->twoTabs) * L->NPIXp * 4;
  L->PS = (SMEM_CAP - fixed) / L->patchBytes;
  int ps_max = env_int("DLWPCS_TC_PS", 3);
  if (ps_max > MAX_PS) ps_max = MAX_PS;
  if (L->PS > ps_max) L->PS = ps_max;
  if (L->PS < 1) return "input patch does not fit shared memory";
  L->AS = (2 * best * L->CoutP <= 512) ? 2 : 1;
  if (env_int("DLWPCS_TC_AS", 2) < 2) L->AS = 1;
  int cols = 32;
  while (cols < L->AS * best * L->CoutP) cols *= 2;
  L->tmemCols = cols;
  L->smemBytes = fixed + L->PS * L->patchBytes;
  return nullptr;
}

// ---- patch tables: for every position of the linearised virtual face, the physical source pixel (within one batch
// element of the source tensor, in its own resolution) or -1 for zero fill.  One table per (geometry, sampling mode).
struct TabKey {
  int dev, n, halo, Hin, Win, Wv, G, pt0, pt1, pt2, pl, mode;
  bool operator<(const TabKey &o) const { return memcmp(this, &o, sizeof(TabKey)) < 0; }
};
std::mutex g_tab_mu;
std::map<TabKey, int32_t *> g_tabs;

}  // namespace

const int32_t *get_patch_table(const Geometry &g, int Wv, int G, int n, int halo, int mode) {
  TabKey key;
  memset(&key, 0, sizeof(key));
  if (cudaGetDevice(&key.dev) != cudaSuccess) {
    set_error("cudaGetDevice failed");
    return nullptr;
  }
  key.n = n; key.halo = halo; key.Hin = g.Hin; key.Win = g.Win; key.Wv = Wv; key.G = G;
  key.pt0 = g.pt[0]; key.pt1 = g.pt[1]; key.pt2 = g.pt[2]; key.pl = g.pl; key.mode = mode;
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto it = g_tabs.find(key);
  if (it != g_tabs.end()) return it->second;
  std::vector<int32_t> lut;
  if (halo > 0) build_pad_lut(n, halo, lut);
  std::vector<int32_t> tab((size_t)6 * G);
  for (int f = 0; f < 6; ++f) {
    const int pt = g.pt[face_group_host(f)];
    for (int q = 0; q < G; ++q) {
      const int rv = q / Wv, cv = q % Wv, r = rv - pt, c = cv - g.pl;
      int32_t val = -1;
      if (r >= 0 && r < g.Hin && c >= 0 && c < g.Win) {
        const int s = halo > 0 ? lut[((size_t)f * g.Hin + r) * g.Win + c] : (f * g.Hin + r) * g.Win + c;
        const int sf = s / (n * n), si = (s % (n * n)) / n, sj = s % n;
        if (mode == DLWPCS_SRC_SAME) val = s;
        else if (mode == DLWPCS_SRC_UP2) val = (sf * (n / 2) + si / 2) * (n / 2) + sj / 2;
        else val = (sf * 2 * n + 2 * si) * 2 * n + 2 * sj;
      }
      tab[(size_t)f * G + q] = val;
    }
  }
  int32_t *dev = nullptr;
  // plain cudaMalloc + synchronous copy: never inside a stream capture -- callers warm the cache first
  cudaError_t e = cudaMalloc(&dev, tab.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemcpy(dev, tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("patch table upload failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  g_tabs[key] = dev;
  return dev;
}

namespace {

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int num_sms() {
  static int sms[kMaxDevices] = {};
  const int dev = current_device_index();
  if (!sms[dev] && (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0))
    sms[dev] = 148;
  return sms[dev];
}

unsigned long long *g_trace = nullptr;      // [TRACE_LAUNCHES][TRACE_CTAS][8]
int g_trace_launches = 0;
}  // namespace
unsigned long long *tc_trace_next(int grid);
namespace {

int launch_tc(TcP &P, cudaStream_t st) {
  const TcPlan &L = P.pl_;
  typedef void (*kern_t)(const TcP);
#define TC_ROW(MB, D) {conv_tc_kernel<MB, 1, D>, conv_tc_kernel<MB, 2, D>, conv_tc_kernel<MB, 3, D>, conv_tc_kernel<MB, 4, D>}
  static const kern_t kerns[2][4][4] = {{TC_ROW(1, false), TC_ROW(2, false), TC_ROW(3, false), TC_ROW(4, false)},
                                        {TC_ROW(1, true), TC_ROW(2, true), TC_ROW(3, true), TC_ROW(4, true)}};
#undef TC_ROW
  static bool attr_set[kMaxDevices][2][4][4] = {};      // the shared-memory opt-in is per device
  CS_CHECK(L.MB >= 1 && L.MB <= 4 && L.KC % 16 == 0 && L.KC <= 64, "internal: bad tile plan");
  const int ki = L.MB - 1, kj = L.KC / 16 - 1, dev_i = current_device_index();
  const int dyn = P.tile_ctr ? 1 : 0;
  const kern_t kern = kerns[dyn][ki][kj];
  if (!attr_set[dev_i][dyn][ki][kj]) {
    CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP));
    attr_set[dev_i][dyn][ki][kj] = true;
  }
  const long long ntiles = 6LL * P.batch * L.tpf;
  CS_CHECK(ntiles < (1LL << 30), "batch too large");
  int grid = num_sms();
  if (grid > ntiles) grid = (int)ntiles;
  // CTA c walks tiles c, c + grid, ...: when the tiles of a face differ in size (e.g. 5 m-blocks = 3 + 2) and grid and
  // tiles-per-face share a factor, some CTAs would only ever see the large tiles -- keep the two coprime (the dynamic
  // schedule of a chained launch hands tiles out on demand and uses every SM)
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  while (!P.tile_ctr && grid > 1 && L.nmb % L.MB != 0 && gcd(grid, L.tpf) != 1) --grid;
  static const int timing = env_int("DLWPCS_TC_TIMING", 0);
  static const int knock = env_int("DLWPCS_TC_KNOCK", 0);
  P.knock = knock;
  static long long *dbg = nullptr;
  if (timing) {
    if (!dbg) CS_CUDA(cudaMalloc(&dbg, 16 * sizeof(long long)));
    CS_CUDA(cudaMemset(dbg, 0, 16 * sizeof(long long)));
    P.dbg = dbg;
  }
  P.trace = tc_trace_next(grid);
  static const int pdl = env_int("DLWPCS_TC_PDL", 1);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = (size_t)L.smemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  CS_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  if (timing) {
    long long h[16];
    CS_CUDA(cudaDeviceSynchronize());
    CS_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr,
            "[tc timing] cin=%d cout=%d n=%d MB=%d tpf=%d PS=%d AS=%d res=%d smem=%d tiles/cta=%.1f | tile %d of the mid "
            "CTA, cycles since kernel start: loader wait %lld got %lld issued %lld | mma wait %lld "
            "start %lld issued %lld (next start %lld) | epi wait %lld ready %lld end %lld\n",
            P.cin, P.cout, P.n, L.MB, L.tpf, L.PS, L.AS, L.resident, L.smemBytes, (double)ntiles / grid, DBG_K,
            h[1] - h[0], h[2] - h[0], h[3] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[11] - h[0],
            h[8] - h[0], h[9] - h[0], h[10] - h[0]);
  }
  return 0;
}

}  // namespace

// trace slot of the next traced launch (nullptr when tracing is off / the grid does not fit); shared with cs_tc_rs.cu
unsigned long long *tc_trace_next(int grid) {
  static const int tracing = env_int("DLWPCS_TC_TRACE", 0);
  if (!tracing || grid > TRACE_CTAS) return nullptr;
  if (!g_trace) {
    if (cudaMalloc(&g_trace, sizeof(unsigned long long) * TRACE_LAUNCHES * TRACE_CTAS * 8) != cudaSuccess) return nullptr;
    cudaMemset(g_trace, 0, sizeof(unsigned long long) * TRACE_LAUNCHES * TRACE_CTAS * 8);
  }
  unsigned long long *p = g_trace + (size_t)(g_trace_launches % TRACE_LAUNCHES) * TRACE_CTAS * 8;
  ++g_trace_launches;
  return p;
}

int tc_trace_read(unsigned long long *host_out, int max_launches, int *n_launches, int reset) {
  const int n = g_trace_launches < TRACE_LAUNCHES ? g_trace_launches : TRACE_LAUNCHES;
  if (n_launches) *n_launches = n;
  const int m = n < max_launches ? n : max_launches;
  if (m > 0 && host_out && g_trace) {
    CS_CUDA(cudaDeviceSynchronize());
    CS_CUDA(cudaMemcpy(host_out, g_trace, sizeof(unsigned long long) * (size_t)m * TRACE_CTAS * 8, cudaMemcpyDeviceToHost));
  }
  if (reset) g_trace_launches = 0;
  return 0;
}

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
  // 0: forward image for dlwpcs_conv2d_fwd (the row-streamed kernel's where it applies), 1: transposed (dgrad),
  // 2: forward image of the classic kernel whatever the layer (chained launches)
  if (transposed == 0 && rs_eligible(d, g)) return rs_packed_weight_bytes(d, g);
  if (transposed == 2) transposed = 0;
  TcPlan L;
  const char *r = make_plan(d, g, transposed ? d->cout : d->cin, transposed ? d->cin : d->cout, &L);
  if (r) {
    set_error("bf16 tensor-core path does not support this configuration: %s", r);
    return -1;
  }
  return 3 * L.groupBytes + 3LL * L.CoutP * 4;
}

// Fill the image records of one layer (forward image in `packed`, transposed in `packed_t`; either may be null); returns
// the number of records appended or -1 (error set).
static int pack_images_of(const dlwpcs_conv_desc *d, const Geometry &g, const dlwpcs_conv_weights *w, int src_cin, int src_cout,
                          void *packed, void *packed_t, int classic_forward, PackImage *im) {
  if (!(src_cin >= 1 && src_cin <= d->cin && src_cout >= 1 && src_cout <= d->cout)) {
    set_error("source kernel shape (%d, %d) does not fit the descriptor's (%d, %d)", src_cin, src_cout, d->cin, d->cout);
    return -1;
  }
  int n = 0;
  for (int t = 0; t < 2; ++t) {
    void *out = t ? packed_t : packed;
    if (!out) continue;
    PackImage &I = im[n];
    memset(&I, 0, sizeof(I));
    I.w_eq = w->w_eq; I.w_pol = w->w_pol; I.w_np = d->independent_north_pole ? w->w_np : nullptr;
    I.b_eq = d->use_bias ? w->b_eq : nullptr; I.b_pol = d->use_bias ? w->b_pol : nullptr;
    I.b_np = (d->use_bias && d->independent_north_pole) ? w->b_np : nullptr;
    I.out = (uint8_t *)out;
    I.kh = d->kh; I.kw = d->kw; I.cin = d->cin; I.cout = d->cout; I.scin = src_cin; I.scout = src_cout;
    I.flip = d->flip_north_pole;
    if (t == 0 && !classic_forward && rs_eligible(d, g)) {
      // the forward image goes to the row-streamed kernel's layout; the transposed one (if any) stays classic
      I.kind = PACK_RS_FWD;
      if (!rs_pack_params(d, g, &I.CinP, &I.CoutP, &I.groupElems)) return -1;
      I.KC = I.CinP;
    } else {
      TcPlan L;
      const char *r = make_plan(d, g, t ? d->cout : d->cin, t ? d->cin : d->cout, &L);
      if (r) {
        set_error("bf16 tensor-core path does not support this configuration: %s", r);
        return -1;
      }
      I.kind = t ? PACK_CLASSIC_T : PACK_CLASSIC_FWD;
      I.CinP = L.CinP; I.CoutP = L.CoutP; I.KC = L.KC;
      I.groupElems = L.groupBytes / 2;
    }
    ++n;
  }
  return n;
}

static int launch_pack(const PackBatch &B, cudaStream_t st) {
  if (B.n == 0) return 0;
  long long most = 0;
  for (int i = 0; i < B.n; ++i) {
    const long long t = 3 * B.im[i].groupElems + 3LL * B.im[i].CoutP;
    most = t > most ? t : most;
  }
  long long blocks = (most + 255) / 256;
  if (blocks > 296) blocks = 296;                        // grid-stride: two waves of CTAs per image are plenty
  pack_batch_kernel<<<dim3((unsigned)blocks, (unsigned)B.n), 256, 0, st>>>(B);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int tc_pack_weights2(const dlwpcs_conv_desc *d, const Geometry &g, const dlwpcs_conv_weights *w, int src_cin,
                     int src_cout, void *packed, void *packed_t, cudaStream_t st, int classic_forward) {
  CS_CHECK(packed || packed_t, "nothing to pack");
  PackBatch B;
  B.n = pack_images_of(d, g, w, src_cin, src_cout, packed, packed_t, classic_forward, B.im);
  if (B.n < 0) return 1;
  return launch_pack(B, st);
}

// Every layer of a network in one launch (or a few: PACK_MAX_IMAGES images per launch)
int tc_pack_batch(const dlwpcs_pack_item *items, int n, cudaStream_t st) {
  PackBatch B;
  B.n = 0;
  for (int i = 0; i < n; ++i) {
    const dlwpcs_pack_item &it = items[i];
    Geometry g;
    if (int rc = derive_geometry(it.desc, &g)) return rc;
    if (B.n + 2 > PACK_MAX_IMAGES) {
      if (int rc = launch_pack(B, st)) return rc;
      B.n = 0;
    }
    const int k = pack_images_of(it.desc, g, &it.w, it.src_cin, it.src_cout, it.packed, it.packed_t, 0, B.im + B.n);
    if (k < 0) return 1;
    B.n += k;
  }
  return launch_pack(B, st);
}

int tc_pack_weights(const dlwpcs_conv_desc *d, const Geometry &g, const dlwpcs_conv_weights *w, int transposed,
                    void *packed, cudaStream_t st) {
  return tc_pack_weights2(d, g, w, d->cin, d->cout, transposed == 1 ? nullptr : packed, transposed == 1 ? packed : nullptr,
                          st, transposed == 2);
}

uint32_t tc_chain_target(const dlwpcs_conv_desc *d, const Geometry &g) {
  TcPlan L;
  if (make_plan(d, g, d->cin, d->cout, &L)) return 0;
  return 6u * (uint32_t)L.tpf;      // the counter of a sample moves once per completed tile
}

int tc_conv_fwd(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed,
                void *y, const dlwpcs_chain *chain, cudaStream_t st) {
  if (!chain && rs_eligible(d, g)) return rs_conv_fwd(d, g, x0, x1, packed, y, st);
  TcP P;
  memset(&P, 0, sizeof(P));
  const char *r = make_plan(d, g, d->cin, d->cout, &P.pl_);
  CS_CHECK(r == nullptr, "bf16 tensor-core path does not support this configuration: %s", r);
  TcPlan &L = P.pl_;
  if (chain) {
    CS_CHECK(chain->tile_counter != nullptr, "chained launch needs a tile counter");
    CS_CHECK(chain->dep == nullptr || chain->dep_target > 0, "chained launch: dependency counters without a target");
    P.dep = chain->dep;
    P.dep_target = chain->dep_target;
    P.done = chain->done;
    P.tile_ctr = chain->tile_counter;
    P.chain_err = chain->error_flag;
  }
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
  P.bias = reinterpret_cast<const float *>(P.wpack + 3 * L.groupBytes);
  P.y = y;
  P.y_f32 = d->y_dtype == DLWPCS_F32;
  P.mask_y = nullptr;
  P.batch = d->batch; P.n = d->n; P.Hout = g.Hout; P.Wout = g.Wout;
  P.cin = d->cin; P.cout = d->cout; P.c0 = d->c0; P.c1 = d->c1; P.mode0 = d->mode0; P.mode1 = d->mode1;
  P.kw = d->kw; P.dh = d->dil_h; P.dw = d->dil_w;
  P.act = d->act; P.slope = d->act_slope; P.maxv = d->act_max;
  L.vec = (d->c0 % 8 == 0) && (d->c1 % 8 == 0) && aligned16(x0) && (d->c1 == 0 || aligned16(x1));
  return launch_tc(P, st);
}

// dgrad on the same kernel: dx_ext[r,c,ci] = sum_{u,v,o} dy[r + pt - u*dh, c + pl - v*dw, o] * W[u,v,ci,o] is a convolution
// over dy (scaled by the activation derivative taken from the forward output) with the taps rotated 180 degrees and the
// channels swapped (weights packed with transposed = 1), zero fill outside dy, followed by the adjoint of the halo gather.
int tc_conv_dgrad(const dlwpcs_conv_desc *d, const Geometry &g, const void *dy, const void *y, const void *packed_t,
                  void *dx, void *workspace, const void *x_in, int in_act, float in_slope, float in_max, cudaStream_t st) {
  dlwpcs_conv_desc dd = *d;
  dd.cin = d->cout; dd.cout = d->cin; dd.c0 = d->cout; dd.c1 = 0;
  dd.mode0 = dd.mode1 = DLWPCS_SRC_SAME;
  dd.halo = 0; dd.n = g.Hout;
  dd.x_dtype = dd.y_dtype = DLWPCS_BF16;
  dd.act = DLWPCS_ACT_NONE; dd.use_bias = 0;
  Geometry gg;
  gg.Hin = g.Hout; gg.Win = g.Wout; gg.Hout = g.Hin; gg.Wout = g.Win;
  for (int k = 0; k < 3; ++k) gg.pt[k] = (d->kh - 1) * d->dil_h - g.pt[k];
  gg.pl = (d->kw - 1) * d->dil_w - g.pl;
  gg.taps = g.taps;
  TcP P;
  memset(&P, 0, sizeof(P));
  const char *r = make_plan(&dd, gg, d->cout, d->cin, &P.pl_);
  CS_CHECK(r == nullptr, "bf16 tensor-core dgrad does not support this configuration: %s", r);
  TcPlan &L = P.pl_;
  P.x0 = (const __nv_bfloat16 *)dy;
  P.x1 = nullptr;
  P.tab0 = get_patch_table(gg, L.Wv, L.G, g.Hout, 0, DLWPCS_SRC_SAME);
  if (!P.tab0) return 3;
  P.tab1 = P.tab0;
  P.ppb0 = P.ppb1 = 6 * g.Hout * g.Wout;
  P.wpack = (const uint8_t *)packed_t;
  P.bias = reinterpret_cast<const float *>(P.wpack + 3 * L.groupBytes);      // zeros in a transposed pack
  P.y = d->halo > 0 ? workspace : dx;
  P.y_f32 = 0;
  P.mask_y = d->act != DLWPCS_ACT_NONE ? y : nullptr;
  P.mask_f32 = 0;
  P.mask_act = d->act; P.mask_slope = d->act_slope; P.mask_max = d->act_max;
  P.batch = d->batch; P.n = g.Hout; P.Hout = gg.Hout; P.Wout = gg.Wout;
  P.cin = d->cout; P.cout = d->cin; P.c0 = d->cout; P.c1 = 0;
  P.mode0 = P.mode1 = DLWPCS_SRC_SAME;
  P.kw = d->kw; P.dh = d->dil_h; P.dw = d->dil_w;
  P.act = DLWPCS_ACT_NONE;
  L.vec = (d->cout % 8 == 0) && aligned16(dy) && (!P.mask_y || aligned16(y));
  if (int rc = launch_tc(P, st)) return rc;
  if (d->halo > 0)
    return dlwpcs_pad_bwd_act(workspace, x_in, dx, d->batch, d->n, d->cin, d->halo, in_act, in_slope, in_max, DLWPCS_BF16, st);
  return 0;
}

}  // namespace dlwpcs
