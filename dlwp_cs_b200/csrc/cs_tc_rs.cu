// tcgen05 (sm_100a) bf16 implicit-GEMM kernel of the cubed-sphere convolution, ROW-STREAMED with the three kernel rows
// stacked along N and the shift-add done by the tensor core itself ("slot ring").
//
// Why: an SS-mode tcgen05.mma re-reads its whole A tile from shared memory (tools/umma_probe.cu: M128 K16 costs
// max(N/2, (4 KB + N*32 B)/128 B) cycles), so the classic kernel (cs_tc.cu: one M128 x N=Cout MMA per tap) is capped at 38 %
// of the tensor peak for 32 output channels -- nine A reads per output block.  Here one MMA serves the three kernel rows:
//
//   geometry   the images (faces) of one weight group are concatenated ROW-WISE: position p = image * Wv + column over the
//              zero/halo padded width Wv; a *strip* is 128 consecutive positions = the 128 TMEM lanes.  Lanes whose column
//              is >= Wout (2 of every Wv) or past the last image are computed and dropped.
//   A          one padded input row y of the strip: 130 positions x Cin, pixel-major with the UMMA 32/64/128-byte swizzle,
//              gathered through the same patch tables as the classic kernel (halo exchange, pole rotation, pooling,
//              up-sampling, concatenation); the kernel column v is a +v position offset of the A start address.
//   B          per (K block, kernel column v): [W[2,v] | W[1,v] | W[0,v]] stacked along N (3 * CoutP columns).
//   D          a ring of NS = 512 / CoutP accumulator *slots* in tensor memory, one per OUTPUT row.  The MMAs of input row
//              y accumulate into the three adjacent slots of output rows y-2, y-1, y (kernel rows 2, 1, 0); the MMAs of row
//              y+1 into y-1, y, y+1; ...  tcgen05.mma instructions execute in issue order also when their accumulator
//              column ranges overlap only partially (tools/umma_ring_probe.cu, exact on the B200), so the vertical shift-add
//              costs nothing: a third of the A reads, and the epilogue reads finished sums exactly as before.
//              Every MMA accumulates: the slots start at zero and the epilogue warp that drains a slot clears it again
//              (tcgen05.st, no shared-memory traffic).
//   pipeline   loaders (8 warps; warp w gathers the rows m = w (mod 8) of the CTA's row sequence, a whole row per turn, its
//              patch-table entries prefetched one turn ahead by 4-byte cp.async) -> MMA issuer (1 warp, descriptors in uniform
//              registers) -> slot ring -> epilogue (8 warps: bias, capped leaky ReLU, bf16, per-warp compaction in shared
//              memory, coalesced stores; two warps per lane quarter take alternate rows), all mbarrier-driven; the weights of
//              the CTA's face group stay resident in shared memory.
//   work       every CTA gets a contiguous range of (strip, output row) pairs of equal COST, cut on the host and passed in
//              the kernel parameters (rs_work_cuts; units = the parts of a range inside one strip; a unit of h output rows
//              streams h + 2 input rows).
//   modes      MODE 1: the 1x1 layer that follows runs as a second MMA inside the epilogue (dlwpcs_conv2d_fwd_head);
//              MODE 2: the 2x2 mean of the output is written as a second tensor (dlwpcs_conv2d_fwd_pool).
//
// Reference semantics: DLWP/custom.py:921-1002 (CubeSphereConv2D.call) with the preceding CubeSpherePadding2D
// (custom.py:1198-1308) and the U-Net's pool / upsample / concatenate (Azure/train_cs.py:197-199, 282-299) folded into the
// load stage, bias + capped leaky ReLU in the epilogue.  Same products as cs_tc.cu; the float32 sums associate differently.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cs_common.cuh"
#include "cs_ptx.cuh"

namespace dlwpcs {

namespace {

constexpr int RS_THREADS = 576, RS_EPI = 256;
constexpr int RS_EPI_WARP0 = 8, RS_TMA_WARP = 16, RS_MMA_WARP = 17;
constexpr int RS_MAXSLOTS = 16, RS_MAXSTAGES = 16;
constexpr int RS_NPOS = 130;       // 128 lanes + (kw - 1) positions of overhang
constexpr int RS_NPIXP = 136;      // rows of one stage (multiple of 8)
constexpr int RS_SMEM_CAP = 227 * 1024;
constexpr int RS_MAXGRID = 160;

struct RsPlan {
  int CinP, KC16, S, logS, RB, cprLog, swzMask, layoutType;
  int CoutP, NT, NS;                 // padded output channels, stacked N = 3 * CoutP, accumulator slots
  int Wv, Hv, G;                     // padded face, patch-table entries per face
  int stageBytes, RSn;               // bytes per input-row stage, stages
  int unitBytes, groupBytes;         // weights of one (kernel column) unit / of one face group
  int stgBytes;                      // output staging per epilogue warp
  int off_w, off_zero, off_misc, off_bias, off_pix, off_pos, off_px, off_stg, smemBytes;
  // fused 1x1 head (conv_rs_kernel<.., 1>): a second MMA per output row inside the epilogue
  int head, CoutP2, Kh, hwBytes, off_hw, off_hbias, off_a2;     // Kh = K of the head's MMA = its padded input channels
};

struct RsP {
  const __nv_bfloat16 *x0, *x1;
  const int32_t *tab0, *tab1;        // [6][G] physical source pixel within one batch element, -1 = zero
  const uint8_t *wpack;
  const float *bias;                 // [3][CoutP]
  __nv_bfloat16 *y;
  const uint8_t *hw;                 // head: packed 1x1 weights [3][k8][CoutP2][8] (the classic kernel's image)
  const float *hbias;                // head: [3][CoutP2]
  __nv_bfloat16 *y2;                 // head output (B,6,Hout,Wout,cout2); y is not written
  __nv_bfloat16 *ypool;              // pool mode: second output (B,6,Hout/2,Wout/2,cout) = AveragePooling3D((1,2,2)) of y
  int cout2, act2;
  float slope2, maxv2;
  int batch, n, Hout, Wout;
  int cin, cout, c0, c1, mode0, mode1, ppb0, ppb1;
  int act;
  float slope, maxv;
  int knock;                         // bottleneck analysis (DLWPCS_RS_KNOCK): 1 no gathers, 2 no MMAs, 4 no epilogue math / stores
  int cut_s[RS_MAXGRID + 1], cut_y[RS_MAXGRID + 1];     // CTA c works on [(cut_s[c], cut_y[c]), (cut_s[c+1], cut_y[c+1]))
  unsigned long long *trace;         // optional per-CTA %globaltimer stamps (DLWPCS_TC_TRACE=1), same slots as cs_tc.cu
  unsigned *err;                     // watchdog flag (a barrier wait that never completes traps instead of hanging the GPU)
  RsPlan L;
};

// cross-kernel timeline (DLWPCS_TC_TRACE=1, tools/trace_step.py): slot 0 kernel entry, 1 prologue done, 2 loaders past the
// grid dependency, 3 first input row landed, 4 first output row complete, 5 last epilogue done, 6 kernel exit, 7 units
__device__ __forceinline__ void rs_trace(const RsP &P, int slot, bool who) {
  if (P.trace && who) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.trace[blockIdx.x * 8 + slot] = t;
  }
}

// Bounded barrier wait: ~2^26 polls (seconds) then flag + trap.  A pipeline bug must not hang the GPU box.
__device__ __forceinline__ void rs_wait(uint32_t bar, uint32_t parity, unsigned *err, int code) {
  uint32_t done = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && it > (1u << 24)) {
      if (err) {
        atomicCAS(err, 0u, (unsigned)code | (parity << 4) | ((bar & 0xFFFu) << 8) | (blockIdx.x << 20));
        __threadfence_system();
      }
      __trap();
    }
  }
}
// Warp-uniform variant for the MMA issuer: the loop lives in one opaque asm statement (a C++ spin loop makes the compiler
// treat everything after it as divergent and the MMA descriptors leave the uniform registers, see cs_tc.cu).
__device__ __forceinline__ void rs_wait_uniform(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
#ifdef RS_DEBUG_WAITS
#define RS_WAIT_MMA(bar, parity, code) rs_wait(bar, parity, P.err, code)
#else
#define RS_WAIT_MMA(bar, parity, code) rs_wait_uniform(bar, parity)
#endif

// Work of a CTA: a contiguous range of (strip, output row) pairs, [ (s0, y0), (s1, y1) ), cut on the host (rs_conv_fwd) and
// passed in the kernel parameters, so that everything the MMA issuer derives from it is provably warp-uniform (constant
// bank + blockIdx): its descriptors then live in uniform registers instead of being converted for every instruction.
struct RsWork {
  int s, y0;                         // current unit: strip, first output row
  int s_end, y_end;                  // end of the range
  int ns0, ns01;                     // strips of group 0, of groups 0 + 1
  int Hout, eqL, polL;               // output rows per strip, positions of the equatorial / of a polar group
};
struct RsUnit {
  int grp, sl, y0, y1, Lg;           // face group, strip within the group, output rows [y0, y1), positions in the group
};
__device__ __forceinline__ bool rs_more(const RsWork &W) { return W.s < W.s_end || (W.s == W.s_end && W.y0 < W.y_end); }
__device__ __forceinline__ RsUnit rs_unit(const RsWork &W) {
  RsUnit u;
  u.y0 = W.y0;
  u.y1 = W.s == W.s_end ? W.y_end : W.Hout;
  if (W.s < W.ns0) { u.grp = 0; u.sl = W.s; u.Lg = W.eqL; }
  else if (W.s < W.ns01) { u.grp = 1; u.sl = W.s - W.ns0; u.Lg = W.polL; }
  else { u.grp = 2; u.sl = W.s - W.ns01; u.Lg = W.polL; }
  return u;
}
__device__ __forceinline__ void rs_advance(RsWork &W, const RsUnit &u) {
  if (u.y1 == W.Hout) { ++W.s; W.y0 = 0; }
  else W.y0 = u.y1;
}
// image index within a group -> (batch element, face)
__device__ __forceinline__ void rs_image(int grp, int i, int &b, int &f) {
  if (grp == 0) { b = i >> 2; f = i & 3; }
  else { b = i; f = 3 + grp; }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
// MODE: 0 plain, 1 fused 1x1 head (section 4.7 of DESIGN.md), 2 second output = 2x2 mean of the output (pooled copy)
template <int KC16T, int MODE>
__global__ void __launch_bounds__(RS_THREADS, 1) conv_rs_kernel(const __grid_constant__ RsP P) {
  constexpr bool HEAD = MODE == 1;
  extern __shared__ uint8_t smem_raw[];
  const RsPlan &L = P.L;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t rows0 = base, wbase = base + L.off_w, misc = base + L.off_misc;
  // barriers (8 bytes each): rfull[16] rempty[16] sfull[16] sempty[16] wfull wempty; tensor-memory slot
  const uint32_t bar_rfull = misc, bar_rempty = misc + 128, bar_sfull = misc + 256, bar_sempty = misc + 384,
                 bar_wfull = misc + 512, bar_wempty = misc + 520, tmem_slot = misc + 528, bar_hfull = misc + 544;
  float *s_bias = reinterpret_cast<float *>(gen + L.off_bias);
  int *s_pix = reinterpret_cast<int *>(gen + L.off_pix);
  const uint32_t stg0 = base + L.off_stg;

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  rs_trace(P, 0, tid == 0);

  RsWork W0;
  W0.s = P.cut_s[blockIdx.x]; W0.y0 = P.cut_y[blockIdx.x];
  W0.s_end = P.cut_s[blockIdx.x + 1]; W0.y_end = P.cut_y[blockIdx.x + 1];
  W0.ns0 = (4 * P.batch * L.Wv + 127) >> 7;
  W0.ns01 = W0.ns0 + ((P.batch * L.Wv + 127) >> 7);
  W0.Hout = P.Hout; W0.eqL = 4 * P.batch * L.Wv; W0.polL = P.batch * L.Wv;

  if (tid == 0) {
    for (int i = 0; i < L.RSn; ++i) {
      mbar_init(bar_rfull + 8 * i, 32);                   // the 32 lanes of the loader warp that owns the row
      mbar_init(bar_rempty + 8 * i, 1);
    }
    for (int i = 0; i < L.NS; ++i) {
      mbar_init(bar_sfull + 8 * i, 1);
      mbar_init(bar_sempty + 8 * i, 4);                   // one arrival per epilogue warp of the row (4 lane quarters)
    }
    mbar_init(bar_wfull, 1);
    mbar_init(bar_wempty, 1);
    mbar_init(bar_hfull, 1);
    mbar_init(bar_hfull + 8, 1);
    fence_mbar_init();
  }
  __syncthreads();                       // the barriers exist: loaders and the weight producer go on from here
  // Tensor memory concerns the MMA issuer and the epilogue warps only (named barrier 5, 288 threads): allocation, then every
  // accumulator slot is cleared -- it is cleared again by the epilogue warp that drains it (tcgen05.st, no shared-memory
  // traffic), so every MMA accumulates and the issue loop has no first-touch special case.  The loaders meanwhile build
  // their tables and wait for the grid dependency.
  uint32_t tmem_base = 0;
  if (warp >= RS_EPI_WARP0) {
    if (warp == RS_MMA_WARP) tmem_alloc(tmem_slot, 512u);
    if (warp == RS_MMA_WARP || warp < RS_EPI_WARP0 + 8) {
      tc_fence_before();
      asm volatile("bar.sync 5, 288;" ::: "memory");
      tc_fence_after();
      tmem_base = *reinterpret_cast<volatile uint32_t *>(gen + (tmem_slot - base));
      if (warp < RS_EPI_WARP0 + 8) {
        const uint32_t lanes = (uint32_t)((warp & 3) * 32) << 16;
        for (int c = ((warp - RS_EPI_WARP0) >> 2) * 16; c < L.NS * L.CoutP; c += 32) tmem_st16_fill(tmem_base + lanes + (uint32_t)c, 0u);
        tmem_st_wait();
      }
      tc_fence_before();
      asm volatile("bar.sync 5, 288;" ::: "memory");
      tc_fence_after();
    }
  }
  // programmatic dependent launch (see cs_tc.cu): whatever a previous kernel may have written -- activations (loaders),
  // packed weights (TMA lane), bias (epilogue) -- is read behind griddepcontrol.wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  rs_trace(P, 1, tid == 0);

  if (warp == RS_TMA_WARP) {
    // ===== weight producer: the whole stacked weight set of a face group, once per group change =====
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int prev = -1, eph = 0;
      for (RsWork W = W0; rs_more(W);) {
        const RsUnit u = rs_unit(W);
        if (u.grp != prev) {
          const bool prev_was_none = prev < 0;
          if (prev >= 0) { rs_wait(bar_wempty, eph, P.err, 1); eph ^= 1; }
          prev = u.grp;
          const uint8_t *wg = P.wpack + (size_t)u.grp * L.groupBytes;
          // the head's weights of all three face groups stay resident (loaded with the first main set): the epilogue's MMAs
          // of a group's last rows may still be in flight when the main weights of the next group arrive
          const bool hload = HEAD && prev_was_none;
          mbar_expect_tx(bar_wfull, (uint32_t)(L.groupBytes + (hload ? 3 * L.hwBytes : 0)));
          for (int o = 0; o < L.groupBytes; o += 32768) {
            const int bytes = min(32768, L.groupBytes - o);
            tma_bulk_g2s(wbase + (uint32_t)o, wg + o, (uint32_t)bytes, bar_wfull);
          }
          if (hload) tma_bulk_g2s(base + (uint32_t)L.off_hw, P.hw, (uint32_t)(3 * L.hwBytes), bar_wfull);
        }
        rs_advance(W, u);
      }
    }
  } else if (warp == RS_MMA_WARP) {
    // ===== MMA issuer: whole-warp uniform loops, one elected lane per instruction =====
    const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 24);
    const uint32_t coutp = (uint32_t)L.CoutP, NS = (uint32_t)L.NS;
    // A: swizzled K-major rows of RB bytes: SBO = 8 rows, LBO unused (1), layout code in bits 61..63, version 1
    const uint64_t a_fix = ((uint64_t)(((uint32_t)(8 * L.RB) >> 4) | (1u << 14) | ((uint32_t)L.layoutType << 29)) << 32) |
                           (1ull << 16);
    // B: un-swizzled K-major core matrices [k8][n][8]: SBO = 128 B (next 8 columns), LBO = NT*16 B (next k8)
    const uint64_t b_fix = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((uint64_t)((uint32_t)L.NT & 0x3FFFu) << 16);
    const uint32_t rb16 = (uint32_t)L.RB >> 4, unit16 = (uint32_t)L.unitBytes >> 4, b_jstep = (uint32_t)(2 * L.NT * 16) >> 4;
    const uint64_t b_w = b_fix | (wbase >> 4);
    // ring positions kept incrementally (no divisions in the row loop): row stage + parity; slot of the output row that is
    // first touched by the current input row (zs, with the parity its "drained" barrier is waited with) and slot of the
    // output row that the current input row completes (cs)
    uint32_t stage = 0, sph = 0, zs = 0, zp = 1, cs = 0;
    int prev = -1, fph = 0;
    bool traced = false;
    for (RsWork W = W0; rs_more(W);) {
      const RsUnit u = rs_unit(W);
      const int H = u.y1 - u.y0, grp = u.grp;
      if (grp != prev) {
        RS_WAIT_MMA(bar_wfull, fph, 4);
        fph ^= 1;
        prev = grp;
      }
      const uint32_t s0 = zs;              // slot of the unit's first output row
      cs = zs;
#pragma unroll 1
      for (int yi = 0; yi < H + 2; ++yi) {
        if (!(P.knock & 16)) RS_WAIT_MMA(bar_rfull + 8 * stage, sph, 5);
        if (yi < H && !(P.knock & 8)) RS_WAIT_MMA(bar_sempty + 8 * zs, zp, 6);   // first touch of this slot: it must have been drained
        tc_fence_after();
        if (P.trace && !traced) { rs_trace(P, 3, lane == 0); traced = true; }
        fence_proxy_async();               // the loaders' generic-proxy writes, acquired through the barrier, before the MMAs' reads
        const uint64_t a_row = a_fix | ((rows0 + stage * (uint32_t)L.stageBytes) >> 4);
        // parts j = 0..2 of the stacked N (kernel rows 2, 1, 0) -> output rows yi - 2 + j, those inside [0, H)
        const int jlo = yi >= 2 ? 0 : 2 - yi, jhi = (H + 1 - yi) < 2 ? (H + 1 - yi) : 2;
        const uint32_t t0 = yi >= 2 ? cs : s0;
        const int np = jhi - jlo + 1;
        const int np1 = (int)(NS - t0) < np ? (int)(NS - t0) : np;
#pragma unroll 1
        for (int seg = 0; seg < 2; ++seg) {
          const int j0 = seg == 0 ? jlo : jlo + np1, cnt = seg == 0 ? np1 : np - np1;
          if (cnt <= 0 || (P.knock & 2)) break;
          const uint32_t d = tmem_base + (seg == 0 ? t0 : 0u) * coutp;
          const uint32_t idesc = idesc0 | ((((uint32_t)cnt * coutp) >> 3) << 17);
          const uint64_t b_seg = b_w + (uint64_t)(((uint32_t)j0 * coutp * 16u) >> 4);
#pragma unroll
          for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int j = 0; j < KC16T; ++j)
              umma_bf16_elect(d, a_row + (uint64_t)((uint32_t)v * rb16 + (uint32_t)j * 2u),
                              b_seg + (uint64_t)((uint32_t)v * unit16 + (uint32_t)j * b_jstep), idesc, 1u);
        }
        if (!(P.knock & 16)) umma_commit_elect(bar_rempty + 8 * stage);                      // the row stage may be refilled
        if (++stage == (uint32_t)L.RSn) { stage = 0; sph ^= 1u; }
        if (yi >= 2) {                                                  // output row yi - 2 is complete
          if (!(P.knock & 8)) umma_commit_elect(bar_sfull + 8 * cs);
          if (++cs == NS) cs = 0;
        }
        if (yi < H && ++zs == NS) { zs = 0; zp ^= 1u; }
      }
      rs_advance(W, u);
      if (rs_more(W) && rs_unit(W).grp != grp) umma_commit_elect(bar_wempty);
    }
  } else if (warp >= RS_EPI_WARP0 && warp < RS_EPI_WARP0 + 8) {
    // ===== epilogue: slot -> registers -> bias / activation -> bf16 -> per-warp compaction -> coalesced stores.  Warp w
    // reads TMEM lanes 32*(w%4)..+31; the two warps of a lane quarter take alternate output rows. =====
    const int quarter = warp & 3, half = (warp - RS_EPI_WARP0) >> 2;
    const uint32_t stg = stg0 + (uint32_t)(warp - RS_EPI_WARP0) * L.stgBytes;
    int *pixw = s_pix + (warp - RS_EPI_WARP0) * 32;
    const uint32_t rowB = (uint32_t)P.cout * 2u;
    const bool act_fast = P.act == DLWPCS_ACT_CAPPED_LEAKY_RELU && P.slope >= 0.f && P.slope <= 1.f;
    const uint32_t cpr = rowB >> 4;
    const int cprLog = (cpr & (cpr - 1)) == 0 ? 31 - __clz(cpr) : -1;
    const uint32_t smask = (cpr & (cpr - 1)) == 0 ? min(cpr, 8u) - 1u : 0u;
    // XOR key of a staged row: rows shorter than 128 bytes share a 128-byte bank line, so the key advances once per line
    // (8 consecutive rows then hit 8 different 16-byte bank groups; keyed by the row itself, 64-byte rows collide two-way)
    const uint32_t kshift = rowB >= 128u ? 0u : (rowB == 64u ? 1u : (rowB == 32u ? 2u : 3u));
    const uint32_t NS = (uint32_t)L.NS;
    asm volatile("griddepcontrol.wait;" ::: "memory");        // the bias sits behind the packed weights
    for (int i = tid - RS_EPI_WARP0 * 32; i < 3 * L.CoutP; i += RS_EPI) s_bias[i] = P.bias[i];
    asm volatile("bar.sync 2, %0;" ::"n"(RS_EPI) : "memory");
    int nunits = 0;
    bool etraced = false;
    if constexpr (HEAD) {
    // ---- fused 1x1 head: the activated bf16 row is written as the A tile of a second MMA (K = CoutP, N = CoutP2) against
    // the resident head weights; its accumulators are read back one turn later, while the tensor core works on the next row
    const uint32_t RBh = (uint32_t)L.CoutP * 2u, hmask = (RBh >> 4) - 1u;
    const uint32_t a2 = base + (uint32_t)L.off_a2 + (uint32_t)half * (128u * RBh);
    const float *s_hbias = reinterpret_cast<const float *>(gen + L.off_hbias);
    const uint32_t d2 = tmem_base + (uint32_t)(L.NS * L.CoutP + half * L.CoutP2);
    const uint32_t bar_h = bar_hfull + 8u * (uint32_t)half;
    const uint32_t rowB2 = (uint32_t)P.cout2 * 2u, cpr2 = rowB2 >> 4;
    const int cprLog2 = (cpr2 & (cpr2 - 1)) == 0 ? 31 - __clz(cpr2) : -1;
    const uint32_t smask2 = (cpr2 & (cpr2 - 1)) == 0 ? min(cpr2, 8u) - 1u : 0u;
    const uint32_t kshift2 = rowB2 >= 128u ? 0u : (rowB2 == 64u ? 1u : (rowB2 == 32u ? 2u : 3u));
    const uint64_t a2desc = (((uint64_t)(((8u * RBh) >> 4) | (1u << 14) | ((RBh == 128u ? 2u : 4u) << 29)) << 32) | (1ull << 16)) |
                            (a2 >> 4);
    const uint64_t b2fix = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((uint64_t)((uint32_t)L.CoutP2 & 0x3FFFu) << 16);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 24) | (((uint32_t)L.CoutP2 >> 3) << 17);
    const bool act2_fast = P.act2 == DLWPCS_ACT_CAPPED_LEAKY_RELU && P.slope2 >= 0.f && P.slope2 <= 1.f;
    for (int i = tid - RS_EPI_WARP0 * 32; i < 3 * L.CoutP2; i += RS_EPI) const_cast<float *>(s_hbias)[i] = P.hbias[i];
    asm volatile("bar.sync 2, %0;" ::"n"(RS_EPI) : "memory");
    uint32_t slot = 0, eph = 0, odd = 0, hph = 0;
    for (RsWork W = W0; rs_more(W) && !(P.knock & 8);) {
      const RsUnit u = rs_unit(W);
      const int H = u.y1 - u.y0;
      const float *bias = s_bias + u.grp * L.CoutP;
      const float *bias2 = s_hbias + u.grp * L.CoutP2;
      const uint64_t b2desc = b2fix | ((base + (uint32_t)L.off_hw + (uint32_t)(u.grp * L.hwBytes)) >> 4);
      const int p = u.sl * 128 + quarter * 32 + lane;
      const int img = p / L.Wv, cv = p - img * L.Wv;
      const bool ok = p < u.Lg && cv < P.Wout;
      int b, f;
      rs_image(u.grp, img, b, f);
      const int opix0 = ((b * 6 + f) * P.Hout + u.y0) * P.Wout + cv;
      const unsigned okmask = __ballot_sync(0xffffffffu, ok);
      const int nvalid = __popc(okmask);
      const uint32_t prow = (uint32_t)__popc(okmask & ((1u << lane) - 1u));
      __syncwarp();
      if (ok) pixw[prow] = opix0;
      __syncwarp();
      const uint32_t srow = stg + prow * rowB2;
      const uint32_t total = (uint32_t)nvalid * rowB2;
      const uint32_t arow = a2 + (uint32_t)(quarter * 32 + lane) * RBh;
      const uint32_t aswz = (((uint32_t)(quarter * 32 + lane) * RBh) >> 7) & hmask;
      int pend = -1;                          // output row (within the unit) whose head accumulators are still in tensor memory
      // read the head accumulators of row `po`, finish (bias, activation, bf16), compact and store
      auto finish = [&](int po) {
        rs_wait(bar_h, hph, P.err, 7);
        hph ^= 1u;
        tc_fence_after();
        for (int n0 = 0; n0 < L.CoutP2; n0 += 16) {
          uint32_t v[16];
          tmem_ld16(d2 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)n0, v);
          tmem_ld_wait();
          if (!ok) continue;
          float r[16];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 bv = *reinterpret_cast<const float4 *>(bias2 + n0 + 4 * k4);
            r[4 * k4 + 0] = __uint_as_float(v[4 * k4 + 0]) + bv.x;
            r[4 * k4 + 1] = __uint_as_float(v[4 * k4 + 1]) + bv.y;
            r[4 * k4 + 2] = __uint_as_float(v[4 * k4 + 2]) + bv.z;
            r[4 * k4 + 3] = __uint_as_float(v[4 * k4 + 3]) + bv.w;
          }
          if (act2_fast) {
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = fminf(fmaxf(r[e], P.slope2 * r[e]), P.maxv2);
          } else if (P.act2 != DLWPCS_ACT_NONE) {
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = act_apply(r[e], P.act2, P.slope2, P.maxv2);
          }
          if (n0 + 8 <= P.cout2)
            st_shared16(srow + ((((uint32_t)n0 >> 3) ^ ((prow >> kshift2) & smask2)) << 4),
                        make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7])));
          if (n0 + 16 <= P.cout2)
            st_shared16(srow + (((((uint32_t)n0 >> 3) + 1u) ^ ((prow >> kshift2) & smask2)) << 4),
                        make_uint4(pack_bf16x2(r[8], r[9]), pack_bf16x2(r[10], r[11]), pack_bf16x2(r[12], r[13]), pack_bf16x2(r[14], r[15])));
        }
        tc_fence_before();
        __syncwarp();
        const uint32_t rowoff = (uint32_t)(po * P.Wout);
        for (uint32_t off = (uint32_t)lane * 16u; off < total; off += 512u) {
          uint32_t row, ch;
          if (cprLog2 >= 0) { row = off >> (4 + cprLog2); ch = (off >> 4) & (cpr2 - 1u); }
          else { row = off / rowB2; ch = (off - row * rowB2) >> 4; }
          uint4 q;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                       : "r"(stg + row * rowB2 + ((ch ^ ((row >> kshift2) & smask2)) << 4)));
          uint8_t *gdst = reinterpret_cast<uint8_t *>(P.y2) + ((size_t)((uint32_t)pixw[row] + rowoff)) * rowB2 + (ch << 4);
          *reinterpret_cast<uint4 *>(gdst) = q;
        }
        __syncwarp();
      };
#pragma unroll 1
      for (int o = 0; o < H; ++o) {
        const uint32_t slot_o = slot, eph_o = eph, mine = ((int)odd == half);
        odd ^= 1u;
        if (++slot == NS) { slot = 0; eph ^= 1u; }
        if (!mine) continue;
        if (pend >= 0) finish(pend);          // frees this half's A tile and head accumulators
        rs_wait(bar_sfull + 8 * slot_o, eph_o, P.err, 2);
        tc_fence_after();
        if (P.trace && !etraced) { rs_trace(P, 4, tid == RS_EPI_WARP0 * 32); etraced = true; }
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot_o * (uint32_t)L.CoutP;
        for (int n0 = 0; n0 < L.CoutP; n0 += 32) {
          uint32_t v[32];
          const bool two = n0 + 16 < L.CoutP;
          tmem_ld16(trow + (uint32_t)n0, v);
          if (two) tmem_ld16(trow + (uint32_t)n0 + 16u, v + 16);
          tmem_ld_wait();
          tmem_st16_fill(trow + (uint32_t)n0, 0u);
          if (two) tmem_st16_fill(trow + (uint32_t)n0 + 16u, 0u);
          if (n0 + 32 >= L.CoutP) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sempty + 8 * slot_o);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const int nb = n0 + 16 * h;
            float r[16];
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const float4 bv = *reinterpret_cast<const float4 *>(bias + nb + 4 * k4);
              r[4 * k4 + 0] = __uint_as_float(v[16 * h + 4 * k4 + 0]) + bv.x;
              r[4 * k4 + 1] = __uint_as_float(v[16 * h + 4 * k4 + 1]) + bv.y;
              r[4 * k4 + 2] = __uint_as_float(v[16 * h + 4 * k4 + 2]) + bv.z;
              r[4 * k4 + 3] = __uint_as_float(v[16 * h + 4 * k4 + 3]) + bv.w;
            }
            if (act_fast) {
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = fminf(fmaxf(r[e], P.slope * r[e]), P.maxv);
            } else if (P.act != DLWPCS_ACT_NONE) {
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = act_apply(r[e], P.act, P.slope, P.maxv);
            }
            // every lane writes its row of the A tile (rows of dropped lanes only feed their own, dropped, outputs)
            st_shared16(arow + ((((uint32_t)nb >> 3) ^ aswz) << 4),
                        make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7])));
            st_shared16(arow + (((((uint32_t)nb >> 3) + 1u) ^ aswz) << 4),
                        make_uint4(pack_bf16x2(r[8], r[9]), pack_bf16x2(r[10], r[11]), pack_bf16x2(r[12], r[13]), pack_bf16x2(r[14], r[15])));
          }
        }
        fence_proxy_async();
        tc_fence_before();
        if (half == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
        else asm volatile("bar.sync 4, 128;" ::: "memory");
        if (quarter == 0) {
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j * 16 < L.Kh)
              umma_bf16_elect(d2, a2desc + (uint64_t)(2u * (uint32_t)j), b2desc + (uint64_t)((uint32_t)j * (uint32_t)(2 * L.CoutP2)), idesc2,
                              j > 0 ? 1u : 0u);
          umma_commit_elect(bar_h);
        }
        pend = o;
      }
      if (pend >= 0) finish(pend);
      rs_advance(W, u);
      ++nunits;
    }
    } else if constexpr (MODE == 2) {
    // ---- pool mode (32 padded output channels): besides the layer's output, its 2x2 mean (AveragePooling3D((1,2,2)),
    // train_cs.py:197) is written as a second tensor, so that the next layer reads a plain source (a quarter of the bytes, and
    // through asynchronous copies instead of registers).  The two warps of a lane quarter take alternate ROW PAIRS and keep
    // both staged rows (two staging buffers per warp); lanes 2i, 2i+1 hold horizontally adjacent pixels (Wv is even), valid
    // lanes come in such pairs, so the pooled pixel j is the mean of the compacted rows 2j, 2j+1 of both buffers -- taken
    // over the bf16-rounded outputs in the order (r,c), (r,c+1), (r+1,c), (r+1,c+1) and rounded to bf16 once, like the
    // classic kernel's pooled load.
    int *pixp = s_pix + 256 + (warp - RS_EPI_WARP0) * 32;
    const int Hp = P.Hout >> 1, Wp = P.Wout >> 1;
    const uint32_t stgA = stg0 + (uint32_t)(2 * (warp - RS_EPI_WARP0)) * L.stgBytes, stgB = stgA + (uint32_t)L.stgBytes;
    uint32_t slot = 0, eph = 0, nrow = 0;
    for (RsWork W = W0; rs_more(W) && !(P.knock & 8);) {
      const RsUnit u = rs_unit(W);
      const int H = u.y1 - u.y0;                  // even, and u.y0 is even (host cuts)
      const float *bias = s_bias + u.grp * L.CoutP;
      const int p = u.sl * 128 + quarter * 32 + lane;
      const int img = p / L.Wv, cv = p - img * L.Wv;
      const bool ok = p < u.Lg && cv < P.Wout;
      int b, f;
      rs_image(u.grp, img, b, f);
      const int opix0 = ((b * 6 + f) * P.Hout + u.y0) * P.Wout + cv;
      const unsigned okmask = __ballot_sync(0xffffffffu, ok);
      const int nvalid = __popc(okmask);
      const uint32_t prow = (uint32_t)__popc(okmask & ((1u << lane) - 1u));
      __syncwarp();
      if (ok) pixw[prow] = opix0;
      if (ok && !(lane & 1)) pixp[prow >> 1] = ((b * 6 + f) * Hp + (u.y0 >> 1)) * Wp + (cv >> 1);
      __syncwarp();
      const uint32_t total = (uint32_t)nvalid * rowB;
      const uint32_t items = (uint32_t)(nvalid >> 1) * cpr;          // 16-byte chunks of the pooled row
#pragma unroll 1
      for (int o = 0; o < H; ++o) {
        const uint32_t slot_o = slot, eph_o = eph, mine = ((int)((nrow >> 1) & 1u) == half);
        ++nrow;
        if (++slot == NS) { slot = 0; eph ^= 1u; }
        if (!mine) continue;
        rs_wait(bar_sfull + 8 * slot_o, eph_o, P.err, 2);
        tc_fence_after();
        if (P.trace && !etraced) { rs_trace(P, 4, tid == RS_EPI_WARP0 * 32); etraced = true; }
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot_o * (uint32_t)L.CoutP;
        const uint32_t sb = (o & 1) ? stgB : stgA, srow = sb + prow * rowB;
        uint32_t v[32];
        tmem_ld16(trow, v);
        tmem_ld16(trow + 16u, v + 16);
        tmem_ld_wait();
        tmem_st16_fill(trow, 0u);
        tmem_st16_fill(trow + 16u, 0u);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sempty + 8 * slot_o);
        if (ok) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float r[16];
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const float4 bv = *reinterpret_cast<const float4 *>(bias + 16 * h + 4 * k4);
              r[4 * k4 + 0] = __uint_as_float(v[16 * h + 4 * k4 + 0]) + bv.x;
              r[4 * k4 + 1] = __uint_as_float(v[16 * h + 4 * k4 + 1]) + bv.y;
              r[4 * k4 + 2] = __uint_as_float(v[16 * h + 4 * k4 + 2]) + bv.z;
              r[4 * k4 + 3] = __uint_as_float(v[16 * h + 4 * k4 + 3]) + bv.w;
            }
            if (act_fast) {
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = fminf(fmaxf(r[e], P.slope * r[e]), P.maxv);
            } else if (P.act != DLWPCS_ACT_NONE) {
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = act_apply(r[e], P.act, P.slope, P.maxv);
            }
            const int nb = 16 * h;
            if (nb + 8 <= P.cout)
              st_shared16(srow + ((((uint32_t)nb >> 3) ^ ((prow >> kshift) & smask)) << 4),
                          make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7])));
            if (nb + 16 <= P.cout)
              st_shared16(srow + (((((uint32_t)nb >> 3) + 1u) ^ ((prow >> kshift) & smask)) << 4),
                          make_uint4(pack_bf16x2(r[8], r[9]), pack_bf16x2(r[10], r[11]), pack_bf16x2(r[12], r[13]), pack_bf16x2(r[14], r[15])));
          }
        }
        __syncwarp();
        {
          const uint32_t rowoff = (uint32_t)(o * P.Wout);
          for (uint32_t off = (uint32_t)lane * 16u; off < total; off += 512u) {
            uint32_t row, ch;
            if (cprLog >= 0) { row = off >> (4 + cprLog); ch = (off >> 4) & (cpr - 1u); }
            else { row = off / rowB; ch = (off - row * rowB) >> 4; }
            uint4 q;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                         : "r"(sb + row * rowB + ((ch ^ ((row >> kshift) & smask)) << 4)));
            uint8_t *gdst = reinterpret_cast<uint8_t *>(P.y) + ((size_t)((uint32_t)pixw[row] + rowoff)) * rowB + (ch << 4);
            *reinterpret_cast<uint4 *>(gdst) = q;
          }
        }
        if (!(o & 1)) continue;           // (the other buffer is written next; this one is read again below)
        // second row of the pair: one 16-byte chunk (8 channels) of one pooled pixel per item, straight to HBM
        const uint32_t rowoffp = (uint32_t)((o >> 1) * Wp);
        for (uint32_t it = (uint32_t)lane; it < items; it += 32u) {
          uint32_t j, ch;
          if (cprLog >= 0) { j = it >> cprLog; ch = it & (cpr - 1u); }
          else { j = it / cpr; ch = it - j * cpr; }
          const uint32_t r0 = 2u * j, r1 = r0 + 1u;
          const uint32_t o0 = r0 * rowB + ((ch ^ ((r0 >> kshift) & smask)) << 4), o1 = r1 * rowB + ((ch ^ ((r1 >> kshift) & smask)) << 4);
          uint4 q[4];
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q[0].x), "=r"(q[0].y), "=r"(q[0].z), "=r"(q[0].w) : "r"(stgA + o0));
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q[1].x), "=r"(q[1].y), "=r"(q[1].z), "=r"(q[1].w) : "r"(stgA + o1));
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q[2].x), "=r"(q[2].y), "=r"(q[2].z), "=r"(q[2].w) : "r"(stgB + o0));
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q[3].x), "=r"(q[3].y), "=r"(q[3].z), "=r"(q[3].w) : "r"(stgB + o1));
          float a[8], t[8];
          unpack_bf16x8(q[0], a);
#pragma unroll
          for (int s4 = 1; s4 < 4; ++s4) {
            unpack_bf16x8(q[s4], t);
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] += t[e];
          }
          const uint4 pl = make_uint4(pack_bf16x2(0.25f * a[0], 0.25f * a[1]), pack_bf16x2(0.25f * a[2], 0.25f * a[3]),
                                      pack_bf16x2(0.25f * a[4], 0.25f * a[5]), pack_bf16x2(0.25f * a[6], 0.25f * a[7]));
          uint8_t *gdst = reinterpret_cast<uint8_t *>(P.ypool) + ((size_t)((uint32_t)pixp[j] + rowoffp)) * rowB + (ch << 4);
          *reinterpret_cast<uint4 *>(gdst) = pl;
        }
        __syncwarp();                     // both buffers are rewritten by this warp's next pair
      }
      rs_advance(W, u);
      ++nunits;
    }
    } else {
    uint32_t slot = 0, eph = 0, odd = 0;      // slot / barrier parity / row parity of the next output row of this CTA
    for (RsWork W = W0; rs_more(W) && !(P.knock & 8);) {
      const RsUnit u = rs_unit(W);
      const int H = u.y1 - u.y0;
      const float *bias = s_bias + u.grp * L.CoutP;
      // this lane's position: image and column
      const int p = u.sl * 128 + quarter * 32 + lane;
      const int img = p / L.Wv, cv = p - img * L.Wv;
      const bool ok = p < u.Lg && cv < P.Wout;
      int b, f;
      rs_image(u.grp, img, b, f);
      const int opix0 = ((b * 6 + f) * P.Hout + u.y0) * P.Wout + cv;       // output pixel of this lane in the unit's first row
      const unsigned okmask = __ballot_sync(0xffffffffu, ok);
      const int nvalid = __popc(okmask);
      const uint32_t prow = (uint32_t)__popc(okmask & ((1u << lane) - 1u));   // compacted row of this lane
      __syncwarp();
      if (ok) pixw[prow] = opix0;
      __syncwarp();
      const uint32_t srow = stg + prow * rowB;
      const uint32_t total = (uint32_t)nvalid * rowB;
#pragma unroll 1
      for (int o = 0; o < H; ++o) {
        const uint32_t slot_o = slot, eph_o = eph, mine = ((int)odd == half);
        odd ^= 1u;
        if (++slot == NS) { slot = 0; eph ^= 1u; }
        if (!mine) continue;
        rs_wait(bar_sfull + 8 * slot_o, eph_o, P.err, 2);
        tc_fence_after();
        if (P.trace && !etraced) { rs_trace(P, 4, tid == RS_EPI_WARP0 * 32); etraced = true; }
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot_o * (uint32_t)L.CoutP;
        for (int n0 = 0; n0 < L.CoutP; n0 += 32) {
          uint32_t v[32];
          const bool two = n0 + 16 < L.CoutP;
          tmem_ld16(trow + (uint32_t)n0, v);
          if (two) tmem_ld16(trow + (uint32_t)n0 + 16u, v + 16);
          tmem_ld_wait();
          tmem_st16_fill(trow + (uint32_t)n0, 0u);          // clear what was read (the next use of the slot accumulates)
          if (two) tmem_st16_fill(trow + (uint32_t)n0 + 16u, 0u);
          if (n0 + 32 >= L.CoutP) {          // everything of the slot is in registers: hand it back to the MMA issuer
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sempty + 8 * slot_o);
          }
          if (!ok || (P.knock & 4)) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const int nb = n0 + 16 * h;
            float r[16];
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const float4 bv = *reinterpret_cast<const float4 *>(bias + nb + 4 * k4);
              r[4 * k4 + 0] = __uint_as_float(v[16 * h + 4 * k4 + 0]) + bv.x;
              r[4 * k4 + 1] = __uint_as_float(v[16 * h + 4 * k4 + 1]) + bv.y;
              r[4 * k4 + 2] = __uint_as_float(v[16 * h + 4 * k4 + 2]) + bv.z;
              r[4 * k4 + 3] = __uint_as_float(v[16 * h + 4 * k4 + 3]) + bv.w;
            }
            if (act_fast) {
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = fminf(fmaxf(r[e], P.slope * r[e]), P.maxv);
            } else if (P.act != DLWPCS_ACT_NONE) {
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = act_apply(r[e], P.act, P.slope, P.maxv);
            }
            if (nb + 8 <= P.cout)
              st_shared16(srow + ((((uint32_t)nb >> 3) ^ ((prow >> kshift) & smask)) << 4),
                          make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7])));
            if (nb + 16 <= P.cout)
              st_shared16(srow + (((((uint32_t)nb >> 3) + 1u) ^ ((prow >> kshift) & smask)) << 4),
                          make_uint4(pack_bf16x2(r[8], r[9]), pack_bf16x2(r[10], r[11]), pack_bf16x2(r[12], r[13]), pack_bf16x2(r[14], r[15])));
          }
        }
        __syncwarp();
        // rows of one image are contiguous in HBM; consecutive lanes copy consecutive 16-byte chunks
        const uint32_t rowoff = (uint32_t)(o * P.Wout);
        for (uint32_t off = (uint32_t)lane * 16u; off < total && !(P.knock & 4); off += 512u) {
          uint32_t row, ch;
          if (cprLog >= 0) { row = off >> (4 + cprLog); ch = (off >> 4) & (cpr - 1u); }
          else { row = off / rowB; ch = (off - row * rowB) >> 4; }
          uint4 q;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                       : "r"(stg + row * rowB + ((ch ^ ((row >> kshift) & smask)) << 4)));
          uint8_t *gdst = reinterpret_cast<uint8_t *>(P.y) + ((size_t)((uint32_t)pixw[row] + rowoff)) * rowB + (ch << 4);
          *reinterpret_cast<uint4 *>(gdst) = q;
        }
        __syncwarp();                                 // the staging buffer is rewritten by this warp's next row
      }
      rs_advance(W, u);
      ++nunits;
    }
    }
    rs_trace(P, 5, tid == RS_EPI_WARP0 * 32);
    if (P.trace && tid == RS_EPI_WARP0 * 32) P.trace[blockIdx.x * 8 + 7] = (unsigned long long)nunits;
  } else if (warp < 8) {
    // ===== loaders: loader warp w gathers the input rows m = w (mod NW) of this CTA's row sequence, one whole row (130
    // positions x Cin) per turn: consecutive lanes copy consecutive 16-byte chunks of consecutive positions.  Every lane owns
    // a fixed channel chunk (32 % S == 0), i.e. a fixed source tensor / sampling mode.  NW = min(8, stages): with fewer than
    // 8 stages every stage belongs to ONE warp, so a waiter is never more than one barrier phase ahead (parity waits). =====
    const int NW = L.RSn < 8 ? L.RSn : 8;
    const int chunk = lane & (L.S - 1), c = chunk * 8;
    const bool first = c < P.c0;
    const __nv_bfloat16 *src = first ? P.x0 : P.x1;
    const int C = first ? P.c0 : P.c1, cc = first ? c : c - P.c0, mode = first ? P.mode0 : P.mode1;
    const int ppb = first ? P.ppb0 : P.ppb1;
    const bool chan_ok = c < P.cin;
    const int pstep = 32 >> L.logS, l0 = lane >> L.logS;
    const uint32_t cw = (uint32_t)chunk & ((1u << L.cprLog) - 1u);
    const bool pooled = mode == DLWPCS_SRC_POOL2;
    const int w2 = P.n * 2;
    int2 *ptab = reinterpret_cast<int2 *>(gen + L.off_pos) + warp * RS_NPIXP;       // this warp's copy: {table offset, batch element}
    // The patch-table entries of a row (source pixel of every position, for each of the two sources' sampling modes) are
    // prefetched into shared memory one turn ahead with 4-byte asynchronous copies: no L2 round trip in front of the gathers.
    const bool two_tabs = P.tab1 != P.tab0;
    const uint32_t pxb = base + (uint32_t)L.off_px + (uint32_t)warp * (4u * RS_NPIXP * 4u);     // [2 buffers][2 tables][136]
    const int *pxg = reinterpret_cast<const int *>(gen + L.off_px) + warp * (4 * RS_NPIXP);
    const int mysel = (first || !two_tabs) ? 0 : RS_NPIXP;
    auto prefetch = [&](int y, int bufi) {
      for (int l = lane; l < RS_NPOS; l += 32) {
        const int to = ptab[l].x;
        const uint32_t dst = pxb + (uint32_t)(bufi * 2 * RS_NPIXP + l) * 4u;
        if (to >= 0) {
          cp_async4(dst, P.tab0 + to + y * L.Wv);
          if (two_tabs) cp_async4(dst + RS_NPIXP * 4u, P.tab1 + to + y * L.Wv);
        } else {
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(-1) : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + RS_NPIXP * 4u), "r"(-1) : "memory");
        }
      }
      cp_async_commit();
    };
    bool waited = false;                  // griddepcontrol.wait sits behind the first table prefetch (the tables are static)
    uint32_t m = 0;                       // rows of this CTA's sequence before the current unit
    for (RsWork W = W0; rs_more(W) && !(P.knock & 16) && warp < NW;) {
      const RsUnit u = rs_unit(W);
      const int H = u.y1 - u.y0;
      // first row of the unit that is this warp's
      int yi = (warp + NW - (int)(m % (uint32_t)NW)) % NW;
      int cur = 0;
      if (yi < H + 2) {
        __syncwarp();
        for (int l = lane; l < RS_NPOS; l += 32) {
          const int p = u.sl * 128 + l;
          const int img = p / L.Wv, cv = p - img * L.Wv;
          int b, f;
          rs_image(u.grp, img, b, f);
          ptab[l] = make_int2(p < u.Lg ? f * L.G + cv : -1, b);
        }
        __syncwarp();
        prefetch(u.y0 + yi, 0);
      }
      if (!waited) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        rs_trace(P, 2, tid == 0);
        waited = true;
      }
#pragma unroll 1
      for (; yi < H + 2; yi += NW) {
        const uint32_t mm = m + (uint32_t)yi, stage = mm % (uint32_t)L.RSn, sph = (mm / (uint32_t)L.RSn) & 1u;
        __syncwarp();                                     // everybody has finished reading the buffer that is refilled now
        if (yi + NW < H + 2) prefetch(u.y0 + yi + NW, cur ^ 1);
        else cp_async_commit();
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        rs_wait(bar_rempty + 8 * stage, sph ^ 1u, P.err, 3);
        const uint32_t sbase = rows0 + stage * (uint32_t)L.stageBytes;
        const int *pxrow = pxg + cur * 2 * RS_NPIXP + mysel;
        cur ^= 1;
        if (!(P.knock & 1)) {
#pragma unroll 4
          for (int l = l0; l < RS_NPOS; l += pstep) {
            const int px = chan_ok ? pxrow[l] : -1;
            const uint32_t ro = (uint32_t)l * (uint32_t)L.RB;
            const uint32_t dst = sbase + ro + ((cw ^ ((ro >> 7) & (uint32_t)L.swzMask)) << 4);
            if (!pooled) {
              const __nv_bfloat16 *g = px >= 0 ? src + ((size_t)(ptab[l].y * ppb + px) * C + cc) : P.x0;
              cp_async16(dst, g, px >= 0 ? 16u : 0u);
            } else {
              // 2x2 mean (AveragePooling3D((1,2,2)), train_cs.py:197) through registers; rounded to bf16 once, like a
              // stored pooled tensor
              uint4 o4 = make_uint4(0, 0, 0, 0);
              if (px >= 0) {
                const __nv_bfloat16 *g = src + ((size_t)(ptab[l].y * ppb + px) * C + cc);
                const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(g));
                const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(g + C));
                const uint4 v2 = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)w2 * C));
                const uint4 v3 = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)(w2 + 1) * C));
                float a[8], t[8];
                unpack_bf16x8(v0, a);
                unpack_bf16x8(v1, t);
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] += t[q];
                unpack_bf16x8(v2, t);
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] += t[q];
                unpack_bf16x8(v3, t);
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] += t[q];
                o4 = make_uint4(pack_bf16x2(0.25f * a[0], 0.25f * a[1]), pack_bf16x2(0.25f * a[2], 0.25f * a[3]),
                                pack_bf16x2(0.25f * a[4], 0.25f * a[5]), pack_bf16x2(0.25f * a[6], 0.25f * a[7]));
              }
              st_shared16(dst, o4);
            }
          }
        }
        // the row is published when this lane's copies have landed (register-path stores are already visible)
        fence_proxy_async();
        cp_async_mbar_arrive(bar_rfull + 8 * stage);
      }
      m += (uint32_t)(H + 2);
      rs_advance(W, u);
    }
  }
  tc_fence_before();
  __syncthreads();
  rs_trace(P, 6, tid == 0);
  if (warp == RS_MMA_WARP) tmem_dealloc(tmem_base, 512u);
}

int rs_env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

// nullptr when the row-streamed kernel serves the layer, otherwise the reason it does not
const char *rs_make_plan(const dlwpcs_conv_desc *d, const Geometry &g, RsPlan *L, const dlwpcs_conv_desc *dh = nullptr,
                         bool pool = false) {
  if (d->kh != 3 || d->kw != 3) return "3x3 kernels only";
  if (d->stride_h != 1 || d->stride_w != 1 || d->dil_h != 1 || d->dil_w != 1) return "stride / dilation 1 only";
  if (d->x_dtype != DLWPCS_BF16 || d->y_dtype != DLWPCS_BF16) return "bf16 in, bf16 out";
  if (d->cout % 8 || d->c0 % 8 || d->c1 % 8) return "channel counts must be multiples of 8";
  // a 2x2-pooled source goes through registers (four loads + mean per chunk); with one warp per row too few bytes are in
  // flight -- the classic kernel's 256-thread gather is faster (measured: 34.8 vs 42 us for 32 -> 64 at 24 x 24, batch 64)
  if ((d->mode0 == DLWPCS_SRC_POOL2 || (d->c1 > 0 && d->mode1 == DLWPCS_SRC_POOL2)) && !rs_env_int("DLWPCS_RS_POOL", 0))
    return "pooled sources stay on the classic kernel";
  L->CinP = (d->cin + 15) / 16 * 16;
  if (L->CinP > 32) L->CinP = (d->cin + 63) / 64 * 64;
  if (L->CinP > 64) return "more than 64 input channels";
  L->KC16 = L->CinP / 16;
  L->S = L->CinP / 8;
  L->logS = L->S == 2 ? 1 : (L->S == 4 ? 2 : 3);
  L->RB = L->CinP * 2;
  L->cprLog = L->RB == 128 ? 3 : (L->RB == 64 ? 2 : 1);
  L->swzMask = (1 << L->cprLog) - 1;
  L->layoutType = L->RB == 128 ? 2 : (L->RB == 64 ? 4 : 6);
  L->CoutP = (d->cout + 15) / 16 * 16;
  if (L->CoutP < 32) L->CoutP = 32;
  L->NT = 3 * L->CoutP;
  if (L->NT > 256) return "more than 80 output channels";
  L->head = dh ? 1 : 0;
  L->CoutP2 = 0;
  L->Kh = 0;
  L->hwBytes = 0;
  if (dh) {
    // fused 1x1 head: CubeSphereConv2D(cout2, 1) on this layer's output, which is then never written
    if (dh->kh != 1 || dh->kw != 1 || dh->stride_h != 1 || dh->stride_w != 1 || dh->halo != 0 || dh->same) return "head must be a plain 1x1 convolution";
    if (dh->cin != d->cout || dh->c1 != 0 || dh->mode0 != DLWPCS_SRC_SAME) return "head must read this layer's output and nothing else";
    if (dh->batch != d->batch || dh->n != g.Hout || g.Hout != g.Wout) return "head geometry differs";
    if (dh->x_dtype != DLWPCS_BF16 || dh->y_dtype != DLWPCS_BF16 || dh->cout % 8) return "head: bf16, output channels a multiple of 8";
    if (L->CoutP != 32 && L->CoutP != 64) return "head needs 32 or 64 (padded) channels in between";
    L->CoutP2 = (dh->cout + 15) / 16 * 16;
    if (L->CoutP2 > 64) return "head with more than 64 output channels";
    // the head's classic image pads its input channels like cs_tc.cu's make_plan: 16 / 32, then multiples of 64
    L->Kh = (dh->cin + 15) / 16 * 16;
    if (L->Kh > 32) L->Kh = (dh->cin + 63) / 64 * 64;
    if (L->Kh > L->CoutP) return "head input channels exceed the padded channels in between";
    L->hwBytes = L->Kh * L->CoutP2 * 2;
  }
  if (pool) {
    // second output = 2x2 mean: lane pairs are pixel pairs only for even widths, row pairs stay on one epilogue warp
    if (dh) return "pool mode and a fused head exclude each other";
    if (L->CoutP != 32) return "pool mode serves layers with up to 32 output channels";
    if ((g.Hout & 1) || (g.Wout & 1)) return "pool mode needs even face edges";
  }
  L->NS = (512 - 2 * L->CoutP2) / L->CoutP;
  L->NS &= ~1;                                       // the two epilogue halves alternate rows: an even ring keeps slot <-> half fixed
  if (L->NS > RS_MAXSLOTS) L->NS = RS_MAXSLOTS;
  L->Wv = g.Wout + 2;
  L->Hv = g.Hout + 2;
  L->G = (L->Hv * L->Wv + 3) / 4 * 4;
  L->stageBytes = (RS_NPIXP * L->RB + 1023) / 1024 * 1024;
  L->unitBytes = L->CinP * L->NT * 2;
  L->groupBytes = 3 * L->unitBytes;
  L->stgBytes = 32 * (dh ? dh->cout : d->cout) * 2;
  const int stgPerWarp = pool ? 2 : 1;              // pool mode keeps both rows of a pair staged
  int off = 0;
  L->off_w = 0;        // filled below, after the row stages
  const int wB = (L->groupBytes + 1023) / 1024 * 1024;
  const int zeroB = 0;
  // per loader warp: {table offset, batch element} of the strip's positions + two double-buffered rows of table entries
  const int posB = 8 * RS_NPIXP * 8 + 8 * 4 * RS_NPIXP * 4;
  const int headB = dh ? (3 * L->hwBytes + 127) / 128 * 128 + (3 * L->CoutP2 * 4 + 127) / 128 * 128 + 2 * 128 * L->CoutP * 2 + 1024 : 0;
  const int fixed = wB + (zeroB + 127) / 128 * 128 + 1024 + (3 * L->CoutP * 4 + 127) / 128 * 128 + 2 * 8 * 32 * 4 + posB + headB +
                    8 * stgPerWarp * L->stgBytes + 1024 /* alignment slack */;
  int rsn = (RS_SMEM_CAP - fixed) / L->stageBytes;
  const int want = rs_env_int("DLWPCS_RS_STAGES", RS_MAXSTAGES);
  if (rsn > want) rsn = want;
  if (rsn > RS_MAXSTAGES) rsn = RS_MAXSTAGES;
  if (rsn < 4) return "stacked weights leave no room for the input-row ring";
  L->RSn = rsn;
  off = rsn * L->stageBytes;
  L->off_a2 = off;                                   // the head's two A tiles stay 1024-byte aligned (swizzle period)
  if (dh) off += 2 * 128 * L->CoutP * 2;
  L->off_w = off; off += wB;
  L->off_zero = off; off += (zeroB + 127) / 128 * 128;
  L->off_misc = off; off += 1024;
  L->off_bias = off; off += (3 * L->CoutP * 4 + 127) / 128 * 128;
  L->off_pix = off; off += 2 * 8 * 32 * 4;
  L->off_pos = off; off += 8 * RS_NPIXP * 8;
  L->off_px = off; off += 8 * 4 * RS_NPIXP * 4;
  L->off_hw = off;
  if (dh) off += (3 * L->hwBytes + 127) / 128 * 128;
  L->off_hbias = off;
  if (dh) off += (3 * L->CoutP2 * 4 + 127) / 128 * 128;
  L->off_stg = off; off += 8 * stgPerWarp * L->stgBytes;
  L->smemBytes = off + 1024;
  return nullptr;
}

// Work split of the row-streamed kernel: CTA c works on the (strip, output row) pairs [(cut_s[c], cut_y[c]), (cut_s[c+1],
// cut_y[c+1])).  Equal COST shares: a unit of h output rows streams h + 2 input rows, and a change of face group reloads the
// weights and drains the pipeline (~3 rows).  Host only.
void rs_work_cuts(int batch, int Wv, int Hout, int grid, int *cut_s, int *cut_y, int quantum = 1) {
  const long long ns0 = (4LL * batch * Wv + 127) / 128, ns1 = (1LL * batch * Wv + 127) / 128;
  const long long strips = ns0 + 2 * ns1, R = strips * Hout;
  const double unit_cost = 2.0, group_cost = 3.0;
  // the total cost depends on the number of units, which depends on the cuts: two passes (the first with an upper bound)
  double total = (double)R + unit_cost * (double)(strips + grid - 1) + group_cost * 2.0;
  for (int pass = 0; pass < 3; ++pass) {
    const double per = total / grid;
    long long pos = 0;
    double acc = 0.0;
    for (int c = 0; c < grid; ++c) {
      cut_s[c] = (int)(pos / Hout);
      cut_y[c] = (int)(pos % Hout);
      const double target = per * (c + 1);
      bool first = true;
      while (pos < R) {
        const long long sidx = pos / Hout;
        const int y = (int)(pos % Hout);
        double enter = (first || y == 0) ? unit_cost : 0.0;
        if (y == 0 && !first && (sidx == ns0 || sidx == ns0 + ns1)) enter += group_cost;
        const int left = Hout - y;
        int x = (int)(target - acc - enter + 0.5);
        x -= x % quantum;                                  // pool mode: units start and end on even rows
        if (c == grid - 1) x = left;                       // the last CTA takes what remains
        if (x > left) x = left;
        if (x < left) {                                    // the share ends inside this strip
          if (x < quantum && !first) break;                // not worth a new unit: stop at the boundary behind us
          if (x < quantum) x = quantum;                    // a CTA with work left never goes empty-handed
        }
        acc += enter + x;
        pos += x;
        first = false;
        if (x < left) break;
      }
    }
    total = acc;                                           // what the cuts actually cost
  }
  cut_s[grid] = (int)strips;
  cut_y[grid] = 0;
}

bool rs_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int rs_num_sms() {
  static int sms[kMaxDevices] = {};
  const int dev = current_device_index();
  if (!sms[dev] && (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0))
    sms[dev] = 148;
  return sms[dev];
}

unsigned *g_rs_err[kMaxDevices] = {};        // device pointers of ...
unsigned *g_rs_err_host[kMaxDevices] = {};   // ... mapped pinned host words: a watchdog code survives the trapped context

}  // namespace

int rs_debug_cuts(const dlwpcs_conv_desc *d, const Geometry &g, int grid, int *cut_s, int *cut_y) {
  RsPlan L;
  if (rs_make_plan(d, g, &L) || grid < 1 || grid > RS_MAXGRID) return 0;
  rs_work_cuts(d->batch, L.Wv, g.Hout, grid, cut_s, cut_y);
  return grid;
}

bool rs_eligible(const dlwpcs_conv_desc *d, const Geometry &g) {
  static const int enabled = rs_env_int("DLWPCS_RS", 1);
  if (!enabled) return false;
  RsPlan L;
  return rs_make_plan(d, g, &L) == nullptr;
}

int64_t rs_packed_weight_bytes(const dlwpcs_conv_desc *d, const Geometry &g) {
  RsPlan L;
  if (rs_make_plan(d, g, &L)) return -1;
  return 3LL * L.groupBytes + 3LL * L.CoutP * 4;
}

// layout parameters of the row-streamed kernel's packed image (cs_pack.cuh: pack_rs_element), for the pack launch in cs_tc.cu
bool rs_pack_params(const dlwpcs_conv_desc *d, const Geometry &g, int *CinP, int *CoutP, long long *groupElems) {
  RsPlan L;
  const char *r = rs_make_plan(d, g, &L);
  if (r) {
    set_error("row-streamed kernel does not support this configuration: %s", r);
    return false;
  }
  *CinP = L.CinP;
  *CoutP = L.CoutP;
  *groupElems = L.groupBytes / 2;
  return true;
}

bool rs_head_eligible(const dlwpcs_conv_desc *d, const Geometry &g, const dlwpcs_conv_desc *dh) {
  static const int enabled = rs_env_int("DLWPCS_RS", 1) && rs_env_int("DLWPCS_RS_HEAD", 1);
  if (!enabled) return false;
  RsPlan L;
  return rs_make_plan(d, g, &L, dh) == nullptr;
}

static int rs_launch(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed, void *y,
                     const dlwpcs_conv_desc *dh, const void *packed_h, void *y2, void *ypool, cudaStream_t st);

int rs_conv_fwd(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed, void *y,
                cudaStream_t st) {
  return rs_launch(d, g, x0, x1, packed, y, nullptr, nullptr, nullptr, nullptr, st);
}

// 3x3 layer + 1x1 head in one launch; packed_h = the head's classic packed image (dlwpcs_pack_weights of the 1x1 layer)
int rs_conv_fwd_head(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed,
                     const dlwpcs_conv_desc *dh, const void *packed_h, void *y2, cudaStream_t st) {
  return rs_launch(d, g, x0, x1, packed, nullptr, dh, packed_h, y2, nullptr, st);
}

bool rs_pool_eligible(const dlwpcs_conv_desc *d, const Geometry &g) {
  static const int enabled = rs_env_int("DLWPCS_RS", 1) && rs_env_int("DLWPCS_RS_POOLOUT", 1);
  if (!enabled) return false;
  RsPlan L;
  return rs_make_plan(d, g, &L, nullptr, true) == nullptr;
}

// the layer's output AND its 2x2 mean (the next layer's pooled input) from one launch
int rs_conv_fwd_pool(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed, void *y,
                     void *ypool, cudaStream_t st) {
  return rs_launch(d, g, x0, x1, packed, y, nullptr, nullptr, nullptr, ypool, st);
}

static int rs_launch(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed, void *y,
                     const dlwpcs_conv_desc *dh, const void *packed_h, void *y2, void *ypool, cudaStream_t st) {
  {
    const int dv = current_device_index();
    if (g_rs_err_host[dv] && *(volatile unsigned *)g_rs_err_host[dv]) {
      const unsigned c = *(volatile unsigned *)g_rs_err_host[dv];
      set_error("row-streamed kernel: a pipeline wait timed out (role %u, parity %u, barrier offset 0x%x, CTA %u)", c & 15u,
                (c >> 4) & 1u, (c >> 8) & 0xFFFu, c >> 20);
      fprintf(stderr, "[dlwpcs] row-streamed kernel watchdog: role %u parity %u barrier 0x%x CTA %u\n", c & 15u, (c >> 4) & 1u,
              (c >> 8) & 0xFFFu, c >> 20);
      return 4;
    }
  }
  RsP P;
  memset(&P, 0, sizeof(P));
  const char *r = rs_make_plan(d, g, &P.L, dh, ypool != nullptr);
  CS_CHECK(r == nullptr, "row-streamed kernel does not support this configuration: %s", r);
  const RsPlan &L = P.L;
  CS_CHECK(!ypool || rs_aligned16(ypool), "bf16 tensors must be 16-byte aligned");
  P.ypool = (__nv_bfloat16 *)ypool;
  CS_CHECK(rs_aligned16(x0) && (d->c1 == 0 || rs_aligned16(x1)) && rs_aligned16(dh ? y2 : y) && rs_aligned16(packed) &&
               (!dh || rs_aligned16(packed_h)),
           "bf16 tensors must be 16-byte aligned");
  if (dh) {
    P.hw = (const uint8_t *)packed_h;
    P.hbias = reinterpret_cast<const float *>(P.hw + 3LL * L.hwBytes);
    P.y2 = (__nv_bfloat16 *)y2;
    P.cout2 = dh->cout; P.act2 = dh->act; P.slope2 = dh->act_slope; P.maxv2 = dh->act_max;
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
  P.bias = reinterpret_cast<const float *>(P.wpack + 3LL * L.groupBytes);
  P.y = (__nv_bfloat16 *)y;
  P.batch = d->batch; P.n = d->n; P.Hout = g.Hout; P.Wout = g.Wout;
  P.cin = d->cin; P.cout = d->cout; P.c0 = d->c0; P.c1 = d->c1; P.mode0 = d->mode0; P.mode1 = d->mode1;
  P.act = d->act; P.slope = d->act_slope; P.maxv = d->act_max;
  if (d->batch == 0) return 0;
  const int dev_i = current_device_index();
  if (!g_rs_err[dev_i]) {
    CS_CUDA(cudaHostAlloc((void **)&g_rs_err_host[dev_i], sizeof(unsigned), cudaHostAllocMapped));
    *g_rs_err_host[dev_i] = 0u;
    CS_CUDA(cudaHostGetDevicePointer((void **)&g_rs_err[dev_i], g_rs_err_host[dev_i], 0));
  }
  P.err = g_rs_err[dev_i];
  typedef void (*kern_t)(const RsP);
  static const kern_t kerns[3][4] = {{conv_rs_kernel<1, 0>, conv_rs_kernel<2, 0>, nullptr, conv_rs_kernel<4, 0>},
                                     {conv_rs_kernel<1, 1>, conv_rs_kernel<2, 1>, nullptr, conv_rs_kernel<4, 1>},
                                     {conv_rs_kernel<1, 2>, conv_rs_kernel<2, 2>, nullptr, conv_rs_kernel<4, 2>}};
  static bool attr_set[kMaxDevices][3][4] = {};
  const int hi = dh ? 1 : (ypool ? 2 : 0);
  const kern_t kern = kerns[hi][L.KC16 - 1];
  CS_CHECK(kern != nullptr, "internal: bad K block");
  if (!attr_set[dev_i][hi][L.KC16 - 1]) {
    CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM_CAP));
    attr_set[dev_i][hi][L.KC16 - 1] = true;
  }
  const long long strips = (4LL * d->batch * L.Wv + 127) / 128 + 2 * ((1LL * d->batch * L.Wv + 127) / 128);
  const long long R = strips * g.Hout;
  CS_CHECK(R < (1LL << 30), "batch too large");
  static const int knock = rs_env_int("DLWPCS_RS_KNOCK", 0);
  P.knock = knock;
  // equal shares of the (strip, output row) sequence; a cut closer than `snap` rows to a strip boundary moves onto it (a
  // unit of h output rows streams h + 2 input rows)
  int grid = rs_num_sms();
  if (grid > RS_MAXGRID) grid = RS_MAXGRID;
  const long long min_rows = 4;
  if ((long long)grid * min_rows > R) grid = (int)((R + min_rows - 1) / min_rows);
  if (grid < 1) grid = 1;
  rs_work_cuts(d->batch, L.Wv, g.Hout, grid, P.cut_s, P.cut_y, ypool ? 2 : 1);
  P.trace = tc_trace_next(grid);
  static const int pdl = rs_env_int("DLWPCS_TC_PDL", 1);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(RS_THREADS);
  cfg.dynamicSmemBytes = (size_t)L.smemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  CS_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  return 0;
}

}  // namespace dlwpcs
