// float32 CUDA-core kernels of the cubed-sphere convolution: the 1e-5-parity path, the general-configuration path
// (any kernel size / stride / dilation / 'same') and, in round 1, the backward pass.  One smem-patch direct convolution
// kernel serves forward and dgrad; wgrad is a split-K reduction with a deterministic second pass.
//
// Reference semantics: DLWP/custom.py:921-1002 (CubeSphereConv2D.call), 1198-1308 (CubeSpherePadding2D.call),
// Azure/train_cs.py:197-199, 282-299 (pool / upsample / concat / capped leaky ReLU around every conv).
#include <string.h>
#include "cs_common.cuh"

namespace dlwpcs {

namespace {

constexpr int CK = 8;       // input channels staged per pass
constexpr int COT = 32;     // output channels per CTA
constexpr int PXT = 4;      // pixels per thread
constexpr int COQ = 8;      // output channels per thread

struct ConvP {
  const float *x0, *x1, *mask_y, *w, *bias;
  float *y;
  const int32_t *lut;
  int B, srcH, srcW;        // face dims of the un-resampled logical source (== Hin - 2*halo)
  int Hin, Win, Hout, Wout;
  int cin, cout, c0, c1, mode0, mode1;
  int kh, kw, sh, sw, dh, dw;
  int pt[3], pl;
  int act;
  float slope, maxv;
  int TH, TW, PH, PW, PHWp, tiles_x, tw4, npg;
  int vec4;
  int mask_act;
  float mask_slope, mask_max;
};

__device__ __forceinline__ float act_apply(float v, int act, float slope, float maxv) {
  if (act == DLWPCS_ACT_CAPPED_LEAKY_RELU) v = v < 0.f ? slope * v : fminf(v, maxv);
  return v;
}
__device__ __forceinline__ float act_grad_from_y(float y, int act, float slope, float maxv) {
  if (act == DLWPCS_ACT_CAPPED_LEAKY_RELU) return y < 0.f ? slope : (y < maxv ? 1.f : 0.f);
  return 1.f;
}

// Sample `nch` (1 or 4) consecutive channels starting at cc of one source at logical pixel (f,i,j).
template <int NCH>
__device__ __forceinline__ void sample_src(const float *__restrict__ src, int C, int mode, long long b, int f, int i,
                                           int j, int sH, int sW, int cc, float *out) {
  if (mode == DLWPCS_SRC_SAME) {
    const float *p = src + (((b * 6 + f) * sH + i) * (long long)sW + j) * C + cc;
    if (NCH == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
      out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else {
      out[0] = __ldg(p);
    }
  } else if (mode == DLWPCS_SRC_UP2) {
    const int h2 = sH >> 1, w2 = sW >> 1;
    const float *p = src + (((b * 6 + f) * h2 + (i >> 1)) * (long long)w2 + (j >> 1)) * C + cc;
    if (NCH == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
      out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else {
      out[0] = __ldg(p);
    }
  } else {  // POOL2: mean of the 2x2 block of a source with twice the edge
    const int h2 = sH * 2, w2 = sW * 2;
    const float *p = src + (((b * 6 + f) * h2 + 2 * i) * (long long)w2 + 2 * j) * C + cc;
#pragma unroll
    for (int k = 0; k < NCH; ++k) out[k] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const float *q = p + ((long long)a * w2 + bb) * C;
        if (NCH == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4 *>(q));
          out[0] += v.x; out[1] += v.y; out[2] += v.z; out[3] += v.w;
        } else {
          out[0] += __ldg(q);
        }
      }
#pragma unroll
    for (int k = 0; k < NCH; ++k) out[k] *= 0.25f;
  }
}

// grid: (tiles, ceil(cout/COT), B*6); block: npg * (COT/COQ) threads
__global__ void __launch_bounds__(256) conv_fp32_kernel(const __grid_constant__ ConvP P) {
  extern __shared__ float smem[];
  const int taps = P.kh * P.kw;
  float *s_patch = smem;                              // [CK][PHWp]
  float *s_w = s_patch + CK * P.PHWp;                 // [taps][CK][COT]
  int *s_src = reinterpret_cast<int *>(s_w + taps * CK * COT);   // [PH*PW] packed (f<<28 | i<<14 | j) or -1

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int tile = blockIdx.x, ty0 = (tile / P.tiles_x) * P.TH, tx0 = (tile % P.tiles_x) * P.TW;
  const int co0 = blockIdx.y * COT;
  const int bf = blockIdx.z, f = bf % 6;
  const long long b = bf / 6;
  const int grp = f < 4 ? 0 : f - 3;
  const int r_base = ty0 * P.sh - P.pt[grp], c_base = tx0 * P.sw - P.pl;
  const int phw = P.PH * P.PW;

  for (int q = tid; q < phw; q += nthr) {
    const int r = r_base + q / P.PW, c = c_base + q % P.PW;
    int code = -1;
    if (r >= 0 && r < P.Hin && c >= 0 && c < P.Win) {
      int sf = f, si = r, sj = c;
      if (P.lut) {
        const int s = __ldg(P.lut + (f * P.Hin + r) * P.Win + c);
        const int hw = P.srcH * P.srcW;
        sf = s / hw;
        const int rem = s - sf * hw;
        si = rem / P.srcW;
        sj = rem - si * P.srcW;
      }
      code = (sf << 28) | (si << 14) | sj;
    }
    s_src[q] = code;
  }

  const int cg = tid / P.npg, pg = tid - cg * P.npg;      // warp-mostly-uniform cout group, consecutive pixel groups
  const int py = pg / P.tw4, px = pg - py * P.tw4;         // thread's pixels: (py, px + q*tw4), q < PXT
  float acc[PXT][COQ];
#pragma unroll
  for (int q = 0; q < PXT; ++q)
#pragma unroll
    for (int k = 0; k < COQ; ++k) acc[q][k] = 0.f;

  const float *wg = P.w + (long long)grp * taps * P.cin * P.cout;

  for (int cc0 = 0; cc0 < P.cin; cc0 += CK) {
    __syncthreads();
    // ---- stage the input patch chunk: CK channels of every patch pixel
    if (P.vec4) {
      for (int idx = tid; idx < phw * 2; idx += nthr) {
        const int q = idx >> 1, quad = idx & 1, c = cc0 + quad * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const int code = s_src[q];
        if (code >= 0 && c < P.cin) {
          const int sf = code >> 28, si = (code >> 14) & 0x3fff, sj = code & 0x3fff;
          if (c < P.c0) sample_src<4>(P.x0, P.c0, P.mode0, b, sf, si, sj, P.srcH, P.srcW, c, v);
          else sample_src<4>(P.x1, P.c1, P.mode1, b, sf, si, sj, P.srcH, P.srcW, c - P.c0, v);
          if (P.mask_y) {
            const float4 m = __ldg(reinterpret_cast<const float4 *>(
                P.mask_y + (((b * 6 + sf) * P.srcH + si) * (long long)P.srcW + sj) * P.c0 + c));
            v[0] *= act_grad_from_y(m.x, P.mask_act, P.mask_slope, P.mask_max);
            v[1] *= act_grad_from_y(m.y, P.mask_act, P.mask_slope, P.mask_max);
            v[2] *= act_grad_from_y(m.z, P.mask_act, P.mask_slope, P.mask_max);
            v[3] *= act_grad_from_y(m.w, P.mask_act, P.mask_slope, P.mask_max);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) s_patch[(quad * 4 + k) * P.PHWp + q] = v[k];
      }
    } else {
      for (int idx = tid; idx < phw * CK; idx += nthr) {
        const int q = idx / CK, k = idx - q * CK, c = cc0 + k;
        float v = 0.f;
        const int code = s_src[q];
        if (code >= 0 && c < P.cin) {
          const int sf = code >> 28, si = (code >> 14) & 0x3fff, sj = code & 0x3fff;
          if (c < P.c0) sample_src<1>(P.x0, P.c0, P.mode0, b, sf, si, sj, P.srcH, P.srcW, c, &v);
          else sample_src<1>(P.x1, P.c1, P.mode1, b, sf, si, sj, P.srcH, P.srcW, c - P.c0, &v);
          if (P.mask_y)
            v *= act_grad_from_y(__ldg(P.mask_y + (((b * 6 + sf) * P.srcH + si) * (long long)P.srcW + sj) * P.c0 + c),
                                 P.mask_act, P.mask_slope, P.mask_max);
        }
        s_patch[k * P.PHWp + q] = v;
      }
    }
    // ---- stage the weight chunk [tap][CK][COT]
    for (int idx = tid; idx < taps * CK * COT; idx += nthr) {
      const int o = idx % COT, k = (idx / COT) % CK, t = idx / (COT * CK);
      const int c = cc0 + k, co = co0 + o;
      s_w[idx] = (c < P.cin && co < P.cout) ? __ldg(wg + ((long long)t * P.cin + c) * P.cout + co) : 0.f;
    }
    __syncthreads();
    if (pg < P.npg && cg < COT / COQ) {
      for (int k = 0; k < CK; ++k) {
        const float *pk = s_patch + k * P.PHWp;
        for (int u = 0; u < P.kh; ++u) {
          const int prow = (py * P.sh + u * P.dh) * P.PW;
          for (int v = 0; v < P.kw; ++v) {
            const float4 w0 = *reinterpret_cast<const float4 *>(s_w + ((u * P.kw + v) * CK + k) * COT + cg * COQ);
            const float4 w1 = *reinterpret_cast<const float4 *>(s_w + ((u * P.kw + v) * CK + k) * COT + cg * COQ + 4);
#pragma unroll
            for (int q = 0; q < PXT; ++q) {
              const float a = pk[prow + (px + q * P.tw4) * P.sw + v * P.dw];
              acc[q][0] = fmaf(a, w0.x, acc[q][0]); acc[q][1] = fmaf(a, w0.y, acc[q][1]);
              acc[q][2] = fmaf(a, w0.z, acc[q][2]); acc[q][3] = fmaf(a, w0.w, acc[q][3]);
              acc[q][4] = fmaf(a, w1.x, acc[q][4]); acc[q][5] = fmaf(a, w1.y, acc[q][5]);
              acc[q][6] = fmaf(a, w1.z, acc[q][6]); acc[q][7] = fmaf(a, w1.w, acc[q][7]);
            }
          }
        }
      }
    }
  }

  if (pg >= P.npg || cg >= COT / COQ) return;
  const int oy = ty0 + py;
  if (oy >= P.Hout) return;
  const int cob = co0 + cg * COQ;
  float bias[COQ];
#pragma unroll
  for (int k = 0; k < COQ; ++k) bias[k] = (P.bias && cob + k < P.cout) ? __ldg(P.bias + grp * P.cout + cob + k) : 0.f;
#pragma unroll
  for (int q = 0; q < PXT; ++q) {
    const int ox = tx0 + px + q * P.tw4;
    if (ox >= P.Wout) continue;
    float *yp = P.y + (((b * 6 + f) * P.Hout + oy) * (long long)P.Wout + ox) * P.cout + cob;
    float o[COQ];
#pragma unroll
    for (int k = 0; k < COQ; ++k) o[k] = act_apply(acc[q][k] + bias[k], P.act, P.slope, P.maxv);
    if ((P.cout & 3) == 0 && cob + COQ <= P.cout) {
      *reinterpret_cast<float4 *>(yp) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4 *>(yp + 4) = make_float4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
      for (int k = 0; k < COQ; ++k)
        if (cob + k < P.cout) yp[k] = o[k];
    }
  }
}

// ---- weight packing ---------------------------------------------------------------------------------------------
// forward:    packed[g][u][v][ci][co] (+ bias[g][co] appended);   dgrad: packed[g][u''][v''][co][ci], taps rotated.
__global__ void pack_fp32_kernel(const float *__restrict__ w_eq, const float *__restrict__ w_pol,
                                 const float *__restrict__ w_np, const float *__restrict__ b_eq,
                                 const float *__restrict__ b_pol, const float *__restrict__ b_np, float *__restrict__ out,
                                 int kh, int kw, int cin, int cout, int flip, int transposed, long long nw) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long per = (long long)kh * kw * cin * cout;
  if (i < nw) {
    const int g = (int)(i / per);
    long long r = i % per;
    int u, v, ci, co;
    if (!transposed) {
      co = (int)(r % cout); r /= cout;
      ci = (int)(r % cin); r /= cin;
      v = (int)(r % kw); u = (int)(r / kw);
    } else {
      ci = (int)(r % cin); r /= cin;
      co = (int)(r % cout); r /= cout;
      v = kw - 1 - (int)(r % kw); u = kh - 1 - (int)(r / kw);
    }
    const float *src = g == 0 ? w_eq : (g == 1 ? w_pol : (w_np ? w_np : w_pol));
    const int us = (g == 2 && flip) ? kh - 1 - u : u;
    out[i] = src[(((long long)us * kw + v) * cin + ci) * cout + co];
  } else if (!transposed && i < nw + 3LL * cout) {
    const int j = (int)(i - nw), g = j / cout, co = j % cout;
    const float *src = g == 0 ? b_eq : (g == 1 ? b_pol : (b_np ? b_np : b_pol));
    out[i] = src ? src[co] : 0.f;
  }
}

void choose_tiles(int Hout, int Wout, int *TH, int *TW) {
  int tw = ((Wout + 3) / 4) * 4;
  if (tw > 64) {
    const int nt = (Wout + 63) / 64;
    tw = (((Wout + nt - 1) / nt) + 3) / 4 * 4;
  }
  int thmax = 256 / tw;
  if (thmax < 1) thmax = 1;
  const int nty = (Hout + thmax - 1) / thmax;
  *TH = (Hout + nty - 1) / nty;
  *TW = tw;
}

int launch_conv(ConvP &P, cudaStream_t st) {
  choose_tiles(P.Hout, P.Wout, &P.TH, &P.TW);
  P.tiles_x = (P.Wout + P.TW - 1) / P.TW;
  const int tiles_y = (P.Hout + P.TH - 1) / P.TH;
  P.tw4 = P.TW / PXT;
  P.npg = P.TH * P.tw4;
  P.PH = (P.TH - 1) * P.sh + (P.kh - 1) * P.dh + 1;
  P.PW = (P.TW - 1) * P.sw + (P.kw - 1) * P.dw + 1;
  P.PHWp = (P.PH * P.PW) | 1;
  const int threads = P.npg * (COT / COQ);
  CS_CHECK(threads <= 256, "internal: tile too large (%d threads)", threads);
  CS_CHECK(P.srcH < 16384 && P.srcW < 16384, "face edge too large");
  const size_t smem = sizeof(float) * ((size_t)CK * P.PHWp + (size_t)P.kh * P.kw * CK * COT) +
                      sizeof(int) * (size_t)P.PH * P.PW;
  CS_CHECK(smem <= 200 * 1024, "kernel window too large for the float32 kernel (%zu bytes of shared memory)", smem);
  static bool big_smem[kMaxDevices] = {};     // the opt-in is per device (function attributes live in the context)
  const int dev = current_device_index();
  if (smem > 48 * 1024 && !big_smem[dev]) {
    CS_CUDA(cudaFuncSetAttribute(conv_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    big_smem[dev] = true;
  }
  const long long gz = (long long)P.B * 6;
  CS_CHECK(gz <= 65535, "batch too large for one launch (B*6 = %lld > 65535)", gz);
  dim3 grid(P.tiles_x * tiles_y, (P.cout + COT - 1) / COT, (unsigned)gz);
  conv_fp32_kernel<<<grid, threads, smem, st>>>(P);
  CS_CUDA(cudaGetLastError());
  return 0;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int fp32_pack_weights(const dlwpcs_conv_desc *d, const dlwpcs_conv_weights *w, int transposed, float *packed,
                      cudaStream_t st) {
  const long long nw = 3LL * d->kh * d->kw * d->cin * d->cout;
  const long long total = transposed ? nw : nw + 3LL * d->cout;
  const int block = 256;
  pack_fp32_kernel<<<(unsigned)((total + block - 1) / block), block, 0, st>>>(
      w->w_eq, w->w_pol, d->independent_north_pole ? w->w_np : nullptr, d->use_bias ? w->b_eq : nullptr,
      d->use_bias ? w->b_pol : nullptr, (d->use_bias && d->independent_north_pole) ? w->b_np : nullptr, packed, d->kh,
      d->kw, d->cin, d->cout, d->flip_north_pole, transposed, nw);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int fp32_conv_fwd(const dlwpcs_conv_desc *d, const Geometry &g, const float *x0, const float *x1, const float *packed,
                  float *y, cudaStream_t st) {
  ConvP P;
  memset(&P, 0, sizeof(P));
  P.x0 = x0; P.x1 = x1; P.mask_y = nullptr;
  P.w = packed;
  P.bias = d->use_bias ? packed + 3LL * g.taps * d->cin * d->cout : nullptr;
  P.y = y;
  P.lut = nullptr;
  if (d->halo > 0) {
    const HaloTables *t = get_halo_tables(d->n, d->halo);
    if (!t) return 3;
    P.lut = t->lut;
  }
  P.B = d->batch; P.srcH = P.srcW = d->n;
  P.Hin = g.Hin; P.Win = g.Win; P.Hout = g.Hout; P.Wout = g.Wout;
  P.cin = d->cin; P.cout = d->cout; P.c0 = d->c0; P.c1 = d->c1; P.mode0 = d->mode0; P.mode1 = d->mode1;
  P.kh = d->kh; P.kw = d->kw; P.sh = d->stride_h; P.sw = d->stride_w; P.dh = d->dil_h; P.dw = d->dil_w;
  P.pt[0] = g.pt[0]; P.pt[1] = g.pt[1]; P.pt[2] = g.pt[2]; P.pl = g.pl;
  P.act = d->act; P.slope = d->act_slope; P.maxv = d->act_max;
  P.vec4 = (d->c0 % 4 == 0) && (d->c1 % 4 == 0) && aligned16(x0) && (d->c1 == 0 || aligned16(x1));
  return launch_conv(P, st);
}

// dgrad: dx_ext[r,c,ci] = sum_{u,v,o} dy[r + pt - u*dh, c + pl - v*dw, o] * W[u,v,ci,o]   (stride 1)
// == the same direct convolution over dy with the taps rotated 180 degrees, cin/cout swapped and zero fill outside dy,
// followed by the adjoint of the halo gather.
int fp32_conv_dgrad(const dlwpcs_conv_desc *d, const Geometry &g, const float *dy, const float *y,
                    const float *packed_t, float *dx, void *workspace, cudaStream_t st) {
  ConvP P;
  memset(&P, 0, sizeof(P));
  P.x0 = dy; P.x1 = nullptr;
  P.mask_y = d->act != DLWPCS_ACT_NONE ? y : nullptr;
  P.mask_act = d->act; P.mask_slope = d->act_slope; P.mask_max = d->act_max;
  P.w = packed_t; P.bias = nullptr;
  P.y = d->halo > 0 ? (float *)workspace : dx;
  P.lut = nullptr;
  P.B = d->batch; P.srcH = g.Hout; P.srcW = g.Wout;
  P.Hin = g.Hout; P.Win = g.Wout; P.Hout = g.Hin; P.Wout = g.Win;
  P.cin = d->cout; P.cout = d->cin; P.c0 = d->cout; P.c1 = 0; P.mode0 = DLWPCS_SRC_SAME; P.mode1 = DLWPCS_SRC_SAME;
  P.kh = d->kh; P.kw = d->kw; P.sh = 1; P.sw = 1; P.dh = d->dil_h; P.dw = d->dil_w;
  for (int k = 0; k < 3; ++k) P.pt[k] = (d->kh - 1) * d->dil_h - g.pt[k];
  P.pl = (d->kw - 1) * d->dil_w - g.pl;
  P.act = DLWPCS_ACT_NONE;
  P.vec4 = (d->cout % 4 == 0) && aligned16(dy) && (!P.mask_y || aligned16(y));
  if (int rc = launch_conv(P, st)) return rc;
  if (d->halo > 0) return dlwpcs_pad_bwd(workspace, dx, d->batch, d->n, d->cin, d->halo, DLWPCS_F32, st);
  return 0;
}

// ---- wgrad --------------------------------------------------------------------------------------------------------
namespace {

constexpr int WCK = 8;     // input channels per CTA
constexpr int WCO = 32;    // output channels per CTA
constexpr int WTAPS = 9;   // taps accumulated per CTA pass

struct WgradP {
  const void *x, *dy, *mask_y;   // float32 or bfloat16 (in_bf16)
  int in_bf16;
  const int32_t *lut;
  float *ws;       // [6*strips][taps][cin][cout]
  float *ws_b;     // [6*strips][cout]
  int B, n, Hin, Win, Ho, Wo, cin, cout, kh, kw, dh, dw, pt[3], pl;
  int TH, TW, PH, PW, tiles_x, strips, cin_chunks, tap_blocks;
  int act;
  float slope, maxv;
};

__device__ __forceinline__ float ld_in(const void *p, long long i, int bf16) {
  return bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p)[i]) : __ldg(reinterpret_cast<const float *>(p) + i);
}

// grid: (cin_chunks*tap_blocks, ceil(cout/WCO), 6*strips); block 256 = (8 input channels) x (32 output channels)
__global__ void __launch_bounds__(256) wgrad_fp32_kernel(const __grid_constant__ WgradP P) {
  extern __shared__ float smem[];
  float *s_patch = smem;                               // [PH*PW][WCK]
  float *s_dy = s_patch + P.PH * P.PW * WCK;           // [TH*TW][WCO]
  int *s_src = reinterpret_cast<int *>(s_dy + P.TH * P.TW * WCO);
  const int tid = threadIdx.x;
  const int c = tid >> 5, o = tid & 31;
  const int chunk = blockIdx.x % P.cin_chunks, tb = blockIdx.x / P.cin_chunks;
  const int cc0 = chunk * WCK, co0 = blockIdx.y * WCO, t0 = tb * WTAPS;
  const int f = blockIdx.z / P.strips, strip = blockIdx.z % P.strips;
  const int grp = f < 4 ? 0 : f - 3;
  const int ty0 = (strip / P.tiles_x) * P.TH, tx0 = (strip % P.tiles_x) * P.TW;
  const int r_base = ty0 - P.pt[grp], c_base = tx0 - P.pl;
  const int phw = P.PH * P.PW, taps = P.kh * P.kw;

  for (int q = tid; q < phw; q += 256) {
    const int r = r_base + q / P.PW, cc = c_base + q % P.PW;
    int code = -1;
    if (r >= 0 && r < P.Hin && cc >= 0 && cc < P.Win) code = P.lut ? __ldg(P.lut + (f * P.Hin + r) * P.Win + cc)
                                                                     : (f * P.Hin + r) * P.Win + cc;
    s_src[q] = code;
  }
  // per-thread tap offsets into the patch
  int toff[WTAPS];
#pragma unroll
  for (int t = 0; t < WTAPS; ++t) {
    const int tt = t0 + t < taps ? t0 + t : 0;
    toff[t] = ((tt / P.kw) * P.dh * P.PW + (tt % P.kw) * P.dw) * WCK + c;
  }
  float acc[WTAPS];
#pragma unroll
  for (int t = 0; t < WTAPS; ++t) acc[t] = 0.f;
  float bsum = 0.f;
  const bool do_bias = (blockIdx.x == 0) && (c == 0);
  const long long src_px = 6LL * P.n * P.n;

  for (int b = 0; b < P.B; ++b) {
    __syncthreads();
    for (int idx = tid; idx < phw * WCK; idx += 256) {
      const int q = idx / WCK, k = idx - q * WCK;
      const int code = s_src[q];
      float v = 0.f;
      if (code >= 0 && cc0 + k < P.cin) v = ld_in(P.x, ((long long)b * src_px + code) * P.cin + cc0 + k, P.in_bf16);
      s_patch[idx] = v;
    }
    for (int idx = tid; idx < P.TH * P.TW * WCO; idx += 256) {
      const int q = idx / WCO, k = idx - q * WCO;
      const int oy = ty0 + q / P.TW, ox = tx0 + q % P.TW;
      float v = 0.f;
      if (oy < P.Ho && ox < P.Wo && co0 + k < P.cout) {
        const long long off = ((((long long)b * 6 + f) * P.Ho + oy) * P.Wo + ox) * P.cout + co0 + k;
        v = ld_in(P.dy, off, P.in_bf16);
        if (P.mask_y) v *= act_grad_from_y(ld_in(P.mask_y, off, P.in_bf16), P.act, P.slope, P.maxv);
      }
      s_dy[idx] = v;
    }
    __syncthreads();
    for (int i = 0; i < P.TH; ++i)
      for (int j = 0; j < P.TW; ++j) {
        const float dyv = s_dy[(i * P.TW + j) * WCO + o];
        const float *pp = s_patch + (i * P.PW + j) * WCK;
#pragma unroll
        for (int t = 0; t < WTAPS; ++t) acc[t] = fmaf(pp[toff[t]], dyv, acc[t]);
        bsum += dyv;
      }
  }
  const long long z = blockIdx.z;
  if (cc0 + c < P.cin && co0 + o < P.cout) {
#pragma unroll
    for (int t = 0; t < WTAPS; ++t)
      if (t0 + t < taps) P.ws[((z * taps + t0 + t) * P.cin + cc0 + c) * P.cout + co0 + o] = acc[t];
  }
  if (do_bias && co0 + o < P.cout) P.ws_b[z * P.cout + co0 + o] = bsum;
}

// second pass: fixed-order sum over the strips of each face group, un-flip / merge the north-pole share
__global__ void wgrad_reduce_kernel(const float *__restrict__ ws, const float *__restrict__ ws_b, float *dw_eq,
                                    float *dw_pol, float *dw_np, float *db_eq, float *db_pol, float *db_np, int strips,
                                    int kh, int kw, int cin, int cout, int flip) {
  const long long per = (long long)kh * kw * cin * cout;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < per) {
    const int u = (int)(i / ((long long)kw * cin * cout));
    const long long rest = i % ((long long)kw * cin * cout);
    auto face_sum = [&](int f, long long idx) {
      float s = 0.f;
      for (int k = 0; k < strips; ++k) s += ws[((long long)(f * strips + k)) * per + idx];
      return s;
    };
    float e = 0.f;
    for (int f = 0; f < 4; ++f) e += face_sum(f, i);
    dw_eq[i] = e;
    const float south = face_sum(4, i);
    const long long i5 = flip ? ((long long)(kh - 1 - u) * kw * cin * cout + rest) : i;   // packed tap holding source row u
    const float north = face_sum(5, i5);
    if (dw_np) {
      dw_pol[i] = south;
      dw_np[i] = north;
    } else {
      dw_pol[i] = south + north;
    }
  } else if (db_eq && i < per + cout) {
    const int co = (int)(i - per);
    auto face_sum = [&](int f) {
      float s = 0.f;
      for (int k = 0; k < strips; ++k) s += ws_b[(long long)(f * strips + k) * cout + co];
      return s;
    };
    float e = 0.f;
    for (int f = 0; f < 4; ++f) e += face_sum(f);
    db_eq[co] = e;
    const float south = face_sum(4), north = face_sum(5);
    if (db_np) {
      db_pol[co] = south;
      db_np[co] = north;
    } else {
      db_pol[co] = south + north;
    }
  }
}

void wgrad_tiles(const Geometry &g, int *TH, int *TW) {
  int tw = g.Wout;
  if (tw > 32) {
    const int nt = (g.Wout + 31) / 32;
    tw = (g.Wout + nt - 1) / nt;
  }
  int thmax = 192 / tw;
  if (thmax < 1) thmax = 1;
  const int nty = (g.Hout + thmax - 1) / thmax;
  *TH = (g.Hout + nty - 1) / nty;
  *TW = tw;
}

}  // namespace

int64_t fp32_wgrad_workspace_bytes(const dlwpcs_conv_desc *d, const Geometry &g) {
  int TH, TW;
  wgrad_tiles(g, &TH, &TW);
  const int64_t strips = (int64_t)((g.Hout + TH - 1) / TH) * ((g.Wout + TW - 1) / TW);
  return 6 * strips * ((int64_t)g.taps * d->cin * d->cout + d->cout) * 4;
}

int fp32_conv_wgrad(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *dy, const void *y,
                    const dlwpcs_conv_wgrads *out, void *workspace, cudaStream_t st) {
  WgradP P;
  memset(&P, 0, sizeof(P));
  P.x = x0; P.dy = dy;
  P.in_bf16 = d->x_dtype == DLWPCS_BF16;
  P.mask_y = d->act != DLWPCS_ACT_NONE ? y : nullptr;
  P.act = d->act; P.slope = d->act_slope; P.maxv = d->act_max;
  P.lut = nullptr;
  if (d->halo > 0) {
    const HaloTables *t = get_halo_tables(d->n, d->halo);
    if (!t) return 3;
    P.lut = t->lut;
  }
  P.B = d->batch; P.n = d->n; P.Hin = g.Hin; P.Win = g.Win; P.Ho = g.Hout; P.Wo = g.Wout;
  P.cin = d->cin; P.cout = d->cout; P.kh = d->kh; P.kw = d->kw; P.dh = d->dil_h; P.dw = d->dil_w;
  P.pt[0] = g.pt[0]; P.pt[1] = g.pt[1]; P.pt[2] = g.pt[2]; P.pl = g.pl;
  wgrad_tiles(g, &P.TH, &P.TW);
  P.tiles_x = (g.Wout + P.TW - 1) / P.TW;
  P.strips = ((g.Hout + P.TH - 1) / P.TH) * P.tiles_x;
  P.PH = P.TH + (d->kh - 1) * d->dil_h;
  P.PW = P.TW + (d->kw - 1) * d->dil_w;
  P.cin_chunks = (d->cin + WCK - 1) / WCK;
  P.tap_blocks = (g.taps + WTAPS - 1) / WTAPS;
  P.ws = (float *)workspace;
  P.ws_b = P.ws + 6LL * P.strips * g.taps * d->cin * d->cout;
  const size_t smem = sizeof(float) * ((size_t)P.PH * P.PW * WCK + (size_t)P.TH * P.TW * WCO) +
                      sizeof(int) * (size_t)P.PH * P.PW;
  CS_CHECK(smem <= 200 * 1024, "kernel window too large for the wgrad kernel (%zu bytes of shared memory)", smem);
  static bool big_smem[kMaxDevices] = {};
  const int dev = current_device_index();
  if (smem > 48 * 1024 && !big_smem[dev]) {
    CS_CUDA(cudaFuncSetAttribute(wgrad_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    big_smem[dev] = true;
  }
  dim3 grid(P.cin_chunks * P.tap_blocks, (d->cout + WCO - 1) / WCO, 6 * P.strips);
  wgrad_fp32_kernel<<<grid, 256, smem, st>>>(P);
  CS_CUDA(cudaGetLastError());
  const long long total = (long long)g.taps * d->cin * d->cout + d->cout;
  wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      P.ws, P.ws_b, out->dw_eq, out->dw_pol, d->independent_north_pole ? out->dw_np : nullptr,
      d->use_bias ? out->db_eq : nullptr, d->use_bias ? out->db_pol : nullptr,
      (d->use_bias && d->independent_north_pole) ? out->db_np : nullptr, P.strips, d->kh, d->kw, d->cin, d->cout,
      d->flip_north_pole);
  CS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dlwpcs
