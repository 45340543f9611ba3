// Weight packing shared by the classic (cs_tc.cu) and the row-streamed (cs_tc_rs.cu) tensor-core kernels: one packed image =
// the HWIO float32 kernels of one CubeSphereConv2D (custom.py:880-914) rounded to bf16 in exactly the shared-memory layout of
// the kernel that reads it, per face group (0 equatorial, 1 south pole, 2 north pole: rows flipped when flip_north_pole --
// custom.py:969/995), followed by the float32 biases [3][CoutP].  The source kernels (kh, kw, scin, scout) are zero-extended
// to the descriptor's (cin, cout).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace dlwpcs {

enum { PACK_CLASSIC_FWD = 0, PACK_CLASSIC_T = 1, PACK_RS_FWD = 2 };

struct PackImage {
  const float *w_eq, *w_pol, *w_np, *b_eq, *b_pol, *b_np;
  uint8_t *out;
  int kind;                      // PACK_*
  int CinP, CoutP, KC;           // padded channels of the GEMM this image feeds (classic transposed: K = cout, N = cin)
  long long groupElems;          // bf16 elements per face group; the image holds 3 * groupElems + 3 * CoutP floats' worth
  int kh, kw, cin, cout, scin, scout, flip;
};
constexpr int PACK_MAX_IMAGES = 24;
struct PackBatch {
  int n;
  PackImage im[PACK_MAX_IMAGES];
};

// classic image: packed[g][kc][tap][k8][n][8]; transposed (dgrad): K runs over the forward cout, N over the forward cin, taps
// rotated 180 degrees, biases zero
__device__ __forceinline__ void pack_classic_element(const PackImage &S, long long i) {
  const int transposed = S.kind == PACK_CLASSIC_T;
  const long long groupElems = S.groupElems;
  const int CoutP = S.CoutP, KC = S.KC, kh = S.kh, kw = S.kw;
  const int taps = kh * kw;
  if (i < 3 * groupElems) {
    const int g = (int)(i / groupElems);
    long long r = i % groupElems;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % CoutP); r /= CoutP;
    const int k8 = (int)(r % (KC / 8)); r /= (KC / 8);
    const int tap = (int)(r % taps);
    const int kc = (int)(r / taps);
    const int k = kc * KC + k8 * 8 + e;
    int u = tap / kw, v = tap % kw;
    const int gemm_cin = transposed ? S.cout : S.cin, gemm_cout = transposed ? S.cin : S.cout;
    float val = 0.f;
    if (k < gemm_cin && n < gemm_cout) {
      const float *src = g == 0 ? S.w_eq : (g == 1 ? S.w_pol : (S.w_np ? S.w_np : S.w_pol));
      if (transposed) { u = kh - 1 - u; v = kw - 1 - v; }
      const int us = (g == 2 && S.flip) ? kh - 1 - u : u;
      const int ci = transposed ? n : k, co = transposed ? k : n;
      if (ci < S.scin && co < S.scout) val = src[(((long long)us * kw + v) * S.scin + ci) * S.scout + co];
    }
    reinterpret_cast<__nv_bfloat16 *>(S.out)[i] = __float2bfloat16_rn(val);
  } else if (i < 3 * groupElems + 3LL * CoutP) {
    const int j = (int)(i - 3 * groupElems), g = j / CoutP, co = j % CoutP;
    const float *src = g == 0 ? S.b_eq : (g == 1 ? S.b_pol : (S.b_np ? S.b_np : S.b_pol));
    float *bo = reinterpret_cast<float *>(S.out + 3 * groupElems * 2);
    bo[j] = (src && !transposed && co < S.scout) ? src[co] : 0.f;
  }
}

// row-streamed image (3x3): packed[g][v][k8][n = j*CoutP + o][8], part j = kernel row 2 - j
__device__ __forceinline__ void pack_rs_element(const PackImage &S, long long i) {
  const long long groupElems = S.groupElems;
  const int CoutP = S.CoutP, CinP = S.CinP, NT = 3 * CoutP;
  if (i < 3 * groupElems) {
    const int g = (int)(i / groupElems);
    long long r = i % groupElems;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % NT); r /= NT;
    const int k8 = (int)(r % (CinP / 8));
    const int v = (int)(r / (CinP / 8));
    const int j = n / CoutP, o = n - j * CoutP, k = k8 * 8 + e, u = 2 - j;
    float val = 0.f;
    if (k < S.cin && o < S.cout && k < S.scin && o < S.scout) {
      const float *src = g == 0 ? S.w_eq : (g == 1 ? S.w_pol : (S.w_np ? S.w_np : S.w_pol));
      const int us = (g == 2 && S.flip) ? 2 - u : u;
      val = src[(((long long)us * 3 + v) * S.scin + k) * S.scout + o];
    }
    reinterpret_cast<__nv_bfloat16 *>(S.out)[i] = __float2bfloat16_rn(val);
  } else if (i < 3 * groupElems + 3LL * CoutP) {
    const int jj = (int)(i - 3 * groupElems), g = jj / CoutP, co = jj % CoutP;
    const float *src = g == 0 ? S.b_eq : (g == 1 ? S.b_pol : (S.b_np ? S.b_np : S.b_pol));
    float *bo = reinterpret_cast<float *>(S.out + 3 * groupElems * 2);
    bo[jj] = (src && co < S.scout) ? src[co] : 0.f;
  }
}

__device__ __forceinline__ void pack_element(const PackImage &S, long long i) {
  if (S.kind == PACK_RS_FWD) pack_rs_element(S, i);
  else pack_classic_element(S, i);
}

}  // namespace dlwpcs
