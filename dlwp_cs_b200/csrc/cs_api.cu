// libdlwpcs C ABI: error reporting, halo tables, geometry, standalone padding / activation kernels and the dispatch of
// the convolution entry points onto the fp32 (cs_fp32.cu) and tcgen05 bf16 (cs_tc.cu) kernels.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <tuple>
#include <initializer_list>
#include "cs_common.cuh"

namespace dlwpcs {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------------------------------
// Halo index table.  Restates the two-stage composition of CubeSpherePadding2D.call (reference DLWP/custom.py:1198-1308)
// as index arithmetic: stage 1 fills rows from the polar / equatorial neighbours (1203-1249), stage 2 fills columns
// reading the row-padded faces, and -- for the two polar faces -- rows p:2p / -2p:-p of the fully padded faces 3 and 1
// (1291-1301).  Verified against the reference code itself in tests/test_lib_cpu.py.
// ---------------------------------------------------------------------------------------------------------------------
void build_pad_lut(int n, int p, std::vector<int32_t> &lut) {
  const int H = n + 2 * p;
  auto src = [n](int f, int i, int j) { return f * n * n + i * n + j; };
  // stage 1: s1[f][r][c], r in [0,H), c in [0,n)
  std::vector<int32_t> s1((size_t)6 * H * n);
  for (int f = 0; f < 6; ++f)
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < n; ++c) {
        int v;
        if (r >= p && r < n + p) {
          v = src(f, r - p, c);
        } else if (r < p) {  // top halo, a = r
          const int a = r;
          switch (f) {
            case 0: v = src(4, n - p + a, c); break;
            case 1: v = src(4, n - 1 - c, n - p + a); break;
            case 2: v = src(4, p - 1 - a, n - 1 - c); break;
            case 3: v = src(4, c, p - 1 - a); break;
            case 4: v = src(2, p - 1 - a, n - 1 - c); break;
            default: v = src(0, n - p + a, c); break;
          }
        } else {  // bottom halo
          const int a = r - (n + p);
          switch (f) {
            case 0: v = src(5, a, c); break;
            case 1: v = src(5, c, n - 1 - a); break;
            case 2: v = src(5, n - 1 - a, n - 1 - c); break;
            case 3: v = src(5, n - 1 - c, a); break;
            case 4: v = src(0, a, c); break;
            default: v = src(2, n - 1 - a, n - 1 - c); break;
          }
        }
        s1[((size_t)f * H + r) * n + c] = v;
      }
  lut.assign((size_t)6 * H * H, 0);
  auto S1 = [&](int f, int r, int c) { return s1[((size_t)f * H + r) * n + c]; };
  auto OUT = [&](int f, int r, int c) -> int32_t & { return lut[((size_t)f * H + r) * H + c]; };
  for (int f = 0; f < 4; ++f)  // periodic equatorial belt
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < H; ++c) {
        if (c < p) OUT(f, r, c) = S1((f + 3) % 4, r, n - p + c);
        else if (c < n + p) OUT(f, r, c) = S1(f, r, c - p);
        else OUT(f, r, c) = S1((f + 1) % 4, r, c - (n + p));
      }
  for (int r = 0; r < H; ++r)
    for (int c = 0; c < H; ++c) {
      if (c < p) {  // left halo column a = c
        OUT(4, r, c) = OUT(3, 2 * p - 1 - c, r);
        OUT(5, r, c) = OUT(3, H - 2 * p + c, H - 1 - r);
      } else if (c < n + p) {
        OUT(4, r, c) = S1(4, r, c - p);
        OUT(5, r, c) = S1(5, r, c - p);
      } else {
        const int a = c - (n + p);
        OUT(4, r, c) = OUT(1, p + a, H - 1 - r);
        OUT(5, r, c) = OUT(1, H - p - 1 - a, r);
      }
    }
}

namespace {
std::mutex g_mu;
std::map<std::tuple<int, int, int>, HaloTables> g_tables;
}  // namespace

const HaloTables *get_halo_tables(int n, int p) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("cudaGetDevice failed");
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_tuple(dev, n, p);
  auto it = g_tables.find(key);
  if (it != g_tables.end()) return &it->second;
  std::vector<int32_t> lut;
  build_pad_lut(n, p, lut);
  const int H = n + 2 * p;
  const int nsrc = 6 * n * n, npad = 6 * H * H;
  std::vector<int32_t> start(nsrc + 1, 0), items(npad);
  for (int q = 0; q < npad; ++q) start[lut[q] + 1]++;
  for (int s = 0; s < nsrc; ++s) start[s + 1] += start[s];
  std::vector<int32_t> fill(start.begin(), start.end() - 1);
  for (int q = 0; q < npad; ++q) items[fill[lut[q]]++] = q;  // ascending q within each source: deterministic order
  HaloTables t;
  t.n = n;
  t.p = p;
  // Plain cudaMalloc + synchronous copies: must not run inside a stream capture -- callers warm the cache first.
  cudaError_t e = cudaMalloc(&t.lut, sizeof(int32_t) * npad);
  if (e == cudaSuccess) e = cudaMalloc(&t.inv_start, sizeof(int32_t) * (nsrc + 1));
  if (e == cudaSuccess) e = cudaMalloc(&t.inv_items, sizeof(int32_t) * npad);
  if (e == cudaSuccess) e = cudaMemcpy(t.lut, lut.data(), sizeof(int32_t) * npad, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(t.inv_start, start.data(), sizeof(int32_t) * (nsrc + 1), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t.inv_items, items.data(), sizeof(int32_t) * npad, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("halo table upload failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  auto res = g_tables.emplace(key, t);
  return &res.first->second;
}

static int out_edge(int edge_in, int k, int stride, int dilation, int same) {
  const int keff = (k - 1) * dilation + 1;
  const int o = same ? edge_in : edge_in - keff + 1;
  return (o + stride - 1) / stride;
}

int derive_geometry(const dlwpcs_conv_desc *d, Geometry *g) {
  CS_CHECK(d != nullptr, "null descriptor");
  CS_CHECK(d->batch >= 0 && d->n > 0, "bad batch/n (%d,%d)", d->batch, d->n);
  CS_CHECK(d->halo >= 0 && d->halo <= d->n, "halo %d out of range for face edge %d", d->halo, d->n);
  CS_CHECK(d->cin > 0 && d->cout > 0, "bad channel counts (%d,%d)", d->cin, d->cout);
  CS_CHECK(d->kh > 0 && d->kw > 0 && d->stride_h > 0 && d->stride_w > 0 && d->dil_h > 0 && d->dil_w > 0,
           "kernel_size / strides / dilation_rate must be positive");
  CS_CHECK(d->c0 + d->c1 == d->cin && d->c0 > 0 && d->c1 >= 0, "c0 + c1 (%d+%d) must equal cin (%d)", d->c0, d->c1,
           d->cin);
  CS_CHECK(d->mode0 >= 0 && d->mode0 <= 2 && d->mode1 >= 0 && d->mode1 <= 2, "bad source mode");
  if (d->mode0 == DLWPCS_SRC_UP2 || (d->c1 > 0 && d->mode1 == DLWPCS_SRC_UP2))
    CS_CHECK(d->n % 2 == 0, "nearest-upsampled source needs an even face edge");
  g->Hin = g->Win = d->n + 2 * d->halo;
  const int keh = (d->kh - 1) * d->dil_h + 1, kew = (d->kw - 1) * d->dil_w + 1;
  g->Hout = out_edge(g->Hin, d->kh, d->stride_h, d->dil_h, d->same);
  g->Wout = out_edge(g->Win, d->kw, d->stride_w, d->dil_w, d->same);
  CS_CHECK(g->Hout > 0 && g->Wout > 0, "kernel (%dx%d, dilation %dx%d) larger than the face (%d)", d->kh, d->kw,
           d->dil_h, d->dil_w, g->Hin);
  int pt = 0, pl = 0;
  if (d->same) {
    const int th = (g->Hout - 1) * d->stride_h + keh - g->Hin, tw = (g->Wout - 1) * d->stride_w + kew - g->Win;
    pt = (th > 0 ? th : 0) / 2;
    pl = (tw > 0 ? tw : 0) / 2;
  }
  g->pt[0] = g->pt[1] = g->pt[2] = pt;
  // custom.py:969/995 reverse the rows of face 5 before and after the conv.  In unflipped coordinates that is the
  // kernel flipped along kh (done at weight packing) with this top offset -- exact for every stride / 'same' / dilation.
  if (d->flip_north_pole) g->pt[2] = (g->Hout - 1) * d->stride_h + (d->kh - 1) * d->dil_h - (g->Hin - 1) - pt;
  g->pl = pl;
  g->taps = d->kh * d->kw;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// standalone padding kernels
// ---------------------------------------------------------------------------------------------------------------------
template <typename V>
__global__ void pad_fwd_kernel(const V *__restrict__ x, V *__restrict__ y, const int32_t *__restrict__ lut, int npad,
                               int nsrc, int vec_per_px, long long total) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int v = (int)(i % vec_per_px);
    const long long q = i / vec_per_px;
    const int qp = (int)(q % npad);
    const long long b = q / npad;
    y[i] = x[(b * nsrc + lut[qp]) * vec_per_px + v];
  }
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void pad_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, const int32_t *__restrict__ inv_start,
                               const int32_t *__restrict__ inv_items, int npad, int nsrc, int c, long long total) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int ch = (int)(i % c);
    const long long s = i / c;
    const int sp = (int)(s % nsrc);
    const long long b = s / nsrc;
    float acc = 0.f;
    for (int k = inv_start[sp]; k < inv_start[sp + 1]; ++k) acc += to_f<T>(dy[(b * npad + inv_items[k]) * c + ch]);
    dx[i] = from_f<T>(acc);
  }
}

// 16-byte version (channel count a multiple of 8 bf16 / 4 fp32): one thread sums one 16-byte chunk of a source pixel over
// the padded positions that read it (1 interior, 2 edge, 3-5 corner: SURVEY.md appendix A), in the table's fixed order
__device__ __forceinline__ float act_grad_from_y(float y, int act, float slope, float maxv);

// mask_y (optional): the tensor whose halo exchange is being differentiated is the OUTPUT of an activated layer; the
// gathered gradient is multiplied by that activation's derivative here, which saves the upstream layer's separate
// act_bwd pass over the same tensor (dlwpcs_conv2d_dgrad_act)
template <typename T>
__global__ void pad_bwd_vec_kernel(const uint4 *__restrict__ dy, uint4 *__restrict__ dx,
                                   const int32_t *__restrict__ inv_start, const int32_t *__restrict__ inv_items, int npad,
                                   int nsrc, int vpp, long long total, const uint4 *__restrict__ mask_y, int act,
                                   float slope, float maxv) {
  constexpr int E = 16 / sizeof(T);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int v = (int)(i % vpp);
    const long long s = i / vpp;
    const int sp = (int)(s % nsrc);
    const long long b = s / nsrc;
    const int k0 = __ldg(inv_start + sp), k1 = __ldg(inv_start + sp + 1);
    if (k1 - k0 == 1 && !mask_y) {          // interior pixel: plain copy
      dx[i] = __ldg(dy + (b * npad + __ldg(inv_items + k0)) * vpp + v);
      continue;
    }
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    for (int k = k0; k < k1; ++k) {
      const uint4 q = __ldg(dy + (b * npad + __ldg(inv_items + k)) * vpp + v);
      const T *t = reinterpret_cast<const T *>(&q);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] += to_f<T>(t[e]);
    }
    if (mask_y) {
      const uint4 m = __ldg(mask_y + i);
      const T *tm = reinterpret_cast<const T *>(&m);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] *= act_grad_from_y(to_f<T>(tm[e]), act, slope, maxv);
    }
    uint4 o;
    T *t = reinterpret_cast<T *>(&o);
#pragma unroll
    for (int e = 0; e < E; ++e) t[e] = from_f<T>(acc[e]);
    dx[i] = o;
  }
}

__device__ __forceinline__ float act_apply(float v, int act, float slope, float maxv) {
  if (act == DLWPCS_ACT_CAPPED_LEAKY_RELU) v = v < 0.f ? slope * v : fminf(v, maxv);
  return v;
}
// derivative expressed through the OUTPUT y (y<0 <=> x<0 for slope>0; y==max <=> x>=max)
__device__ __forceinline__ float act_grad_from_y(float y, int act, float slope, float maxv) {
  if (act == DLWPCS_ACT_CAPPED_LEAKY_RELU) return y < 0.f ? slope : (y < maxv ? 1.f : 0.f);
  return 1.f;
}

template <typename T>
__global__ void act_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, long long n, int act, float slope,
                               float maxv) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = from_f<T>(act_apply(to_f<T>(x[i]), act, slope, maxv));
}
template <typename T>
__global__ void act_bwd_kernel(const T *__restrict__ dy, const T *__restrict__ y, T *__restrict__ dx, long long n,
                               int act, float slope, float maxv) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dx[i] = from_f<T>(to_f<T>(dy[i]) * act_grad_from_y(to_f<T>(y[i]), act, slope, maxv));
}

// 16 bytes per thread (count a multiple of 8 bf16 / 4 fp32, aligned pointers)
template <typename T, bool BWD>
__global__ void act_vec_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ yv, uint4 *__restrict__ out,
                               long long nvec, int act, float slope, float maxv) {
  constexpr int E = 16 / sizeof(T);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nvec; i += stride) {
    const uint4 av = __ldg(a + i);
    uint4 o;
    const T *ap = reinterpret_cast<const T *>(&av);
    T *op = reinterpret_cast<T *>(&o);
    if (BWD) {
      const uint4 y4 = __ldg(yv + i);
      const T *yp = reinterpret_cast<const T *>(&y4);
#pragma unroll
      for (int e = 0; e < E; ++e) op[e] = from_f<T>(to_f<T>(ap[e]) * act_grad_from_y(to_f<T>(yp[e]), act, slope, maxv));
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) op[e] = from_f<T>(act_apply(to_f<T>(ap[e]), act, slope, maxv));
    }
    out[i] = o;
  }
}

template <typename T>
__global__ void mse_kernel(const T *__restrict__ y, const T *__restrict__ t, T *__restrict__ dy, float *loss, long long n,
                           float inv) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (; i < n; i += stride) {
    const float d = to_f<T>(y[i]) - to_f<T>(t[i]);
    acc = fmaf(d, d, acc);
    dy[i] = from_f<T>(2.f * d * inv);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    atomicAdd(loss, s * inv);
  }
}

__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                            float *__restrict__ v, long long n, float lr_t, float b1, float b2, float eps, float gscale) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---------------------------------------------------------------------------------------------------------------------
// insolation of the forced rollout (DLWP/util.py:306-364 as re-evaluated every forecast iteration by
// TimeSeriesEstimator.predict, DLWP/model/extensions.py:272-288).  Mixed precision exactly like the reference's numpy
// expression: the day and the hour angle in float32 (util.py:340, 348, 357), orbit / declination in float64.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) insolation_kernel(T *__restrict__ out, long long npix, int c_total, int c_first,
                                                         int n_sol, const double *__restrict__ sinlat,
                                                         const double *__restrict__ coslat, const float *__restrict__ lon,
                                                         const double *__restrict__ days, int batch, float S) {
  __shared__ double s_sd[8], s_cd[8], s_r2[8];
  __shared__ float s_day[8];
  const int b = blockIdx.y;
  if (threadIdx.x < n_sol) {
    const double PI = 3.141592653589793;
    const double eps = 23.4441 * PI / 180., ecc = 0.016715, om = 282.7 * PI / 180.;
    const double beta = sqrt(1 - ecc * ecc);
    const float d = (float)days[(size_t)threadIdx.x * batch + b];
    const double lambda_m0 = ecc * (1. + beta) * sin(om);
    const float term = __fdiv_rn(__fmul_rn(6.2831855f, __fsub_rn(d, 80.5f)), 365.f);
    const double lambda_m = lambda_m0 + (double)term;
    const double lambda_ = lambda_m + 2. * ecc * sin(lambda_m - om);
    const double dec = asin(sin(eps) * sin(lambda_));
    const double rho = (1. - ecc * ecc) / (1. + ecc * cos(lambda_ - om));
    s_sd[threadIdx.x] = sin(dec);
    s_cd[threadIdx.x] = cos(dec);
    s_r2[threadIdx.x] = 1.0 / (rho * rho);
    s_day[threadIdx.x] = d;
  }
  __syncthreads();
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const double sl = sinlat[p], cl = coslat[p];
    const float lo = __fdiv_rn(lon[p], 360.f);
    for (int n = 0; n < n_sol; ++n) {
      const float h = __fmul_rn(6.2831855f, __fadd_rn(s_day[n], lo));
      const float ch = (float)cos((double)h);              // float32 cosine of the float32 hour angle, like np.cos
      double sol = (double)S * (sl * s_sd[n] - cl * s_cd[n] * (double)ch) * s_r2[n];
      if (sol < 0.) sol = 0.;
      out[((size_t)b * npix + p) * c_total + c_first + n] = from_f<T>((float)sol);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// resampling steps either side of the conv for the TRAINING path (materialised tensors; the rollout engine folds them into
// the conv's load stage): AveragePooling3D((1,2,2)), UpSampling3D((1,2,2)) + concatenate (Azure/train_cs.py:197-198,
// 282-299) and their adjoints.  16 bytes per thread; channel counts multiples of 8 (bf16) / 4 (fp32).
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
struct Vec16 {
  static constexpr int E = 16 / sizeof(T);
  __device__ static void load(const uint4 *p, float *f) {
    const uint4 q = __ldg(p);
    const T *t = reinterpret_cast<const T *>(&q);
#pragma unroll
    for (int e = 0; e < E; ++e) f[e] = to_f<T>(t[e]);
  }
  __device__ static uint4 pack(const float *f) {
    uint4 o;
    T *t = reinterpret_cast<T *>(&o);
#pragma unroll
    for (int e = 0; e < E; ++e) t[e] = from_f<T>(f[e]);
    return o;
  }
};

// y (BF, n, n, vpp) <- mean of the 2x2 blocks of x (BF, 2n, 2n, vpp);  bwd: dx (BF, 2n, 2n, vpp) <- 0.25 * dy (BF, n, n, vpp)
template <typename T, bool BWD>
__global__ void pool2_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int n, int vpp, long long total) {
  constexpr int E = Vec16<T>::E;
  const int no = BWD ? 2 * n : n;                          // edge of the output
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int v = (int)(i % vpp);
    long long q = i / vpp;
    const int c = (int)(q % no); q /= no;
    const int r = (int)(q % no);
    const long long bf = q / no;
    float a[E];
    if (BWD) {
      Vec16<T>::load(in + ((bf * n + (r >> 1)) * n + (c >> 1)) * vpp + v, a);
#pragma unroll
      for (int e = 0; e < E; ++e) a[e] *= 0.25f;
    } else {
      float t[E];
      const uint4 *src = in + ((bf * 2 * n + 2 * r) * 2 * n + 2 * c) * vpp + v;
      Vec16<T>::load(src, a);
      Vec16<T>::load(src + vpp, t);
#pragma unroll
      for (int e = 0; e < E; ++e) a[e] += t[e];
      Vec16<T>::load(src + (long long)2 * n * vpp, t);
#pragma unroll
      for (int e = 0; e < E; ++e) a[e] += t[e];
      Vec16<T>::load(src + (long long)(2 * n + 1) * vpp, t);
#pragma unroll
      for (int e = 0; e < E; ++e) a[e] = 0.25f * (a[e] + t[e]);
    }
    out[i] = Vec16<T>::pack(a);
  }
}

// t (BF, n, n, va + vb) <- [nearest-upsampled a (BF, n/2, n/2, va) | b (BF, n, n, vb)]
template <typename T>
__global__ void up2cat_fwd_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b, uint4 *__restrict__ t, int n,
                                  int va, int vb, long long total) {
  const int vt = va + vb, h = n >> 1;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int v = (int)(i % vt);
    long long q = i / vt;
    const int c = (int)(q % n); q /= n;
    const int r = (int)(q % n);
    const long long bf = q / n;
    t[i] = v < va ? __ldg(a + ((bf * h + (r >> 1)) * h + (c >> 1)) * va + v)
                  : __ldg(b + ((bf * n + r) * n + c) * vb + (v - va));
  }
}

// float32 -> three-segment bf16 split for the float32-accurate tensor-core convolution (dlwpcs_split3): every value x is
// written as hi = bf16(x) and lo = bf16(x - hi) in the channel layout [hi(C) | lo(C) | hi(C)], so that ONE bf16 GEMM with
// the weights stacked as [w_hi ; w_hi ; w_lo] along K accumulates x_hi*w_hi + x_lo*w_hi + x_hi*w_lo in float32 (the
// dropped x_lo*w_lo term is 2^-16 relative).  The resampling of the U-Net (2x2 mean / nearest up-sampling of source a,
// concatenation with source b) is done here in float32, BEFORE the split.  One thread per pixel and 8 channels.
__global__ void split3_kernel(const float *__restrict__ a, const float *__restrict__ b, uint4 *__restrict__ out, int n,
                              int ca, int cb, int mode, long long total) {
  const int C = ca + cb, vc = C >> 3, va = ca >> 3;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int v = (int)(i % vc);
    long long q = i / vc;
    const int c = (int)(q % n); q /= n;
    const int r = (int)(q % n);
    const long long bf = q / n;
    float x[8];
    if (v < va) {
      if (mode == DLWPCS_SRC_POOL2) {
        const int w2 = 2 * n;
        const float *src = a + (((bf * w2 + 2 * r) * w2 + 2 * c) * (long long)ca + 8 * v);
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const float4 t0 = __ldg(reinterpret_cast<const float4 *>(src + ((long long)dy * w2 + dx) * ca));
            const float4 t1 = __ldg(reinterpret_cast<const float4 *>(src + ((long long)dy * w2 + dx) * ca) + 1);
            x[0] += t0.x; x[1] += t0.y; x[2] += t0.z; x[3] += t0.w; x[4] += t1.x; x[5] += t1.y; x[6] += t1.z; x[7] += t1.w;
          }
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] *= 0.25f;
      } else {
        const int h = mode == DLWPCS_SRC_UP2 ? n >> 1 : n;
        const int rr = mode == DLWPCS_SRC_UP2 ? r >> 1 : r, cc = mode == DLWPCS_SRC_UP2 ? c >> 1 : c;
        const float4 *src = reinterpret_cast<const float4 *>(a + (((bf * h + rr) * h + cc) * (long long)ca + 8 * v));
        const float4 t0 = __ldg(src), t1 = __ldg(src + 1);
        x[0] = t0.x; x[1] = t0.y; x[2] = t0.z; x[3] = t0.w; x[4] = t1.x; x[5] = t1.y; x[6] = t1.z; x[7] = t1.w;
      }
    } else {
      const float4 *src = reinterpret_cast<const float4 *>(b + (((bf * n + r) * n + c) * (long long)cb + 8 * (v - va)));
      const float4 t0 = __ldg(src), t1 = __ldg(src + 1);
      x[0] = t0.x; x[1] = t0.y; x[2] = t0.z; x[3] = t0.w; x[4] = t1.x; x[5] = t1.y; x[6] = t1.z; x[7] = t1.w;
    }
    __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      hi[e] = __float2bfloat16_rn(x[e]);
      lo[e] = __float2bfloat16_rn(x[e] - __bfloat162float(hi[e]));
    }
    uint4 *row = out + (((bf * n + r) * n + c) * 3LL * vc);
    const uint4 h4 = *reinterpret_cast<const uint4 *>(hi);
    row[v] = h4;
    row[vc + v] = *reinterpret_cast<const uint4 *>(lo);
    row[2 * vc + v] = h4;
  }
}

// adjoint: da (BF, n/2, n/2, va) <- sum of the 2x2 blocks of dt[..., :va];  db (BF, n, n, vb) <- dt[..., va:]
template <typename T>
__global__ void up2cat_bwd_kernel(const uint4 *__restrict__ dt, uint4 *__restrict__ da, uint4 *__restrict__ db, int n,
                                  int va, int vb, long long total_a, long long total) {
  constexpr int E = Vec16<T>::E;
  const int vt = va + vb, h = n >> 1;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    if (i < total_a) {
      const int v = (int)(i % va);
      long long q = i / va;
      const int c = (int)(q % h); q /= h;
      const int r = (int)(q % h);
      const long long bf = q / h;
      const uint4 *src = dt + ((bf * n + 2 * r) * n + 2 * c) * vt + v;
      float s[E], u[E];
      Vec16<T>::load(src, s);
      Vec16<T>::load(src + vt, u);
#pragma unroll
      for (int e = 0; e < E; ++e) s[e] += u[e];
      Vec16<T>::load(src + (long long)n * vt, u);
#pragma unroll
      for (int e = 0; e < E; ++e) s[e] += u[e];
      Vec16<T>::load(src + (long long)(n + 1) * vt, u);
#pragma unroll
      for (int e = 0; e < E; ++e) s[e] += u[e];
      da[i] = Vec16<T>::pack(s);
    } else {
      const long long j = i - total_a;
      const int v = (int)(j % vb);
      const long long px = j / vb;
      db[j] = __ldg(dt + px * vt + va + v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// device-side data feed: ArrayDataGenerator.generate (DLWP/model/generators.py:872-984; convolutional, channels_last, no
// sequence) as one transposing gather.  The training array (time, varlev, P) stays in HBM in the reference's layout (one
// plane of P = 6*N*N pixels per variable); a batch is read plane-wise (coalesced along the pixels), transposed in shared
// memory and written channels_last (coalesced along the channels).  blockIdx.z = 0: predictors, 1: targets.
// ---------------------------------------------------------------------------------------------------------------------
struct FeedP {
  const float *array, *insol, *consts;     // (T, V, P), (T, P) or null, (n_const, P) or null
  const long long *samples;                // (B)
  const int *in_vars, *out_vars;           // variable indices of the input / output selections
  void *x, *y;                             // (B, P, cx), (B, P, cy)
  long long P;
  int V, v_in, v_out, t_in, t_out, interval, n_const, has_sol, cx, cy, out_bf16, z0;
};

__global__ void __launch_bounds__(256) feed_gather_kernel(const __grid_constant__ FeedP F) {
  extern __shared__ float tile[];          // [32 pixels][C + 1]
  const bool tgt = blockIdx.z + F.z0 == 1;
  const int C = tgt ? F.cy : F.cx, ld = C + 1;
  const long long p0 = (long long)blockIdx.x * 32;
  const int b = blockIdx.y, px = threadIdx.x & 31, cl = threadIdx.x >> 5;
  const long long s = F.samples[b];
  const int per_t = F.v_in + F.has_sol;
  for (int c = cl; c < C; c += 8) {
    const float *plane;
    if (tgt) {
      const int t = c / F.v_out, v = c - t * F.v_out;
      plane = F.array + ((s + (long long)F.interval * (F.t_in + t)) * F.V + F.out_vars[v]) * F.P;
    } else if (c < F.t_in * per_t) {
      const int t = c / per_t, v = c - t * per_t;
      const long long time = s + (long long)t * F.interval;
      plane = v < F.v_in ? F.array + (time * F.V + F.in_vars[v]) * F.P : F.insol + time * F.P;
    } else {
      plane = F.consts + (long long)(c - F.t_in * per_t) * F.P;
    }
    tile[px * ld + c] = p0 + px < F.P ? __ldg(plane + p0 + px) : 0.f;
  }
  __syncthreads();
  const int npx = (int)((F.P - p0) < 32 ? (F.P - p0) : 32);
  void *out = tgt ? F.y : F.x;
  const long long base = ((long long)b * F.P + p0) * C;
  for (int i = threadIdx.x; i < npx * C; i += 256) {
    const float v = tile[(i / C) * ld + (i % C)];
    if (F.out_bf16) reinterpret_cast<__nv_bfloat16 *>(out)[base + i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float *>(out)[base + i] = v;
  }
}

}  // namespace dlwpcs

using namespace dlwpcs;

extern "C" {

int dlwpcs_version(void) { return DLWPCS_VERSION; }
const char *dlwpcs_last_error(void) { return g_err; }

int dlwpcs_conv_out_edge(int edge_in, int k, int stride, int dilation, int same) {
  if (edge_in <= 0 || k <= 0 || stride <= 0 || dilation <= 0) return -1;
  return out_edge(edge_in, k, stride, dilation, same);
}

int dlwpcs_pad_lut_host(int n, int p, int32_t *lut) {
  CS_CHECK(n > 0 && p >= 0 && p <= n && lut != nullptr, "bad arguments to dlwpcs_pad_lut_host (n=%d, p=%d)", n, p);
  std::vector<int32_t> v;
  build_pad_lut(n, p, v);
  memcpy(lut, v.data(), v.size() * sizeof(int32_t));
  return 0;
}

static int elem_size(int dtype) { return dtype == DLWPCS_F32 ? 4 : (dtype == DLWPCS_BF16 ? 2 : 0); }

int dlwpcs_pad_fwd(const void *x, void *y, int batch, int n, int c, int p, int dtype, void *stream) {
  const int es = elem_size(dtype);
  CS_CHECK(es != 0, "unsupported dtype %d", dtype);
  CS_CHECK(batch >= 0 && n > 0 && c > 0 && p >= 0 && p <= n, "bad arguments to dlwpcs_pad_fwd");
  if (batch == 0) return 0;
  CS_CHECK(x && y, "null tensor pointer");
  const HaloTables *t = get_halo_tables(n, p);
  if (!t) return 3;
  cudaStream_t st = (cudaStream_t)stream;
  const int H = n + 2 * p, npad = 6 * H * H, nsrc = 6 * n * n;
  const int row_bytes = c * es;
  const bool al16 = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  if (row_bytes % 16 == 0 && al16) {
    const int vpp = row_bytes / 16;
    const long long total = (long long)batch * npad * vpp;
    pad_fwd_kernel<uint4><<<grid_for(total, 256), 256, 0, st>>>((const uint4 *)x, (uint4 *)y, t->lut, npad, nsrc, vpp,
                                                                  total);
  } else if (row_bytes % 4 == 0) {
    const int vpp = row_bytes / 4;
    const long long total = (long long)batch * npad * vpp;
    pad_fwd_kernel<uint32_t><<<grid_for(total, 256), 256, 0, st>>>((const uint32_t *)x, (uint32_t *)y, t->lut, npad,
                                                                     nsrc, vpp, total);
  } else {
    const int vpp = row_bytes / 2;
    const long long total = (long long)batch * npad * vpp;
    pad_fwd_kernel<uint16_t><<<grid_for(total, 256), 256, 0, st>>>((const uint16_t *)x, (uint16_t *)y, t->lut, npad,
                                                                     nsrc, vpp, total);
  }
  CS_CUDA(cudaGetLastError());
  return 0;
}

int dlwpcs_pad_bwd(const void *dy, void *dx, int batch, int n, int c, int p, int dtype, void *stream) {
  return dlwpcs_pad_bwd_act(dy, nullptr, dx, batch, n, c, p, DLWPCS_ACT_NONE, 0.f, 0.f, dtype, stream);
}

int dlwpcs_pad_bwd_act(const void *dy, const void *y_in, void *dx, int batch, int n, int c, int p, int act, float slope,
                       float maxv, int dtype, void *stream) {
  CS_CHECK(elem_size(dtype) != 0, "unsupported dtype %d", dtype);
  CS_CHECK(batch >= 0 && n > 0 && c > 0 && p >= 0 && p <= n, "bad arguments to dlwpcs_pad_bwd");
  if (act == DLWPCS_ACT_NONE) y_in = nullptr;
  if (batch == 0) return 0;
  CS_CHECK(dy && dx, "null tensor pointer");
  const HaloTables *t = get_halo_tables(n, p);
  if (!t) return 3;
  cudaStream_t st = (cudaStream_t)stream;
  const int H = n + 2 * p, npad = 6 * H * H, nsrc = 6 * n * n;
  const long long total = (long long)batch * nsrc * c;
  const int epv = 16 / (int)elem_size(dtype);       // elements per 16-byte chunk
  if (c % epv == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(y_in) & 15) == 0) {
    const int vpp = c / epv;
    const long long tv = total / epv;
    if (dtype == DLWPCS_F32)
      pad_bwd_vec_kernel<float><<<grid_for(tv, 256), 256, 0, st>>>((const uint4 *)dy, (uint4 *)dx, t->inv_start,
                                                                     t->inv_items, npad, nsrc, vpp, tv, (const uint4 *)y_in,
                                                                     act, slope, maxv);
    else
      pad_bwd_vec_kernel<__nv_bfloat16><<<grid_for(tv, 256), 256, 0, st>>>((const uint4 *)dy, (uint4 *)dx, t->inv_start,
                                                                             t->inv_items, npad, nsrc, vpp, tv,
                                                                             (const uint4 *)y_in, act, slope, maxv);
    CS_CUDA(cudaGetLastError());
    return 0;
  }
  if (dtype == DLWPCS_F32)
    pad_bwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float *)dy, (float *)dx, t->inv_start,
                                                                  t->inv_items, npad, nsrc, c, total);
  else
    pad_bwd_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(
        (const __nv_bfloat16 *)dy, (__nv_bfloat16 *)dx, t->inv_start, t->inv_items, npad, nsrc, c, total);
  CS_CUDA(cudaGetLastError());
  if (y_in) return dlwpcs_act_bwd(dx, y_in, dx, total, act, slope, maxv, dtype, stream);      // odd channel counts
  return 0;
}

int dlwpcs_act_fwd(const void *x, void *y, int64_t count, int act, float slope, float maxv, int dtype, void *stream) {
  CS_CHECK(elem_size(dtype) != 0 && x && y && count >= 0, "bad arguments to dlwpcs_act_fwd");
  if (count == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLWPCS_F32)
    act_fwd_kernel<float><<<grid_for(count, 256), 256, 0, st>>>((const float *)x, (float *)y, count, act, slope, maxv);
  else
    act_fwd_kernel<__nv_bfloat16><<<grid_for(count, 256), 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y,
                                                                          count, act, slope, maxv);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int dlwpcs_act_bwd(const void *dy, const void *y, void *dx, int64_t count, int act, float slope, float maxv, int dtype,
                   void *stream) {
  CS_CHECK(elem_size(dtype) != 0 && dy && y && dx && count >= 0, "bad arguments to dlwpcs_act_bwd");
  if (count == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int epv = 16 / (int)elem_size(dtype);
  if (count % epv == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
    const long long nv = count / epv;
    if (dtype == DLWPCS_F32)
      act_vec_kernel<float, true><<<grid_for(nv, 256), 256, 0, st>>>((const uint4 *)dy, (const uint4 *)y, (uint4 *)dx, nv, act,
                                                                       slope, maxv);
    else
      act_vec_kernel<__nv_bfloat16, true><<<grid_for(nv, 256), 256, 0, st>>>((const uint4 *)dy, (const uint4 *)y, (uint4 *)dx,
                                                                               nv, act, slope, maxv);
    CS_CUDA(cudaGetLastError());
    return 0;
  }
  if (dtype == DLWPCS_F32)
    act_bwd_kernel<float><<<grid_for(count, 256), 256, 0, st>>>((const float *)dy, (const float *)y, (float *)dx, count,
                                                                  act, slope, maxv);
  else
    act_bwd_kernel<__nv_bfloat16><<<grid_for(count, 256), 256, 0, st>>>(
        (const __nv_bfloat16 *)dy, (const __nv_bfloat16 *)y, (__nv_bfloat16 *)dx, count, act, slope, maxv);
  CS_CUDA(cudaGetLastError());
  return 0;
}

static bool vec_ok(int dtype, std::initializer_list<int> channels, std::initializer_list<const void *> ptrs) {
  const int epv = 16 / (int)elem_size(dtype);
  for (int c : channels)
    if (c % epv) return false;
  for (const void *p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return false;
  return true;
}

int dlwpcs_pool2(const void *in, void *out, int batch, int n, int c, int backward, int dtype, void *stream) {
  CS_CHECK(elem_size(dtype) != 0 && in && out && batch >= 0 && n > 0 && c > 0, "bad arguments to dlwpcs_pool2");
  CS_CHECK(vec_ok(dtype, {c}, {in, out}), "dlwpcs_pool2 needs 16-byte aligned tensors and a channel count that fills 16-byte chunks");
  if (batch == 0) return 0;
  const int vpp = c / (16 / (int)elem_size(dtype)), no = backward ? 2 * n : n;
  const long long total = (long long)batch * 6 * no * no * vpp;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(total, 256);
  if (dtype == DLWPCS_F32) {
    if (backward) pool2_kernel<float, true><<<g, 256, 0, st>>>((const uint4 *)in, (uint4 *)out, n, vpp, total);
    else pool2_kernel<float, false><<<g, 256, 0, st>>>((const uint4 *)in, (uint4 *)out, n, vpp, total);
  } else {
    if (backward) pool2_kernel<__nv_bfloat16, true><<<g, 256, 0, st>>>((const uint4 *)in, (uint4 *)out, n, vpp, total);
    else pool2_kernel<__nv_bfloat16, false><<<g, 256, 0, st>>>((const uint4 *)in, (uint4 *)out, n, vpp, total);
  }
  CS_CUDA(cudaGetLastError());
  return 0;
}

int dlwpcs_up2cat_fwd(const void *a, const void *b, void *t, int batch, int n, int ca, int cb, int dtype, void *stream) {
  CS_CHECK(elem_size(dtype) != 0 && a && b && t && batch >= 0 && n > 0 && n % 2 == 0 && ca > 0 && cb > 0,
           "bad arguments to dlwpcs_up2cat_fwd");
  CS_CHECK(vec_ok(dtype, {ca, cb}, {a, b, t}), "dlwpcs_up2cat_fwd needs 16-byte aligned tensors and channel counts that fill 16-byte chunks");
  if (batch == 0) return 0;
  const int epv = 16 / (int)elem_size(dtype), va = ca / epv, vb = cb / epv;
  const long long total = (long long)batch * 6 * n * n * (va + vb);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLWPCS_F32)
    up2cat_fwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const uint4 *)a, (const uint4 *)b, (uint4 *)t, n, va, vb, total);
  else
    up2cat_fwd_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>((const uint4 *)a, (const uint4 *)b, (uint4 *)t, n, va, vb,
                                                                            total);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int dlwpcs_split3(const float *a, int ca, int mode_a, const float *b, int cb, void *out, int batch, int n, void *stream) {
  CS_CHECK(a && out && batch >= 0 && n > 0 && ca > 0 && cb >= 0 && (cb == 0 || b), "bad arguments to dlwpcs_split3");
  CS_CHECK(ca % 8 == 0 && cb % 8 == 0, "dlwpcs_split3: channel counts must be multiples of 8 (got %d + %d)", ca, cb);
  CS_CHECK(mode_a == DLWPCS_SRC_SAME || mode_a == DLWPCS_SRC_POOL2 || (mode_a == DLWPCS_SRC_UP2 && n % 2 == 0),
           "dlwpcs_split3: bad sampling mode %d for face edge %d", mode_a, n);
  CS_CHECK(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
           "dlwpcs_split3 needs 16-byte aligned tensors");
  if (batch == 0) return 0;
  const long long total = (long long)batch * 6 * n * n * ((ca + cb) / 8);
  split3_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a, b, (uint4 *)out, n, ca, cb, mode_a, total);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int dlwpcs_up2cat_bwd(const void *dt, void *da, void *db, int batch, int n, int ca, int cb, int dtype, void *stream) {
  CS_CHECK(elem_size(dtype) != 0 && dt && da && db && batch >= 0 && n > 0 && n % 2 == 0 && ca > 0 && cb > 0,
           "bad arguments to dlwpcs_up2cat_bwd");
  CS_CHECK(vec_ok(dtype, {ca, cb}, {dt, da, db}), "dlwpcs_up2cat_bwd needs 16-byte aligned tensors and channel counts that fill 16-byte chunks");
  if (batch == 0) return 0;
  const int epv = 16 / (int)elem_size(dtype), va = ca / epv, vb = cb / epv;
  const long long total_a = (long long)batch * 6 * (n / 2) * (n / 2) * va;
  const long long total = total_a + (long long)batch * 6 * n * n * vb;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLWPCS_F32)
    up2cat_bwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const uint4 *)dt, (uint4 *)da, (uint4 *)db, n, va, vb, total_a,
                                                                    total);
  else
    up2cat_bwd_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>((const uint4 *)dt, (uint4 *)da, (uint4 *)db, n, va, vb,
                                                                            total_a, total);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int dlwpcs_feed_gather(const float *array, const float *insolation, const float *constants, const int64_t *samples,
                       const int32_t *in_vars, const int32_t *out_vars, void *x, void *y, int batch, int64_t npix,
                       int n_var, int v_in, int v_out, int t_in, int t_out, int interval, int n_const, int dtype,
                       void *stream) {
  // x == NULL or y == NULL: only the other part is produced (sequence mode asks for the targets of later steps alone)
  CS_CHECK(elem_size(dtype) != 0 && array && samples && in_vars && out_vars && (x || y), "null pointer / bad dtype in dlwpcs_feed_gather");
  CS_CHECK(batch >= 0 && npix > 0 && n_var > 0 && v_in > 0 && v_out > 0 && t_in > 0 && t_out > 0 && interval > 0 &&
               n_const >= 0 && (n_const == 0 || constants),
           "bad arguments to dlwpcs_feed_gather");
  if (batch == 0) return 0;
  FeedP F;
  F.array = array; F.insol = insolation; F.consts = constants;
  F.samples = (const long long *)samples; F.in_vars = in_vars; F.out_vars = out_vars;
  F.x = x; F.y = y; F.P = npix; F.V = n_var; F.v_in = v_in; F.v_out = v_out; F.t_in = t_in; F.t_out = t_out;
  F.interval = interval; F.n_const = n_const; F.has_sol = insolation ? 1 : 0;
  F.cx = t_in * (v_in + F.has_sol) + n_const;
  F.cy = t_out * v_out;
  F.out_bf16 = dtype == DLWPCS_BF16;
  const int cmax = F.cx > F.cy ? F.cx : F.cy;
  CS_CHECK(cmax <= 368, "more than 368 channels per sample (shared-memory transpose tile)");
  F.z0 = x ? 0 : 1;
  dim3 grid((unsigned)((npix + 31) / 32), (unsigned)batch, (x && y) ? 2 : 1);
  feed_gather_kernel<<<grid, 256, (size_t)32 * (cmax + 1) * sizeof(float), (cudaStream_t)stream>>>(F);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int dlwpcs_insolation(void *out, int dtype, int batch, int64_t npix, int c_total, int c_first, int n_sol,
                      const double *sinlat, const double *coslat, const float *lon, const double *days, float S,
                      void *stream) {
  CS_CHECK(elem_size(dtype) != 0 && out && sinlat && coslat && lon && days, "bad arguments to dlwpcs_insolation");
  CS_CHECK(batch >= 0 && npix > 0 && n_sol >= 1 && n_sol <= 8 && c_first >= 0 && c_first + n_sol <= c_total,
           "bad channel range for dlwpcs_insolation");
  if (batch == 0) return 0;
  dim3 grid((unsigned)((npix + 255) / 256 > 64 ? 64 : (npix + 255) / 256), (unsigned)batch);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLWPCS_F32)
    insolation_kernel<float><<<grid, 256, 0, st>>>((float *)out, npix, c_total, c_first, n_sol, sinlat, coslat, lon, days,
                                                  batch, S);
  else
    insolation_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16 *)out, npix, c_total, c_first, n_sol, sinlat, coslat,
                                                          lon, days, batch, S);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int dlwpcs_mse_loss_grad(const void *y, const void *t, void *dy, float *loss_accum, int64_t count, float inv_count,
                         int dtype, void *stream) {
  CS_CHECK(elem_size(dtype) != 0 && y && t && dy && loss_accum && count >= 0, "bad arguments to dlwpcs_mse_loss_grad");
  if (count == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLWPCS_F32)
    mse_kernel<float><<<grid_for(count, 256), 256, 0, st>>>((const float *)y, (const float *)t, (float *)dy, loss_accum,
                                                             count, inv_count);
  else
    mse_kernel<__nv_bfloat16><<<grid_for(count, 256), 256, 0, st>>>(
        (const __nv_bfloat16 *)y, (const __nv_bfloat16 *)t, (__nv_bfloat16 *)dy, loss_accum, count, inv_count);
  CS_CUDA(cudaGetLastError());
  return 0;
}

// step counter on the device: graph-replayable (a captured launch cannot carry a changing host scalar)
__global__ void adam_dev_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                float *__restrict__ v, long long n, float lr, float b1, float b2, float eps, float gscale,
                                const int32_t *__restrict__ step) {
  const double t = (double)(*step);
  const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}
__global__ void incr_kernel(int32_t *c) { *c += 1; }

int dlwpcs_trace_read(unsigned long long *host_out, int max_launches, int *n_launches, int reset) {
  return tc_trace_read(host_out, max_launches, n_launches, reset);
}

int dlwpcs_adam_step_dev(float *param, const float *grad, float *m, float *v, int64_t count, float lr, float beta1,
                         float beta2, float eps, int32_t *step_counter, float grad_scale, void *stream) {
  CS_CHECK(param && grad && m && v && step_counter && count >= 0, "bad arguments to dlwpcs_adam_step_dev");
  incr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_counter);
  if (count > 0)
    adam_dev_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, count, lr, beta1, beta2, eps,
                                                                             grad_scale, step_counter);
  CS_CUDA(cudaGetLastError());
  return 0;
}

int dlwpcs_adam_step(float *param, const float *grad, float *m, float *v, int64_t count, float lr, float beta1,
                     float beta2, float eps, int step, float grad_scale, void *stream) {
  CS_CHECK(param && grad && m && v && count >= 0 && step >= 1, "bad arguments to dlwpcs_adam_step");
  if (count == 0) return 0;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  adam_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, count, (float)lr_t, beta1, beta2,
                                                                       eps, grad_scale);
  CS_CUDA(cudaGetLastError());
  return 0;
}

static int check_common(const dlwpcs_conv_desc *d, Geometry *g) {
  if (int rc = derive_geometry(d, g)) return rc;
  CS_CHECK(elem_size(d->x_dtype) && elem_size(d->y_dtype), "unsupported dtype");
  return 0;
}

static int64_t fp32_packed_floats(const dlwpcs_conv_desc *d, int transposed) {
  const int64_t w = 3LL * d->kh * d->kw * d->cin * d->cout;
  return transposed ? w : w + 3LL * d->cout;
}

int64_t dlwpcs_packed_weight_bytes(const dlwpcs_conv_desc *d, int transposed) {
  Geometry g;
  if (check_common(d, &g)) return -1;
  if (d->x_dtype == DLWPCS_BF16) return tc_packed_weight_bytes(d, g, transposed);
  return fp32_packed_floats(d, transposed == 1) * 4;
}

int dlwpcs_pack_weights(const dlwpcs_conv_desc *d, const dlwpcs_conv_weights *w, int transposed, void *packed,
                        void *stream) {
  Geometry g;
  if (int rc = check_common(d, &g)) return rc;
  CS_CHECK(w && packed && w->w_eq && w->w_pol, "null weights");
  CS_CHECK(!d->independent_north_pole || w->w_np, "independent_north_pole needs w_np");
  CS_CHECK(transposed == 1 || !d->use_bias || (w->b_eq && w->b_pol && (!d->independent_north_pole || w->b_np)),
           "use_bias needs biases");
  if (d->x_dtype == DLWPCS_BF16) return tc_pack_weights(d, g, w, transposed, packed, (cudaStream_t)stream);
  return fp32_pack_weights(d, w, transposed == 1, (float *)packed, (cudaStream_t)stream);
}

int dlwpcs_pack_weights2(const dlwpcs_conv_desc *d, const dlwpcs_conv_weights *w, int src_cin, int src_cout, void *packed,
                         void *packed_t, void *stream) {
  Geometry g;
  if (int rc = check_common(d, &g)) return rc;
  CS_CHECK(d->x_dtype == DLWPCS_BF16, "dlwpcs_pack_weights2 serves the bf16 tensor-core path");
  CS_CHECK(w && w->w_eq && w->w_pol && (packed || packed_t), "null weights");
  CS_CHECK(!d->independent_north_pole || w->w_np, "independent_north_pole needs w_np");
  CS_CHECK(!packed || !d->use_bias || (w->b_eq && w->b_pol && (!d->independent_north_pole || w->b_np)),
           "use_bias needs biases");
  return tc_pack_weights2(d, g, w, src_cin, src_cout, packed, packed_t, (cudaStream_t)stream);
}

int dlwpcs_pack_weights_batch(const dlwpcs_pack_item *items, int n_items, void *stream) {
  CS_CHECK(items != nullptr && n_items >= 0, "null items");
  for (int i = 0; i < n_items; ++i) {
    const dlwpcs_pack_item &it = items[i];
    Geometry g;
    CS_CHECK(it.desc != nullptr, "item %d: null descriptor", i);
    if (int rc = check_common(it.desc, &g)) return rc;
    CS_CHECK(it.desc->x_dtype == DLWPCS_BF16, "dlwpcs_pack_weights_batch serves the bf16 tensor-core path");
    CS_CHECK(it.w.w_eq && it.w.w_pol && (it.packed || it.packed_t), "item %d: null weights", i);
    CS_CHECK(!it.desc->independent_north_pole || it.w.w_np, "item %d: independent_north_pole needs w_np", i);
    CS_CHECK(!it.packed || !it.desc->use_bias || (it.w.b_eq && it.w.b_pol && (!it.desc->independent_north_pole || it.w.b_np)),
             "item %d: use_bias needs biases", i);
  }
  return tc_pack_batch(items, n_items, (cudaStream_t)stream);
}

int dlwpcs_conv2d_fwd(const dlwpcs_conv_desc *d, const void *x0, const void *x1, const void *packed_w, void *y,
                      void *stream) {
  Geometry g;
  if (int rc = check_common(d, &g)) return rc;
  if (d->batch == 0) return 0;
  CS_CHECK(x0 && packed_w && y && (d->c1 == 0 || x1), "null tensor pointer");
  if (d->x_dtype == DLWPCS_BF16) {
    const char *why = "";
    CS_CHECK(tc_supported(d, g, &why), "bf16 tensor-core path does not support this configuration: %s", why);
    return tc_conv_fwd(d, g, x0, x1, packed_w, y, nullptr, (cudaStream_t)stream);
  }
  CS_CHECK(d->y_dtype == DLWPCS_F32, "float32 input requires float32 output");
  return fp32_conv_fwd(d, g, (const float *)x0, (const float *)x1, (const float *)packed_w, (float *)y,
                       (cudaStream_t)stream);
}

int dlwpcs_rs_work_cuts(const dlwpcs_conv_desc *d, int grid, int32_t *cut_s, int32_t *cut_y) {
  Geometry g;
  if (!d || !cut_s || !cut_y || check_common(d, &g)) return 0;
  return rs_debug_cuts(d, g, grid, cut_s, cut_y);
}

int dlwpcs_conv2d_pool_fusable(const dlwpcs_conv_desc *d) {
  Geometry g;
  if (!d || check_common(d, &g)) return 0;
  return rs_pool_eligible(d, g) ? 1 : 0;
}

int dlwpcs_conv2d_fwd_pool(const dlwpcs_conv_desc *d, const void *x0, const void *x1, const void *packed_w, void *y,
                           void *y_pool, void *stream) {
  Geometry g;
  if (int rc = check_common(d, &g)) return rc;
  CS_CHECK(rs_pool_eligible(d, g), "this layer cannot write its pooled output itself (dlwpcs_conv2d_pool_fusable)");
  if (d->batch == 0) return 0;
  CS_CHECK(x0 && packed_w && y && y_pool && (d->c1 == 0 || x1), "null tensor pointer");
  return rs_conv_fwd_pool(d, g, x0, x1, packed_w, y, y_pool, (cudaStream_t)stream);
}

int dlwpcs_conv2d_head_fusable(const dlwpcs_conv_desc *d, const dlwpcs_conv_desc *head) {
  Geometry g, gh;
  if (!d || !head || check_common(d, &g) || check_common(head, &gh)) return 0;
  return rs_head_eligible(d, g, head) ? 1 : 0;
}

int dlwpcs_conv2d_fwd_head(const dlwpcs_conv_desc *d, const void *x0, const void *x1, const void *packed_w,
                           const dlwpcs_conv_desc *head, const void *head_packed_w, void *y_head, void *stream) {
  Geometry g, gh;
  if (int rc = check_common(d, &g)) return rc;
  CS_CHECK(head != nullptr, "null head descriptor");
  if (int rc = check_common(head, &gh)) return rc;
  CS_CHECK(rs_head_eligible(d, g, head), "this pair of layers cannot be fused (dlwpcs_conv2d_head_fusable)");
  if (d->batch == 0) return 0;
  CS_CHECK(x0 && packed_w && head_packed_w && y_head && (d->c1 == 0 || x1), "null tensor pointer");
  return rs_conv_fwd_head(d, g, x0, x1, packed_w, head, head_packed_w, y_head, (cudaStream_t)stream);
}

int dlwpcs_conv2d_fwd_chained(const dlwpcs_conv_desc *d, const void *x0, const void *x1, const void *packed_w, void *y,
                              const dlwpcs_chain *chain, void *stream) {
  Geometry g;
  if (int rc = check_common(d, &g)) return rc;
  CS_CHECK(chain != nullptr, "null chain descriptor");
  CS_CHECK(d->x_dtype == DLWPCS_BF16, "chained launches exist on the bf16 tensor-core path only");
  if (d->batch == 0) return 0;
  CS_CHECK(x0 && packed_w && y && (d->c1 == 0 || x1), "null tensor pointer");
  const char *why = "";
  CS_CHECK(tc_supported(d, g, &why), "bf16 tensor-core path does not support this configuration: %s", why);
  return tc_conv_fwd(d, g, x0, x1, packed_w, y, chain, (cudaStream_t)stream);
}

uint32_t dlwpcs_chain_target(const dlwpcs_conv_desc *d) {
  Geometry g;
  if (check_common(d, &g) || d->x_dtype != DLWPCS_BF16) return 0;
  return tc_chain_target(d, g);
}

static int check_bwd(const dlwpcs_conv_desc *d, Geometry *g) {
  if (int rc = check_common(d, g)) return rc;
  CS_CHECK(d->x_dtype == d->y_dtype, "backward needs activations and outputs of one dtype (float32 or bfloat16)");
  CS_CHECK(d->stride_h == 1 && d->stride_w == 1, "backward is implemented for strides == 1 only (all cubed-sphere "
           "models in the reference use stride 1: Azure/train_cs.py:200-207)");
  CS_CHECK(d->c1 == 0 && d->mode0 == DLWPCS_SRC_SAME, "backward takes a single un-resampled input source");
  return 0;
}

int64_t dlwpcs_dgrad_workspace_bytes(const dlwpcs_conv_desc *d) {
  Geometry g;
  if (check_bwd(d, &g)) return -1;
  return d->halo > 0 ? (int64_t)d->batch * 6 * g.Hin * g.Win * d->cin * elem_size(d->x_dtype) : 0;
}

int dlwpcs_conv2d_dgrad(const dlwpcs_conv_desc *d, const void *dy, const void *y, const void *packed_w_t, void *dx,
                        void *workspace, void *stream) {
  Geometry g;
  if (int rc = check_bwd(d, &g)) return rc;
  CS_CHECK(dy && packed_w_t && dx && (d->halo == 0 || workspace), "null pointer");
  CS_CHECK(d->act == DLWPCS_ACT_NONE || y, "activation derivative needs the forward output y");
  if (d->batch == 0) return 0;
  if (d->x_dtype == DLWPCS_BF16)
    return tc_conv_dgrad(d, g, dy, y, packed_w_t, dx, workspace, nullptr, DLWPCS_ACT_NONE, 0.f, 0.f, (cudaStream_t)stream);
  return fp32_conv_dgrad(d, g, (const float *)dy, (const float *)y, (const float *)packed_w_t, (float *)dx, workspace,
                         (cudaStream_t)stream);
}

int dlwpcs_conv2d_dgrad_act(const dlwpcs_conv_desc *d, const void *dy, const void *y, const void *packed_w_t, void *dx,
                            void *workspace, const void *x_in, int in_act, float in_slope, float in_max, void *stream) {
  Geometry g;
  if (int rc = check_bwd(d, &g)) return rc;
  CS_CHECK(in_act == DLWPCS_ACT_NONE || x_in, "the input activation's derivative needs the layer input x_in");
  if (in_act == DLWPCS_ACT_NONE || d->batch == 0) return dlwpcs_conv2d_dgrad(d, dy, y, packed_w_t, dx, workspace, stream);
  if (d->x_dtype == DLWPCS_BF16 && d->halo > 0) {
    CS_CHECK(dy && packed_w_t && dx && workspace, "null pointer");
    CS_CHECK(d->act == DLWPCS_ACT_NONE || y, "activation derivative needs the forward output y");
    return tc_conv_dgrad(d, g, dy, y, packed_w_t, dx, workspace, x_in, in_act, in_slope, in_max, (cudaStream_t)stream);
  }
  // no halo scatter-add to fuse into (1x1 layers) or the float32 kernels: plain dgrad, then the mask in place
  if (int rc = dlwpcs_conv2d_dgrad(d, dy, y, packed_w_t, dx, workspace, stream)) return rc;
  return dlwpcs_act_bwd(dx, x_in, dx, (int64_t)d->batch * 6 * d->n * d->n * d->cin, in_act, in_slope, in_max, d->x_dtype,
                        stream);
}

int64_t dlwpcs_wgrad_workspace_bytes(const dlwpcs_conv_desc *d) {
  Geometry g;
  if (check_bwd(d, &g)) return -1;
  if (d->x_dtype == DLWPCS_BF16 && tc_wgrad_supported(d, g)) return tc_wgrad_workspace_bytes(d, g);
  return fp32_wgrad_workspace_bytes(d, g);
}

int dlwpcs_conv2d_wgrad(const dlwpcs_conv_desc *d, const void *x0, const void *dy, const void *y,
                        const dlwpcs_conv_wgrads *out, void *workspace, void *stream) {
  Geometry g;
  if (int rc = check_bwd(d, &g)) return rc;
  CS_CHECK(x0 && dy && out && workspace && out->dw_eq && out->dw_pol, "null pointer");
  CS_CHECK(!d->independent_north_pole || out->dw_np, "independent_north_pole needs dw_np");
  CS_CHECK(!d->use_bias || (out->db_eq && out->db_pol && (!d->independent_north_pole || out->db_np)),
           "use_bias needs bias gradient buffers");
  CS_CHECK(d->act == DLWPCS_ACT_NONE || y, "activation derivative needs the forward output y");
  CS_CHECK(d->batch > 0, "wgrad needs a non-empty batch");
  // bf16 activations: tcgen05 kernel (cs_wgrad_tc.cu); float32 (and shapes outside its reach): CUDA-core kernel
  if (d->x_dtype == DLWPCS_BF16 && tc_wgrad_supported(d, g))
    return tc_conv_wgrad(d, g, x0, dy, y, out, workspace, (cudaStream_t)stream);
  return fp32_conv_wgrad(d, g, x0, dy, y, out, workspace,
                         (cudaStream_t)stream);
}

int dlwpcs_conv2d_fwd_host(const dlwpcs_conv_desc *d, const dlwpcs_conv_weights *w, const void *x_host, void *y_host) {
  Geometry g;
  if (int rc = check_common(d, &g)) return rc;
  CS_CHECK(w && x_host && y_host, "null pointer");
  CS_CHECK(d->c1 == 0 && d->mode0 == DLWPCS_SRC_SAME, "host entry point takes one un-resampled input");
  const int kk = d->kh * d->kw;
  const size_t wbytes = (size_t)kk * d->cin * d->cout * 4, bbytes = (size_t)d->cout * 4;
  const size_t xbytes = (size_t)d->batch * 6 * d->n * d->n * d->cin * elem_size(d->x_dtype);
  const size_t ybytes = (size_t)d->batch * 6 * g.Hout * g.Wout * d->cout * elem_size(d->y_dtype);
  const int64_t pbytes = dlwpcs_packed_weight_bytes(d, 0);
  if (pbytes < 0) return 1;
  if (d->halo > 0 && !get_halo_tables(d->n, d->halo)) return 3;
  char *dev = nullptr;
  const size_t wtot = 3 * wbytes + 3 * bbytes;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t total = al(wtot) + al((size_t)pbytes) + al(xbytes) + al(ybytes);
  CS_CUDA(cudaMalloc(&dev, total));
  char *dw = dev, *dp = dw + al(wtot), *dxp = dp + al((size_t)pbytes), *dyp = dxp + al(xbytes);
  dlwpcs_conv_weights dv = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int rc = 0;
  cudaError_t e = cudaSuccess;
  auto up = [&](const float *src, size_t off, size_t bytes) -> const float * {
    if (!src) return nullptr;
    if (e == cudaSuccess) e = cudaMemcpy(dw + off, src, bytes, cudaMemcpyHostToDevice);
    return (const float *)(dw + off);
  };
  dv.w_eq = up(w->w_eq, 0, wbytes);
  dv.w_pol = up(w->w_pol, wbytes, wbytes);
  dv.w_np = up(w->w_np, 2 * wbytes, wbytes);
  dv.b_eq = up(w->b_eq, 3 * wbytes, bbytes);
  dv.b_pol = up(w->b_pol, 3 * wbytes + bbytes, bbytes);
  dv.b_np = up(w->b_np, 3 * wbytes + 2 * bbytes, bbytes);
  if (e == cudaSuccess) e = cudaMemcpy(dxp, x_host, xbytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("host->device copy failed: %s", cudaGetErrorString(e));
    rc = 2;
  }
  if (!rc) rc = dlwpcs_pack_weights(d, &dv, 0, dp, nullptr);
  if (!rc) rc = dlwpcs_conv2d_fwd(d, dxp, nullptr, dp, dyp, nullptr);
  if (!rc) {
    e = cudaMemcpy(y_host, dyp, ybytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
      set_error("device->host copy failed: %s", cudaGetErrorString(e));
      rc = 2;
    }
  }
  cudaFree(dev);
  return rc;
}

}  // extern "C"
