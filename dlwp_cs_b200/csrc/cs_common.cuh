// Shared declarations of libdlwpcs (internal).  See include/dlwpcs.h for the public C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <vector>
#include "dlwpcs.h"

namespace dlwpcs {

void set_error(const char *fmt, ...);

#define CS_CHECK(cond, ...)          \
  do {                               \
    if (!(cond)) {                   \
      ::dlwpcs::set_error(__VA_ARGS__); \
      return 1;                      \
    }                                \
  } while (0)

#define CS_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::dlwpcs::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));             \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

// ---- halo tables (CubeSpherePadding2D as an index map), cached per (device, n, p) --------------------------------
struct HaloTables {
  int n, p;
  int32_t *lut;        // device [6*(n+2p)^2]  padded position -> flat source pixel f*n*n + i*n + j
  int32_t *inv_start;  // device [6*n*n + 1]   CSR over source pixels
  int32_t *inv_items;  // device [6*(n+2p)^2]  padded positions reading each source pixel
};
void build_pad_lut(int n, int p, std::vector<int32_t> &lut);
const HaloTables *get_halo_tables(int n, int p);  // nullptr on error (message set)

// ---- geometry derived from a conv descriptor ----------------------------------------------------------------------
struct Geometry {
  int Hin, Win;     // edge the conv sees (n + 2*halo)
  int Hout, Wout;
  int pt[3];        // zero rows above, per face group {equatorial, south pole, north pole}; may be negative (offset)
  int pl;           // zero columns left
  int taps;
};
int derive_geometry(const dlwpcs_conv_desc *d, Geometry *g);

static inline int face_group_host(int f) { return f < 4 ? 0 : f - 3; }

// per-device one-time state (opt-in shared-memory attributes, SM counts) is indexed by the current device
constexpr int kMaxDevices = 64;
static inline int current_device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

// fp32 path (cs_fp32.cu)
int fp32_pack_weights(const dlwpcs_conv_desc *d, const dlwpcs_conv_weights *w, int transposed, float *packed,
                      cudaStream_t st);
int fp32_conv_fwd(const dlwpcs_conv_desc *d, const Geometry &g, const float *x0, const float *x1, const float *packed,
                  float *y, cudaStream_t st);
int fp32_conv_dgrad(const dlwpcs_conv_desc *d, const Geometry &g, const float *dy, const float *y,
                    const float *packed_t, float *dx, void *workspace, cudaStream_t st);
int fp32_conv_wgrad(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *dy, const void *y,
                    const dlwpcs_conv_wgrads *out, void *workspace, cudaStream_t st);
int64_t fp32_wgrad_workspace_bytes(const dlwpcs_conv_desc *d, const Geometry &g);

// tensor-core path (cs_tc.cu)
int64_t tc_packed_weight_bytes(const dlwpcs_conv_desc *d, const Geometry &g, int transposed);
int tc_pack_weights(const dlwpcs_conv_desc *d, const Geometry &g, const dlwpcs_conv_weights *w, int transposed,
                    void *packed, cudaStream_t st);
int tc_pack_weights2(const dlwpcs_conv_desc *d, const Geometry &g, const dlwpcs_conv_weights *w, int src_cin,
                     int src_cout, void *packed, void *packed_t, cudaStream_t st, int classic_forward = 0);
int tc_pack_batch(const dlwpcs_pack_item *items, int n, cudaStream_t st);
int tc_conv_fwd(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed,
                void *y, const dlwpcs_chain *chain, cudaStream_t st);
uint32_t tc_chain_target(const dlwpcs_conv_desc *d, const Geometry &g);
bool tc_supported(const dlwpcs_conv_desc *d, const Geometry &g, const char **why);
int tc_conv_dgrad(const dlwpcs_conv_desc *d, const Geometry &g, const void *dy, const void *y, const void *packed_t,
                  void *dx, void *workspace, const void *x_in, int in_act, float in_slope, float in_max, cudaStream_t st);

int tc_trace_read(unsigned long long *host_out, int max_launches, int *n_launches, int reset);
unsigned long long *tc_trace_next(int grid);

// row-streamed kernel (cs_tc_rs.cu): 3x3, stride 1, bf16 in / out, <= 64 input and <= 80 output channels; the three kernel
// rows are stacked along N.  Its packed-weight image differs from the classic one; dlwpcs_conv2d_fwd uses it whenever
// rs_eligible() says so (DLWPCS_RS=0 disables it).
bool rs_eligible(const dlwpcs_conv_desc *d, const Geometry &g);
int64_t rs_packed_weight_bytes(const dlwpcs_conv_desc *d, const Geometry &g);
bool rs_pack_params(const dlwpcs_conv_desc *d, const Geometry &g, int *CinP, int *CoutP, long long *groupElems);
int rs_conv_fwd(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed, void *y,
                cudaStream_t st);
// 3x3 layer + the 1x1 CubeSphereConv2D that is its only consumer (the output layer of every cubed-sphere network,
// Azure/train_cs.py:228) in one launch: a second MMA per output row inside the epilogue; the 3x3 layer's output never
// reaches HBM.  packed_h = the head's classic packed image.
// the layer's output and its 2x2 mean (AveragePooling3D((1,2,2)), train_cs.py:197: the next layer's input) from one launch
bool rs_pool_eligible(const dlwpcs_conv_desc *d, const Geometry &g);
int rs_conv_fwd_pool(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed, void *y,
                     void *ypool, cudaStream_t st);
int rs_debug_cuts(const dlwpcs_conv_desc *d, const Geometry &g, int grid, int *cut_s, int *cut_y);
bool rs_head_eligible(const dlwpcs_conv_desc *d, const Geometry &g, const dlwpcs_conv_desc *dh);
int rs_conv_fwd_head(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *x1, const void *packed,
                     const dlwpcs_conv_desc *dh, const void *packed_h, void *y2, cudaStream_t st);

// patch table of a linearised virtual face (width Wv, G entries per face): physical source pixel or -1; cached per device
const int32_t *get_patch_table(const Geometry &g, int Wv, int G, int n, int halo, int mode);

// tensor-core wgrad (cs_wgrad_tc.cu)
bool tc_wgrad_supported(const dlwpcs_conv_desc *d, const Geometry &g);
int64_t tc_wgrad_workspace_bytes(const dlwpcs_conv_desc *d, const Geometry &g);
int tc_conv_wgrad(const dlwpcs_conv_desc *d, const Geometry &g, const void *x0, const void *dy, const void *y,
                  const dlwpcs_conv_wgrads *out, void *workspace, cudaStream_t st);

}  // namespace dlwpcs
