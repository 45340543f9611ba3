/*
 * dlwpcs.h -- C ABI of the B200-native cubed-sphere convolution engine (libdlwpcs.so).
 *
 * The reference (jweyn/DLWP-CS) has no FFI: its "operator interface" for this path is two Keras layer classes in
 * DLWP/custom.py.  Each entry point below names the reference code it replaces (file:line relative to the reference
 * repository).  All tensors are channels_last, face-major:  (batch, 6, edge, edge, channels), faces 4/5 = south/north
 * pole (custom.py:759-763).  Pointers are plain device pointers unless the function name ends in `_host`; `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Nothing here allocates caller-visible memory; the
 * library keeps small per-device caches (halo index tables).  Every function returns 0 on success, non-zero on error;
 * dlwpcs_last_error() then returns a static, thread-local message.
 */
#ifndef DLWPCS_H_
#define DLWPCS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLWPCS_VERSION 1

enum { DLWPCS_F32 = 0, DLWPCS_BF16 = 1 };
enum { DLWPCS_ACT_NONE = 0, DLWPCS_ACT_CAPPED_LEAKY_RELU = 1 };       /* keras ReLU(negative_slope, max_value) */
/* how a fused input source is sampled relative to the convolution's own resolution (edge n):                          */
enum { DLWPCS_SRC_SAME = 0,      /* source edge n                                                                      */
       DLWPCS_SRC_POOL2 = 1,     /* source edge 2n, 2x2 mean   == AveragePooling3D((1,2,2))   Azure/train_cs.py:197     */
       DLWPCS_SRC_UP2 = 2 };     /* source edge n/2, nearest   == UpSampling3D((1,2,2))       Azure/train_cs.py:198     */

/* One CubeSphereConv2D call (custom.py:824-869 constructor arguments + 921-1002 call), optionally with the preceding
 * CubeSpherePadding2D (custom.py:1082-1308) folded into the load stage (`halo` = p) and the U-Net's
 * pool / upsample / concatenate (Azure/train_cs.py:282-299) folded into the input sampling.                          */
typedef struct dlwpcs_conv_desc {
  int32_t batch;
  int32_t n;                 /* face edge of the tensor the layer pair sees BEFORE the halo exchange                   */
  int32_t halo;              /* p of the fused CubeSpherePadding2D; 0 = input is used as is                            */
  int32_t cin, cout;
  int32_t kh, kw;            /* kernel_size                                                                             */
  int32_t stride_h, stride_w;
  int32_t dil_h, dil_w;      /* dilation_rate                                                                           */
  int32_t same;              /* 0 = padding 'valid', 1 = 'same' (TensorFlow zero padding, extra on bottom/right)        */
  int32_t flip_north_pole;   /* custom.py:965-996                                                                       */
  int32_t independent_north_pole;
  int32_t use_bias;
  int32_t act;               /* DLWPCS_ACT_*; applied after bias (the ReLU layer the U-Net applies to every conv)       */
  float act_slope, act_max;
  int32_t x_dtype, y_dtype;  /* DLWPCS_F32 / DLWPCS_BF16 (bf16 activations select the tcgen05 tensor-core kernel)       */
  int32_t c0, mode0;         /* first  input source: channels [0,c0)      and its DLWPCS_SRC_* sampling mode           */
  int32_t c1, mode1;         /* second input source: channels [c0,c0+c1)  (c1 = 0: none); c0 + c1 == cin               */
} dlwpcs_conv_desc;

/* Weights in the reference's logical layout (custom.py:880-914): HWIO float32 kernels (kh,kw,cin,cout) and (cout,)
 * biases.  w_np / b_np only with independent_north_pole; biases only with use_bias.                                  */
typedef struct dlwpcs_conv_weights {
  const float *w_eq, *w_pol, *w_np;
  const float *b_eq, *b_pol, *b_np;
} dlwpcs_conv_weights;

int dlwpcs_version(void);
const char *dlwpcs_last_error(void);

/* Output face size of a conv: keras conv_output_length as used by compute_output_shape, custom.py:1004-1030.
 * `edge_in` is the edge the conv sees (n + 2*halo).                                                                  */
int dlwpcs_conv_out_edge(int edge_in, int k, int stride, int dilation, int same);

/* Halo index table of CubeSpherePadding2D(p) (custom.py:1198-1308 two-stage composition): lut[f][r][c] = flat source
 * index f'*n*n + i*n + j.  Host-only, no GPU needed.                                                                  */
int dlwpcs_pad_lut_host(int n, int p, int32_t *lut /* [6][n+2p][n+2p] */);

/* Standalone CubeSpherePadding2D.call (custom.py:1198-1308): x (B,6,n,n,c) -> y (B,6,n+2p,n+2p,c).                     */
int dlwpcs_pad_fwd(const void *x, void *y, int batch, int n, int c, int p, int dtype, void *stream);
/* Its adjoint (TF autodiff of the slices/concats): dx[src] = sum of dy over every padded position reading src.
 * Deterministic gather over the inverse table (no float atomics).                                                     */
int dlwpcs_pad_bwd(const void *dy, void *dx, int batch, int n, int c, int p, int dtype, void *stream);
/* The same adjoint followed by the derivative of the activation that produced the un-padded tensor (y_in = that tensor,
 * i.e. the forward input of the halo exchange): dx = scatter_add(dy) * act'(y_in).                                     */
int dlwpcs_pad_bwd_act(const void *dy, const void *y_in, void *dx, int batch, int n, int c, int p, int act, float slope,
                       float maxv, int dtype, void *stream);

/* Weight packing: HWIO float32 -> the per-face-group layouts the kernels read.  `packed` must hold
 * dlwpcs_packed_weight_bytes(desc, transposed) bytes.  transposed = 0: forward, in the layout of whichever kernel
 * dlwpcs_conv2d_fwd runs for the descriptor (3x3 stride-1 bf16 layers with <= 64 input and <= 80 output channels go to the
 * row-streamed kernel, whose image stacks the three kernel rows along N); 1: dgrad (taps rotated 180 degrees, cin/cout
 * swapped); 2: forward image for dlwpcs_conv2d_fwd_chained (always the classic kernel's image).  Group 0 = equatorial, 1 = south pole, 2 = north pole (polar or independent kernel, rows flipped
 * when flip_north_pole -- equivalent to custom.py:969/995 for every stride).                           */
int64_t dlwpcs_packed_weight_bytes(const dlwpcs_conv_desc *d, int transposed);
int dlwpcs_pack_weights(const dlwpcs_conv_desc *d, const dlwpcs_conv_weights *w, int transposed, void *packed,
                        void *stream);
/* bf16 path: forward and transposed (dgrad) images in ONE launch (either pointer may be NULL), from kernels of the logical
 * shape (kh, kw, src_cin, src_cout) zero-extended to the descriptor's (cin, cout) -- the channel padding of the bf16 path
 * (multiples of 8 for 16-byte gathers) then needs no padded copies of the parameters.                                  */
int dlwpcs_pack_weights2(const dlwpcs_conv_desc *d, const dlwpcs_conv_weights *w, int src_cin, int src_cout, void *packed,
                         void *packed_t, void *stream);

/* bf16 path: the images of MANY layers in one launch -- what a training step does after every optimizer update (one launch
 * instead of one or two per layer).  Per item: the layer's descriptor (batch / n are not used), its parameters, the logical
 * kernel shape as for dlwpcs_pack_weights2, and the destination(s); either of packed / packed_t may be NULL.            */
typedef struct dlwpcs_pack_item {
  const dlwpcs_conv_desc *desc;
  dlwpcs_conv_weights w;
  int32_t src_cin, src_cout;
  void *packed, *packed_t;
} dlwpcs_pack_item;
int dlwpcs_pack_weights_batch(const dlwpcs_pack_item *items, int n_items, void *stream);

/* CubeSphereConv2D.call (+ optional fused padding / sampling / bias / activation).
 *   x0, x1 : the input source(s), dtype d->x_dtype;   y : (B,6,Hout,Wout,cout), dtype d->y_dtype.                     */
int dlwpcs_conv2d_fwd(const dlwpcs_conv_desc *d, const void *x0, const void *x1, const void *packed_w, void *y,
                      void *stream);

/* Backward of the same (what TF autodiff derives from custom.py:921-1002 and 1198-1308), stride 1 only, float32:
 *   dgrad: dy (B,6,Hout,Wout,cout) [, y: forward output, to apply the activation derivative] -> dx (B,6,n,n,cin),
 *          including the halo scatter-add.  `workspace` holds dlwpcs_dgrad_workspace_bytes(d) bytes.
 *   wgrad: x0 (B,6,n,n,cin), dy [, y] -> HWIO float32 gradients in the reference's parameter order, `workspace` holds
 *          dlwpcs_wgrad_workspace_bytes(d) bytes.                                                                     */
int64_t dlwpcs_dgrad_workspace_bytes(const dlwpcs_conv_desc *d);
int dlwpcs_conv2d_dgrad(const dlwpcs_conv_desc *d, const void *dy, const void *y, const void *packed_w_t, void *dx,
                        void *workspace, void *stream);
/* dgrad whose result is additionally multiplied by the derivative of the activation that produced the layer's INPUT
 * (x_in = forward input of this layer = output of the previous ReLU(0.1, 10) layer, Azure/train_cs.py:199): the mask is
 * applied inside the halo scatter-add, so the previous layer's backward receives dL/d(pre-activation) directly and needs
 * no pass of its own over the tensor.  in_act = DLWPCS_ACT_NONE: identical to dlwpcs_conv2d_dgrad.                     */
int dlwpcs_conv2d_dgrad_act(const dlwpcs_conv_desc *d, const void *dy, const void *y, const void *packed_w_t, void *dx,
                            void *workspace, const void *x_in, int in_act, float in_slope, float in_max, void *stream);
int64_t dlwpcs_wgrad_workspace_bytes(const dlwpcs_conv_desc *d);
typedef struct dlwpcs_conv_wgrads {
  float *dw_eq, *dw_pol, *dw_np;
  float *db_eq, *db_pol, *db_np;
} dlwpcs_conv_wgrads;
int dlwpcs_conv2d_wgrad(const dlwpcs_conv_desc *d, const void *x0, const void *dy, const void *y,
                        const dlwpcs_conv_wgrads *g, void *workspace, void *stream);

/* Elementwise / resampling steps either side of the conv (Azure/train_cs.py:197-199), standalone, float32 or bf16.   */
int dlwpcs_act_fwd(const void *x, void *y, int64_t count, int act, float slope, float maxv, int dtype, void *stream);
int dlwpcs_act_bwd(const void *dy, const void *y, void *dx, int64_t count, int act, float slope, float maxv,
                   int dtype, void *stream);

/* Materialised resampling steps of the training path (the rollout folds them into the conv's load stage), float32 or
 * bf16, 16 bytes per thread (channel counts multiples of 4 / 8, 16-byte aligned tensors):
 *   dlwpcs_pool2      AveragePooling3D((1,2,2)) train_cs.py:197: backward = 0: in (B,6,2n,2n,c) -> out (B,6,n,n,c), mean
 *                     of each 2x2 block; backward = 1 (adjoint): in (B,6,n,n,c) -> out (B,6,2n,2n,c) = 0.25 * in.
 *   dlwpcs_up2cat_fwd UpSampling3D((1,2,2)) + concatenate train_cs.py:198, 293, 299: t (B,6,n,n,ca+cb) =
 *                     [nearest-upsampled a (B,6,n/2,n/2,ca) | b (B,6,n,n,cb)];  _bwd: da = 2x2 block sums of dt[..., :ca],
 *                     db = dt[..., ca:].                                                                               */
int dlwpcs_pool2(const void *in, void *out, int batch, int n, int c, int backward, int dtype, void *stream);
int dlwpcs_up2cat_fwd(const void *a, const void *b, void *t, int batch, int n, int ca, int cb, int dtype, void *stream);
int dlwpcs_up2cat_bwd(const void *dt, void *da, void *db, int batch, int n, int ca, int cb, int dtype, void *stream);

/* Training-step helpers on flat float32 buffers (the engine keeps all parameters / gradients of a model in one buffer so
 * that data-parallel training needs a single all-reduce).
 *   dlwpcs_mse_loss_grad: keras 'mse' (Azure/train_cs.py:424): adds sum((y-t)^2) * inv_count to *loss_accum (device
 *     float, caller zeroes it) and writes dy = 2 (y-t) * inv_count.  y / t / dy float32 or bfloat16 (dtype), same shape.
 *   dlwpcs_adam_step: Keras Adam (DLWPFunctional.build_model, DLWP/model/models.py:353-377) with the gradient scaled by
 *     grad_scale first (1/world_size after the all-reduce):  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *     p -= lr sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps),  t = step (1-based).                                         */
int dlwpcs_mse_loss_grad(const void *y, const void *t, void *dy, float *loss_accum, int64_t count, float inv_count,
                         int dtype, void *stream);
int dlwpcs_adam_step(float *param, const float *grad, float *m, float *v, int64_t count, float lr, float beta1,
                     float beta2, float eps, int step, float grad_scale, void *stream);
/* The same update with the step counter t in device memory: *step_counter is incremented first, then used.  No host
 * scalar changes between calls, so the whole training step can be replayed from a CUDA graph.                         */
int dlwpcs_adam_step_dev(float *param, const float *grad, float *m, float *v, int64_t count, float lr, float beta1,
                         float beta2, float eps, int32_t *step_counter, float grad_scale, void *stream);

/* Insolation channels of the forced rollout: DLWP/util.py:306-364 (`insolation`) as TimeSeriesEstimator.predict
 * re-evaluates it before every forecast iteration (DLWP/model/extensions.py:272-288), written in place into channels
 * [c_first, c_first + n_sol) of a channels_last tensor out (batch, npix, c_total) -- the engine's forcing buffer -- so the
 * autoregressive loop never leaves the device.  sinlat / coslat: sin / cos of the latitude of every pixel (float64, npix),
 * lon: longitude in degrees, 0-360 (float32, npix), days: day of year (util.py:301-303) of every sample for each of the
 * n_sol input time steps, float64 [n_sol][batch].  Same mixed precision as the reference's numpy expression.          */
int dlwpcs_insolation(void *out, int dtype, int batch, int64_t npix, int c_total, int c_first, int n_sol,
                      const double *sinlat, const double *coslat, const float *lon, const double *days, float S,
                      void *stream);

/* Device-side data feed: ArrayDataGenerator.generate (DLWP/model/generators.py:872-984; convolutional model,
 * channels_last, no sequence) for one batch of sample indices, with the training array resident in HBM in the reference's
 * own layout.  array (T, n_var, npix) float32, insolation (T, npix) or NULL, constants (n_const, npix) or NULL, all device;
 * samples (batch) int64 start times; in_vars / out_vars: variable indices of the input / output selections (device int32).
 *   x (batch, npix, t_in*(v_in + has_insolation) + n_const):  channel t*(v_in+1)+v = array[s + t*interval, in_vars[v]],
 *       insolation last within each time step (generators.py:880-899), constants appended (train_cs.py:396-407);
 *   y (batch, npix, t_out*v_out):  channel t*v_out+v = array[s + interval*(t_in + t), out_vars[v]]  (generators.py:943-946).
 * dtype = element type of x and y (float32 or bf16); x or y may be NULL to produce only the other one.
 *                                                                   */
int dlwpcs_feed_gather(const float *array, const float *insolation, const float *constants, const int64_t *samples,
                       const int32_t *in_vars, const int32_t *out_vars, void *x, void *y, int batch, int64_t npix,
                       int n_var, int v_in, int v_out, int t_in, int t_out, int interval, int n_const, int dtype,
                       void *stream);

/* Host-buffer entry point: the call a reference-side binding makes with numpy arrays.  Copies x (and weights) to the
 * device, runs pad(halo)+conv, copies y back; synchronous.                                                            */
int dlwpcs_conv2d_fwd_host(const dlwpcs_conv_desc *d, const dlwpcs_conv_weights *w, const void *x_host, void *y_host);

/* float32-accurate convolution on the tensor cores (the reference's arithmetic is float32: custom.py:928,947,968,979):
 * a float32 activation x is split into hi = bf16(x), lo = bf16(x - hi) and laid out as [hi(C) | lo(C) | hi(C)] per pixel;
 * with the weights stacked [w_hi ; w_hi ; w_lo] along the input-channel axis, ONE launch of the bf16 tensor-core kernel
 * (dlwpcs_conv2d_fwd with x_dtype = BF16, y_dtype = F32, cin = 3C) accumulates x_hi*w_hi + x_lo*w_hi + x_hi*w_lo in float32
 * in tensor memory -- 16 mantissa bits per operand, the dropped lo*lo term is 2^-16 relative.  dlwpcs_split3 produces that
 * layout from float32 tensors and performs the U-Net's resampling (Azure/train_cs.py:197-198, 293, 299) in float32 before
 * the split: source a (batch,6,na,na,ca) sampled with mode_a (DLWPCS_SRC_SAME: na = n, _POOL2: na = 2n, 2x2 mean, _UP2:
 * na = n/2, nearest), optionally concatenated with b (batch,6,n,n,cb); out (batch,6,n,n,3*(ca+cb)) bf16.  ca, cb multiples
 * of 8.                                                                                                                */
int dlwpcs_split3(const float *a, int ca, int mode_a, const float *b, int cb, void *out, int batch, int n, void *stream);

/* Chained launch of the forward convolution (bf16 tensor-core path) for back-to-back layers of a network (the layer
 * sequence of Azure/train_cs.py:233-388 inside the rollout loop of models.py:446-454): instead of waiting for the whole
 * previous kernel, every tile waits until the launch that produced its input has completed that SAMPLE, and tiles are
 * handed out on demand, so the pipeline fill of layer k+1 runs on the SMs layer k has already left.  The launches must be
 * enqueued back to back on one stream; `dep` = the `done` counters handed to the launch that produced x0 (NULL: plain
 * stream order, e.g. after a non-chained kernel), `dep_target` = dlwpcs_chain_target(desc of that launch).  All counters
 * are caller-owned device memory, zero before the first launch of a chain replay; error_flag (optional) is set to 1 if a
 * dependency did not resolve within ~2 s (mis-wired chain) instead of hanging the device.                              */
typedef struct dlwpcs_chain {
  const uint32_t *dep;
  uint32_t dep_target;
  uint32_t *done;            /* [batch] counters this launch raises, or NULL                                            */
  uint32_t *tile_counter;    /* one counter: the launch's dynamic tile scheduler                                        */
  uint32_t *error_flag;
} dlwpcs_chain;
int dlwpcs_conv2d_fwd_chained(const dlwpcs_conv_desc *d, const void *x0, const void *x1, const void *packed_w, void *y,
                              const dlwpcs_chain *chain, void *stream);
uint32_t dlwpcs_chain_target(const dlwpcs_conv_desc *d);

/* A 3x3 CubeSphereConv2D (with its fused padding / sampling / bias / activation) followed by the 1x1 CubeSphereConv2D that
 * is its ONLY consumer -- the output layer of every cubed-sphere network, Azure/train_cs.py:228 after :225-227 -- in one
 * launch: the activated bf16 output row of the first layer becomes the A operand of a second tcgen05.mma inside the
 * epilogue, and never reaches HBM.  Same arithmetic as the two separate launches (one bf16 rounding in between, float32
 * accumulation of the 1x1 products in the same order): bit-identical results.
 *   dlwpcs_conv2d_head_fusable: 1 when the pair qualifies (3x3 stride-1 bf16 layer served by the row-streamed kernel with
 *       32 or 64 padded output channels; head: 1x1, stride 1, no halo, single un-resampled source of d->cout channels,
 *       bf16, <= 64 output channels in multiples of 8), else 0.
 *   head_packed_w: dlwpcs_pack_weights(head, ..., transposed = 0).  y_head: (B,6,Hout,Wout,head->cout) bf16.           */
int dlwpcs_conv2d_head_fusable(const dlwpcs_conv_desc *d, const dlwpcs_conv_desc *head);
int dlwpcs_conv2d_fwd_head(const dlwpcs_conv_desc *d, const void *x0, const void *x1, const void *packed_w,
                           const dlwpcs_conv_desc *head, const void *head_packed_w, void *y_head, void *stream);

/* A CubeSphereConv2D whose output is also read through AveragePooling3D((1,2,2)) (Azure/train_cs.py:197: the encoder layers
 * in front of a pooling, e.g. conv_2d_1_2 -> conv_2d_2) writes that 2x2 mean itself, as a second tensor y_pool
 * (B,6,Hout/2,Wout/2,cout) bf16: the pooled layer then reads a plain source -- a quarter of the bytes, through asynchronous
 * copies.  The mean is taken over the bf16-rounded outputs in float32 and rounded to bf16 once, exactly what the fused
 * pooled load (DLWPCS_SRC_POOL2) computes.  dlwpcs_conv2d_pool_fusable: 1 for 3x3 stride-1 bf16 layers served by the
 * row-streamed kernel with <= 32 output channels and even face edges.                                                    */
int dlwpcs_conv2d_pool_fusable(const dlwpcs_conv_desc *d);
int dlwpcs_conv2d_fwd_pool(const dlwpcs_conv_desc *d, const void *x0, const void *x1, const void *packed_w, void *y,
                           void *y_pool, void *stream);

/* Host only (no GPU needed; diagnostics / tests): the work split of the row-streamed kernel for a layer it serves.  The
 * (strip, output row) sequence of the layer -- strips of 128 positions over the row-wise concatenated faces of one weight
 * group, equatorial first -- is cut into `grid` contiguous ranges of equal cost; CTA c works on [(cut_s[c], cut_y[c]),
 * (cut_s[c+1], cut_y[c+1])).  Both arrays hold grid + 1 entries.  Returns grid, or 0 when the layer is not served by that
 * kernel or grid is out of range (1..160).                                                                             */
int dlwpcs_rs_work_cuts(const dlwpcs_conv_desc *d, int grid, int32_t *cut_s, int32_t *cut_y);

/* Diagnostics (no reference counterpart): with DLWPCS_TC_TRACE=1 in the environment every launch of the tensor-core
 * convolution kernel records, per CTA, eight %globaltimer values (ns): 0 kernel entry, 1 prologue done, 2 loaders past
 * the grid dependency, 3 first input patch in shared memory, 4 first accumulators complete, 5 last epilogue done,
 * 6 kernel exit, 7 = tiles walked.  Copies the records of up to max_launches launches (160 CTAs x 8 values each, launch
 * order) to host_out, stores the number recorded in *n_launches; reset != 0 restarts the recording.                   */
int dlwpcs_trace_read(unsigned long long *host_out, int max_launches, int *n_launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* DLWPCS_H_ */
