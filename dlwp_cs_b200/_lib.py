"""ctypes binding of libdlwpcs.so (include/dlwpcs.h).  torch supplies device memory and the current stream only.

There is no CPU fallback: if the shared library is missing or a CUDA tensor is not supplied, the calls raise.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libdlwpcs.so')

F32, BF16 = 0, 1
ACT_NONE, ACT_CAPPED_LEAKY_RELU = 0, 1
SRC_SAME, SRC_POOL2, SRC_UP2 = 0, 1, 2


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('batch', 'n', 'halo', 'cin', 'cout', 'kh', 'kw', 'stride_h', 'stride_w',
                                               'dil_h', 'dil_w', 'same', 'flip_north_pole', 'independent_north_pole',
                                               'use_bias', 'act')] + \
               [('act_slope', ctypes.c_float), ('act_max', ctypes.c_float)] + \
               [(n, ctypes.c_int32) for n in ('x_dtype', 'y_dtype', 'c0', 'mode0', 'c1', 'mode1')]


class ConvWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('w_eq', 'w_pol', 'w_np', 'b_eq', 'b_pol', 'b_np')]


class PackItem(ctypes.Structure):
    _fields_ = [('desc', ctypes.POINTER(ConvDesc)), ('w', ConvWeights), ('src_cin', ctypes.c_int32),
                ('src_cout', ctypes.c_int32), ('packed', ctypes.c_void_p), ('packed_t', ctypes.c_void_p)]


class Chain(ctypes.Structure):
    _fields_ = [('dep', ctypes.c_void_p), ('dep_target', ctypes.c_uint32), ('done', ctypes.c_void_p),
                ('tile_counter', ctypes.c_void_p), ('error_flag', ctypes.c_void_p)]


class ConvWgrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('dw_eq', 'dw_pol', 'dw_np', 'db_eq', 'db_pol', 'db_np')]


_lib = None


def load():
    """Load the shared library (once).  Raises ImportError with build instructions if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('libdlwpcs.so not found at %s -- build it with `python -m dlwp_cs_b200.build` '
                          '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    dp, wp, gp = ctypes.POINTER(ConvDesc), ctypes.POINTER(ConvWeights), ctypes.POINTER(ConvWgrads)
    sigs = {
        'dlwpcs_version': (i32, []),
        'dlwpcs_last_error': (ctypes.c_char_p, []),
        'dlwpcs_conv_out_edge': (i32, [i32, i32, i32, i32, i32]),
        'dlwpcs_pad_lut_host': (i32, [i32, i32, vp]),
        'dlwpcs_pad_fwd': (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
        'dlwpcs_pad_bwd': (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
        'dlwpcs_packed_weight_bytes': (i64, [dp, i32]),
        'dlwpcs_pack_weights': (i32, [dp, wp, i32, vp, vp]),
        'dlwpcs_conv2d_fwd': (i32, [dp, vp, vp, vp, vp, vp]),
        'dlwpcs_dgrad_workspace_bytes': (i64, [dp]),
        'dlwpcs_conv2d_dgrad': (i32, [dp, vp, vp, vp, vp, vp, vp]),
        'dlwpcs_wgrad_workspace_bytes': (i64, [dp]),
        'dlwpcs_conv2d_wgrad': (i32, [dp, vp, vp, vp, gp, vp, vp]),
        'dlwpcs_act_fwd': (i32, [vp, vp, i64, i32, f32, f32, i32, vp]),
        'dlwpcs_act_bwd': (i32, [vp, vp, vp, i64, i32, f32, f32, i32, vp]),
        'dlwpcs_conv2d_fwd_host': (i32, [dp, wp, vp, vp]),
        'dlwpcs_mse_loss_grad': (i32, [vp, vp, vp, vp, i64, f32, i32, vp]),
        'dlwpcs_adam_step': (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, f32, vp]),
        'dlwpcs_pool2': (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
        'dlwpcs_up2cat_fwd': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
        'dlwpcs_up2cat_bwd': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
        'dlwpcs_feed_gather': (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
        'dlwpcs_insolation': (i32, [vp, i32, i32, i64, i32, i32, i32, vp, vp, vp, vp, f32, vp]),
        'dlwpcs_adam_step_dev': (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, f32, vp]),
        'dlwpcs_trace_read': (i32, [vp, i32, ctypes.POINTER(ctypes.c_int), i32]),
        'dlwpcs_conv2d_fwd_chained': (i32, [dp, vp, vp, vp, vp, ctypes.POINTER(Chain), vp]),
        'dlwpcs_chain_target': (ctypes.c_uint32, [dp]),
        'dlwpcs_split3': (i32, [vp, i32, i32, vp, i32, vp, i32, i32, vp]),
        'dlwpcs_pack_weights2': (i32, [dp, wp, i32, i32, vp, vp, vp]),
        'dlwpcs_pad_bwd_act': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, i32, vp]),
        'dlwpcs_conv2d_dgrad_act': (i32, [dp, vp, vp, vp, vp, vp, vp, i32, f32, f32, vp]),
        'dlwpcs_conv2d_head_fusable': (i32, [dp, dp]),
        'dlwpcs_rs_work_cuts': (i32, [dp, i32, vp, vp]),
        'dlwpcs_conv2d_pool_fusable': (i32, [dp]),
        'dlwpcs_conv2d_fwd_pool': (i32, [dp, vp, vp, vp, vp, vp, vp]),
        'dlwpcs_pack_weights_batch': (i32, [ctypes.POINTER(PackItem), i32, vp]),
        'dlwpcs_conv2d_fwd_head': (i32, [dp, vp, vp, vp, dp, vp, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError here == header and library out of sync
        fn.restype, fn.argtypes = res, args
    if lib.dlwpcs_version() != 1:
        raise ImportError('libdlwpcs.so version mismatch')
    _lib = lib
    return lib


EXPORTED = ('dlwpcs_version', 'dlwpcs_last_error', 'dlwpcs_conv_out_edge', 'dlwpcs_pad_lut_host', 'dlwpcs_pad_fwd',
            'dlwpcs_pad_bwd', 'dlwpcs_packed_weight_bytes', 'dlwpcs_pack_weights', 'dlwpcs_conv2d_fwd',
            'dlwpcs_dgrad_workspace_bytes', 'dlwpcs_conv2d_dgrad', 'dlwpcs_wgrad_workspace_bytes',
            'dlwpcs_conv2d_wgrad', 'dlwpcs_act_fwd', 'dlwpcs_act_bwd', 'dlwpcs_conv2d_fwd_host',
            'dlwpcs_mse_loss_grad', 'dlwpcs_adam_step', 'dlwpcs_adam_step_dev', 'dlwpcs_insolation', 'dlwpcs_pool2',
            'dlwpcs_up2cat_fwd', 'dlwpcs_up2cat_bwd', 'dlwpcs_feed_gather', 'dlwpcs_trace_read',
            'dlwpcs_conv2d_fwd_chained', 'dlwpcs_chain_target', 'dlwpcs_split3',
            'dlwpcs_pack_weights2', 'dlwpcs_pad_bwd_act', 'dlwpcs_conv2d_dgrad_act', 'dlwpcs_conv2d_head_fusable',
            'dlwpcs_conv2d_fwd_head', 'dlwpcs_rs_work_cuts', 'dlwpcs_pack_weights_batch', 'dlwpcs_conv2d_pool_fusable',
            'dlwpcs_conv2d_fwd_pool')


class DlwpcsError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise DlwpcsError(load().dlwpcs_last_error().decode() or 'libdlwpcs error %d' % rc)


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise DlwpcsError('unsupported dtype %s (float32 or bfloat16)' % (dt,))


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    """Every tensor handed to the library must live on the CURRENT CUDA device: the kernels are launched on that device's
    current stream and the library caches its halo / patch tables per current device (call ``torch.cuda.set_device`` --
    or wrap the call in ``torch.cuda.device(t.device)`` -- in a one-process-per-GPU launcher)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DlwpcsError('dlwp_cs_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU path'
                              % t.device.type)
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise DlwpcsError('tensor on cuda:%d but the current device is cuda:%d -- call torch.cuda.set_device(%d) first'
                              % (t.device.index, cur, t.device.index))


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def pad_lut(n, p):
    """Halo index table (6, n+2p, n+2p) int32 -- host only."""
    h = n + 2 * p
    out = np.empty((6, h, h), dtype=np.int32)
    check(load().dlwpcs_pad_lut_host(n, p, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def conv_out_edge(edge_in, k, stride, dilation, same):
    r = load().dlwpcs_conv_out_edge(edge_in, k, stride, dilation, int(bool(same)))
    if r < 0:
        raise DlwpcsError('bad arguments to conv_out_edge')
    return r


def make_desc(batch, n, cin, cout, kernel_size=(3, 3), strides=(1, 1), dilation=(1, 1), halo=0, same=False,
              flip_north_pole=True, independent_north_pole=False, use_bias=True, act=ACT_NONE, act_slope=0.1,
              act_max=10.0, x_dtype=F32, y_dtype=F32, c0=None, mode0=SRC_SAME, c1=0, mode1=SRC_SAME):
    d = ConvDesc()
    d.batch, d.n, d.halo, d.cin, d.cout = batch, n, halo, cin, cout
    d.kh, d.kw = kernel_size
    d.stride_h, d.stride_w = strides
    d.dil_h, d.dil_w = dilation
    d.same = int(bool(same))
    d.flip_north_pole = int(bool(flip_north_pole))
    d.independent_north_pole = int(bool(independent_north_pole))
    d.use_bias = int(bool(use_bias))
    d.act, d.act_slope, d.act_max = act, act_slope, act_max
    d.x_dtype, d.y_dtype = x_dtype, y_dtype
    d.c0 = cin - c1 if c0 is None else c0
    d.mode0, d.c1, d.mode1 = mode0, c1, mode1
    return d


def copy_desc(d, **changes):
    """A copy of a conv descriptor with some fields replaced."""
    c = ConvDesc()
    ctypes.memmove(ctypes.byref(c), ctypes.byref(d), ctypes.sizeof(ConvDesc))
    for k, v in changes.items():
        setattr(c, k, v)
    return c


def out_shape(d):
    hin = d.n + 2 * d.halo
    return (d.batch, 6, conv_out_edge(hin, d.kh, d.stride_h, d.dil_h, d.same),
            conv_out_edge(hin, d.kw, d.stride_w, d.dil_w, d.same), d.cout)


def pack_weights(d, w_eq, w_pol, w_np=None, b_eq=None, b_pol=None, b_np=None, transposed=False):
    """HWIO float32 parameters -> the packed device buffer the kernels read (uint8 tensor).  transposed: False / 0 forward
    (for conv2d_fwd), True / 1 dgrad, 2 forward for conv2d_fwd_chained (see include/dlwpcs.h)."""
    lib = load()
    ws = [w_eq, w_pol, w_np, b_eq, b_pol, b_np]
    require_cuda(*ws)
    ws = [None if t is None else t.detach().to(torch.float32).contiguous() for t in ws]
    nbytes = lib.dlwpcs_packed_weight_bytes(ctypes.byref(d), int(transposed))
    if nbytes < 0:
        check(1)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w_eq.device)
    cw = ConvWeights(*[t.data_ptr() if t is not None else None for t in ws])
    check(lib.dlwpcs_pack_weights(ctypes.byref(d), ctypes.byref(cw), int(transposed), ptr(packed), stream_ptr()))
    return packed


def pack_weights2(d, w_eq, w_pol, w_np=None, b_eq=None, b_pol=None, b_np=None, forward=True, transposed=True):
    """bf16 path: (packed, packed_t) in one launch; the kernels may have fewer input / output channels than the descriptor
    (they are zero-extended, see include/dlwpcs.h)."""
    lib = load()
    ws = [w_eq, w_pol, w_np, b_eq, b_pol, b_np]
    require_cuda(*ws)
    ws = [None if t is None else (t.detach() if (t.dtype == torch.float32 and t.is_contiguous())
                                   else t.detach().to(torch.float32).contiguous()) for t in ws]
    outs = []
    for want, tr in ((forward, 0), (transposed, 1)):
        if not want:
            outs.append(None)
            continue
        nbytes = lib.dlwpcs_packed_weight_bytes(ctypes.byref(d), tr)
        if nbytes < 0:
            check(1)
        outs.append(torch.empty(nbytes, dtype=torch.uint8, device=w_eq.device))
    cw = ConvWeights(*[t.data_ptr() if t is not None else None for t in ws])
    check(lib.dlwpcs_pack_weights2(ctypes.byref(d), ctypes.byref(cw), int(w_eq.shape[2]), int(w_eq.shape[3]), ptr(outs[0]),
                                   ptr(outs[1]), stream_ptr()))
    return outs[0], outs[1]


def pack_weights_batch(entries, out=None):
    """dlwpcs_pack_weights_batch: the bf16 images of many layers in ONE launch.  entries: list of (d, (w_eq, w_pol, w_np,
    b_eq, b_pol, b_np), want_forward, want_transposed); the kernels may have fewer channels than the descriptor (zero
    extension).  out: the list a previous call returned (its buffers are reused).  -> list of (packed, packed_t)."""
    lib = load()
    items = (PackItem * len(entries))()
    keep, res = [], []
    for i, (d, ws, want_f, want_t) in enumerate(entries):
        require_cuda(*ws)
        ws = [None if t is None else (t.detach() if (t.dtype == torch.float32 and t.is_contiguous())
                                       else t.detach().to(torch.float32).contiguous()) for t in ws]
        keep.append(ws)
        bufs = []
        for j, (want, tr) in enumerate(((want_f, 0), (want_t, 1))):
            if not want:
                bufs.append(None)
                continue
            nbytes = lib.dlwpcs_packed_weight_bytes(ctypes.byref(d), tr)
            if nbytes < 0:
                check(1)
            old = out[i][j] if out is not None else None
            bufs.append(old if (old is not None and old.numel() == nbytes) else
                        torch.empty(nbytes, dtype=torch.uint8, device=ws[0].device))
        items[i].desc = ctypes.pointer(d)
        items[i].w = ConvWeights(*[t.data_ptr() if t is not None else None for t in ws])
        items[i].src_cin, items[i].src_cout = int(ws[0].shape[2]), int(ws[0].shape[3])
        items[i].packed = bufs[0].data_ptr() if bufs[0] is not None else None
        items[i].packed_t = bufs[1].data_ptr() if bufs[1] is not None else None
        res.append((bufs[0], bufs[1]))
    check(lib.dlwpcs_pack_weights_batch(items, len(entries), stream_ptr()))
    return res


def conv2d_fwd(d, x0, x1, packed, out=None):
    require_cuda(x0, x1, packed)
    shp = out_shape(d)
    ydt = torch.float32 if d.y_dtype == F32 else torch.bfloat16
    y = out if out is not None else torch.empty(shp, dtype=ydt, device=x0.device)
    check(load().dlwpcs_conv2d_fwd(ctypes.byref(d), ptr(x0), ptr(x1), ptr(packed), ptr(y), stream_ptr()))
    return y


def rs_work_cuts(d, grid):
    """Work split of the row-streamed kernel (host only): list of (strip, row) cut points, grid + 1 entries, or None when
    the layer is not served by that kernel."""
    cs = (ctypes.c_int32 * (grid + 1))()
    cy = (ctypes.c_int32 * (grid + 1))()
    if load().dlwpcs_rs_work_cuts(ctypes.byref(d), grid, cs, cy) != grid:
        return None
    return [(int(cs[i]), int(cy[i])) for i in range(grid + 1)]


def conv2d_pool_fusable(d):
    """True when the layer `d` can write the 2x2 mean of its output itself (dlwpcs_conv2d_fwd_pool)."""
    return bool(load().dlwpcs_conv2d_pool_fusable(ctypes.byref(d)))


def conv2d_fwd_pool(d, x0, x1, packed, out=None, out_pool=None):
    """dlwpcs_conv2d_fwd_pool: -> (y, 2x2 mean of y), both bf16."""
    require_cuda(x0, x1, packed, out, out_pool)
    shp = out_shape(d)
    y = out if out is not None else torch.empty(shp, dtype=torch.bfloat16, device=x0.device)
    yp = out_pool if out_pool is not None else torch.empty((shp[0], 6, shp[2] // 2, shp[3] // 2, shp[4]), dtype=torch.bfloat16,
                                                            device=x0.device)
    check(load().dlwpcs_conv2d_fwd_pool(ctypes.byref(d), ptr(x0), ptr(x1), ptr(packed), ptr(y), ptr(yp), stream_ptr()))
    return y, yp


def conv2d_head_fusable(d, head):
    """True when `head` (a 1x1 CubeSphereConv2D reading only the output of `d`) can run inside the epilogue of `d`."""
    return bool(load().dlwpcs_conv2d_head_fusable(ctypes.byref(d), ctypes.byref(head)))


def conv2d_fwd_head(d, x0, x1, packed, head, head_packed, out=None):
    """dlwpcs_conv2d_fwd_head: the 3x3 layer `d` and its 1x1 consumer `head` in one launch; returns the head's output."""
    require_cuda(x0, x1, packed, head_packed, out)
    y = out if out is not None else torch.empty(out_shape(head), dtype=torch.bfloat16, device=x0.device)
    check(load().dlwpcs_conv2d_fwd_head(ctypes.byref(d), ptr(x0), ptr(x1), ptr(packed), ctypes.byref(head), ptr(head_packed),
                                        ptr(y), stream_ptr()))
    return y


def split3(a, mode_a=SRC_SAME, b=None, out=None):
    """float32 (B,6,na,na,Ca) [+ (B,6,n,n,Cb)] -> bf16 (B,6,n,n,3*(Ca+Cb)) in the [hi | lo | hi] layout of the float32-accurate
    tensor-core convolution; `a` is resampled (2x2 mean / nearest up-sampling) in float32 first.  See include/dlwpcs.h."""
    require_cuda(a, b, out)
    if a.dtype != torch.float32 or (b is not None and b.dtype != torch.float32):
        raise DlwpcsError('split3 takes float32 tensors')
    a = a.contiguous()
    b = None if b is None else b.contiguous()
    bs, _, na, _, ca = a.shape
    n = na if mode_a == SRC_SAME else (na // 2 if mode_a == SRC_POOL2 else na * 2)
    cb = 0 if b is None else b.shape[-1]
    if b is not None and (b.shape[0] != bs or b.shape[2] != n):
        raise DlwpcsError('split3: second source %r does not match %d x %d faces of batch %d' % (tuple(b.shape), n, n, bs))
    shape = (bs, 6, n, n, 3 * (ca + cb))
    if out is None:
        out = torch.empty(shape, dtype=torch.bfloat16, device=a.device)
    elif tuple(out.shape) != shape or out.dtype != torch.bfloat16 or not out.is_contiguous():
        raise DlwpcsError('split3: output buffer must be contiguous bfloat16 of shape %r' % (shape,))
    check(load().dlwpcs_split3(ptr(a), ca, mode_a, ptr(b), cb, ptr(out), bs, n, stream_ptr()))
    return out


def split3_weights(w, cin_pad=0):
    """HWIO float32 kernel (kh,kw,Cin,Cout) -> (kh,kw,3*(Cin+cin_pad),Cout) float32 holding [w_hi ; w_hi ; w_lo] along the
    input-channel axis (all values bf16-representable), the weight side of the [hi | lo | hi] activation layout."""
    w = w.detach().to(torch.float32)
    if cin_pad:
        w = torch.nn.functional.pad(w, (0, 0, 0, cin_pad))
    hi = w.to(torch.bfloat16).to(torch.float32)
    lo = (w - hi).to(torch.bfloat16).to(torch.float32)
    return torch.cat([hi, hi, lo], dim=2).contiguous()


def chain_target(d):
    """Value the per-sample completion counters of a chained launch of `d` reach (tiles per sample x epilogue warps)."""
    v = load().dlwpcs_chain_target(ctypes.byref(d))
    if v == 0:
        raise DlwpcsError('chained launches need a bf16 tensor-core configuration')
    return int(v)


def conv2d_fwd_chained(d, x0, x1, packed, out, dep, dep_target, done, tile_counter, error_flag=None):
    """dlwpcs_conv2d_fwd_chained: dep / done are int32 device tensors of `batch` counters (dep may be None: plain stream
    order), tile_counter / error_flag 1-element int32 device tensors; all zero before the chain starts."""
    require_cuda(x0, x1, packed, out, dep, done, tile_counter, error_flag)
    ch = Chain(dep.data_ptr() if dep is not None else None, int(dep_target) if dep is not None else 0,
               done.data_ptr() if done is not None else None, tile_counter.data_ptr(),
               error_flag.data_ptr() if error_flag is not None else None)
    check(load().dlwpcs_conv2d_fwd_chained(ctypes.byref(d), ptr(x0), ptr(x1), ptr(packed), ptr(out), ctypes.byref(ch),
                                           stream_ptr()))
    return out


def conv2d_dgrad(d, dy, y, packed_t, x_in=None, in_act=None):
    """in_act = (code, slope, max) with x_in = the layer's forward input: the result is dL/d(pre-activation of the layer
    that produced x_in) (dlwpcs_conv2d_dgrad_act)."""
    require_cuda(dy, y, packed_t, x_in)
    lib = load()
    nbytes = lib.dlwpcs_dgrad_workspace_bytes(ctypes.byref(d))
    if nbytes < 0:
        check(1)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dy.device)
    dx = torch.empty((d.batch, 6, d.n, d.n, d.cin), dtype=torch.float32 if d.x_dtype == F32 else torch.bfloat16,
                     device=dy.device)
    if in_act is not None and in_act[0] != ACT_NONE:
        check(lib.dlwpcs_conv2d_dgrad_act(ctypes.byref(d), ptr(dy), ptr(y), ptr(packed_t), ptr(dx), ptr(ws), ptr(x_in),
                                          int(in_act[0]), float(in_act[1]), float(in_act[2]), stream_ptr()))
    else:
        check(lib.dlwpcs_conv2d_dgrad(ctypes.byref(d), ptr(dy), ptr(y), ptr(packed_t), ptr(dx), ptr(ws), stream_ptr()))
    return dx


def conv2d_wgrad(d, x0, dy, y, outs=None):
    """Returns (dw_eq, dw_pol, dw_np|None, db_eq|None, db_pol|None, db_np|None), float32 HWIO.  outs: optional tuple of six
    pre-allocated contiguous float32 tensors (None where the layer has no such parameter) to write into -- e.g. views of a
    flat gradient buffer."""
    require_cuda(x0, dy, y)
    lib = load()
    nbytes = lib.dlwpcs_wgrad_workspace_bytes(ctypes.byref(d))
    if nbytes < 0:
        check(1)
    dev = x0.device
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    wshape = (d.kh, d.kw, d.cin, d.cout)
    inp, bias = bool(d.independent_north_pole), bool(d.use_bias)
    shapes = (wshape, wshape, wshape if inp else None, (d.cout,) if bias else None, (d.cout,) if bias else None,
              (d.cout,) if (bias and inp) else None)
    if outs is not None:
        for given, shp in zip(outs, shapes):
            if (given is None) != (shp is None) or (given is not None and (
                    tuple(given.shape) != tuple(shp) or given.dtype != torch.float32 or not given.is_contiguous())):
                raise DlwpcsError('conv2d_wgrad: output buffers do not match the layer')
        require_cuda(*outs)
    else:
        outs = tuple(None if shp is None else torch.empty(shp, dtype=torch.float32, device=dev) for shp in shapes)
    g = ConvWgrads(*[t.data_ptr() if t is not None else None for t in outs])
    check(lib.dlwpcs_conv2d_wgrad(ctypes.byref(d), ptr(x0), ptr(dy), ptr(y), ctypes.byref(g), ptr(ws), stream_ptr()))
    return outs


def pad_fwd(x, p):
    require_cuda(x)
    b, six, n, n2, c = x.shape
    y = torch.empty((b, 6, n + 2 * p, n + 2 * p, c), dtype=x.dtype, device=x.device)
    check(load().dlwpcs_pad_fwd(ptr(x), ptr(y), b, n, c, p, dtype_code(x.dtype), stream_ptr()))
    return y


def pad_bwd(dy, p):
    require_cuda(dy)
    b, six, h, h2, c = dy.shape
    n = h - 2 * p
    dx = torch.empty((b, 6, n, n, c), dtype=dy.dtype, device=dy.device)
    check(load().dlwpcs_pad_bwd(ptr(dy), ptr(dx), b, n, c, p, dtype_code(dy.dtype), stream_ptr()))
    return dx


def act_fwd(x, act, slope, maxv):
    require_cuda(x)
    y = torch.empty_like(x)
    check(load().dlwpcs_act_fwd(ptr(x), ptr(y), x.numel(), act, slope, maxv, dtype_code(x.dtype), stream_ptr()))
    return y


def act_bwd(dy, y, act, slope, maxv):
    require_cuda(dy, y)
    dx = torch.empty_like(dy)
    check(load().dlwpcs_act_bwd(ptr(dy), ptr(y), ptr(dx), dy.numel(), act, slope, maxv, dtype_code(dy.dtype),
                                stream_ptr()))
    return dx


def conv2d_fwd_host(d, x, w_eq, w_pol, w_np=None, b_eq=None, b_pol=None, b_np=None):
    """numpy in / numpy out through the host-buffer C entry point (H2D + kernels + D2H inside the call)."""
    lib = load()
    arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in
            (w_eq, w_pol, w_np, b_eq, b_pol, b_np)]
    cw = ConvWeights(*[a.ctypes.data if a is not None else None for a in arrs])
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty(out_shape(d), dtype=np.float32)
    check(lib.dlwpcs_conv2d_fwd_host(ctypes.byref(d), ctypes.byref(cw), x.ctypes.data_as(ctypes.c_void_p),
                                     y.ctypes.data_as(ctypes.c_void_p)))
    return y


def mse_loss_grad(y, t, loss_accum, scale=1.0):
    """keras 'mse': adds scale * mean((y-t)^2) to the 1-element float32 device tensor loss_accum, returns its gradient
    with respect to y (scale = the output's loss weight, 1/S for the S outputs of a multi-step model, train_cs.py:424-426)."""
    require_cuda(y, t, loss_accum)
    if tuple(t.shape) != tuple(y.shape) or t.dtype != y.dtype:
        raise DlwpcsError('mse_loss_grad: target %r %s does not match the model output %r %s'
                          % (tuple(t.shape), t.dtype, tuple(y.shape), y.dtype))
    if loss_accum.dtype != torch.float32 or loss_accum.numel() != 1:
        raise DlwpcsError('mse_loss_grad: loss_accum must be a 1-element float32 tensor')
    y, t = y.contiguous(), t.contiguous()
    dy = torch.empty_like(y)
    n = y.numel()
    check(load().dlwpcs_mse_loss_grad(ptr(y), ptr(t), ptr(dy), ptr(loss_accum), n, float(scale) / max(n, 1),
                                      dtype_code(y.dtype), stream_ptr()))
    return dy


def adam_step(param, grad, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """In-place Keras-Adam update of a flat float32 parameter buffer."""
    require_cuda(param, grad, m, v)
    for t in (param, grad, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise DlwpcsError('adam_step needs contiguous float32 buffers')
    check(load().dlwpcs_adam_step(ptr(param), ptr(grad), ptr(m), ptr(v), param.numel(), lr, beta1, beta2, eps, int(step),
                                  grad_scale, stream_ptr()))


def adam_step_dev(param, grad, m, v, lr, beta1, beta2, eps, step_counter, grad_scale=1.0):
    """Keras-Adam update with the step counter in a 1-element int32 device tensor (incremented by the call): replayable
    from a CUDA graph."""
    require_cuda(param, grad, m, v, step_counter)
    for t in (param, grad, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise DlwpcsError('adam_step_dev needs contiguous float32 buffers')
    if step_counter.dtype != torch.int32 or step_counter.numel() != 1:
        raise DlwpcsError('step_counter must be a 1-element int32 tensor')
    check(load().dlwpcs_adam_step_dev(ptr(param), ptr(grad), ptr(m), ptr(v), param.numel(), lr, beta1, beta2, eps,
                                      ptr(step_counter), grad_scale, stream_ptr()))


def insolation(out, c_first, n_sol, sinlat, coslat, lon, days, S=1.0):
    """Write the insolation channels [c_first, c_first + n_sol) of the channels_last tensor out (B, ..., C) in place.
    sinlat / coslat float64 (npix,), lon float32 (npix,) degrees, days float64 (n_sol, B) day of year."""
    require_cuda(out, sinlat, coslat, lon, days)
    b, c = out.shape[0], out.shape[-1]
    npix = out.numel() // max(b * c, 1)
    if (sinlat.dtype, coslat.dtype, lon.dtype, days.dtype) != (torch.float64, torch.float64, torch.float32, torch.float64):
        raise DlwpcsError('insolation: sinlat / coslat / days must be float64 and lon float32')
    if sinlat.numel() != npix or coslat.numel() != npix or lon.numel() != npix or tuple(days.shape) != (n_sol, b):
        raise DlwpcsError('insolation: geometry / days arrays do not match the output tensor')
    if not out.is_contiguous():
        raise DlwpcsError('insolation: the output tensor must be contiguous')
    check(load().dlwpcs_insolation(ptr(out), dtype_code(out.dtype), b, npix, c, c_first, n_sol, ptr(sinlat), ptr(coslat),
                                   ptr(lon), ptr(days.contiguous()), S, stream_ptr()))
    return out


def trace_read(max_launches=256, reset=True):
    """Per-CTA %globaltimer records of the last tensor-core conv launches (needs DLWPCS_TC_TRACE=1 in the environment
    before the library is loaded): uint64 array (launches, 160, 8), see include/dlwpcs.h."""
    out = np.zeros((max_launches, 160, 8), dtype=np.uint64)
    n = ctypes.c_int(0)
    check(load().dlwpcs_trace_read(out.ctypes.data_as(ctypes.c_void_p), max_launches, ctypes.byref(n), int(reset)))
    return out[:min(n.value, max_launches)]


def resample_vec_ok(*tensors):
    """True when the 16-byte resampling kernels apply: CUDA, contiguous, float32 / bf16, channels fill 16-byte chunks."""
    for t in tensors:
        if not (t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16)):
            return False
        if t.shape[-1] % (16 // t.element_size()) or t.data_ptr() % 16:
            return False
    return True


def pool2(x, backward=False):
    """AveragePooling3D((1,2,2)) on (B,6,2n,2n,C) -> (B,6,n,n,C); backward=True: its adjoint (B,6,n,n,C) -> (B,6,2n,2n,C)."""
    require_cuda(x)
    b, _, e, _, c = x.shape
    n = e if backward else e // 2
    out = torch.empty((b, 6, 2 * n, 2 * n, c) if backward else (b, 6, n, n, c), dtype=x.dtype, device=x.device)
    check(load().dlwpcs_pool2(ptr(x), ptr(out), b, n, c, int(backward), dtype_code(x.dtype), stream_ptr()))
    return out


def up2cat_fwd(a, b):
    require_cuda(a, b)
    bs, _, n, _, cb = b.shape
    ca = a.shape[-1]
    t = torch.empty((bs, 6, n, n, ca + cb), dtype=b.dtype, device=b.device)
    check(load().dlwpcs_up2cat_fwd(ptr(a), ptr(b), ptr(t), bs, n, ca, cb, dtype_code(b.dtype), stream_ptr()))
    return t


def up2cat_bwd(dt, ca):
    require_cuda(dt)
    bs, _, n, _, ct = dt.shape
    cb = ct - ca
    da = torch.empty((bs, 6, n // 2, n // 2, ca), dtype=dt.dtype, device=dt.device)
    db = torch.empty((bs, 6, n, n, cb), dtype=dt.dtype, device=dt.device)
    check(load().dlwpcs_up2cat_bwd(ptr(dt), ptr(da), ptr(db), bs, n, ca, cb, dtype_code(dt.dtype), stream_ptr()))
    return da, db
