"""Raw (non-autograd) device ops: torch tensors in, torch tensors out, arithmetic in libmsg_b200.so.

This is the torch-facing side of the C-ABI: it checks devices/dtypes the way the reference's pybind
shims do (``CHECK_CUDA``, multi_stylegan/op_static/fused_bias_act.cpp:13-14), forces contiguity like
the reference launchers (fused_bias_act_kernel.cu:58-60), allocates outputs and workspaces through
torch's caching allocator and launches on torch's current stream.  Nothing here falls back to
PyTorch arithmetic.
"""
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvDesc

import ctypes

_DTYPES = {torch.float32: _lib.MSG_F32, torch.float64: _lib.MSG_F64}

# engine selection for the conv primitives (tests flip this to cross-check the two engines)
conv_flags = _lib.CONV_AUTO
# The tcgen05 engine works on channels-last activations; with this on (default) the conv primitives
# convert NCHW inputs to torch.channels_last and return channels-last outputs (same logical shape).
conv_channels_last = True


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (multi_stylegan_b200 has no CPU path)" % name)


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return ctypes.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(t: torch.Tensor):
    """torch's current stream on the tensor's device as a raw cudaStream_t (host-side cost matters: the train step issues
    ~600 of our launches per iteration)."""
    if _raw_stream is not None:
        idx = t.device.index
        return ctypes.c_void_p(_raw_stream(idx if idx is not None else torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _on_device(object):
    """Device guard that costs nothing when the tensor already lives on the current device (the usual case)."""
    __slots__ = ("guard",)

    def __init__(self, device: torch.device) -> None:
        idx = device.index
        self.guard = None if (idx is None or idx == torch.cuda.current_device()) else torch.cuda.device(device)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            return self.guard.__exit__(*exc)
        return False


def _aligned(t: torch.Tensor) -> torch.Tensor:
    """Contiguous and 16-byte aligned (TMA / 128-bit accesses)."""
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def _is_cl(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def _act(t: torch.Tensor):
    """Activation operand of a conv primitive -> (dense tensor, layout code)."""
    if conv_channels_last or (_is_cl(t) and not t.is_contiguous()):
        t = t.contiguous(memory_format=torch.channels_last)
        if t.data_ptr() % 16:
            t = t.clone(memory_format=torch.channels_last)
        return t, _lib.LAYOUT_NHWC
    return _aligned(t), _lib.LAYOUT_NCHW


def _empty_act(shape, layout, device) -> torch.Tensor:
    fmt = torch.channels_last if layout == _lib.LAYOUT_NHWC else torch.contiguous_format
    return torch.empty(shape, dtype=torch.float32, device=device, memory_format=fmt)


def _dtype_code(t: torch.Tensor, what: str) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError("%s: dtype %s not supported (float32 / float64 only)" % (what, t.dtype))


def _workspace(nbytes: int, device) -> Tuple[Optional[torch.Tensor], Optional[ctypes.c_void_p]]:
    if nbytes <= 0:
        return None, None
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return ws, ctypes.c_void_p(ws.data_ptr())


# ---------------------------------------------------------------------------------------------------
# fused_bias_act — same signature as the reference's fused_act_cuda.fused_bias_act
# (multi_stylegan/op_static/fused_bias_act.cpp:11-20)
# ---------------------------------------------------------------------------------------------------
def _cl_only(t: torch.Tensor) -> bool:
    """4-D tensor stored channels-last (and not also plain-contiguous): memory order [B, H, W, C]."""
    return t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def fused_bias_act(input: torch.Tensor, bias: torch.Tensor, refer: torch.Tensor, act: int, grad: int,
                   alpha: float, scale: float) -> torch.Tensor:
    """Channels-last inputs are processed in place of their layout (bias index = memory index % C) and the
    result keeps the layout; anything else is made contiguous like the reference launcher does."""
    _require_cuda(input, "input")
    _require_cuda(bias, "bias")
    cl = _cl_only(input) and input.data_ptr() % 16 == 0
    fmt = torch.channels_last if cl else torch.contiguous_format
    x = input if cl else _aligned(input)
    b = bias.contiguous()
    ref = refer
    if refer.numel():
        ref = refer.contiguous(memory_format=fmt) if refer.dim() == 4 else refer.contiguous()
        if ref.data_ptr() % 16:
            ref = ref.clone(memory_format=fmt if refer.dim() == 4 else torch.contiguous_format)
    dt = _dtype_code(x, "fused_bias_act")
    if b.numel() and b.dtype != x.dtype:
        raise RuntimeError("fused_bias_act: bias dtype %s != input dtype %s" % (b.dtype, x.dtype))
    if ref.numel() and (ref.dtype != x.dtype or ref.shape != x.shape):
        raise RuntimeError("fused_bias_act: refer must match input in dtype and size")
    if cl:
        step_b = 1
    else:
        step_b = 1
        for i in range(2, x.dim()):
            step_b *= x.size(i)
    out = torch.empty_like(x)
    with _on_device(x.device):
        rc = _lib.lib().msg_fused_bias_act(_ptr(out), _ptr(x), _ptr(b), _ptr(ref), int(act), int(grad),
                                           float(alpha), float(scale), x.numel(), step_b, b.numel(), dt,
                                           _stream(x))
    _lib.check(rc, "fused_bias_act")
    return out


def fused_bias_act_bwd(grad_output: torch.Tensor, out: torch.Tensor, alpha: float, scale: float,
                       channels: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(grad_input, grad_bias) of lrelu(x + b) * scale in one pass over the data
    (reference: op_static/fused_act.py:31-40 = kernel + separate ATen sum).  Follows the layout of `out`."""
    _require_cuda(grad_output, "grad_output")
    _require_cuda(out, "out")
    if out.shape != grad_output.shape or out.dtype != grad_output.dtype:
        raise RuntimeError("fused_bias_act_bwd: out must match grad_output")
    cl = _cl_only(out) and out.data_ptr() % 16 == 0
    if cl:
        ref = out
        g = grad_output.contiguous(memory_format=torch.channels_last)
        if g.data_ptr() % 16:
            g = g.clone(memory_format=torch.channels_last)
    else:
        g = _aligned(grad_output)
        ref = _aligned(out)
    dt = _dtype_code(g, "fused_bias_act_bwd")
    if g.dim() < 2 or g.size(1) != channels:
        raise RuntimeError("fused_bias_act_bwd: dim 1 of grad_output must be the bias dimension")
    step_b = 1
    if not cl:
        for i in range(2, g.dim()):
            step_b *= g.size(i)
    dx = torch.empty_like(g)
    db = torch.empty(channels, dtype=g.dtype, device=g.device)
    L = _lib.lib()
    with _on_device(g.device):
        nbytes = L.msg_fused_bias_act_bwd_workspace(g.numel(), step_b, channels, dt)
        ws, wsp = _workspace(nbytes, g.device)
        rc = L.msg_fused_bias_act_bwd(_ptr(dx), ctypes.c_void_p(db.data_ptr()), _ptr(g), _ptr(ref), float(alpha),
                                      float(scale), g.numel(), max(step_b, 1), channels, wsp, nbytes, dt,
                                      _stream(g))
    _lib.check(rc, "fused_bias_act_bwd")
    return dx, db


# ---------------------------------------------------------------------------------------------------
# upfirdn2d — same signature as the reference's upfirdn2d_cuda.upfirdn2d
# (multi_stylegan/op_static/upfirdn2d.cpp:12-22); input is [major, in_h, in_w, minor]
# ---------------------------------------------------------------------------------------------------
def upfirdn2d(input: torch.Tensor, kernel: torch.Tensor, up_x: int, up_y: int, down_x: int, down_y: int,
              pad_x0: int, pad_x1: int, pad_y0: int, pad_y1: int) -> torch.Tensor:
    _require_cuda(input, "input")
    _require_cuda(kernel, "kernel")
    if input.dim() != 4 or kernel.dim() != 2:
        raise RuntimeError("upfirdn2d: input must be [major, h, w, minor] and kernel [kh, kw]")
    x = _aligned(input)
    k = kernel.contiguous()
    dt = _dtype_code(x, "upfirdn2d")
    if k.dtype != x.dtype:
        k = k.to(x.dtype)
    major, in_h, in_w, minor = x.shape
    kh, kw = k.shape
    L = _lib.lib()
    out_h = L.msg_upfirdn2d_out_size(in_h, up_y, down_y, pad_y0, pad_y1, kh)
    out_w = L.msg_upfirdn2d_out_size(in_w, up_x, down_x, pad_x0, pad_x1, kw)
    if out_h < 0 or out_w < 0:
        raise RuntimeError("upfirdn2d: negative output size (%d, %d)" % (out_h, out_w))
    out = torch.empty((major, out_h, out_w, minor), dtype=x.dtype, device=x.device)
    with _on_device(x.device):
        rc = L.msg_upfirdn2d(_ptr(out), _ptr(x), _ptr(k), major, in_h, in_w, minor, kh, kw, up_x, up_y, down_x,
                             down_y, pad_x0, pad_x1, pad_y0, pad_y1, dt, _stream(x))
    _lib.check(rc, "upfirdn2d")
    return out


def blur_noise_bias_act(x: torch.Tensor, kernel: torch.Tensor, pad: Sequence[int], noise: Optional[torch.Tensor],
                        noise_w: Optional[torch.Tensor], bias: Optional[torch.Tensor], slope: float,
                        gain: float) -> torch.Tensor:
    """lrelu(fir(x) + noise_w * noise + bias[c]) * gain in the channels-last blur kernel (up = down = 1).
    x [B,C,H,W] (C % 4 == 0), kernel [kh,kw] <= 4x4, pad (x0, x1, y0, y1), noise [B or 1, 1, OH, OW]."""
    _check_f32(x, "x")
    _require_cuda(kernel, "kernel")
    x = _cl(x)
    k = kernel.contiguous().to(torch.float32)
    B, C, H, W = x.shape
    kh, kw = k.shape
    px0, px1, py0, py1 = [int(v) for v in pad]
    L = _lib.lib()
    out_h = L.msg_upfirdn2d_out_size(H, 1, 1, py0, py1, kh)
    out_w = L.msg_upfirdn2d_out_size(W, 1, 1, px0, px1, kw)
    if out_h < 0 or out_w < 0:
        raise RuntimeError("blur_noise_bias_act: negative output size")
    out = torch.empty((B, C, out_h, out_w), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    nbs = 0
    if noise is not None:
        _check_f32(noise, "noise")
        noise, noise_w = _aligned(noise), _aligned(noise_w)
        if noise.numel() == B * out_h * out_w:
            nbs = out_h * out_w
        elif noise.numel() != out_h * out_w:
            raise RuntimeError("blur_noise_bias_act: noise must be [B or 1, 1, OH, OW]")
    if bias is not None:
        bias = _aligned(bias)
    with _on_device(x.device):
        rc = L.msg_upfirdn2d_bias_act(_ptr(out), _ptr(x), _ptr(k), B, H, W, C, kh, kw, px0, px1, py0, py1, _ptr(noise),
                                      _ptr(noise_w) if noise is not None else None, nbs, _ptr(bias), 1, float(slope),
                                      float(gain), _stream(x))
    _lib.check(rc, "upfirdn2d_bias_act")
    return out


# ---------------------------------------------------------------------------------------------------
# conv primitives (fp32).  w is [O, C, kh, kw] (shared) or [B, O, C, kh, kw] (one filter bank per sample,
# the reference's groups=B trick, multi_stylegan_generator.py:390-411).
# ---------------------------------------------------------------------------------------------------
def _pair(v) -> Tuple[int, int]:
    if isinstance(v, (tuple, list)):
        return int(v[0]), int(v[1])
    return int(v), int(v)


def _conv_desc(B, C, H, W, O, kh, kw, stride, padding, per_sample, layout, w_transposed=False) -> ConvDesc:
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    if H + 2 * ph < kh or W + 2 * pw < kw:
        raise RuntimeError("conv2d: kernel (%d,%d) larger than padded input (%d,%d)" % (kh, kw, H + 2 * ph, W + 2 * pw))
    d = ConvDesc()
    d.B, d.C, d.H, d.W, d.O, d.kh, d.kw = B, C, H, W, O, kh, kw
    d.stride_h, d.stride_w, d.pad_h, d.pad_w = sh, sw, ph, pw
    d.OH = (H + 2 * ph - kh) // sh + 1
    d.OW = (W + 2 * pw - kw) // sw + 1
    d.w_batch_stride = O * C * kh * kw if per_sample else 0
    d.layout = layout
    d.w_transposed = 1 if w_transposed else 0
    return d


def _check_f32(t: torch.Tensor, name: str) -> None:
    _require_cuda(t, name)
    if t.dtype != torch.float32:
        raise RuntimeError("conv2d: %s must be float32, got %s" % (name, t.dtype))


def _w_dims(w: torch.Tensor, B: int, w_transposed: bool = False):
    """(per_sample, O, C, kh, kw) of a filter tensor [O,C,kh,kw] / [B,O,C,kh,kw]; with w_transposed the two channel
    dimensions are stored the other way round ([C,O,kh,kw]: conv_transpose2d's weight layout)."""
    if w.dim() == 5:
        if w.size(0) != B:
            raise RuntimeError("conv2d: per-sample weight batch %d != input batch %d" % (w.size(0), B))
        a, b = w.size(1), w.size(2)
        return (True, b, a, w.size(3), w.size(4)) if w_transposed else (True, a, b, w.size(3), w.size(4))
    if w.dim() == 4:
        a, b = w.size(0), w.size(1)
        return (False, b, a, w.size(2), w.size(3)) if w_transposed else (False, a, b, w.size(2), w.size(3))
    raise RuntimeError("conv2d: weight must be [O,C,kh,kw] or [B,O,C,kh,kw]")


def conv2d_forward(x: torch.Tensor, w: torch.Tensor, stride=1, padding=0, alpha: float = 1.0,
                   bias: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                   noise_w: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None, act: bool = False,
                   slope: float = 0.2, gain: float = 1.0, w_transposed: bool = False,
                   x2: Optional[torch.Tensor] = None, col_scale: Optional[torch.Tensor] = None,
                   out2_scale: Optional[torch.Tensor] = None):
    """y = epilogue(alpha * conv(x, w)); the optional epilogue (noise, bias, leaky ReLU, residual add, gain) runs
    inside the conv kernel (msg_conv_epilogue, include/msg_b200.h).  With `x2` the convolution is applied to the channel
    concatenation [x | x2] without materialising it (msg_conv2d_forward_cat2; see cat2_supported).
    `col_scale` [B or 1, O] multiplies the accumulator per sample and output channel (the demodulation factor of the
    shared-weight modulated convolution); with `out2_scale` [B or 1, O] the call returns (y, y * out2_scale)."""
    _check_f32(x, "x")
    _check_f32(w, "w")
    x, layout = _act(x)
    w = _aligned(w)
    c1 = x.shape[1]
    if x2 is not None:
        _check_f32(x2, "x2")
        x2, layout2 = _act(x2)
        if layout2 != layout or x2.shape[0] != x.shape[0] or x2.shape[2:] != x.shape[2:]:
            raise RuntimeError("conv2d: x and x2 must share batch, spatial size and memory format")
    B, _, H, W = x.shape
    C = c1 + (x2.shape[1] if x2 is not None else 0)
    per_sample, O, Cw, kh, kw = _w_dims(w, B, w_transposed)
    if Cw != C:
        raise RuntimeError("conv2d: weight has %d input channels, input has %d" % (Cw, C))
    d = _conv_desc(B, C, H, W, O, kh, kw, stride, padding, per_sample, layout, w_transposed)
    y = _empty_act((B, O, d.OH, d.OW), layout, x.device)
    fused = (bias is not None or noise is not None or add is not None or act or gain != 1.0 or col_scale is not None
             or out2_scale is not None)
    ep = None
    keep = []
    y2 = None
    if fused:
        ep = _lib.ConvEpilogue()
        ep.act, ep.slope, ep.gain = (1 if act else 0), float(slope), float(gain)
        if col_scale is not None:
            _check_f32(col_scale, "col_scale")
            col_scale = _aligned(col_scale)
            if col_scale.numel() not in (O, B * O):
                raise RuntimeError("conv2d: col_scale must be [B or 1, %d]" % O)
            ep.col_scale = col_scale.data_ptr()
            ep.col_scale_batch_stride = O if col_scale.numel() == B * O else 0
        if out2_scale is not None:
            _check_f32(out2_scale, "out2_scale")
            out2_scale = _aligned(out2_scale)
            if out2_scale.numel() not in (O, B * O):
                raise RuntimeError("conv2d: out2_scale must be [B or 1, %d]" % O)
            y2 = torch.empty_like(y)
            ep.y2, ep.y2_scale = y2.data_ptr(), out2_scale.data_ptr()
            ep.y2_scale_batch_stride = O if out2_scale.numel() == B * O else 0
        keep += [col_scale, out2_scale]
        if bias is not None:
            _check_f32(bias, "bias")
            bias = _aligned(bias)
            if bias.numel() != O:
                raise RuntimeError("conv2d: epilogue bias must have %d elements" % O)
            ep.bias = bias.data_ptr()
        if noise is not None:
            _check_f32(noise, "noise")
            noise, noise_w = _aligned(noise), _aligned(noise_w)
            if noise.numel() == d.OH * d.OW:
                ep.noise_batch_stride = 0
            elif noise.numel() == B * d.OH * d.OW:
                ep.noise_batch_stride = d.OH * d.OW
            else:
                raise RuntimeError("conv2d: epilogue noise must be [B or 1, 1, OH, OW]")
            ep.noise, ep.noise_w = noise.data_ptr(), noise_w.data_ptr()
        if add is not None:
            _check_f32(add, "add")
            if add.shape != y.shape:
                raise RuntimeError("conv2d: epilogue add must have the output's shape")
            add = add.contiguous(memory_format=torch.channels_last if layout == _lib.LAYOUT_NHWC
                                 else torch.contiguous_format)
            if add.data_ptr() % 16:
                add = add.clone()
            ep.add = add.data_ptr()
        keep += [bias, noise, noise_w, add]
    L = _lib.lib()
    with _on_device(x.device):
        nbytes = L.msg_conv2d_workspace(ctypes.byref(d), 0, conv_flags)
        ws, wsp = _workspace(nbytes, x.device)
        if x2 is None:
            rc = L.msg_conv2d_forward_fused(_ptr(y), _ptr(x), _ptr(w), ctypes.byref(d), float(alpha),
                                            ctypes.byref(ep) if ep is not None else None, wsp, nbytes, conv_flags,
                                            _stream(x))
        else:
            rc = L.msg_conv2d_forward_cat2(_ptr(y), _ptr(x), int(c1), _ptr(x2), _ptr(w), ctypes.byref(d), float(alpha),
                                           ctypes.byref(ep) if ep is not None else None, wsp, nbytes, conv_flags,
                                           _stream(x))
    _lib.check(rc, "conv2d_forward")
    del keep
    if out2_scale is not None:
        return y, y2
    return y


def cat2_supported(x: torch.Tensor, x2: torch.Tensor, stride=1) -> bool:
    """Whether conv2d_forward(x, w, x2=x2) can read the concatenation in place (msg_conv2d_forward_cat2): tcgen05
    engine, channels-last CUDA tensors, stride 1, both channel counts multiples of 32."""
    s = stride if isinstance(stride, int) else stride[0]
    return (conv_channels_last and conv_flags != _lib.CONV_FORCE_SIMT and x.is_cuda and x2.is_cuda and s == 1
            and x.shape[1] % 32 == 0 and x2.shape[1] % 32 == 0 and x.dtype == torch.float32 and x2.dtype == torch.float32
            and tensor_core_path_available())


def conv2d_dgrad(dy: torch.Tensor, w: torch.Tensor, in_hw: Sequence[int], stride=1, padding=0,
                 alpha: float = 1.0, w_transposed: bool = False, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx of conv2d(x, w) given dy; also conv_transpose2d(dy, w) with output size in_hw.  `add` (shape of dx) is summed
    into the result inside the kernel's epilogue (gradient accumulation without a separate pass)."""
    _check_f32(dy, "dy")
    _check_f32(w, "w")
    dy, layout = _act(dy)
    w = _aligned(w)
    B, O, OH, OW = dy.shape
    per_sample, Ow, C, kh, kw = _w_dims(w, B, w_transposed)
    if Ow != O:
        raise RuntimeError("conv2d_dgrad: weight has %d output channels, dy has %d" % (Ow, O))
    H, W = int(in_hw[0]), int(in_hw[1])
    d = _conv_desc(B, C, H, W, O, kh, kw, stride, padding, per_sample, layout, w_transposed)
    if (d.OH, d.OW) != (OH, OW):
        raise RuntimeError("conv2d_dgrad: dy spatial size (%d,%d) inconsistent with input size (%d,%d)" % (OH, OW, H, W))
    dx = _empty_act((B, C, H, W), layout, dy.device)
    if add is not None:
        _check_f32(add, "add")
        if tuple(add.shape) != (B, C, H, W):
            raise RuntimeError("conv2d_dgrad: `add` must have the shape of dx")
        add = add.contiguous(memory_format=torch.channels_last if layout == _lib.LAYOUT_NHWC else torch.contiguous_format)
        if add.data_ptr() % 16:
            add = add.clone()
    L = _lib.lib()
    with _on_device(dy.device):
        nbytes = L.msg_conv2d_workspace(ctypes.byref(d), 1, conv_flags)
        ws, wsp = _workspace(nbytes, dy.device)
        rc = L.msg_conv2d_dgrad_acc(_ptr(dx), _ptr(dy), _ptr(w), ctypes.byref(d), float(alpha), _ptr(add), wsp, nbytes,
                                    conv_flags, _stream(dy))
    _lib.check(rc, "conv2d_dgrad")
    return dx


def conv2d_dgrad_act_bwd(dy: torch.Tensor, w: torch.Tensor, ref: torch.Tensor, padding=0, alpha: float = 1.0,
                         slope: float = 0.2, gain: float = 1.0, want_dbias: bool = True):
    """(g_pre, dbias) = activation backward of the layer that produced `ref`, applied to the stride-1 dgrad of the NEXT
    layer inside that dgrad's epilogue:  g_pre = alpha * conv^T(dy, w) * (ref > 0 ? 1 : slope) * gain,  dbias = sum over
    (b, y, x) of g_pre.  Returns None when the shape is not eligible (the caller then runs conv2d_dgrad +
    noise_bias_act_cl_bwd): tcgen05 engine, channels-last, shared weights."""
    if w.dim() != 4 or not conv_channels_last or not dy.is_cuda:
        return None
    _check_f32(dy, "dy")
    _check_f32(w, "w")
    _check_f32(ref, "ref")
    dy, layout = _act(dy)
    w = _aligned(w)
    B, O, OH, OW = dy.shape
    per_sample, Ow, C, kh, kw = _w_dims(w, B, False)
    if Ow != O:
        raise RuntimeError("conv2d_dgrad_act_bwd: weight has %d output channels, dy has %d" % (Ow, O))
    H, W = int(ref.shape[2]), int(ref.shape[3])
    if tuple(ref.shape) != (B, C, H, W):
        raise RuntimeError("conv2d_dgrad_act_bwd: `ref` must have the shape of dx")
    d = _conv_desc(B, C, H, W, O, kh, kw, 1, padding, per_sample, layout, False)
    if (d.OH, d.OW) != (OH, OW):
        raise RuntimeError("conv2d_dgrad_act_bwd: dy spatial size inconsistent with `ref`")
    L = _lib.lib()
    if layout != _lib.LAYOUT_NHWC or not L.msg_conv2d_dgrad_mask_supported(ctypes.byref(d), conv_flags):
        return None
    ref = ref.contiguous(memory_format=torch.channels_last)
    if ref.data_ptr() % 16:
        return None
    dx = _empty_act((B, C, H, W), layout, dy.device)
    dbias = torch.empty(C, device=dy.device, dtype=torch.float32) if want_dbias else None
    with _on_device(dy.device):
        nbytes = L.msg_conv2d_workspace(ctypes.byref(d), 1, conv_flags)
        ws, wsp = _workspace(nbytes, dy.device)
        rc = L.msg_conv2d_dgrad_mask(_ptr(dx), _ptr(dbias), _ptr(dy), _ptr(w), ctypes.byref(d), float(alpha), _ptr(ref),
                                     float(slope), float(gain), wsp, nbytes, conv_flags, _stream(dy))
    _lib.check(rc, "conv2d_dgrad_act_bwd")
    return dx, dbias


def conv2d_wgrad(dy: torch.Tensor, x: torch.Tensor, khw: Sequence[int], stride=1, padding=0,
                 per_sample: bool = False, alpha: float = 1.0, w_transposed: bool = False) -> torch.Tensor:
    _check_f32(dy, "dy")
    _check_f32(x, "x")
    x, layout = _act(x)
    dy, layout_dy = _act(dy)
    if layout_dy != layout:
        dy = dy.contiguous(memory_format=torch.channels_last if layout == _lib.LAYOUT_NHWC else torch.contiguous_format)
    B, C, H, W = x.shape
    if dy.size(0) != B:
        raise RuntimeError("conv2d_wgrad: batch mismatch")
    O = dy.size(1)
    kh, kw = int(khw[0]), int(khw[1])
    d = _conv_desc(B, C, H, W, O, kh, kw, stride, padding, per_sample, layout, w_transposed)
    if (d.OH, d.OW) != (dy.size(2), dy.size(3)):
        raise RuntimeError("conv2d_wgrad: dy spatial size inconsistent with x")
    oc = (C, O) if w_transposed else (O, C)
    shape = (B,) + oc + (kh, kw) if per_sample else oc + (kh, kw)
    dw = torch.empty(shape, dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with _on_device(x.device):
        nbytes = L.msg_conv2d_workspace(ctypes.byref(d), 2, conv_flags)
        ws, wsp = _workspace(nbytes, x.device)
        rc = L.msg_conv2d_wgrad(_ptr(dw), _ptr(dy), _ptr(x), ctypes.byref(d), float(alpha), wsp, nbytes,
                                conv_flags, _stream(x))
    _lib.check(rc, "conv2d_wgrad")
    return dw


def conv2d_last_engine() -> str:
    return {0: "none", 1: "simt", 2: "tcgen05"}[_lib.lib().msg_conv2d_last_engine()]


def tensor_core_path_available() -> bool:
    return bool(_lib.lib().msg_tensor_core_path_available())


def profile_enable(on: bool) -> None:
    _lib.lib().msg_profile_enable(1 if on else 0)


def profile_summary():
    """List of dicts, one per distinct tcgen05 conv launch shape (call after torch.cuda.synchronize())."""
    buf = (_lib.ProfileEntry * 256)()
    n = _lib.lib().msg_profile_summary(buf, 256)
    out = []
    for e in buf[:n]:
        if e.launches:
            out.append(dict(kind="pixgemm" if e.kind == 0 else "redgemm", taps=e.taps, k_channels=e.k_channels,
                            n_channels=e.n_channels, pixels=e.pixels, launches=e.launches, ms_total=e.ms_total,
                            flops_per_launch=e.flops_per_launch))
    return out


# ---------------------------------------------------------------------------------------------------
# small-M linears (csrc/linear_ops.cu)
# ---------------------------------------------------------------------------------------------------
def style_mapping_supported(depth: int, K: int) -> bool:
    return bool(_lib.lib().msg_style_mapping_supported(int(depth), int(K)))


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def style_mapping_forward(z: torch.Tensor, weights, biases, alpha: float, slope: float, gain: float, eps: float):
    """(acts [depth, M, K], x0 [M, K]): pixel norm + depth x [linear -> bias + leaky ReLU] in one launch; acts[-1] is
    the mapped latent."""
    _check_f32(z, "z")
    z = z.contiguous()
    M, K = z.shape
    L = len(weights)
    ws = [w.contiguous() for w in weights]
    for w in ws:
        _check_f32(w, "weight")
        if tuple(w.shape) != (K, K):
            raise RuntimeError("style_mapping_forward: weights must be [K, K]")
    acts = torch.empty((L, M, K), device=z.device, dtype=torch.float32)
    x0 = torch.empty((M, K), device=z.device, dtype=torch.float32)
    with _on_device(z.device):
        rc = _lib.lib().msg_style_mapping_forward(_ptr(acts), _ptr(x0), _ptr(z), _ptr_array(ws), _ptr_array(list(biases)),
                                                  L, M, K, float(alpha), float(slope), float(gain), float(eps), _stream(z))
    _lib.check(rc, "style_mapping_forward")
    return acts, x0


def style_mapping_backward(gy: torch.Tensor, acts: torch.Tensor, x0: torch.Tensor, weights, biases, alpha: float,
                           slope: float, gain: float):
    """(dW [depth, K, K], db [depth, K]) of style_mapping_forward."""
    gy = gy.contiguous()
    L, M, K = acts.shape
    ws = [w.contiguous() for w in weights]
    dW = torch.empty((L, K, K), device=gy.device, dtype=torch.float32)
    db = torch.zeros((L, K), device=gy.device, dtype=torch.float32)
    work = torch.empty((L, M, K), device=gy.device, dtype=torch.float32)
    with _on_device(gy.device):
        rc = _lib.lib().msg_style_mapping_backward(_ptr(dW), _ptr(db), _ptr(gy), _ptr(acts), _ptr(x0), _ptr_array(ws),
                                                   _ptr_array(list(biases)), L, M, K, float(alpha), float(slope), float(gain),
                                                   _ptr(work), _stream(gy))
    _lib.check(rc, "style_mapping_backward")
    return dW, db


class LinearGroup(object):
    """Host-side description of a group of linears reading slices of one [M, R] input (msg_linear_item table)."""

    def __init__(self, specs):
        """specs: list of (W [N, K], bias [N] or None, in_off, alpha, beta)."""
        n = len(specs)
        self.items = (_lib.LinearItem * n)()
        self.params = []
        out_off = w_off = b_off = 0
        self.out_slices, self.w_slices, self.b_slices = [], [], []
        slots = {}
        for i, (W, b, in_off, alpha, beta) in enumerate(specs):
            N, K = W.shape
            it = self.items[i]
            it.W, it.bias = W.data_ptr(), (None if b is None else b.data_ptr())
            it.N, it.K, it.in_off, it.out_off, it.w_off, it.b_off = N, K, int(in_off), out_off, w_off, b_off
            it.alpha, it.beta = float(alpha), float(beta)
            self.out_slices.append((out_off, N))
            self.w_slices.append((w_off, N, K))
            self.b_slices.append((b_off, N) if b is not None else None)
            slots.setdefault((int(in_off), K), []).append(i)
            out_off += N
            w_off += N * K
            b_off += N
            self.params.append((W, b))
        self.out_row, self.w_total, self.b_total = out_off, w_off, b_off
        self.max_n = max(W.shape[0] for W, *_ in specs)
        self.max_k = max(W.shape[1] for W, *_ in specs)
        # items of one slot must be adjacent in the table: reorder is not needed when specs arrive slot by slot; in general
        # the slot table lists runs of adjacent items (several runs per slice are summed by the caller — not used here)
        runs = []
        for (in_off, K), idx in slots.items():
            start = idx[0]
            for a, b2 in zip(idx, idx[1:]):
                if b2 != a + 1:
                    raise RuntimeError("LinearGroup: items reading the same input slice must be adjacent")
            runs.append((in_off, K, start, len(idx)))
        self.slots = (_lib.LinearSlot * len(runs))()
        for j, (in_off, K, first, count) in enumerate(runs):
            s = self.slots[j]
            s.in_off, s.K, s.first, s.count = in_off, K, first, count
        self.covered = sorted((in_off, K) for in_off, K, _, _ in runs)
        self.key = tuple((W.data_ptr(), None if b is None else b.data_ptr()) for W, b in self.params)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        M, R = x.shape
        out = torch.empty(M * self.out_row, device=x.device, dtype=torch.float32)      # item-major: [M, N_i] blocks
        with _on_device(x.device):
            rc = _lib.lib().msg_linear_group_forward(_ptr(out), _ptr(x), R, self.items, len(self.items), M,
                                                     self.max_n, self.max_k, _stream(x))
        _lib.check(rc, "linear_group_forward")
        return out

    def backward(self, gout: torch.Tensor, x: torch.Tensor, need_dx: bool):
        M, R = x.shape
        dW = torch.empty(self.w_total, device=x.device, dtype=torch.float32)
        db = torch.zeros(max(self.b_total, 1), device=x.device, dtype=torch.float32)
        dx = None
        if need_dx:
            full = sum(k for _, k in self.covered) == R and all(a[0] + a[1] == b[0] for a, b in zip(self.covered, self.covered[1:]))
            dx = torch.empty((M, R), device=x.device, dtype=torch.float32) if full else torch.zeros((M, R), device=x.device, dtype=torch.float32)
        with _on_device(x.device):
            rc = _lib.lib().msg_linear_group_backward(_ptr(dW), _ptr(db), _ptr(dx), _ptr(gout), _ptr(x), R,
                                                      self.items, len(self.items), self.slots, len(self.slots), M,
                                                      self.max_n, self.max_k, _stream(x))
        _lib.check(rc, "linear_group_backward")
        return dW, db, dx


# ---------------------------------------------------------------------------------------------------
# non-local block passes (csrc/attention_ops.cu); all tensors channels-last [B, C, H, W] views
# ---------------------------------------------------------------------------------------------------
def _cl_empty(B, C, H, W, device):
    return torch.empty((B, H, W, C), device=device, dtype=torch.float32).permute(0, 3, 1, 2)


def nl_split_pool(qkv: torch.Tensor, cq: int, cv: int):
    """qkv [B, cq+cq+cv, H, W] channels-last -> (theta [B,cq,H,W], phi_p [B,cq,H/2,W/2], g_p [B,cv,H/2,W/2], idx)."""
    _check_f32(qkv, "qkv")
    qkv = qkv.contiguous(memory_format=torch.channels_last)
    B, CT, H, W = qkv.shape
    if CT != 2 * cq + cv:
        raise RuntimeError("nl_split_pool: channel count")
    dev = qkv.device
    theta, phi, g = _cl_empty(B, cq, H, W, dev), _cl_empty(B, cq, H // 2, W // 2, dev), _cl_empty(B, cv, H // 2, W // 2, dev)
    idx = torch.empty((B, H // 2, W // 2, (cq + cv) // 4), device=dev, dtype=torch.int32)
    with _on_device(dev):
        rc = _lib.lib().msg_nl_split_pool(_ptr(theta), _ptr(phi), _ptr(g), _ptr(idx), _ptr(qkv), B, H, W, cq, cv, _stream(qkv))
    _lib.check(rc, "nl_split_pool")
    return theta, phi, g, idx


def nl_merge_unpool(dtheta: torch.Tensor, dphi: torch.Tensor, dg: torch.Tensor, idx: torch.Tensor):
    cl = torch.channels_last
    dtheta, dphi, dg = dtheta.contiguous(memory_format=cl), dphi.contiguous(memory_format=cl), dg.contiguous(memory_format=cl)
    B, cq, H, W = dtheta.shape
    cv = dg.shape[1]
    dqkv = _cl_empty(B, 2 * cq + cv, H, W, dtheta.device)
    with _on_device(dtheta.device):
        rc = _lib.lib().msg_nl_merge_unpool(_ptr(dqkv), _ptr(dtheta), _ptr(dphi), _ptr(dg), _ptr(idx), B, H, W, cq, cv,
                                            _stream(dtheta))
    _lib.check(rc, "nl_merge_unpool")
    return dqkv


def softmax_rows_(x: torch.Tensor, n: int) -> torch.Tensor:
    """In-place softmax over runs of n consecutive floats of the (dense) tensor x."""
    rows = x.numel() // n
    with _on_device(x.device):
        rc = _lib.lib().msg_softmax_rows(_ptr(x), rows, int(n), _stream(x))
    _lib.check(rc, "softmax_rows")
    return x


def softmax_rows_bwd_(dp: torch.Tensor, p: torch.Tensor, n: int) -> torch.Tensor:
    rows = dp.numel() // n
    with _on_device(dp.device):
        rc = _lib.lib().msg_softmax_rows_bwd(_ptr(dp), _ptr(p), rows, int(n), _stream(dp))
    _lib.check(rc, "softmax_rows_bwd")
    return dp


def demod_factors_bwd(gd: torch.Tensor, d: torch.Tensor, s: torch.Tensor, wsq: torch.Tensor, W: torch.Tensor, scale: float,
                      need_w: bool = True, need_s: bool = True):
    """(dW [O,C,kh,kw] or None, ds [B,C] or None) of demod_factors."""
    gd, d, s = gd.contiguous(), d.contiguous(), s.contiguous()
    W = W.contiguous()
    O, C = wsq.shape
    B = s.shape[0]
    taps = W.numel() // (O * C)
    dW = torch.empty_like(W) if need_w else None
    ds = torch.empty_like(s) if need_s else None
    with _on_device(W.device):
        rc = _lib.lib().msg_demod_factors_bwd(_ptr(dW), _ptr(ds), _ptr(gd), _ptr(d), _ptr(s), _ptr(wsq), _ptr(W), B, O, C, taps,
                                              float(scale), _stream(W))
    _lib.check(rc, "demod_factors_bwd")
    return dW, ds


def colsum_cl(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """sum over (batch, height, width) of a channels-last [B, C, H, W] activation, times scale -> [C]; deterministic."""
    x = x.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x.shape
    out = torch.empty(C, device=x.device, dtype=torch.float32)
    L = _lib.lib()
    with _on_device(x.device):
        nbytes = L.msg_colsum_workspace(B * H * W, C)
        ws, wsp = _workspace(nbytes, x.device)
        rc = L.msg_colsum_nhwc(_ptr(out), _ptr(x), B * H * W, C, float(scale), wsp, nbytes, _stream(x))
    _lib.check(rc, "colsum_nhwc")
    return out


def dot(a: torch.Tensor, b: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """scale * sum(a * b) for two dense tensors with the same memory layout -> [1]; deterministic."""
    if a.shape != b.shape or a.stride() != b.stride():
        raise RuntimeError("dot: operands must share shape and strides")
    out = torch.empty(1, device=a.device, dtype=torch.float32)
    work = torch.empty(4096, device=a.device, dtype=torch.float32)
    with _on_device(a.device):
        rc = _lib.lib().msg_dot(_ptr(out), _ptr(a), _ptr(b), a.numel(), float(scale), _ptr(work), _stream(a))
    _lib.check(rc, "dot")
    return out


def mbstd_forward(x: torch.Tensor, groups: int, alpha: float) -> torch.Tensor:
    """cat(x, minibatch-stddev plane) for a channels-last [B, C, H, W] activation -> [B, C+1, H, W] channels-last."""
    _check_f32(x, "x")
    x = x.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x.shape
    out = _cl_empty(B, C + 1, H, W, x.device)
    L = _lib.lib()
    with _on_device(x.device):
        nbytes = L.msg_mbstd_workspace(int(groups))
        ws, wsp = _workspace(nbytes, x.device)
        rc = L.msg_mbstd_forward(_ptr(out), _ptr(x), B, C, H * W, int(groups), float(alpha), wsp, nbytes, _stream(x))
    _lib.check(rc, "mbstd_forward")
    return out


def mbstd_backward(gout: torch.Tensor, x: torch.Tensor, groups: int, alpha: float) -> torch.Tensor:
    gout = gout.contiguous(memory_format=torch.channels_last)
    x = x.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x.shape
    gx = _cl_empty(B, C, H, W, x.device)
    L = _lib.lib()
    with _on_device(x.device):
        nbytes = L.msg_mbstd_workspace(int(groups))
        ws, wsp = _workspace(nbytes, x.device)
        rc = L.msg_mbstd_backward(_ptr(gx), _ptr(gout), _ptr(x), B, C, H * W, int(groups), float(alpha), wsp, nbytes, _stream(x))
    _lib.check(rc, "mbstd_backward")
    return gx


def tf32_mma_rate_probe(iters: int, device) -> float:
    """Launch the tensor-core issue-rate probe (csrc/mma_rate.cu) on torch's current stream; returns its FLOPs."""
    flops = ctypes.c_double(0.0)
    dev = torch.device(device)
    with _on_device(dev):
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        stream = ctypes.c_void_p(_raw_stream(idx)) if _raw_stream is not None else ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = _lib.lib().msg_tf32_mma_rate_probe(int(iters), None, ctypes.byref(flops), stream)
    _lib.check(rc, "tf32_mma_rate_probe")
    return float(flops.value)


def launch_count() -> int:
    return int(_lib.lib().msg_launch_count())


# ---------------------------------------------------------------------------------------------------
# small fused kernels
# ---------------------------------------------------------------------------------------------------
def modulate_weights(W: torch.Tensor, s: torch.Tensor, scale: float, demodulate: bool):
    """W [O,C,kh,kw], s [B,C] -> (w_mod [B,O,C,kh,kw], demod [B,O] or None)
    (multi_stylegan/multi_stylegan_generator.py:384-388)."""
    _check_f32(W, "W")
    _check_f32(s, "s")
    W = _aligned(W)
    s = _aligned(s)
    O, C, kh, kw = W.shape
    B = s.size(0)
    if s.shape != (B, C):
        raise RuntimeError("modulate_weights: style must be [B, C]")
    out = torch.empty((B, O, C, kh, kw), dtype=torch.float32, device=W.device)
    dm = torch.empty((B, O), dtype=torch.float32, device=W.device) if demodulate else None
    with _on_device(W.device):
        rc = _lib.lib().msg_modulate_weights(_ptr(out), _ptr(dm), _ptr(W), _ptr(s), B, O, C, kh * kw, float(scale),
                                             1 if demodulate else 0, _stream(W))
    _lib.check(rc, "modulate_weights")
    return out, dm


def modulate_weights_bwd(g: torch.Tensor, W: torch.Tensor, s: torch.Tensor, demod: Optional[torch.Tensor],
                         scale: float, demodulate: bool):
    """(dW [O,C,kh,kw], ds [B,C]) from g = dL/dw_mod [B,O,C,kh,kw] (first-order backward of modulate_weights)."""
    _check_f32(g, "g")
    _check_f32(W, "W")
    _check_f32(s, "s")
    g, W, s = _aligned(g), _aligned(W), _aligned(s)
    O, C, kh, kw = W.shape
    B = s.size(0)
    if g.shape != (B, O, C, kh, kw):
        raise RuntimeError("modulate_weights_bwd: g must be [B,O,C,kh,kw]")
    if demodulate:
        demod = _aligned(demod)
    dW = torch.empty_like(W)
    ds = torch.empty_like(s)
    L = _lib.lib()
    with _on_device(W.device):
        nbytes = L.msg_modulate_weights_bwd_workspace(B, O, C, kh * kw)
        ws, wsp = _workspace(nbytes, W.device)
        rc = L.msg_modulate_weights_bwd(_ptr(dW), _ptr(ds), _ptr(g), _ptr(W), _ptr(s), _ptr(demod) if demodulate else None,
                                        B, O, C, kh * kw, float(scale), 1 if demodulate else 0, wsp, nbytes, _stream(W))
    _lib.check(rc, "modulate_weights_bwd")
    return dW, ds


def noise_bias_act(x: torch.Tensor, noise: Optional[torch.Tensor], noise_w: Optional[torch.Tensor],
                   bias: Optional[torch.Tensor], alpha: float, scale: float) -> torch.Tensor:
    """lrelu(x + noise_w * noise + bias[c]) * scale — multi_stylegan_generator.py:292 fused with
    op_static/fused_act.py:58."""
    _check_f32(x, "x")
    x = _aligned(x)
    B, C = x.size(0), x.size(1)
    HW = x.numel() // max(B * C, 1)
    nbs = 0
    if noise is not None:
        _check_f32(noise, "noise")
        noise = _aligned(noise)
        if noise.numel() == HW:
            nbs = 0
        elif noise.numel() == B * HW:
            nbs = HW
        else:
            raise RuntimeError("noise_bias_act: noise must be [B or 1, 1, H, W]")
        noise_w = _aligned(noise_w)
    if bias is not None:
        bias = _aligned(bias)
    out = torch.empty_like(x)
    with _on_device(x.device):
        rc = _lib.lib().msg_noise_bias_act(_ptr(out), _ptr(x), _ptr(noise), _ptr(noise_w) if noise is not None else None,
                                           _ptr(bias), B, C, HW, nbs, float(alpha), float(scale), _stream(x))
    _lib.check(rc, "noise_bias_act")
    return out


def _cl(t: torch.Tensor) -> torch.Tensor:
    """Dense channels-last, 16-byte aligned (no copy when the producer was one of our kernels)."""
    t = t.contiguous(memory_format=torch.channels_last)
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.channels_last)
    return t


def noise_bias_act_cl(x: torch.Tensor, ref: Optional[torch.Tensor], noise: Optional[torch.Tensor],
                      noise_w: Optional[torch.Tensor], bias: Optional[torch.Tensor], alpha: float,
                      scale: float) -> torch.Tensor:
    """Channels-last StyledConv2d epilogue (multi_stylegan_generator.py:292 + op_static/fused_act.py:58):
    v = x + noise_w * noise + bias[c];  ref None: lrelu(v) * scale;  else (ref > 0 ? v : alpha v) * scale.
    x [B,C,H,W] (C % 4 == 0), noise [B or 1, 1, H, W]."""
    _check_f32(x, "x")
    x = _cl(x)
    B, C, H, W = x.shape
    if C % 4:
        raise RuntimeError("noise_bias_act_cl: channel count must be a multiple of 4")
    rows = B * H * W
    if ref is not None:
        _check_f32(ref, "ref")
        ref = _cl(ref)
        if ref.shape != x.shape:
            raise RuntimeError("noise_bias_act_cl: ref must match x")
    period = 1
    if noise is not None:
        _check_f32(noise, "noise")
        noise = _aligned(noise)
        if noise.numel() not in (H * W, rows):
            raise RuntimeError("noise_bias_act_cl: noise must be [B or 1, 1, H, W]")
        period = noise.numel()
        noise_w = _aligned(noise_w)
    if bias is not None:
        bias = _aligned(bias)
        if bias.numel() != C:
            raise RuntimeError("noise_bias_act_cl: bias must have C elements")
    out = torch.empty_like(x)
    with _on_device(x.device):
        rc = _lib.lib().msg_noise_bias_act_nhwc(_ptr(out), _ptr(x), _ptr(ref), _ptr(noise),
                                                _ptr(noise_w) if noise is not None else None, _ptr(bias), rows, C,
                                                period, float(alpha), float(scale), _stream(x))
    _lib.check(rc, "noise_bias_act_nhwc")
    return out


def noise_bias_act_cl_bwd(grad_output: torch.Tensor, out: torch.Tensor, noise: Optional[torch.Tensor], alpha: float,
                          scale: float, want_param_grads: bool = True):
    """(dx, dbias [C], dnoise_w [1] or None) in one pass + a tiny deterministic reduction; with
    want_param_grads=False only dx (no reduction)."""
    _check_f32(grad_output, "grad_output")
    _check_f32(out, "out")
    g = _cl(grad_output)
    ref = _cl(out)
    B, C, H, W = g.shape
    rows = B * H * W
    period = 1
    dnw = None
    if noise is not None and want_param_grads:
        noise = _aligned(noise)
        period = noise.numel()
        dnw = torch.empty(1, dtype=torch.float32, device=g.device)
    dx = torch.empty_like(g)
    db = torch.empty(C, dtype=torch.float32, device=g.device) if want_param_grads else None
    L = _lib.lib()
    with _on_device(g.device):
        nbytes = L.msg_noise_bias_act_nhwc_bwd_workspace(rows, C)
        ws, wsp = _workspace(nbytes, g.device)
        rc = L.msg_noise_bias_act_nhwc_bwd(_ptr(dx), ctypes.c_void_p(db.data_ptr()) if db is not None else None,
                                           ctypes.c_void_p(dnw.data_ptr()) if dnw is not None else None, _ptr(g),
                                           _ptr(ref), _ptr(noise), rows, C, period, float(alpha), float(scale), wsp,
                                           nbytes, _stream(g))
    _lib.check(rc, "noise_bias_act_nhwc_bwd")
    return dx, db, dnw


# ---------------------------------------------------------------------------------------------------
# shared-weight form of the modulated convolution (include/msg_b200.h: msg_demod_factors, msg_styled_act_bwd,
# msg_upfirdn2d_bias_act_mod)
# ---------------------------------------------------------------------------------------------------
def demod_factors(W: torch.Tensor, s: torch.Tensor, scale: float):
    """W [O,C,kh,kw], s [B,C] -> (d [B,O] = rsqrt(scale^2 sum_c s^2 sum_t W^2 + 1e-8), wsq [O,C] = sum_t W^2)
    (multi_stylegan_generator.py:386-388)."""
    _check_f32(W, "W")
    _check_f32(s, "s")
    W, s = _aligned(W), _aligned(s)
    O, C, kh, kw = W.shape
    B = s.size(0)
    if s.shape != (B, C):
        raise RuntimeError("demod_factors: style must be [B, C]")
    d = torch.empty((B, O), dtype=torch.float32, device=W.device)
    wsq = torch.empty((O, C), dtype=torch.float32, device=W.device)
    with _on_device(W.device):
        rc = _lib.lib().msg_demod_factors(_ptr(d), _ptr(wsq), _ptr(W), _ptr(s), B, O, C, kh * kw, float(scale), _stream(W))
    _lib.check(rc, "demod_factors")
    return d, wsq


def _bc_stride(t: Optional[torch.Tensor], B: int, C: int, name: str) -> int:
    if t is None:
        return 0
    if t.numel() == B * C:
        return C
    if t.numel() == C:
        return 0
    raise RuntimeError("%s must be [B or 1, %d]" % (name, C))


def styled_act_bwd(g_out: Optional[torch.Tensor], g_out2: Optional[torch.Tensor], out: torch.Tensor,
                   col_scale: Optional[torch.Tensor], out2_scale: Optional[torch.Tensor],
                   noise: Optional[torch.Tensor], slope: float, gain: float):
    """(g_pre [B,C,H,W] channels-last, sums [4,B,C]) — see msg_styled_act_bwd in include/msg_b200.h."""
    _check_f32(out, "out")
    out = _cl(out)
    B, C, H, W = out.shape
    if C % 4:
        raise RuntimeError("styled_act_bwd: channel count must be a multiple of 4")
    if g_out is None and g_out2 is None:
        raise RuntimeError("styled_act_bwd: at least one incoming gradient is needed")
    if g_out is not None:
        _check_f32(g_out, "g_out")
        g_out = _cl(g_out)
    if g_out2 is not None:
        _check_f32(g_out2, "g_out2")
        g_out2 = _cl(g_out2)
        if out2_scale is None:
            raise RuntimeError("styled_act_bwd: g_out2 needs out2_scale")
    for t in (g_out, g_out2):
        if t is not None and t.shape != out.shape:
            raise RuntimeError("styled_act_bwd: gradients must have the shape of out")
    if col_scale is not None:
        col_scale = _aligned(col_scale)
    if out2_scale is not None:
        out2_scale = _aligned(out2_scale)
    nbs = 0
    if noise is not None:
        _check_f32(noise, "noise")
        noise = _aligned(noise)
        if noise.numel() == B * H * W and B > 1:
            nbs = H * W
        elif noise.numel() != H * W:
            raise RuntimeError("styled_act_bwd: noise must be [B or 1, 1, H, W]")
    g_pre = torch.empty_like(out)
    sums = torch.empty((4, B, C), dtype=torch.float32, device=out.device)
    L = _lib.lib()
    with _on_device(out.device):
        nbytes = L.msg_styled_act_bwd_workspace(B, H * W, C)
        ws, wsp = _workspace(nbytes, out.device)
        rc = L.msg_styled_act_bwd(_ptr(g_pre), _ptr(sums), _ptr(g_out), _ptr(g_out2), _ptr(out), _ptr(col_scale),
                                  _bc_stride(col_scale, B, C, "col_scale"), _ptr(out2_scale) if g_out2 is not None else None,
                                  _bc_stride(out2_scale, B, C, "out2_scale"), _ptr(noise), nbs, B, H * W, C,
                                  float(slope), float(gain), wsp, nbytes, _stream(out))
    _lib.check(rc, "styled_act_bwd")
    return g_pre, sums


def blur_noise_bias_act_mod(x: torch.Tensor, kernel: torch.Tensor, pad: Sequence[int], col_scale: Optional[torch.Tensor],
                            noise: Optional[torch.Tensor], noise_w: Optional[torch.Tensor], bias: Optional[torch.Tensor],
                            slope: float, gain: float, out2_scale: Optional[torch.Tensor]):
    """(out, out2 or None): out = lrelu(col_scale[b,c] * fir(x) + noise_w * noise + bias[c]) * gain,
    out2 = out * out2_scale[b,c] — msg_upfirdn2d_bias_act_mod."""
    _check_f32(x, "x")
    _require_cuda(kernel, "kernel")
    x = _cl(x)
    k = kernel.contiguous().to(torch.float32)
    B, C, H, W = x.shape
    kh, kw = k.shape
    px0, px1, py0, py1 = [int(v) for v in pad]
    L = _lib.lib()
    out_h = L.msg_upfirdn2d_out_size(H, 1, 1, py0, py1, kh)
    out_w = L.msg_upfirdn2d_out_size(W, 1, 1, px0, px1, kw)
    if out_h < 0 or out_w < 0:
        raise RuntimeError("blur_noise_bias_act_mod: negative output size")
    out = torch.empty((B, C, out_h, out_w), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    out2 = torch.empty_like(out) if out2_scale is not None else None
    nbs = 0
    if noise is not None:
        _check_f32(noise, "noise")
        noise, noise_w = _aligned(noise), _aligned(noise_w)
        if noise.numel() == B * out_h * out_w and B > 1:
            nbs = out_h * out_w
        elif noise.numel() != out_h * out_w:
            raise RuntimeError("blur_noise_bias_act_mod: noise must be [B or 1, 1, OH, OW]")
    if bias is not None:
        bias = _aligned(bias)
    if col_scale is not None:
        col_scale = _aligned(col_scale)
    if out2_scale is not None:
        out2_scale = _aligned(out2_scale)
    with _on_device(x.device):
        rc = L.msg_upfirdn2d_bias_act_mod(_ptr(out), _ptr(out2), _ptr(x), _ptr(k), B, H, W, C, kh, kw, px0, px1, py0, py1,
                                          _ptr(col_scale), _bc_stride(col_scale, B, C, "col_scale"), _ptr(noise),
                                          _ptr(noise_w) if noise is not None else None, nbs, _ptr(bias), 1, float(slope),
                                          float(gain), _ptr(out2_scale), _bc_stride(out2_scale, B, C, "out2_scale"),
                                          _stream(x))
    _lib.check(rc, "upfirdn2d_bias_act_mod")
    return out, out2


def affine_warp(x: torch.Tensor, theta: torch.Tensor, mode: int = 0) -> torch.Tensor:
    """Bilinear warp in pixel coordinates; theta [B,2,3] maps output (x,y,1) to input (x,y)."""
    _check_f32(x, "x")
    _check_f32(theta, "theta")
    x = _aligned(x)
    theta = _aligned(theta)
    B, C, H, W = x.shape
    if theta.shape != (B, 2, 3):
        raise RuntimeError("affine_warp: theta must be [B, 2, 3]")
    out = torch.empty_like(x)
    with _on_device(x.device):
        rc = _lib.lib().msg_affine_warp(_ptr(out), _ptr(x), _ptr(theta), B, C, H, W, int(mode), _stream(x))
    _lib.check(rc, "affine_warp")
    return out


def affine_warp_bwd(grad_output: torch.Tensor, theta: torch.Tensor, mode: int = 0) -> torch.Tensor:
    """Gradient of affine_warp w.r.t. its input image."""
    _check_f32(grad_output, "grad_output")
    _check_f32(theta, "theta")
    g = _aligned(grad_output)
    theta = _aligned(theta)
    B, C, H, W = g.shape
    if theta.shape != (B, 2, 3):
        raise RuntimeError("affine_warp_bwd: theta must be [B, 2, 3]")
    dx = torch.empty_like(g)
    with _on_device(g.device):
        rc = _lib.lib().msg_affine_warp_bwd(_ptr(dx), _ptr(g), _ptr(theta), B, C, H, W, int(mode), _stream(g))
    _lib.check(rc, "affine_warp_bwd")
    return dx
