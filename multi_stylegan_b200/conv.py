"""Differentiable conv primitives on top of the C-ABI conv kernels (csrc/conv_*.cu).

A convolution is bilinear in (input, weight), so the three kernels {forward, dgrad, wgrad} are closed
under differentiation: each autograd Function's backward is written with the other two, which gives
gradients of any order — R1 (loss.py:311-316) differentiates the discriminator's input gradient and
path-length regularisation (multi_stylegan_generator.py:193-200) the generator's latent gradient.

Replaces F.conv2d / F.conv_transpose2d at multi_stylegan_generator.py:398,409 and equalized_layer.py:70-73.
Weights are [O, C, kh, kw] (shared) or [B, O, C, kh, kw] (one filter bank per sample = the reference's
``groups=batch`` reshaping)."""
from typing import Tuple

import torch
from torch.autograd import Function

from . import _C


def _pair(v) -> Tuple[int, int]:
    return (int(v[0]), int(v[1])) if isinstance(v, (tuple, list)) else (int(v), int(v))


def _khw(w: torch.Tensor):
    return tuple(w.shape[-2:])


def _cat2_backward(g, x, x2, w, stride, padding, alpha, need_x, need_x2, need_w):
    """Gradients of conv([x | x2], w) w.r.t. x, x2 and w from the gradient g of the (pre-epilogue) convolution output:
    two dgrads against the two channel slices of the filter and two wgrads, all differentiable."""
    c1 = x.shape[1]
    dx = dx2 = dw = None
    if need_x:
        dx = _ConvDgrad.apply(g, w[:, :c1], tuple(x.shape[2:]), stride, padding, alpha)
    if need_x2:
        dx2 = _ConvDgrad.apply(g, w[:, c1:], tuple(x2.shape[2:]), stride, padding, alpha)
    if need_w:
        dw = torch.cat([_ConvWgrad.apply(g, x, _khw(w), stride, padding, False, alpha),
                        _ConvWgrad.apply(g, x2, _khw(w), stride, padding, False, alpha)], dim=1)
    return dx, dx2, dw


class _ConvForward2(Function):
    """alpha * conv([x | x2], w) with the concatenation read in place by the kernel's K loop (shared 4-D filters)."""

    @staticmethod
    def forward(ctx, x, x2, w, stride, padding, alpha):
        ctx.save_for_backward(x, x2, w)
        ctx.stride, ctx.padding, ctx.alpha = stride, padding, alpha
        return _C.conv2d_forward(x, w, stride, padding, alpha=alpha, x2=x2)

    @staticmethod
    def backward(ctx, dy):
        x, x2, w = ctx.saved_tensors
        if _C.conv_channels_last and dy.is_cuda:
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx, dx2, dw = _cat2_backward(dy, x, x2, w, ctx.stride, ctx.padding, ctx.alpha, *ctx.needs_input_grad[:3])
        return dx, dx2, dw, None, None, None


class _ConvForward(Function):
    """`wt`: the filter tensor stores its two channel dimensions transposed ([C, O, kh, kw], the layout of
    conv_transpose2d's weight); every kernel reads / writes that layout in place."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, alpha, wt=False):
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.padding, ctx.alpha, ctx.wt = stride, padding, alpha, wt
        return _C.conv2d_forward(x, w, stride, padding, alpha=alpha, w_transposed=wt)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0] and ctx.needs_input_grad[1] and _C.conv_channels_last and dy.is_cuda:
            # a channel slice of a concatenation's gradient is a strided view: densify it once for both kernels
            dy = dy.contiguous(memory_format=torch.channels_last)
        if ctx.needs_input_grad[0]:
            dx = _ConvDgrad.apply(dy, w, tuple(x.shape[2:]), ctx.stride, ctx.padding, ctx.alpha, ctx.wt)
        if ctx.needs_input_grad[1]:
            dw = _ConvWgrad.apply(dy, x, _khw(w), ctx.stride, ctx.padding, w.dim() == 5, ctx.alpha, ctx.wt)
        return dx, dw, None, None, None, None


class _ConvDgrad(Function):
    """dx = alpha * conv^T(dy, w); as a function of (dy, w) it is the transposed convolution."""

    @staticmethod
    def forward(ctx, dy, w, in_hw, stride, padding, alpha, wt=False):
        ctx.save_for_backward(dy, w)
        ctx.in_hw, ctx.stride, ctx.padding, ctx.alpha, ctx.wt = in_hw, stride, padding, alpha, wt
        return _C.conv2d_dgrad(dy, w, in_hw, stride, padding, alpha=alpha, w_transposed=wt)

    @staticmethod
    def backward(ctx, ddx):
        dy, w = ctx.saved_tensors
        g_dy = g_w = None
        if ctx.needs_input_grad[0]:
            g_dy = _ConvForward.apply(ddx, w, ctx.stride, ctx.padding, ctx.alpha, ctx.wt)
        if ctx.needs_input_grad[1]:
            g_w = _ConvWgrad.apply(dy, ddx, _khw(w), ctx.stride, ctx.padding, w.dim() == 5, ctx.alpha, ctx.wt)
        return g_dy, g_w, None, None, None, None, None


class _ConvWgrad(Function):
    @staticmethod
    def forward(ctx, dy, x, khw, stride, padding, per_sample, alpha, wt=False):
        ctx.save_for_backward(dy, x)
        ctx.stride, ctx.padding, ctx.alpha, ctx.wt = stride, padding, alpha, wt
        return _C.conv2d_wgrad(dy, x, khw, stride, padding, per_sample, alpha=alpha, w_transposed=wt)

    @staticmethod
    def backward(ctx, ddw):
        dy, x = ctx.saved_tensors
        g_dy = g_x = None
        if ctx.needs_input_grad[0]:
            g_dy = _ConvForward.apply(x, ddw, ctx.stride, ctx.padding, ctx.alpha, ctx.wt)
        if ctx.needs_input_grad[1]:
            g_x = _ConvDgrad.apply(dy, ddw, tuple(x.shape[2:]), ctx.stride, ctx.padding, ctx.alpha, ctx.wt)
        return g_dy, g_x, None, None, None, None, None, None


class _ConvBiasAct(Function):
    """out = lrelu(alpha * conv(x, w) + noise_w * noise + bias) * gain with the whole tail in the conv kernel's
    epilogue.  The backward is composed of differentiable pieces (the masked activation backward of
    op_static/fused_act.py, then dgrad / wgrad), so gradients of any order exist."""

    @staticmethod
    def forward(ctx, x, w, noise, noise_w, bias, stride, padding, slope, gain, alpha, x2=None):
        out = _C.conv2d_forward(x, w, stride, padding, alpha=alpha, bias=bias, noise=noise, noise_w=noise_w, act=True,
                                slope=slope, gain=gain, x2=x2)
        ctx.save_for_backward(x, w, out, noise if noise is not None else x.new_empty(0),
                              x2 if x2 is not None else x.new_empty(0))
        ctx.has_noise = noise is not None
        ctx.has_bias = bias is not None
        ctx.has_x2 = x2 is not None
        ctx.stride, ctx.padding, ctx.slope, ctx.gain, ctx.alpha = stride, padding, slope, gain, alpha
        return out

    @staticmethod
    def backward(ctx, gout):
        from .op_static.fused_act import NoiseBiasActBackward
        x, w, out, noise, x2 = ctx.saved_tensors
        noise = noise if ctx.has_noise else None
        # parameter gradients nobody asked for (the discriminator inside the generator step) skip their reduction
        want = (ctx.has_bias and ctx.needs_input_grad[4]) or (ctx.has_noise and ctx.needs_input_grad[3])
        g_pre, g_bias, g_noise_w = NoiseBiasActBackward.apply(gout, out, noise, ctx.slope, ctx.gain, want)
        if not want:
            g_bias = g_noise_w = None
        dx = dw = dx2 = None
        if ctx.has_x2:
            dx, dx2, dw = _cat2_backward(g_pre, x, x2, w, ctx.stride, ctx.padding, ctx.alpha, ctx.needs_input_grad[0],
                                         ctx.needs_input_grad[10], ctx.needs_input_grad[1])
        else:
            if ctx.needs_input_grad[0]:
                dx = _ConvDgrad.apply(g_pre, w, tuple(x.shape[2:]), ctx.stride, ctx.padding, ctx.alpha)
            if ctx.needs_input_grad[1]:
                dw = _ConvWgrad.apply(g_pre, x, tuple(w.shape[-2:]), ctx.stride, ctx.padding, w.dim() == 5, ctx.alpha)
        return (dx, dw, None, g_noise_w if (ctx.has_noise and want) else None,
                g_bias if (ctx.has_bias and want) else None, None, None, None, None, None, dx2)


def conv2d_bias_act(x: torch.Tensor, w: torch.Tensor, bias=None, noise=None, noise_w=None, stride=1, padding=0,
                    negative_slope: float = 0.2, gain: float = 1.0, alpha: float = 1.0, x2=None) -> torch.Tensor:
    """Fused alpha * conv -> (+ noise_w * noise) -> (+ bias[c]) -> leaky ReLU -> * gain (channel count must be a
    multiple of 4 for the channels-last activation-backward kernel).  `x2`: convolve the channel concatenation
    [x | x2] in place (shared filters; _C.cat2_supported)."""
    return _ConvBiasAct.apply(x, w, noise, noise_w, bias, _pair(stride), _pair(padding), negative_slope, gain,
                              float(alpha), x2)


class _ConvAddScale(Function):
    """out = (alpha * conv(x, w) + other) * gain — the residual join of ResNetBlock / NonLocalBlock
    (u_net_2d_discriminator.py:186,381) inside the conv epilogue."""

    @staticmethod
    def forward(ctx, x, w, other, stride, padding, gain, alpha, x2=None):
        ctx.save_for_backward(x, w, x2 if x2 is not None else x.new_empty(0))
        ctx.has_x2 = x2 is not None
        ctx.stride, ctx.padding, ctx.gain, ctx.alpha = stride, padding, gain, alpha
        return _C.conv2d_forward(x, w, stride, padding, alpha=alpha, add=other, gain=gain, x2=x2)

    @staticmethod
    def backward(ctx, gout):
        x, w, x2 = ctx.saved_tensors
        g = gout if ctx.gain == 1.0 else gout * ctx.gain
        dx = dw = dx2 = None
        if ctx.has_x2:
            dx, dx2, dw = _cat2_backward(g, x, x2, w, ctx.stride, ctx.padding, ctx.alpha, ctx.needs_input_grad[0],
                                         ctx.needs_input_grad[7], ctx.needs_input_grad[1])
        else:
            if ctx.needs_input_grad[0]:
                dx = _ConvDgrad.apply(g, w, tuple(x.shape[2:]), ctx.stride, ctx.padding, ctx.alpha)
            if ctx.needs_input_grad[1]:
                dw = _ConvWgrad.apply(g, x, tuple(w.shape[-2:]), ctx.stride, ctx.padding, w.dim() == 5, ctx.alpha)
        return dx, dw, (g if ctx.needs_input_grad[2] else None), None, None, None, None, dx2


def conv2d_add_scale(x: torch.Tensor, w: torch.Tensor, other: torch.Tensor, stride=1, padding=0,
                     gain: float = 1.0, alpha: float = 1.0, x2=None) -> torch.Tensor:
    return _ConvAddScale.apply(x, w, other, _pair(stride), _pair(padding), gain, float(alpha), x2)


class ResBlockFused(Function):
    """The discriminator's residual block (u_net_2d_discriminator.py:174-186)
        out = (lrelu(conv3x3(lrelu(conv3x3(x) + b1)) + b2) + conv1x1(x)) / sqrt(2)
    as three kernels forward (both activations and the join live in conv epilogues; the 1/sqrt(2) is folded into the
    second activation's gain and the residual convolution's alpha) with a hand-written FIRST-ORDER backward: the block
    input feeds the main and the residual path, and the sum of its two gradients is formed in the epilogue of the main
    path's dgrad (`add` operand) instead of a separate pass; the scaling by 1/sqrt(2) never runs as a pass either.
    `x2`: the block is applied to the channel concatenation [x | x2] read in place (U-Net decoder, reference :137).
    R1 (which differentiates this backward) runs the composite Functions above instead (_mode.higher_order_gradients)."""

    @staticmethod
    def forward(ctx, x, x2, w1, b1, w2, b2, wr, a1, a2, ar, slope, g1, g2):
        h1 = _C.conv2d_forward(x, w1, 1, 1, alpha=a1, bias=b1, act=True, slope=slope, gain=g1, x2=x2)
        h2 = _C.conv2d_forward(h1, w2, 1, 1, alpha=a2, bias=b2, act=True, slope=slope, gain=g2)
        out = _C.conv2d_forward(x, wr, 1, 0, alpha=ar, add=h2, x2=x2)
        ctx.save_for_backward(x, x2, w1, w2, wr, h1, h2)
        ctx.cfg = (a1, a2, ar, slope, g1, g2)
        return out

    @staticmethod
    def backward(ctx, gout):
        from ._mode import NO_DOUBLE_BACKWARD
        if torch.is_grad_enabled():
            raise RuntimeError(NO_DOUBLE_BACKWARD)
        x, x2, w1, w2, wr, h1, h2 = ctx.saved_tensors
        a1, a2, ar, slope, g1, g2 = ctx.cfg
        need = ctx.needs_input_grad
        need_x, need_x2, need_w1, need_b1, need_w2, need_b2, need_wr = need[:7]
        hw = tuple(x.shape[2:])
        g2_pre, db2, _ = _C.noise_bias_act_cl_bwd(gout, h2, None, slope, g2, need_b2)
        # first activation's backward inside the dgrad that produces its incoming gradient (one kernel, no pass over gh1)
        fused = _C.conv2d_dgrad_act_bwd(g2_pre, w2, h1, 1, alpha=a2, slope=slope, gain=g1, want_dbias=need_b1)
        if fused is not None:
            g1_pre, db1 = fused
        else:
            gh1 = _C.conv2d_dgrad(g2_pre, w2, hw, 1, 1, alpha=a2)
            g1_pre, db1, _ = _C.noise_bias_act_cl_bwd(gh1, h1, None, slope, g1, need_b1)
        dw2 = _C.conv2d_wgrad(g2_pre, h1, (3, 3), 1, 1, False, alpha=a2) if need_w2 else None
        dx = dx2 = dw1 = dwr = None
        if x2 is None:
            if need_x:
                dx = _C.conv2d_dgrad(g1_pre, w1, hw, 1, 1, alpha=a1, add=_C.conv2d_dgrad(gout, wr, hw, 1, 0, alpha=ar))
            if need_w1:
                dw1 = _C.conv2d_wgrad(g1_pre, x, (3, 3), 1, 1, False, alpha=a1)
            if need_wr:
                dwr = _C.conv2d_wgrad(gout, x, (1, 1), 1, 0, False, alpha=ar)
        else:
            c1 = x.shape[1]
            if need_x:
                dx = _C.conv2d_dgrad(g1_pre, w1[:, :c1], hw, 1, 1, alpha=a1,
                                     add=_C.conv2d_dgrad(gout, wr[:, :c1], hw, 1, 0, alpha=ar))
            if need_x2:
                dx2 = _C.conv2d_dgrad(g1_pre, w1[:, c1:], hw, 1, 1, alpha=a1,
                                      add=_C.conv2d_dgrad(gout, wr[:, c1:], hw, 1, 0, alpha=ar))
            if need_w1:
                dw1 = torch.cat([_C.conv2d_wgrad(g1_pre, x, (3, 3), 1, 1, False, alpha=a1),
                                 _C.conv2d_wgrad(g1_pre, x2, (3, 3), 1, 1, False, alpha=a1)], dim=1)
            if need_wr:
                dwr = torch.cat([_C.conv2d_wgrad(gout, x, (1, 1), 1, 0, False, alpha=ar),
                                 _C.conv2d_wgrad(gout, x2, (1, 1), 1, 0, False, alpha=ar)], dim=1)
        return (dx, dx2, dw1, db1 if need_b1 else None, dw2, db2 if need_b2 else None, dwr,
                None, None, None, None, None, None)


class DownscaleTapFused(Function):
    """(conv(x) + bias, x): the strided convolution that leaves an encoder stage of the U-Net together with the skip
    connection that taps the same tensor (u_net_2d_discriminator.py:76-83,100-104).  Forward: the bias lives in the conv
    epilogue.  Backward (first order): the skip connection's gradient is the `add` operand of the convolution's dgrad, so
    the sum of the two gradients of x never runs as a separate pass."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding, alpha, beta):
        bs = None if b is None else b * beta
        y = _C.conv2d_forward(x, w, stride, padding, alpha=alpha, bias=bs)
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, alpha, beta, b is not None)
        ctx.set_materialize_grads(False)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, gy, gskip):
        from ._mode import NO_DOUBLE_BACKWARD
        if torch.is_grad_enabled():
            raise RuntimeError(NO_DOUBLE_BACKWARD)
        x, w = ctx.saved_tensors
        stride, padding, alpha, beta, has_b = ctx.cfg
        need = ctx.needs_input_grad
        if gy is None:
            return gskip, None, None, None, None, None, None
        if _C.conv_channels_last and gy.is_cuda:
            gy = gy.contiguous(memory_format=torch.channels_last)
        dx = _C.conv2d_dgrad(gy, w, tuple(x.shape[2:]), stride, padding, alpha=alpha, add=gskip) if need[0] else None
        dw = _C.conv2d_wgrad(gy, x, tuple(w.shape[-2:]), stride, padding, False, alpha=alpha) if need[1] else None
        db = _C.colsum_cl(gy, beta) if (has_b and need[2]) else None
        return dx, dw, db, None, None, None, None


def downscale_with_tap(x, w, b, stride, padding, alpha: float, beta: float):
    return DownscaleTapFused.apply(x, w, b, _pair(stride), _pair(padding), float(alpha), float(beta))


def res_block(x, x2, w1, b1, w2, b2, wr, a1: float, a2: float, ar: float, slope: float, g1: float, g2: float):
    return ResBlockFused.apply(x, x2, w1, b1, w2, b2, wr, float(a1), float(a2), float(ar), float(slope), float(g1), float(g2))


def conv2d(x: torch.Tensor, w: torch.Tensor, stride=1, padding=0, alpha: float = 1.0, x2=None) -> torch.Tensor:
    """alpha * conv(x, w); x [B,C,H,W]; w [O,C,kh,kw] or [B,O,C,kh,kw].  `x2`: convolve [x | x2] (see conv2d_bias_act)."""
    if x2 is not None:
        return _ConvForward2.apply(x, x2, w, _pair(stride), _pair(padding), float(alpha))
    return _ConvForward.apply(x, w, _pair(stride), _pair(padding), float(alpha))


def conv_transpose2d(x: torch.Tensor, w: torch.Tensor, stride=1, padding=0, weight_is_conv_layout: bool = False) -> torch.Tensor:
    """x [B,Cin,H,W]; w [Cin,Cout,kh,kw] or [B,Cin,Cout,kh,kw] (torch's conv_transpose2d weight layout, which is the
    layout of the conv whose input-gradient this is).  With weight_is_conv_layout the filters come as
    [(B,) Cout, Cin, kh, kw] (the generator's modulated banks, :384-398) and are read in place instead of being
    transposed and copied.  output_padding = 0."""
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    kh, kw = w.shape[-2:]
    out_hw = ((x.shape[2] - 1) * sh - 2 * ph + kh, (x.shape[3] - 1) * sw - 2 * pw + kw)
    return _ConvDgrad.apply(x, w, out_hw, (sh, sw), (ph, pw), 1.0, bool(weight_is_conv_layout))
