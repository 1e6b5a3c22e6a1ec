"""Shared-weight form of the generator's modulated convolutions (first-order fast path).

The reference builds one filter bank per sample, ``scale * W * s[b] * demod[b]`` (multi_stylegan_generator.py:379-388),
and convolves with ``groups = batch`` (:390-411).  The same numbers come out of

    y[b, o] = demod[b, o] * conv(scale * W, s[b, c] * x[b, c])                      

where every sample shares ONE GEMM weight operand: the style is applied to the activations — by the *previous* layer's
epilogue, which writes ``out * s_next`` as a second output next to ``out`` — and the demodulation factor, the noise,
the bias and the leaky ReLU live in this layer's epilogue.  What that buys: no [B, O, C, kh, kw] tensor is ever built,
transformed or differentiated, and the weight gradient is one batch-reduced GEMM instead of B per-sample ones.

Functions here have hand-written first-order backwards (one sweep `_C.styled_act_bwd` + shared dgrad / wgrad).  They are
NOT differentiable twice: the path-length regulariser (multi_stylegan_generator.py:193-200), which differentiates the
generator's backward, runs the per-sample-weight formulation in multi_stylegan_generator.py, whose every piece is
differentiable to any order (_mode.higher_order_gradients)."""
from typing import Optional, Tuple

import torch
from torch.autograd import Function

from . import _C

from ._mode import NO_DOUBLE_BACKWARD as _NO_DOUBLE


def _check_first_order() -> None:
    if torch.is_grad_enabled():
        raise RuntimeError(_NO_DOUBLE)


class DemodFactors(Function):
    """d[b, o] = rsqrt(scale^2 * sum_c s[b, c]^2 * sum_t W[o, c, t]^2 + 1e-8) — reference :386-388."""

    @staticmethod
    def forward(ctx, W, s, scale):
        d, wsq = _C.demod_factors(W, s, scale)
        ctx.save_for_backward(W, s, d, wsq)
        ctx.scale = scale
        return d

    @staticmethod
    def backward(ctx, gd):
        _check_first_order()
        W, s, d, wsq = ctx.saved_tensors
        # q = gd * d^3 * (-scale^2 / 2);  ds = 2 s (q wsq);  dW = W * 2 (q^T s^2)  — two launches (msg_demod_factors_bwd)
        dW, ds = _C.demod_factors_bwd(gd, d, s, wsq, W, ctx.scale, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dW, ds, None


def _param_grads(ctx, sums, d, nw, bias):
    """(dd, dnw, dbias, ds_next) from the four per-(sample, channel) sums of _C.styled_act_bwd."""
    S1, S2, S3, S4 = sums[0], sums[1], sums[2], sums[3]
    dd = dnw = dbias = None
    if d is not None:
        t = S2
        if nw is not None:
            t = t - nw * S3
        if bias is not None:
            t = t - bias.view(1, -1) * S1
        dd = t / d
    if nw is not None:
        dnw = S3.sum().reshape(1)
    if bias is not None:
        dbias = S1.sum(0)
    return dd, dnw, dbias, S4


class StyledConvFused(Function):
    """(out, out * s_next) = epilogue(d * conv(scale * W, xs)) for a stride-1 modulated convolution whose input already
    carries the style (xs = s * x).  W [O, C, kh, kw] shared, d [B, O], noise [B or 1, 1, H, W], nw [1], bias [O],
    s_next [B, O] or None."""

    @staticmethod
    def forward(ctx, xs, W, d, noise, nw, bias, s_next, stride, padding, slope, gain, scale):
        r = _C.conv2d_forward(xs, W, stride, padding, alpha=scale, bias=bias, noise=noise, noise_w=nw, act=True,
                              slope=slope, gain=gain, col_scale=d, out2_scale=s_next)
        out, out2 = r if s_next is not None else (r, None)
        ctx.save_for_backward(xs, W, d, out, noise, nw, bias, s_next)
        ctx.cfg = (stride, padding, slope, gain, scale)
        ctx.set_materialize_grads(False)     # an unused output must reach backward as None, not as a tensor of zeros
        if out2 is None:
            return out, None
        return out, out2

    @staticmethod
    def backward(ctx, g_out, g_out2):
        _check_first_order()
        xs, W, d, out, noise, nw, bias, s_next = ctx.saved_tensors
        stride, padding, slope, gain, scale = ctx.cfg
        if g_out is None and g_out2 is None:
            return (None,) * 12
        g_pre, sums = _C.styled_act_bwd(g_out, g_out2 if s_next is not None else None, out, d, s_next,
                                        noise if nw is not None else None, slope, gain)
        dd, dnw, dbias, ds_next = _param_grads(ctx, sums, d, nw if noise is not None else None, bias)
        dxs = dW = None
        if ctx.needs_input_grad[0]:
            dxs = _C.conv2d_dgrad(g_pre, W, tuple(xs.shape[2:]), stride, padding, alpha=scale)
        if ctx.needs_input_grad[1]:
            dW = _C.conv2d_wgrad(g_pre, xs, tuple(W.shape[-2:]), stride, padding, False, alpha=scale)
        return (dxs, dW, dd if ctx.needs_input_grad[2] else None, None,
                dnw if (nw is not None and noise is not None and ctx.needs_input_grad[4]) else None,
                dbias if ctx.needs_input_grad[5] else None,
                ds_next if (s_next is not None and g_out2 is not None and ctx.needs_input_grad[6]) else None,
                None, None, None, None, None)


class StyledUpConvFused(Function):
    """The upsampling layer in the same form: 2x2 / stride-2 transposed convolution with shared weights (reference
    :393-401), then ONE FIR pass that applies the x4 blur (:403), the demodulation factor (it commutes with the
    per-channel FIR), noise, bias, leaky ReLU and writes (out, out * s_next)."""

    @staticmethod
    def forward(ctx, xs, W, d, kernel, pad, noise, nw, bias, s_next, stride, padding, slope, gain, scale):
        kh, kw = W.shape[-2:]
        out_hw = ((xs.shape[2] - 1) * stride[0] - 2 * padding[0] + kh, (xs.shape[3] - 1) * stride[1] - 2 * padding[1] + kw)
        y = _C.conv2d_dgrad(xs, W, out_hw, stride, padding, alpha=scale, w_transposed=True)
        out, out2 = _C.blur_noise_bias_act_mod(y, kernel, pad, d, noise, nw, bias, slope, gain, s_next)
        ctx.save_for_backward(xs, W, d, out, noise, nw, bias, s_next, kernel)
        ctx.cfg = (stride, padding, slope, gain, scale, pad, tuple(y.shape))
        ctx.set_materialize_grads(False)
        return out, out2

    @staticmethod
    def backward(ctx, g_out, g_out2):
        _check_first_order()
        xs, W, d, out, noise, nw, bias, s_next, kernel = ctx.saved_tensors
        stride, padding, slope, gain, scale, pad, y_shape = ctx.cfg
        if g_out is None and g_out2 is None:
            return (None,) * 14
        g_pre, sums = _C.styled_act_bwd(g_out, g_out2 if s_next is not None else None, out, d, s_next,
                                        noise if nw is not None else None, slope, gain)
        dd, dnw, dbias, ds_next = _param_grads(ctx, sums, d, nw if noise is not None else None, bias)
        dxs = dW = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            # adjoint of the blur (op_static/upfirdn2d.py:34-45,114-117): flipped taps, g_pad
            kh, kw = kernel.shape
            px0, px1, py0, py1 = pad
            in_h, in_w = y_shape[2], y_shape[3]
            g_pad = (kw - px0 - 1, in_w - out.shape[3] + px0, kh - py0 - 1, in_h - out.shape[2] + py0)
            g_y = _C.upfirdn2d(g_pre.permute(0, 2, 3, 1), torch.flip(kernel, [0, 1]), 1, 1, 1, 1, *g_pad).permute(0, 3, 1, 2)
            if ctx.needs_input_grad[0]:
                dxs = _C.conv2d_forward(g_y, W, stride, padding, alpha=scale, w_transposed=True)
            if ctx.needs_input_grad[1]:
                dW = _C.conv2d_wgrad(xs, g_y, tuple(W.shape[-2:]), stride, padding, False, alpha=scale, w_transposed=True)
        return (dxs, dW, dd if ctx.needs_input_grad[2] else None, None, None, None,
                dnw if (nw is not None and noise is not None and ctx.needs_input_grad[6]) else None,
                dbias if ctx.needs_input_grad[7] else None,
                ds_next if (s_next is not None and g_out2 is not None and ctx.needs_input_grad[8]) else None,
                None, None, None, None, None)


def styled_conv(xs: torch.Tensor, W: torch.Tensor, s: torch.Tensor, scale: float, demodulate: bool,
                noise: Optional[torch.Tensor], noise_w: Optional[torch.Tensor], bias: Optional[torch.Tensor],
                s_next: Optional[torch.Tensor], stride, padding, slope: float, gain: float
                ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    d = DemodFactors.apply(W, s, scale) if demodulate else None
    return StyledConvFused.apply(xs, W, d, noise, noise_w, bias, s_next, stride, padding, slope, gain, scale)


def styled_up_conv(xs: torch.Tensor, W: torch.Tensor, s: torch.Tensor, scale: float, demodulate: bool,
                   kernel: torch.Tensor, pad, noise: Optional[torch.Tensor], noise_w: Optional[torch.Tensor],
                   bias: Optional[torch.Tensor], s_next: Optional[torch.Tensor], stride, padding, slope: float,
                   gain: float) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    d = DemodFactors.apply(W, s, scale) if demodulate else None
    pad4 = (pad[0], pad[1], pad[0], pad[1])
    return StyledUpConvFused.apply(xs, W, d, kernel, pad4, noise, noise_w, bias, s_next, stride, padding, slope, gain, scale)
