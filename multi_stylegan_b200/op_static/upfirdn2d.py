"""upfirdn2d(input, kernel, up, down, pad) with the reference's API and double-differentiability
(multi_stylegan/op_static/upfirdn2d.py:22-153), computed by csrc/upfirdn2d.cu.

The adjoint of an (up, down, pad) FIR is the (down, up, g_pad) FIR with the flipped taps
(upfirdn2d.py:34-45, 114-117); the double backward re-applies the forward configuration (:71-82)."""
import torch
from torch.autograd import Function

from .. import _C


def _fir(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, channels_last):
    """[B, C, H, W] -> [B, C, H', W'] through the native [major, h, w, minor] op: planes (major = B*C, minor = 1,
    exactly the reference's view, upfirdn2d.py:102) or, for channels-last activations, pixels-of-channels
    (major = B, minor = C) without any layout copy."""
    B, C, H, W = x.shape
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
        out = _C.upfirdn2d(x.permute(0, 2, 3, 1), kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
        return out.permute(0, 3, 1, 2)
    out = _C.upfirdn2d(x.reshape(-1, H, W, 1), kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
    return out.view(B, C, out.shape[1], out.shape[2])


class UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size, channels_last):
        up_x, up_y = up
        down_x, down_y = down
        g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1 = g_pad
        grad_input = _fir(grad_output, grad_kernel, down_x, down_y, up_x, up_y,
                          g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1, channels_last)
        ctx.channels_last = channels_last
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.in_size, ctx.out_size = in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gradgrad_out = _fir(gradgrad_input, kernel, ctx.up[0], ctx.up[1], ctx.down[0], ctx.down[1], *ctx.pad,
                            ctx.channels_last)
        return gradgrad_out, None, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        batch, channel, in_h, in_w = input.shape
        ctx.in_size = input.shape
        ctx.channels_last = _C._cl_only(input) and channel % 4 == 0
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        out_h = (in_h * up_y + pad_y0 + pad_y1 - kernel_h) // down_y + 1
        out_w = (in_w * up_x + pad_x0 + pad_x1 - kernel_w) // down_x + 1
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        g_pad_x0 = kernel_w - pad_x0 - 1
        g_pad_y0 = kernel_h - pad_y0 - 1
        g_pad_x1 = in_w * up_x - out_w * down_x + pad_x0 - up_x + 1
        g_pad_y1 = in_h * up_y - out_h * down_y + pad_y0 - up_y + 1
        ctx.g_pad = (g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        return _fir(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, ctx.channels_last)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad,
                                             ctx.g_pad, ctx.in_size, ctx.out_size, ctx.channels_last)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))


class BlurNoiseBiasAct(Function):
    """lrelu(upfirdn2d(x, kernel, pad=pad) + noise_weight * noise + bias) * scale — the blur after an up-convolution and
    the StyledConv2d tail (multi_stylegan_generator.py:403, :289-292, fused_act.py:58) in one pass.  The backward is
    composed of the differentiable activation backward and the FIR adjoint, so any order of gradient exists."""

    @staticmethod
    def forward(ctx, input, kernel, pad, noise, noise_weight, bias, negative_slope, scale):
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        out = _C.blur_noise_bias_act(input, kernel, pad, noise, noise_weight, bias, negative_slope, scale)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]), out, noise if noise is not None else out.new_empty(0))
        ctx.has_noise, ctx.has_bias = noise is not None, bias is not None
        ctx.negative_slope, ctx.scale, ctx.pad = negative_slope, scale, pad
        ctx.in_size, ctx.out_size = input.shape, (out.shape[2], out.shape[3])
        in_h, in_w = input.shape[2], input.shape[3]
        ctx.g_pad = (kernel_w - pad_x0 - 1, in_w - out.shape[3] + pad_x0, kernel_h - pad_y0 - 1,
                     in_h - out.shape[2] + pad_y0)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        from .fused_act import NoiseBiasActBackward
        kernel, grad_kernel, out, noise = ctx.saved_tensors
        noise = noise if ctx.has_noise else None
        g_pre, g_bias, g_noise_w = NoiseBiasActBackward.apply(grad_output, out, noise, ctx.negative_slope, ctx.scale)
        grad_input = None
        if ctx.needs_input_grad[0]:
            grad_input = UpFirDn2dBackward.apply(g_pre, kernel, grad_kernel, (1, 1), (1, 1), ctx.pad, ctx.g_pad,
                                                 ctx.in_size, ctx.out_size, True)
        return (grad_input, None, None, None, g_noise_w if ctx.has_noise else None, g_bias if ctx.has_bias else None,
                None, None)


def blur_noise_bias_leaky_relu(input, kernel, pad, noise, noise_weight, bias, negative_slope=0.2, scale=1.0):
    return BlurNoiseBiasAct.apply(input, kernel, (pad[0], pad[1], pad[0], pad[1]), noise, noise_weight, bias,
                                  negative_slope, scale)
