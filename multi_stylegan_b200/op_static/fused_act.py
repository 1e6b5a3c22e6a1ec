"""FusedLeakyReLU / fused_leaky_relu with the reference's API and double-differentiability
(multi_stylegan/op_static/fused_act.py:22-89), computed by the sm_100a kernels in csrc/fused_bias_act.cu.

Differences in mechanism, not in results: the first-order backward produces grad_input and grad_bias in
one pass (the reference runs the kernel and then a separate ``sum``)."""
import torch
from torch import nn
from torch.autograd import Function

from .. import _C


class FusedLeakyReLUFunctionBackward(Function):
    """grad_input = grad_output * (out > 0 ? 1 : slope) * scale ; grad_bias = sum over all dims but 1.
    Linear in grad_output, so its own backward is the same masked scaling (fused_act.py:44-51)."""

    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        grad_input, grad_bias = _C.fused_bias_act_bwd(grad_output, out, negative_slope, scale, out.shape[1])
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        if gradgrad_input is None:
            gradgrad_input = torch.zeros_like(out)
        if gradgrad_bias is None:
            gradgrad_bias = gradgrad_input.new_empty(0)
        gradgrad_out = _C.fused_bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        empty = input.new_empty(0)
        out = _C.fused_bias_act(input, bias, empty, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return grad_input, grad_bias, None, None


class FusedLeakyReLU(nn.Module):
    """Same constructor, parameter name (``bias``) and defaults as fused_act.py:76-85 (note scale=1.)."""

    def __init__(self, channel, negative_slope=0.2, scale=1.):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)


# ---------------------------------------------------------------------------------------------------
# Noise injection + bias + leaky ReLU as ONE pass (the StyledConv2d epilogue,
# multi_stylegan/multi_stylegan_generator.py:289-292 followed by fused_act.py:58), channels-last kernels.
# Differentiable to any order like the reference's op: the backward is linear in grad_output, so its own
# backward is the masked form of the forward.
# ---------------------------------------------------------------------------------------------------
class NoiseBiasActBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, noise, negative_slope, scale, want_param_grads=True):
        ctx.save_for_backward(out, noise)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input, grad_bias, grad_noise_w = _C.noise_bias_act_cl_bwd(grad_output, out, noise, negative_slope, scale,
                                                                       want_param_grads)
        if not want_param_grads:     # (autograd needs tensors for the declared outputs; these are never used)
            grad_bias = grad_input.new_zeros(0)
        return grad_input, grad_bias, grad_noise_w

    @staticmethod
    def backward(ctx, gg_input, gg_bias, gg_noise_w):
        out, noise = ctx.saved_tensors
        if gg_input is None:
            gg_input = torch.zeros_like(out)
        gg_out = _C.noise_bias_act_cl(gg_input, out, noise if gg_noise_w is not None else None, gg_noise_w, gg_bias,
                                      ctx.negative_slope, ctx.scale)
        return gg_out, None, None, None, None, None


class NoiseBiasAct(Function):
    @staticmethod
    def forward(ctx, input, noise, noise_weight, bias, negative_slope, scale):
        out = _C.noise_bias_act_cl(input, None, noise, noise_weight, bias, negative_slope, scale)
        ctx.save_for_backward(out, noise)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, noise = ctx.saved_tensors
        grad_input, grad_bias, grad_noise_w = NoiseBiasActBackward.apply(grad_output, out, noise, ctx.negative_slope,
                                                                         ctx.scale)
        return grad_input, None, grad_noise_w, grad_bias, None, None


def noise_bias_leaky_relu(input, noise, noise_weight, bias, negative_slope=0.2, scale=1.0):
    """lrelu(input + noise_weight * noise + bias[c]) * scale; noise [B or 1, 1, H, W], noise_weight [1]."""
    return NoiseBiasAct.apply(input, noise, noise_weight, bias, negative_slope, scale)
