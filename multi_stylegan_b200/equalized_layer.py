"""Equalised-learning-rate layers with the reference's constructor arguments, parameter names and
scaling (multi_stylegan/equalized_layer.py:9-74, 210-277): weights ~ N(0,1) scaled at run time by
sqrt(2)/sqrt(fan_in), biases by sqrt(2)/sqrt(C_out).  The convolution runs on the sm_100a kernels
(conv.py); the M = batch linears stay on torch (cuBLAS) exactly as in the reference."""
import math
from typing import Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import conv


def _pair(v) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else tuple(v)


class EqualizedConv2d(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[int, Tuple[int, int]] = 3,
                 stride: Union[int, Tuple[int, int]] = 1, padding: Union[int, Tuple[int, int]] = 1,
                 bias: bool = True) -> None:
        super().__init__()
        self.kernel_size, self.stride, self.padding = _pair(kernel_size), _pair(stride), _pair(padding)
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.scale = math.sqrt(2) / math.sqrt(in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.scale_bias = math.sqrt(2) / math.sqrt(out_channels)

    def extra_repr(self) -> str:
        o, i, kh, kw = self.weight.shape
        return "{}, {}, kernel_size=({}, {}), stride={}, padding={}, bias={}".format(
            i, o, kh, kw, self.stride, self.padding, self.bias is not None)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        output = conv.conv2d(input, self.weight, self.stride, self.padding, alpha=self.scale)   # scale folded into the kernel
        if self.bias is not None:
            output = output + (self.bias * self.scale_bias).view(1, -1, 1, 1)
        return output


class EqualizedLinear(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.scale = math.sqrt(2) / math.sqrt(in_channels)
        self.scale_bias = math.sqrt(2) / math.sqrt(out_channels)

    def extra_repr(self) -> str:
        return "{}, {}, bias={}".format(self.weight.shape[1], self.weight.shape[0], self.bias is not None)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if input.dim() == 2:
            # x (W*scale)^T + b*scale_bias (reference :250-254) as ONE library GEMM: the two constants are its alpha / beta
            if self.bias is not None:
                return torch.addmm(self.bias, input, self.weight.t(), beta=self.scale_bias, alpha=self.scale)
            return torch.mm(input, self.weight.t()) * self.scale
        bias = None if self.bias is None else self.bias * self.scale_bias
        return F.linear(input, self.weight * self.scale, bias)


class PixelwiseNormalization(nn.Module):
    def __init__(self, alpha: float = 1e-8) -> None:
        super().__init__()
        self.alpha = alpha

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return input / torch.sqrt(torch.mean(input ** 2, dim=1, keepdim=True) + self.alpha)
