"""Oracle for the two native ops and the conv primitives (CPU).  TEST INFRASTRUCTURE ONLY."""
from typing import Sequence

import torch
import torch.nn.functional as F


def fused_bias_act(x: torch.Tensor, b: torch.Tensor, ref: torch.Tensor, act: int, grad: int, alpha: float,
                   scale: float) -> torch.Tensor:
    """multi_stylegan/op_static/fused_bias_act_kernel.cu:25-48: bias indexes dim 1 (:29,:67-71),
    empty bias/ref mean absent (:62-63); act*10+grad selects the formula (:36-45); `default` -> y = x."""
    x = x.contiguous()
    v = x
    if b.numel():
        shape = [1, -1] + [1] * (x.dim() - 2)
        v = v + b.reshape(shape)
    code = act * 10 + grad
    if code == 30:
        y = torch.where(v > 0, v, v * alpha)
    elif code == 31:
        y = torch.where(ref.reshape(v.shape) > 0, v, v * alpha)
    elif code in (12, 32):
        y = torch.zeros_like(v)
    else:
        y = v
    return y * scale


def upfirdn2d_out_size(n: int, up: int, down: int, p0: int, p1: int, k: int) -> int:
    """multi_stylegan/op_static/upfirdn2d_kernel.cu:167-168."""
    return (n * up + p0 + p1 - k + down) // down


def upfirdn2d(x: torch.Tensor, kernel: torch.Tensor, up_x: int, up_y: int, down_x: int, down_y: int,
              pad_x0: int, pad_x1: int, pad_y0: int, pad_y1: int) -> torch.Tensor:
    """x [major, in_h, in_w, minor] -> [major, out_h, out_w, minor].
    Semantics of upfirdn2d_kernel.cu:77,85-88,114-129 (and upfirdn2d.py:156-190): zero-insert by `up`,
    pad (negative pad crops), correlate with the flipped taps (true convolution), keep every `down`-th."""
    major, in_h, in_w, minor = x.shape
    kh, kw = kernel.shape
    full_h, full_w = in_h * up_y + pad_y0 + pad_y1, in_w * up_x + pad_x0 + pad_x1
    canvas = x.new_zeros(major, minor, max(full_h, 0), max(full_w, 0))
    src = x.permute(0, 3, 1, 2)
    # position of input sample (iy, ix) on the padded, zero-inserted canvas
    ys = torch.arange(in_h, device=x.device) * up_y + pad_y0
    xs = torch.arange(in_w, device=x.device) * up_x + pad_x0
    my = (ys >= 0) & (ys < full_h)
    mx = (xs >= 0) & (xs < full_w)
    if my.any() and mx.any():
        canvas[:, :, ys[my][:, None], xs[mx][None, :]] = src[:, :, my][:, :, :, mx]
    out_h = upfirdn2d_out_size(in_h, up_y, down_y, pad_y0, pad_y1, kh)
    out_w = upfirdn2d_out_size(in_w, up_x, down_x, pad_x0, pad_x1, kw)
    if out_h <= 0 or out_w <= 0 or full_h < kh or full_w < kw:
        return x.new_zeros(major, max(out_h, 0), max(out_w, 0), minor)
    w = torch.flip(kernel, [0, 1]).to(device=x.device, dtype=x.dtype).view(1, 1, kh, kw)
    y = F.conv2d(canvas.reshape(major * minor, 1, full_h, full_w), w)
    y = y[:, :, ::down_y, ::down_x].reshape(major, minor, out_h, out_w)
    return y.permute(0, 2, 3, 1).contiguous()


def conv2d(x: torch.Tensor, w: torch.Tensor, stride=1, padding=0) -> torch.Tensor:
    """Shared weights [O,C,kh,kw] -> F.conv2d (equalized_layer.py:70-73); per-sample weights
    [B,O,C,kh,kw] -> the reference's groups=B reshaping (multi_stylegan_generator.py:390,406-411)."""
    if w.dim() == 4:
        return F.conv2d(x, w, stride=stride, padding=padding)
    B, O, C, kh, kw = w.shape
    y = F.conv2d(x.reshape(1, B * C, *x.shape[2:]), w.reshape(B * O, C, kh, kw), stride=stride, padding=padding,
                 groups=B)
    return y.view(B, O, y.shape[2], y.shape[3])


def conv_transpose2d(x: torch.Tensor, w: torch.Tensor, stride=1, padding=0) -> torch.Tensor:
    """w [Cin,Cout,kh,kw] or per-sample [B,Cin,Cout,kh,kw] (multi_stylegan_generator.py:393-401)."""
    if w.dim() == 4:
        return F.conv_transpose2d(x, w, stride=stride, padding=padding)
    B, Ci, Co, kh, kw = w.shape
    y = F.conv_transpose2d(x.reshape(1, B * Ci, *x.shape[2:]), w.reshape(B * Ci, Co, kh, kw), stride=stride,
                           padding=padding, groups=B)
    return y.view(B, Co, y.shape[2], y.shape[3])


def conv2d_dgrad(dy: torch.Tensor, w: torch.Tensor, in_hw: Sequence[int], stride=1, padding=0) -> torch.Tensor:
    with torch.enable_grad():
        x = torch.zeros(dy.shape[0], w.shape[-3], in_hw[0], in_hw[1], dtype=dy.dtype, requires_grad=True)
        y = conv2d(x, w.detach(), stride, padding)
        return torch.autograd.grad(y, x, dy.detach())[0]


def conv2d_wgrad(dy: torch.Tensor, x: torch.Tensor, khw: Sequence[int], stride=1, padding=0,
                 per_sample: bool = False) -> torch.Tensor:
    B, C = x.shape[:2]
    O = dy.shape[1]
    shape = (B, O, C, khw[0], khw[1]) if per_sample else (O, C, khw[0], khw[1])
    with torch.enable_grad():
        w = torch.zeros(shape, dtype=x.dtype, requires_grad=True)
        y = conv2d(x.detach(), w, stride, padding)
        return torch.autograd.grad(y, w, dy.detach())[0]


def noise_bias_act(x, noise, noise_w, bias, alpha, scale):
    """multi_stylegan_generator.py:292 then op_static/fused_act.py:58."""
    v = x
    if noise is not None:
        v = v + noise_w * noise
    if bias is not None:
        v = v + bias.view(1, -1, 1, 1)
    return torch.where(v > 0, v, v * alpha) * scale


def noise_bias_act_masked(x, ref, noise, noise_w, bias, alpha, scale):
    """The same epilogue with the leaky-ReLU mask taken from `ref` when given (the form the backward and the
    double backward use: fused_bias_act_kernel.cu:36-45 case 31, extended by the noise term)."""
    v = x
    if noise is not None:
        v = v + noise_w * noise
    if bias is not None:
        v = v + bias.view(1, -1, 1, 1)
    m = v if ref is None else ref
    return torch.where(m > 0, v, v * alpha) * scale


def modulate_weights(W, s, scale, demodulate):
    """multi_stylegan_generator.py:384-388.  W [O,C,kh,kw], s [B,C]."""
    w = scale * W.unsqueeze(0) * s.view(s.shape[0], 1, -1, 1, 1)
    if not demodulate:
        return w, None
    d = torch.rsqrt((w ** 2).sum(dim=[2, 3, 4]) + 1e-8)
    return w * d.view(*d.shape, 1, 1, 1), d


def affine_warp(x: torch.Tensor, theta: torch.Tensor, mode: int = 0) -> torch.Tensor:
    """Bilinear sampling at pixel coordinates theta @ (x, y, 1) via grid_sample(align_corners=True);
    mode 0 = reflection padding, 1 = zeros."""
    B, C, H, W = x.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=x.dtype), torch.arange(W, dtype=x.dtype), indexing="ij")
    base = torch.stack([xs, ys, torch.ones_like(xs)], dim=-1).view(1, H, W, 3)
    src = torch.einsum("bhwk,bjk->bhwj", base.expand(B, H, W, 3), theta)           # [B,H,W,2] pixel coords
    gx = 2 * src[..., 0] / max(W - 1, 1) - 1
    gy = 2 * src[..., 1] / max(H - 1, 1) - 1
    grid = torch.stack([gx, gy], dim=-1)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="reflection" if mode == 0 else "zeros",
                         align_corners=True)
