"""CPU stand-ins for multi_stylegan_b200._C built from oracle.ops — used ONLY by the `oracle_backend`
fixture to test host-side logic without a GPU.  The product never imports this."""
import torch

from oracle import ops

__all__ = ["fused_bias_act", "fused_bias_act_bwd", "upfirdn2d", "conv2d_forward", "conv2d_dgrad", "conv2d_wgrad",
           "modulate_weights", "noise_bias_act", "affine_warp", "noise_bias_act_cl", "noise_bias_act_cl_bwd", "modulate_weights_bwd", "blur_noise_bias_act", "affine_warp_bwd",
           "demod_factors", "styled_act_bwd", "blur_noise_bias_act_mod", "conv2d_dgrad_act_bwd",
           "demod_factors_bwd", "colsum_cl", "dot"]


def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    return ops.fused_bias_act(input, bias, refer, act, grad, alpha, scale)


def fused_bias_act_bwd(grad_output, out, alpha, scale, channels):
    dx = ops.fused_bias_act(grad_output, grad_output.new_empty(0), out, 3, 1, alpha, scale)
    dims = [0] + list(range(2, dx.dim()))
    return dx, dx.sum(dims)


def upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    return ops.upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)


def _wt(w, w_transposed):
    return w.transpose(-4, -3) if w_transposed else w


def conv2d_forward(x, w, stride=1, padding=0, alpha=1.0, bias=None, noise=None, noise_w=None, add=None, act=False,
                   slope=0.2, gain=1.0, w_transposed=False, x2=None, col_scale=None, out2_scale=None):
    if x2 is not None:
        x = torch.cat([x, x2], dim=1)
    v = ops.conv2d(x, _wt(w, w_transposed), stride, padding) * alpha
    if col_scale is not None:
        v = v * col_scale.reshape(-1, v.shape[1], 1, 1)
    if out2_scale is not None:
        y = _tail(v, bias, noise, noise_w, add, act, slope, gain)
        return y, y * out2_scale.reshape(-1, y.shape[1], 1, 1)
    return _tail(v, bias, noise, noise_w, add, act, slope, gain)


def _tail(v, bias, noise, noise_w, add, act, slope, gain):
    if noise is not None:
        v = v + noise_w * noise
    if bias is not None:
        v = v + bias.view(1, -1, 1, 1)
    if act:
        v = torch.where(v > 0, v, v * slope)
    if add is not None:
        v = v + add
    return v * gain


def conv2d_dgrad(dy, w, in_hw, stride=1, padding=0, alpha=1.0, w_transposed=False, add=None):
    dx = ops.conv2d_dgrad(dy, _wt(w, w_transposed), in_hw, stride, padding) * alpha
    return dx if add is None else dx + add


def conv2d_dgrad_act_bwd(dy, w, ref, padding=0, alpha=1.0, slope=0.2, gain=1.0, want_dbias=True):
    dx = ops.conv2d_dgrad(dy, w, tuple(ref.shape[2:]), 1, padding) * alpha
    dx = dx * torch.where(ref > 0, torch.ones_like(ref), torch.full_like(ref, slope)) * gain
    return dx, (dx.sum((0, 2, 3)) if want_dbias else None)


def conv2d_wgrad(dy, x, khw, stride=1, padding=0, per_sample=False, alpha=1.0, w_transposed=False):
    return _wt(ops.conv2d_wgrad(dy, x, khw, stride, padding, per_sample) * alpha, w_transposed)


def modulate_weights(W, s, scale, demodulate):
    return ops.modulate_weights(W, s, scale, demodulate)


def noise_bias_act(x, noise, noise_w, bias, alpha, scale):
    return ops.noise_bias_act(x, noise, noise_w, bias, alpha, scale)


def affine_warp(x, theta, mode=0):
    return ops.affine_warp(x, theta, mode)


def noise_bias_act_cl(x, ref, noise, noise_w, bias, alpha, scale):
    return ops.noise_bias_act_masked(x, ref, noise, noise_w, bias, alpha, scale)


def noise_bias_act_cl_bwd(grad_output, out, noise, alpha, scale, want_param_grads=True):
    dx = ops.noise_bias_act_masked(grad_output, out, None, None, None, alpha, scale)
    if not want_param_grads:
        return dx, None, None
    db = dx.sum([0, 2, 3])
    dnw = None if noise is None else (dx.sum(1, keepdim=True) * noise).sum().reshape(1)
    return dx, db, dnw


def modulate_weights_bwd(g, W, s, demod, scale, demodulate):
    Wd = W.detach().clone().requires_grad_(True)
    sd = s.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        w, _ = ops.modulate_weights(Wd, sd, scale, demodulate)
        dW, ds = torch.autograd.grad(w, (Wd, sd), g)
    return dW, ds


def blur_noise_bias_act(x, kernel, pad, noise, noise_w, bias, slope, gain):
    px0, px1, py0, py1 = pad
    B, C, H, W = x.shape
    y = ops.upfirdn2d(x.reshape(-1, H, W, 1), kernel, 1, 1, 1, 1, px0, px1, py0, py1)
    y = y.view(B, C, y.shape[1], y.shape[2])
    return ops.noise_bias_act_masked(y, None, noise, noise_w, bias, slope, gain)


def affine_warp_bwd(grad_output, theta, mode=0):
    x = torch.zeros_like(grad_output, requires_grad=True)
    with torch.enable_grad():
        y = ops.affine_warp(x, theta, mode)
        return torch.autograd.grad(y, x, grad_output)[0]


def demod_factors(W, s, scale):
    wsq = W.pow(2).sum(dim=(2, 3))
    return torch.rsqrt(scale * scale * (s * s) @ wsq.t() + 1e-8), wsq


def styled_act_bwd(g_out, g_out2, out, col_scale, out2_scale, noise, slope, gain):
    B, C = out.shape[0], out.shape[1]
    bc = lambda t: t.reshape(-1, C, 1, 1)
    gt = torch.zeros_like(out) if g_out is None else g_out
    if g_out2 is not None:
        gt = gt + bc(out2_scale) * g_out2
    pos = out > 0
    gv = gt * torch.where(pos, torch.full_like(out, gain), torch.full_like(out, gain * slope))
    v = out * torch.where(pos, torch.full_like(out, 1.0 / gain), torch.full_like(out, 1.0 / (gain * slope)))
    g_pre = gv if col_scale is None else gv * bc(col_scale)
    nz = torch.zeros(1, 1, out.shape[2], out.shape[3]) if noise is None else noise
    s4 = torch.zeros(B, C) if g_out2 is None else (out * g_out2).sum(dim=(2, 3))
    sums = torch.stack([gv.sum(dim=(2, 3)), (gv * v).sum(dim=(2, 3)), (gv * nz).sum(dim=(2, 3)), s4])
    return g_pre, sums


def blur_noise_bias_act_mod(x, kernel, pad, col_scale, noise, noise_w, bias, slope, gain, out2_scale):
    px0, px1, py0, py1 = pad
    B, C, H, W = x.shape
    y = ops.upfirdn2d(x.reshape(-1, H, W, 1), kernel, 1, 1, 1, 1, px0, px1, py0, py1)
    y = y.view(B, C, y.shape[1], y.shape[2])
    if col_scale is not None:
        y = y * col_scale.reshape(-1, C, 1, 1)
    out = ops.noise_bias_act_masked(y, None, noise, noise_w, bias, slope, gain)
    return out, (None if out2_scale is None else out * out2_scale.reshape(-1, C, 1, 1))


def demod_factors_bwd(gd, d, s, wsq, W, scale, need_w=True, need_s=True):
    q = gd * d * d * d * (-0.5 * scale * scale)
    ds = 2.0 * s * torch.mm(q, wsq) if need_s else None
    dW = W * (2.0 * torch.mm(q.t(), s * s)).view(W.shape[0], W.shape[1], 1, 1) if need_w else None
    return dW, ds


def colsum_cl(x, scale=1.0):
    return x.sum((0, 2, 3)) * scale


def dot(a, b, scale=1.0):
    return ((a * b).sum() * scale).reshape(1)
