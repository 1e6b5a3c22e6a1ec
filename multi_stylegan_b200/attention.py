"""The discriminator's non-local block (u_net_2d_discriminator.py:332-381) on the library's own kernels, first-order form.

    theta, phi, g = conv1x1(x) (x3);  phi, g <- maxpool2x2;  beta = softmax(theta^T phi);  o = conv1x1(g beta^T)
    out = (gamma * o + conv1x1_res(x)) / sqrt(2)

Forward = 5 tcgen05 GEMM launches + 2 memory-bound passes:
  1. the three input convolutions as ONE GEMM with the filters stacked (the block input — 768 channels in the U-Net
     decoder, read in place from its two sources — is read once instead of three times, and its gradient comes out of one
     dgrad instead of three partial gradients that have to be added);
  2. `nl_split_pool`: dense theta + pooled phi, g + argmax;
  3. scores = 1x1 convolution of theta with phi of the same sample as its filter bank (msg_conv2d_forward with
     per-sample weights: [B, HW/4, C/8] is exactly phi's channels-last memory), row softmax in place;
  4. attended = 1x1 convolution of the attention matrix with g as the (transposed) per-sample filter bank;
  5. the output convolution, which also writes gamma/sqrt(2) * o (second epilogue output), and the residual convolution
     whose epilogue adds it.
The backward mirrors it with dgrad / per-sample wgrad launches and the two adjoint passes.  The [B, HW, HW/4] attention
matrix is still materialised (once, plus its gradient) — the reference materialises it too (:370-380); no cuBLAS, no ATen
softmax / max-pool / elementwise kernels remain.  R1 (which differentiates the backward) keeps the composite formulation
(_mode.higher_order_gradients)."""
import math
from typing import Optional

import torch
from torch.autograd import Function

from . import _C
from ._mode import NO_DOUBLE_BACKWARD as _NO_DOUBLE


def _as_filters(pooled: torch.Tensor) -> torch.Tensor:
    """channels-last [B, C, PH, PW] -> its memory viewed as per-sample 1x1 filters [B, PH*PW, C, 1, 1]."""
    B, C, PH, PW = pooled.shape
    return pooled.permute(0, 2, 3, 1).reshape(B, PH * PW, C, 1, 1)


class NonLocalFused(Function):
    @staticmethod
    def forward(ctx, x, x2, w_theta, w_phi, w_g, w_o, w_res, gamma, a_in, a_o, a_res):
        j = 1.0 / math.sqrt(2)
        cq, cv = w_theta.shape[0], w_g.shape[0]
        w_qkv = torch.cat([w_theta, w_phi, w_g], dim=0)
        qkv = _C.conv2d_forward(x, w_qkv, 1, 0, alpha=a_in, x2=x2)
        theta, phi, g, idx = _C.nl_split_pool(qkv, cq, cv)
        del qkv
        nk = phi.shape[2] * phi.shape[3]
        P = _C.conv2d_forward(theta, _as_filters(phi), 1, 0)                       # scores [B, nk, H, W] = rows [HW, nk]
        _C.softmax_rows_(P, nk)
        att = _C.conv2d_forward(P, _as_filters(g), 1, 0, w_transposed=True)        # [B, cv, H, W]
        gj = (gamma.detach() * j).reshape(1, 1).expand(1, w_o.shape[0]).contiguous()
        o, o_scaled = _C.conv2d_forward(att, w_o, 1, 0, alpha=a_o, out2_scale=gj)
        out = _C.conv2d_forward(x, w_res, 1, 0, alpha=a_res, add=o_scaled, x2=x2)
        ctx.save_for_backward(x, x2, w_qkv, w_o, w_res, gamma, theta, phi, g, idx, P, att, o)
        ctx.cfg = (a_in, a_o, a_res, cq, cv)
        return out

    @staticmethod
    def backward(ctx, gout):
        if torch.is_grad_enabled():
            raise RuntimeError(_NO_DOUBLE)
        x, x2, w_qkv, w_o, w_res, gamma, theta, phi, g, idx, P, att, o = ctx.saved_tensors
        a_in, a_o, a_res, cq, cv = ctx.cfg
        need = ctx.needs_input_grad
        j = 1.0 / math.sqrt(2)
        gout = gout.contiguous(memory_format=torch.channels_last)
        hw = tuple(x.shape[2:])
        B = x.shape[0]
        nk = phi.shape[2] * phi.shape[3]
        d_gamma = _C.dot(gout, o, j).reshape(gamma.shape) if need[7] else None
        # gamma / sqrt(2) is a device scalar: it multiplies the (half as wide) results instead of the incoming gradient
        gj = gamma.detach() * j
        datt = _C.conv2d_dgrad(gout, w_o, hw, 1, 0, alpha=a_o) * gj
        dw_o = _C.conv2d_wgrad(gout, att, (1, 1), 1, 0, False, alpha=a_o) * gj if need[5] else None
        dP = _C.conv2d_forward(datt, _as_filters(g), 1, 0)                          # [B, nk, H, W]
        dg = _C.conv2d_wgrad(datt, P, (1, 1), 1, 0, True, w_transposed=True)        # [B, nk, cv, 1, 1]
        _C.softmax_rows_bwd_(dP, P, nk)                                             # dP <- dS
        dtheta = _C.conv2d_forward(dP, _as_filters(phi), 1, 0, w_transposed=True)   # [B, cq, H, W]
        dphi = _C.conv2d_wgrad(dP, theta, (1, 1), 1, 0, True)                       # [B, nk, cq, 1, 1]
        del dP
        ph, pw = phi.shape[2], phi.shape[3]
        dqkv = _C.nl_merge_unpool(dtheta, dphi.view(B, ph, pw, cq).permute(0, 3, 1, 2),
                                  dg.view(B, ph, pw, cv).permute(0, 3, 1, 2), idx)
        dx = dx2 = dw_qkv = dw_res = None
        if x2 is None:
            if need[0]:
                dx = _C.conv2d_dgrad(dqkv, w_qkv, hw, 1, 0, alpha=a_in, add=_C.conv2d_dgrad(gout, w_res, hw, 1, 0, alpha=a_res))
            dw_qkv = _C.conv2d_wgrad(dqkv, x, (1, 1), 1, 0, False, alpha=a_in)
            if need[6]:
                dw_res = _C.conv2d_wgrad(gout, x, (1, 1), 1, 0, False, alpha=a_res)
        else:
            c1 = x.shape[1]
            if need[0]:
                dx = _C.conv2d_dgrad(dqkv, w_qkv[:, :c1], hw, 1, 0, alpha=a_in,
                                     add=_C.conv2d_dgrad(gout, w_res[:, :c1], hw, 1, 0, alpha=a_res))
            if need[1]:
                dx2 = _C.conv2d_dgrad(dqkv, w_qkv[:, c1:], hw, 1, 0, alpha=a_in,
                                      add=_C.conv2d_dgrad(gout, w_res[:, c1:], hw, 1, 0, alpha=a_res))
            dw_qkv = torch.cat([_C.conv2d_wgrad(dqkv, x, (1, 1), 1, 0, False, alpha=a_in),
                                _C.conv2d_wgrad(dqkv, x2, (1, 1), 1, 0, False, alpha=a_in)], dim=1)
            if need[6]:
                dw_res = torch.cat([_C.conv2d_wgrad(gout, x, (1, 1), 1, 0, False, alpha=a_res),
                                    _C.conv2d_wgrad(gout, x2, (1, 1), 1, 0, False, alpha=a_res)], dim=1)
        return (dx, dx2, dw_qkv[:cq] if need[2] else None, dw_qkv[cq:2 * cq] if need[3] else None,
                dw_qkv[2 * cq:] if need[4] else None, dw_o, dw_res, d_gamma, None, None, None)


def non_local_block(x: torch.Tensor, x2: Optional[torch.Tensor], w_theta, w_phi, w_g, w_o, w_res, gamma: torch.Tensor,
                    a_in: float, a_o: float, a_res: float) -> torch.Tensor:
    """a_res already contains the join's 1/sqrt(2)."""
    return NonLocalFused.apply(x, x2, w_theta, w_phi, w_g, w_o, w_res, gamma, float(a_in), float(a_o), float(a_res))


def eligible(x: torch.Tensor, cq: int, cv: int) -> bool:
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
        return False
    H, W = x.shape[2], x.shape[3]
    nk = (H // 2) * (W // 2)
    return cq % 4 == 0 and cv % 4 == 0 and cq >= 4 and cv >= 4 and nk % 4 == 0 and 4 <= nk <= 4096
