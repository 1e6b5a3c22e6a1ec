"""First-order fused forms of the generator's small-M linears (csrc/linear_ops.cu).

* `style_mapping`: PixelwiseNormalization + depth x [EqualizedLinear(bias=False) -> FusedLeakyReLU]
  (multi_stylegan_generator.py:208-235) as one launch forward and one launch backward.
* `style_linears`: the `modulation_mapping` EqualizedLinear of every ModulatedConv2d that reads a per-layer latent
  (:355-361) as one launch forward, two backward (weight / bias gradients, latent gradient).

Both are used only where a first-order backward suffices (see _mode.py); the module-by-module torch formulation
stays in place for everything else (CPU, double backward through the style linears for path-length regularisation)."""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function

from . import _C
from ._mode import NO_DOUBLE_BACKWARD as _NO_DOUBLE


class StyleMappingFn(Function):
    @staticmethod
    def forward(ctx, z, alpha, slope, gain, eps, *params):
        weights, biases = list(params[0::2]), list(params[1::2])
        acts, x0 = _C.style_mapping_forward(z, weights, biases, alpha, slope, gain, eps)
        ctx.save_for_backward(acts, x0, *weights)
        ctx.has_bias = [b is not None for b in biases]
        ctx.biases = biases
        ctx.cfg = (alpha, slope, gain)
        return acts[-1]

    @staticmethod
    def backward(ctx, gy):
        if torch.is_grad_enabled():
            raise RuntimeError(_NO_DOUBLE)
        acts, x0, *weights = ctx.saved_tensors
        alpha, slope, gain = ctx.cfg
        dW, db = _C.style_mapping_backward(gy, acts, x0, weights, ctx.biases, alpha, slope, gain)
        grads: List[Optional[torch.Tensor]] = []
        for l in range(len(weights)):
            grads.append(dW[l] if ctx.needs_input_grad[5 + 2 * l] else None)
            grads.append(db[l] if (ctx.has_bias[l] and ctx.needs_input_grad[6 + 2 * l]) else None)
        return (None, None, None, None, None, *grads)


def style_mapping(z: torch.Tensor, layers: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]], alpha: float,
                  slope: float, gain: float, eps: float) -> torch.Tensor:
    """layers: [(W_l [K, K], act_bias_l [K] or None)]."""
    flat = []
    for W, b in layers:
        flat += [W, b]
    return StyleMappingFn.apply(z, float(alpha), float(slope), float(gain), float(eps), *flat)


_GROUPS: Dict[tuple, "_C.LinearGroup"] = {}


class LinearGroupFn(Function):
    @staticmethod
    def forward(ctx, x, group, *params):
        ctx.group = group
        ctx.save_for_backward(x)
        return group.forward(x)

    @staticmethod
    def backward(ctx, gout):
        if torch.is_grad_enabled():
            raise RuntimeError(_NO_DOUBLE)
        x, = ctx.saved_tensors
        g = ctx.group
        dW, db, dx = g.backward(gout.contiguous(), x, ctx.needs_input_grad[0])
        grads: List[Optional[torch.Tensor]] = []
        for i, (w_off, N, K) in enumerate(g.w_slices):
            grads.append(dW[w_off:w_off + N * K].view(N, K) if ctx.needs_input_grad[2 + 2 * i] else None)
            bs = g.b_slices[i]
            grads.append(db[bs[0]:bs[0] + bs[1]] if (bs is not None and ctx.needs_input_grad[3 + 2 * i]) else None)
        return (dx, None, *grads)


def style_linears(x: torch.Tensor, specs: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor], int, float, float]]
                  ) -> List[torch.Tensor]:
    """x [M, R]; specs: (W [N, K], bias [N] or None, in_off, alpha, beta), items reading the same slice adjacent.
    Returns the outputs as contiguous views [M, N_i] of one flat buffer."""
    key = tuple((W.data_ptr(), tuple(W.shape), None if b is None else b.data_ptr(), int(o), float(a), float(be))
                for W, b, o, a, be in specs)
    group = _GROUPS.get(key)
    if group is None:
        if len(_GROUPS) > 64:
            _GROUPS.clear()
        group = _GROUPS[key] = _C.LinearGroup(specs)
    flat = []
    for W, b, *_ in specs:
        flat += [W, b]
    out = LinearGroupFn.apply(x, group, *flat)
    M = x.shape[0]
    return [out[off * M:(off + n) * M].view(M, n) for off, n in group.out_slices]
