"""``fused_act_cuda`` as the reference binds it (multi_stylegan/op_static/fused_bias_act.cpp:11-21):
``fused_bias_act(input, bias, refer, act, grad, alpha, scale) -> Tensor`` — empty ``bias`` / ``refer`` mean "absent",
the bias indexes dim 1, a new contiguous tensor is returned.  Computed by msg_fused_bias_act (include/msg_b200.h)."""
from multi_stylegan_b200 import _C

__all__ = ["fused_bias_act"]


def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    return _C.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
