"""``upfirdn2d_cuda`` as the reference binds it (multi_stylegan/op_static/upfirdn2d.cpp:12-22):
``upfirdn2d(input[major,h,w,minor], kernel[kh,kw], up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)``
-> ``[major, out_h, out_w, minor]``.  Computed by msg_upfirdn2d (include/msg_b200.h)."""
from multi_stylegan_b200 import _C

__all__ = ["upfirdn2d"]


def upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    return _C.upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
