"""Drop-in stand-ins for the reference's two pre-built CUDA extension modules.

The reference's Python ops do ``import fused_act_cuda`` / ``import upfirdn2d_cuda`` (multi_stylegan/op_static/fused_act.py:8,
upfirdn2d.py:8) — pybind modules compiled from fused_bias_act.cpp:11-21 and upfirdn2d.cpp:12-22.  Putting this directory
on ``sys.path`` (``sys.path.insert(0, multi_stylegan_b200.shims.PATH)``) makes those imports resolve to the modules here,
which have the same function names and argument lists and run on libmsg_b200.so: the reference's own ``op_static`` package
then works unchanged on the B200-native kernels."""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
