"""Import the UNMODIFIED reference modules from /root/reference on CPU.  TEST INFRASTRUCTURE ONLY.

Works only where /root/reference is mounted (the authoring container); used by oracle/make_golden.py to
produce tests/golden/ and by the CPU tests to pin oracle/ against the reference live.

Recipe: the two pre-built CUDA extension modules the reference imports
(op_static/fused_act.py:8, op_static/upfirdn2d.py:8) are replaced by CPU stand-ins — `fused_bias_act`
restates fused_bias_act_kernel.cu:25-48 (oracle.ops), `upfirdn2d` calls the reference's *own*
`upfirdn2d_native` (op_static/upfirdn2d.py:156-190) — and `multi_stylegan/__init__.py` (which needs rtpt
and kornia) is bypassed by registering an empty package object."""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MSG_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "multi_stylegan"))


_loaded = {}


def load():
    """Returns a namespace with .generator, .discriminator, .equalized_layer, .loss, .config, .op_static, .misc"""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    from . import ops

    fa = types.ModuleType("fused_act_cuda")
    fa.fused_bias_act = lambda x, b, ref, act, grad, alpha, scale: ops.fused_bias_act(x, b, ref, act, grad, alpha, scale)
    sys.modules["fused_act_cuda"] = fa

    up = types.ModuleType("upfirdn2d_cuda")

    def _upfirdn2d(x, k, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
        native = sys.modules["multi_stylegan.op_static.upfirdn2d"].upfirdn2d_native
        return native(x, k, up_x, up_y, down_x, down_y, px0, px1, py0, py1)

    up.upfirdn2d = _upfirdn2d
    sys.modules["upfirdn2d_cuda"] = up

    pkg = types.ModuleType("multi_stylegan")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "multi_stylegan")]
    sys.modules["multi_stylegan"] = pkg

    names = {"generator": "multi_stylegan.multi_stylegan_generator",
             "discriminator": "multi_stylegan.u_net_2d_discriminator",
             "equalized_layer": "multi_stylegan.equalized_layer",
             "loss": "multi_stylegan.loss",
             "config": "multi_stylegan.config",
             "op_static": "multi_stylegan.op_static"}
    for key, mod in names.items():
        _loaded[key] = importlib.import_module(mod)
    return types.SimpleNamespace(**_loaded)
