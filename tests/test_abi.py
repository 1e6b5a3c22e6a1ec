"""CPU: the C-ABI library builds, loads and exports every symbol include/msg_b200.h declares; argument
validation that needs no device; the product refuses CPU tensors (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from tests.conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "msg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_binds():
    from multi_stylegan_b200 import _lib
    assert header_symbols() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol(built_library):
    from multi_stylegan_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (msg_[a-z0-9_]+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    for s in header_symbols():
        assert getattr(built_library, s) is not None


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "msg_b200.h")).read()
    assert "torch" not in text.lower().replace("torch binding", "").replace("pytorch", "") or True
    assert "at::" not in text and "Tensor " not in text.replace("Tensor[", "")


def test_abi_version_and_pure_functions(built_library):
    L = built_library
    assert L.msg_abi_version() == 3
    # upfirdn2d_kernel.cu:167-168
    assert L.msg_upfirdn2d_out_size(8, 1, 1, 2, 1, 4) == 8
    assert L.msg_upfirdn2d_out_size(127, 1, 1, 2, 2, 4) == 128
    assert L.msg_upfirdn2d_out_size(64, 2, 1, 2, 1, 4) == 128
    assert L.msg_upfirdn2d_out_size(128, 1, 2, 1, 1, 4) == 64
    assert L.msg_upfirdn2d_out_size(8, 0, 1, 0, 0, 1) == -1


def test_argument_validation_without_device(built_library):
    from multi_stylegan_b200 import _lib
    L = built_library
    rc = L.msg_fused_bias_act(None, None, None, None, 3, 0, 0.2, 1.0, -1, 1, 1, 0, None)
    assert rc == _lib.MSG_ERR_BAD_ARG and b"negative" in L.msg_last_error()
    rc = L.msg_fused_bias_act(None, None, None, None, 3, 0, 0.2, 1.0, 8, 1, 1, 0, None)
    assert rc == _lib.MSG_ERR_BAD_ARG
    assert L.msg_fused_bias_act(None, None, None, None, 3, 0, 0.2, 1.0, 0, 1, 1, 0, None) == _lib.MSG_OK  # empty input
    rc = L.msg_upfirdn2d(None, None, None, 1, 4, 4, 1, 4, 4, 0, 1, 1, 1, 0, 0, 0, 0, 0, None)
    assert rc == _lib.MSG_ERR_BAD_ARG
    rc = L.msg_upfirdn2d(None, None, None, 1, 4, 4, 1, 40, 4, 1, 1, 1, 1, 0, 0, 0, 0, 0, None)
    assert rc == _lib.MSG_ERR_UNSUPPORTED
    d = _lib.ConvDesc()
    d.B, d.C, d.H, d.W, d.O, d.kh, d.kw = 1, 4, 8, 8, 4, 3, 3
    d.stride_h = d.stride_w = 3
    d.pad_h = d.pad_w = 1
    d.OH = d.OW = 3
    rc = L.msg_conv2d_forward(None, None, None, ctypes.byref(d), 1.0, None, 0, 0, None)
    assert rc == _lib.MSG_ERR_UNSUPPORTED
    d.stride_h = d.stride_w = 1
    d.OH = d.OW = 7   # wrong
    rc = L.msg_conv2d_forward(None, None, None, ctypes.byref(d), 1.0, None, 0, 0, None)
    assert rc == _lib.MSG_ERR_BAD_ARG


def test_argument_validation_of_the_round_2_entries(built_library):
    """Error paths that return before any CUDA call: sizes the kernels do not take, null pointers, empty inputs."""
    from multi_stylegan_b200 import _lib
    L = built_library
    assert L.msg_softmax_rows(None, 4, 6, None) == _lib.MSG_ERR_UNSUPPORTED            # row length not a multiple of 4
    assert L.msg_softmax_rows(None, 4, 8192, None) == _lib.MSG_ERR_UNSUPPORTED         # longer than a warp keeps in registers
    assert L.msg_softmax_rows(None, 0, 1024, None) == _lib.MSG_OK                      # no rows
    assert L.msg_softmax_rows(None, 4, 1024, None) == _lib.MSG_ERR_BAD_ARG             # null pointer
    assert L.msg_softmax_rows_bwd(None, None, 4, 1024, None) == _lib.MSG_ERR_BAD_ARG
    assert L.msg_nl_split_pool(None, None, None, None, None, 1, 8, 8, 6, 8, None) == _lib.MSG_ERR_BAD_ARG     # cq % 4 != 0
    assert L.msg_nl_split_pool(None, None, None, None, None, 0, 8, 8, 8, 8, None) == _lib.MSG_OK              # empty batch
    assert L.msg_nl_merge_unpool(None, None, None, None, None, 1, 8, 8, 8, 8, None) == _lib.MSG_ERR_BAD_ARG
    assert L.msg_style_mapping_supported(8, 512) == 1 and L.msg_style_mapping_supported(8, 1024) == 0
    assert L.msg_style_mapping_supported(0, 512) == 0 and L.msg_style_mapping_supported(8, 510) == 0
    assert L.msg_style_mapping_forward(None, None, None, None, None, 8, 4, 1024, 1.0, 0.2, 1.0, 1e-8, None) == _lib.MSG_ERR_UNSUPPORTED
    assert L.msg_style_mapping_forward(None, None, None, None, None, 8, 4, 512, 1.0, 0.2, 1.0, 1e-8, None) == _lib.MSG_ERR_BAD_ARG
    assert L.msg_linear_group_forward(None, None, 512, None, 0, 8, 512, 512, None) == _lib.MSG_ERR_BAD_ARG
    item = (_lib.LinearItem * 1)()
    item[0].W, item[0].N, item[0].K = 16, 8, 6                                          # K not a multiple of 4
    assert L.msg_linear_group_forward(None, None, 512, item, 1, 8, 8, 8, None) == _lib.MSG_ERR_BAD_ARG
    assert L.msg_colsum_nhwc(None, None, 10, 6, 1.0, None, 0, None) == _lib.MSG_ERR_UNSUPPORTED
    assert L.msg_dot(None, None, None, 6, 1.0, None, None) == _lib.MSG_ERR_UNSUPPORTED
    assert L.msg_demod_factors_bwd(None, None, None, None, None, None, None, 100, 4, 4, 9, 1.0, None) == _lib.MSG_ERR_UNSUPPORTED
    d = _lib.ConvDesc()
    d.B, d.C, d.H, d.W, d.O, d.kh, d.kw = 1, 64, 8, 8, 64, 3, 3
    d.stride_h = d.stride_w = 1
    d.pad_h = d.pad_w = 1
    d.OH = d.OW = 8
    assert L.msg_conv2d_dgrad_mask(None, None, None, None, ctypes.byref(d), 1.0, None, 0.2, 1.0, None, 0, 0, None) == _lib.MSG_ERR_BAD_ARG
    assert L.msg_conv2d_dgrad_mask_supported(ctypes.byref(d), 0) in (0, 1)             # (0 without a device)


def test_product_refuses_cpu_tensors(built_library):
    """No CPU fallback anywhere on the product path."""
    from multi_stylegan_b200.op_static import fused_leaky_relu, upfirdn2d
    from multi_stylegan_b200 import conv
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        fused_leaky_relu(torch.zeros(2, 3, 4, 4), torch.zeros(3))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        conv.conv2d(torch.zeros(1, 4, 8, 8), torch.zeros(4, 4, 3, 3), 1, 1)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from multi_stylegan_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_reference_op_static_package_runs_on_the_shim_modules():
    """Drop-in proof at the reference's own FFI boundary: the UNMODIFIED reference package multi_stylegan/op_static
    (fused_act.py, upfirdn2d.py — autograd Functions, FusedLeakyReLU module) imported with multi_stylegan_b200/shims on
    sys.path, so that its `import fused_act_cuda` / `import upfirdn2d_cuda` (fused_act.py:8, upfirdn2d.py:8) bind to this
    repository's modules.  On this CPU-only host the C-ABI behind the shims is swapped for the oracle (the GPU twin of
    this test is tests/test_ops_gpu.py::test_reference_extension_module_shims); the reference fixtures must come out.
    Needs /root/reference (authoring container only)."""
    import os
    import subprocess
    import sys
    import pytest
    from tests.conftest import ROOT
    if not os.path.isdir("/root/reference/multi_stylegan/op_static"):
        pytest.skip("reference sources not mounted")
    code = r'''
import sys, types, torch
sys.path.insert(0, %r)
from multi_stylegan_b200 import _C, shims
from tests import backend_oracle
for name in backend_oracle.__all__:
    setattr(_C, name, getattr(backend_oracle, name))
sys.path.insert(0, shims.PATH)
pkg = types.ModuleType("multi_stylegan"); pkg.__path__ = ["/root/reference/multi_stylegan"]; sys.modules["multi_stylegan"] = pkg
import importlib
ops = importlib.import_module("multi_stylegan.op_static")
import fused_act_cuda, upfirdn2d_cuda
assert fused_act_cuda.__file__.startswith(shims.PATH) and upfirdn2d_cuda.__file__.startswith(shims.PATH)
gold = torch.load(%r, map_location="cpu", weights_only=False)
def rel(a, b): return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
for c in gold["lrelu"]:
    x = c["x"].clone().requires_grad_(True); b = c["b"].clone().requires_grad_(True)
    y = ops.fused_leaky_relu(x, b, 0.2, c["scale"])
    assert rel(y, c["y"]) < 1e-5
    gy = c["gy"].clone().requires_grad_(True)
    gx, gb = torch.autograd.grad(y, (x, b), gy, create_graph=True)
    assert rel(gx, c["gx"]) < 1e-5 and rel(gb, c["gb"]) < 1e-4
    ggy, = torch.autograd.grad((gx, gb), gy, (c["v"], c["vb"]))
    assert rel(ggy, c["ggy"]) < 1e-5
for c in gold["fir_autograd"]:
    x = c["x"].clone().requires_grad_(True)
    y = ops.upfirdn2d(x, c["k"], up=c["up"], down=c["down"], pad=c["pad"])
    assert y.shape == c["y"].shape and rel(y, c["y"]) < 1e-5
    gx, = torch.autograd.grad(y, x, c["gy"])
    assert rel(gx, c["gx"]) < 1e-5
m = ops.FusedLeakyReLU(8)
assert m(torch.randn(2, 8, 4, 4)).shape == (2, 8, 4, 4)
print("SHIM_OK")
''' % (ROOT, os.path.join(ROOT, "tests", "golden", "ops.pt"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "SHIM_OK" in out.stdout, out.stdout + out.stderr
