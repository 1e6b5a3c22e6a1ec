"""GPU: conv forward / dgrad / wgrad through the C-ABI.  The CUDA-core engine must match the fp32 oracle
to 1e-4; the tcgen05 (TF32) engine to 1e-2 relative (north-star tolerance), on the layer shapes of the
generator and discriminator (channels scaled down where the CPU oracle would take minutes)."""
import pytest
import torch

from oracle import ops
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


# (B, C, O, H, W, k, stride, pad, per_sample)
SHAPES = [
    (2, 8, 12, 8, 8, 3, 1, 1, True),        # tiny: CUDA-core territory
    (2, 16, 16, 4, 4, 3, 1, 1, True),       # generator 4x4
    (2, 32, 48, 16, 16, 3, 1, 1, True),     # 16 px rows  (64B swizzle atom)
    (2, 64, 64, 32, 32, 3, 1, 1, True),     # 32 px rows  (128B swizzle atom)
    (2, 64, 128, 64, 64, 3, 1, 1, True),
    (1, 512, 512, 32, 32, 3, 1, 1, True),   # full channel count, one sample
    (2, 64, 3, 32, 32, 1, 1, 0, True),      # tRGB (N=3)
    (2, 6, 32, 64, 64, 3, 1, 1, False),     # first D conv (C=6)
    (2, 33, 64, 32, 32, 3, 1, 1, False),    # mbstd channel count (odd C)
    (2, 64, 48, 64, 64, 1, 1, 0, False),    # theta/phi 1x1 (N=48)
    (2, 32, 32, 64, 64, 3, 2, 0, False),    # D downscale, 64 -> 31
    (2, 32, 32, 128, 128, 3, 2, 0, False),  # 128 -> 63 (row pitch not a multiple of 16 bytes)
    (2, 32, 1, 64, 64, 1, 1, 0, False),     # pixel head (N=1)
    (2, 48, 40, 32, 32, 2, 2, 0, True),     # stride-2 2x2 (adjoint of the generator's up-conv)
]


def make(shape, seed=0):
    B, C, O, H, W, k, s, p, per = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn((B, O, C, k, k) if per else (O, C, k, k), generator=g) / (C * k * k) ** 0.5
    y = ops.conv2d(x, w, s, p)
    dy = torch.randn(y.shape, generator=g)
    return x, w, y, dy


def run_engine(shape, flags, tol, channels_last=True):
    from multi_stylegan_b200 import _C, _lib
    B, C, O, H, W, k, s, p, per = shape
    x, w, y, dy = make(shape)
    old = _C.conv_flags, _C.conv_channels_last
    _C.conv_flags = flags
    _C.conv_channels_last = channels_last
    try:
        got = _C.conv2d_forward(x.to(dev()), w.to(dev()), s, p)
        eng_f = _C.conv2d_last_engine()
        assert got.shape == y.shape and rel_err(got, y) < tol, ("forward", eng_f, rel_err(got, y))
        got = _C.conv2d_dgrad(dy.to(dev()), w.to(dev()), (H, W), s, p)
        eng_d = _C.conv2d_last_engine()
        want = ops.conv2d_dgrad(dy, w, (H, W), s, p)
        assert rel_err(got, want) < tol, ("dgrad", eng_d, rel_err(got, want))
        got = _C.conv2d_wgrad(dy.to(dev()), x.to(dev()), (k, k), s, p, per)
        eng_w = _C.conv2d_last_engine()
        want = ops.conv2d_wgrad(dy, x, (k, k), s, p, per)
        assert rel_err(got, want) < tol, ("wgrad", eng_w, rel_err(got, want))
        torch.cuda.synchronize()
    finally:
        _C.conv_flags, _C.conv_channels_last = old
    return eng_f, eng_d, eng_w


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("channels_last", [False, True])
def test_conv_cuda_core_engine_exact_fp32(built_library, shape, channels_last):
    from multi_stylegan_b200 import _lib
    engines = run_engine(shape, _lib.CONV_FORCE_SIMT, 1e-4, channels_last)
    assert engines == ("simt", "simt", "simt")


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_tensor_core_engine_tf32(built_library, shape):
    """Channels-last activations: every layer shape of G and D must run on tcgen05 (TF32, 1e-2)."""
    from multi_stylegan_b200 import _C, _lib
    if not _C.tensor_core_path_available():
        pytest.skip("not an sm_100 device")
    engines = run_engine(shape, _lib.CONV_FORCE_TC, 1e-2)
    assert engines == ("tcgen05", "tcgen05", "tcgen05"), engines


def test_conv_nchw_inputs_use_cuda_cores(built_library):
    from multi_stylegan_b200 import _lib
    engines = run_engine(SHAPES[3], _lib.CONV_AUTO, 1e-4, channels_last=False)
    assert engines == ("simt", "simt", "simt")


def test_transposed_conv_is_dgrad(built_library):
    """Generator up-conv (multi_stylegan_generator.py:393-401): per-sample 2x2 stride-2 transposed conv."""
    from multi_stylegan_b200 import conv
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 64, 32, 32, generator=g)
    w = torch.randn(2, 64, 48, 2, 2, generator=g) / 16          # [B, Cin, Cout, kh, kw]
    want = ops.conv_transpose2d(x, w, stride=2)
    got = conv.conv_transpose2d(x.to(dev()), w.to(dev()), stride=2)
    assert got.shape == want.shape == (2, 48, 64, 64)
    assert rel_err(got, want) < 1e-2


def test_conv_autograd_double_backward_on_device(built_library):
    import torch.nn.functional as F
    from multi_stylegan_b200 import conv
    torch.manual_seed(0)
    x = torch.randn(2, 32, 32, 32, requires_grad=True)
    w = (torch.randn(48, 32, 3, 3) / 17).requires_grad_(True)

    def run(fn, x, w):
        y = fn(x, w)
        gx, = torch.autograd.grad((y ** 2).sum(), x, create_graph=True)
        gw, = torch.autograd.grad((gx ** 2).sum(), w)
        return y, gx, gw
    want = run(lambda x, w: F.conv2d(x, w, padding=1), x, w)
    xd = x.detach().to(dev()).requires_grad_(True)
    wd = w.detach().to(dev()).requires_grad_(True)
    got = run(lambda x, w: conv.conv2d(x, w, 1, 1), xd, wd)
    for a, b in zip(got, want):
        assert rel_err(a, b) < 2e-2
