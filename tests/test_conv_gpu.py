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
    (2, 32, 64, 160, 160, 3, 1, 1, False),  # >= 148 tiles of 256 pixels: two sub-tiles per CTA, two MMA-issuing warps
    (1, 32, 128, 200, 200, 3, 1, 1, False),
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


EPI_SHAPES = [
    (2, 64, 64, 32, 32, 3, 1, 1, True),      # StyledConv2d 3x3
    (2, 32, 48, 16, 16, 3, 1, 1, True),
    (3, 64, 128, 40, 24, 3, 1, 1, False),    # ragged tile edges, several n-chunks
    (2, 64, 320, 16, 16, 1, 1, 0, False),    # two n-tiles of 256 (second one partial)
    (2, 16, 20, 8, 8, 3, 1, 1, False),       # N < 32: scalar-store epilogue
    (9, 32, 32, 64, 64, 3, 1, 1, False),     # more tiles than fit one round of accumulators per CTA
]


@pytest.mark.parametrize("shape", EPI_SHAPES)
@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_conv_fused_epilogue(built_library, shape, engine):
    """conv + noise + bias + leaky ReLU + residual add + gain in the kernel epilogue == the same ops applied to
    the oracle's convolution (multi_stylegan_generator.py:289-292, fused_act.py:58, u_net_2d_discriminator.py:186)."""
    from multi_stylegan_b200 import _C, _lib
    from tests import backend_oracle
    B, C, O, H, W, k, s, p, per = shape
    x, w, y, dy = make(shape, seed=3)
    g = torch.Generator().manual_seed(9)
    bias = torch.randn(O, generator=g)
    nw = torch.tensor([0.41])
    add = torch.randn(y.shape, generator=g)
    d = dev()
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_TC if engine == "tc" else _lib.CONV_FORCE_SIMT
    tol = 1e-2 if engine == "tc" else 1e-4
    try:
        for noise in (torch.randn(B, 1, y.shape[2], y.shape[3], generator=g), torch.randn(1, 1, y.shape[2], y.shape[3], generator=g), None):
            for kw in (dict(bias=bias, act=True, gain=1.0), dict(act=True, gain=2 ** 0.5), dict(add=add, gain=0.5 ** 0.5),
                       dict(bias=bias, add=add, act=True, gain=0.7)):
                if noise is not None and "add" in kw:
                    continue
                kwargs = dict(kw, noise=noise, noise_w=nw if noise is not None else None)
                want = backend_oracle.conv2d_forward(x, w, s, p, alpha=0.9, **kwargs)
                dk = {k2: (v.to(d) if isinstance(v, torch.Tensor) else v) for k2, v in kwargs.items()}
                got = _C.conv2d_forward(x.to(d), w.to(d), s, p, alpha=0.9, **dk)
                assert got.shape == want.shape and rel_err(got, want) < tol, (engine, kw.keys(), rel_err(got, want))
        torch.cuda.synchronize()
    finally:
        _C.conv_flags = old


def test_conv_bias_act_autograd_on_device(built_library):
    """Fused conv+epilogue Function: forward, gradients and a second-order gradient against torch on the CPU."""
    import torch.nn.functional as F
    from multi_stylegan_b200 import conv
    torch.manual_seed(1)
    x = torch.randn(2, 32, 16, 16, requires_grad=True)
    w = (torch.randn(2, 48, 32, 3, 3) / 17).requires_grad_(True)
    b = torch.randn(48, requires_grad=True)
    nw = torch.tensor([0.3], requires_grad=True)
    noise = torch.randn(2, 1, 16, 16)

    def ref(x, w, b, nw):
        y = F.conv2d(x.reshape(1, -1, 16, 16), w.reshape(-1, 32, 3, 3), padding=1, groups=2).reshape(2, 48, 16, 16)
        return F.leaky_relu(y + nw * noise + b.view(1, -1, 1, 1), 0.2)

    def run(fn, args):
        y = fn(*args)
        gs = torch.autograd.grad((y ** 2).sum(), args, create_graph=True)
        gg = torch.autograd.grad(sum((g ** 2).sum() for g in gs), args[1])[0]
        return (y,) + gs + (gg,)
    want = run(ref, (x, w, b, nw))
    d = dev()
    nd = noise.to(d)
    from multi_stylegan_b200 import _C, _lib
    old = _C.conv_flags
    try:
        # exact-fp32 engine: every order to 1e-3.  TF32 engine: outputs / first-order gradients to 2e-2; the
        # second-order term differentiates through leaky-ReLU masks that TF32 rounding flips, so it is checked in L2.
        for flags, tol in ((_lib.CONV_FORCE_SIMT, 1e-3), (_lib.CONV_AUTO, 2e-2)):
            _C.conv_flags = flags
            args = tuple(t.detach().to(d).requires_grad_(True) for t in (x, w, b, nw))
            got = run(lambda x, w, b, nw: conv.conv2d_bias_act(x, w, bias=b, noise=nd, noise_w=nw, stride=1, padding=1), args)
            for a, r in zip(got[:-1], want[:-1]):
                assert rel_err(a, r) < tol, (flags, rel_err(a, r))
            gg, rg = got[-1].double().cpu(), want[-1].double()
            assert ((gg - rg).norm() / rg.norm()).item() < (1e-3 if flags == _lib.CONV_FORCE_SIMT else 0.1)
    finally:
        _C.conv_flags = old


def test_conv_cta_pair_kernel_subprocess(built_library):
    """The cta_group::2 kernel (default for wide layers with >= two rounds of tile pairs; MSG_B200_TC_VARIANT bit 256,
    read once per process, selects it for every eligible shape) against the oracle: plain, with the fused epilogues,
    per-sample and shared filters, odd tile counts."""
    import os
    import subprocess
    import sys
    from tests.conftest import ROOT
    code = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "from multi_stylegan_b200 import _C, _lib\n"
        "from oracle import ops\n"
        "from tests import backend_oracle\n"
        "_C.conv_flags = _lib.CONV_FORCE_TC\n"
        "g = torch.Generator().manual_seed(0)\n"
        "for (B, C, O, H, W, k, per) in [(2, 64, 320, 40, 24, 3, True), (3, 32, 512, 64, 64, 1, True), (1, 96, 256, 130, 130, 3, False),\n"
        "                                (2, 64, 256, 24, 40, 3, False), (2, 64, 128, 40, 24, 3, False), (3, 32, 96, 64, 64, 3, True)]:\n"
        "    x = torch.randn(B, C, H, W, generator=g)\n"
        "    w = torch.randn((B, O, C, k, k) if per else (O, C, k, k), generator=g) / (C * k * k) ** 0.5\n"
        "    bias = torch.randn(O, generator=g); add = torch.randn(B, O, H, W, generator=g)\n"
        "    noise = torch.randn(B, 1, H, W, generator=g); nw = torch.tensor([0.3])\n"
        "    for kw in (dict(), dict(bias=bias, act=True, gain=1.4, noise=noise, noise_w=nw), dict(add=add, gain=0.7)):\n"
        "        want = backend_oracle.conv2d_forward(x, w, 1, k // 2, alpha=0.9, **kw)\n"
        "        dk = {a: (v.cuda() if isinstance(v, torch.Tensor) else v) for a, v in kw.items()}\n"
        "        got = _C.conv2d_forward(x.cuda(), w.cuda(), 1, k // 2, alpha=0.9, **dk).cpu()\n"
        "        err = ((got - want).abs().max() / want.abs().max()).item()\n"
        "        assert err < 1e-2, (B, C, O, H, W, k, list(kw), err)\n"
        "torch.cuda.synchronize(); print('PAIR_OK')\n" % ROOT)
    env = dict(os.environ, MSG_B200_TC_VARIANT=str(256 + 1024))      # + the opt-in 128-channel pair tiles
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "PAIR_OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("shape", [(2, 48, 40, 32, 32, 2, 2, 0, True), (2, 64, 64, 32, 32, 3, 1, 1, True),
                                   (2, 32, 64, 64, 64, 3, 2, 0, False)])
@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_conv_transposed_weight_layout(built_library, shape, engine):
    """w_transposed: the filters are stored [C, O, kh, kw] (conv_transpose2d's layout) and read / written in place."""
    from multi_stylegan_b200 import _C, _lib
    B, C, O, H, W, k, s, p, per = shape
    x, w, y, dy = make(shape, seed=5)
    wt = w.transpose(-4, -3).contiguous()
    d = dev()
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_TC if engine == "tc" else _lib.CONV_FORCE_SIMT
    tol = 1e-2 if engine == "tc" else 1e-4
    try:
        got = _C.conv2d_forward(x.to(d), wt.to(d), s, p, w_transposed=True)
        assert rel_err(got, y) < tol
        got = _C.conv2d_dgrad(dy.to(d), wt.to(d), (H, W), s, p, w_transposed=True)
        assert rel_err(got, ops.conv2d_dgrad(dy, w, (H, W), s, p)) < tol
        got = _C.conv2d_wgrad(dy.to(d), x.to(d), (k, k), s, p, per, w_transposed=True)
        want = ops.conv2d_wgrad(dy, x, (k, k), s, p, per).transpose(-4, -3)
        assert got.shape == want.shape and rel_err(got, want) < tol
        torch.cuda.synchronize()
    finally:
        _C.conv_flags = old


@pytest.mark.parametrize("shape", [(2, 32, 64, 48, 16, 16, 3, 1), (2, 128, 128, 128, 64, 64, 3, 1), (1, 64, 32, 48, 37, 29, 1, 0),
                                   (2, 256, 128, 128, 32, 32, 1, 0)])
def test_conv_two_source_concat(built_library, shape):
    """msg_conv2d_forward_cat2: conv([x1 | x2], w) read in place == conv of the materialised concatenation
    (u_net_2d_discriminator.py:137,174-186), plain and with the fused epilogues of the decoder's ResNetBlock."""
    from multi_stylegan_b200 import _C
    from tests import backend_oracle
    B, C1, C2, O, H, W, k, p = shape
    g = torch.Generator().manual_seed(5)
    x1, x2 = torch.randn(B, C1, H, W, generator=g), torch.randn(B, C2, H, W, generator=g)
    w = torch.randn(O, C1 + C2, k, k, generator=g) / ((C1 + C2) * k * k) ** 0.5
    bias, add = torch.randn(O, generator=g), torch.randn(B, O, H, W, generator=g)
    d = dev()
    assert _C.cat2_supported(x1.to(d), x2.to(d), 1)
    for kw in (dict(), dict(bias=bias, act=True, gain=2 ** 0.5), dict(add=add, gain=0.5 ** 0.5)):
        want = backend_oracle.conv2d_forward(torch.cat([x1, x2], 1), w, 1, p, alpha=0.8, **kw)
        dk = {k2: (v.to(d) if isinstance(v, torch.Tensor) else v) for k2, v in kw.items()}
        got = _C.conv2d_forward(x1.to(d), w.to(d), 1, p, alpha=0.8, x2=x2.to(d), **dk)
        assert _C.conv2d_last_engine() == "tcgen05"
        assert got.shape == want.shape and rel_err(got, want) < 1e-2, (kw.keys(), rel_err(got, want))
        # bit-identical to the same kernel fed with the materialised concatenation (same K order, same rounding)
        same = _C.conv2d_forward(torch.cat([x1, x2], 1).to(d), w.to(d), 1, p, alpha=0.8, **dk)
        assert torch.equal(got, same)
    torch.cuda.synchronize()


def test_conv_two_source_autograd_on_device(built_library):
    """The three two-source Functions (plain, bias+act, add+scale): outputs, first-order gradients w.r.t. both sources,
    the filter and the residual, and a second-order gradient, against torch ops on the concatenation."""
    import torch.nn.functional as F
    from multi_stylegan_b200 import conv
    torch.manual_seed(2)
    x1 = torch.randn(2, 32, 16, 16, requires_grad=True)
    x2 = torch.randn(2, 64, 16, 16, requires_grad=True)
    w = (torch.randn(48, 96, 3, 3) / 29).requires_grad_(True)
    b = torch.randn(48, requires_grad=True)
    other = torch.randn(2, 48, 16, 16, requires_grad=True)

    def run(fn, args):
        y = fn(*args)
        gs = torch.autograd.grad((y ** 2).sum(), args, create_graph=True)
        gg = torch.autograd.grad(sum((g ** 2).sum() for g in gs), args[2])[0]
        return (y,) + gs + (gg,)
    refs = [lambda x1, x2, w, b, o: F.conv2d(torch.cat([x1, x2], 1), w, padding=1) * 0.7 + 0 * (b.sum() + o.sum()),
            lambda x1, x2, w, b, o: F.leaky_relu(F.conv2d(torch.cat([x1, x2], 1), w, padding=1) * 0.7 + b.view(1, -1, 1, 1), 0.2) * 1.3 + 0 * o.sum(),
            lambda x1, x2, w, b, o: (F.conv2d(torch.cat([x1, x2], 1), w, padding=1) * 0.7 + o) * 0.6 + 0 * b.sum()]
    ours = [lambda x1, x2, w, b, o: conv.conv2d(x1, w, padding=1, alpha=0.7, x2=x2) + 0 * (b.sum() + o.sum()),
            lambda x1, x2, w, b, o: conv.conv2d_bias_act(x1, w, bias=b, padding=1, gain=1.3, alpha=0.7, x2=x2) + 0 * o.sum(),
            lambda x1, x2, w, b, o: conv.conv2d_add_scale(x1, w, o, padding=1, gain=0.6, alpha=0.7, x2=x2) + 0 * b.sum()]
    d = dev()
    for ref, fn in zip(refs, ours):
        want = run(ref, (x1, x2, w, b, other))
        args = tuple(t.detach().to(d).requires_grad_(True) for t in (x1, x2, w, b, other))
        got = run(fn, args)
        for a, r in zip(got[:-1], want[:-1]):
            assert a.shape == r.shape and rel_err(a, r) < 2e-2, rel_err(a, r)
        gg, rg = got[-1].double().cpu(), want[-1].double()
        assert ((gg - rg).norm() / rg.norm()).item() < 0.1


@pytest.mark.parametrize("case", [
    # (B, C, O, H, W, per_sample): image rows of > 64 pixels, N <= 128 -> the row-tap kernel (one activation box per filter
    # row and chunk, the three taps entered at shifted rows of it)
    (2, 64, 128, 70, 130, False),      # ragged right / bottom edges, two tiles per image row
    (1, 128, 128, 129, 256, False),    # odd number of rows: the second row of the last 256-pixel tile is out of the image
    (2, 32, 64, 96, 96, True),         # per-sample filters, N = 64
    (8, 32, 128, 130, 200, False),     # enough tiles for two sub-tiles per CTA
    (3, 96, 100, 67, 65, False),       # N with a tail inside the last 32-channel chunk
    (2, 32, 128, 100, 256, False),     # CTA pairs with one image-row segment per CTA (too few tiles for two)
    (2, 64, 128, 101, 200, True),      # CTA pairs, per-sample filters, odd number of pair tiles per sample
])
def test_conv_row_tap_kernel(built_library, case):
    from multi_stylegan_b200 import _C, _lib
    from tests import backend_oracle
    if not _C.tensor_core_path_available():
        pytest.skip("not an sm_100 device")
    B, C, O, H, W, per = case
    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn((B, O, C, 3, 3) if per else (O, C, 3, 3), generator=g) / (C * 9) ** 0.5
    bias = torch.randn(O, generator=g)
    add = torch.randn(B, O, H, W, generator=g)
    noise = torch.randn(B, 1, H, W, generator=g)
    nw = torch.tensor([0.3])
    d = dev()
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_TC
    try:
        for kw in (dict(), dict(bias=bias, act=True, gain=1.4, noise=noise, noise_w=nw), dict(add=add, gain=0.7)):
            want = backend_oracle.conv2d_forward(x, w, 1, 1, alpha=0.9, **kw)
            dk = {a: (v.to(d) if isinstance(v, torch.Tensor) else v) for a, v in kw.items()}
            got = _C.conv2d_forward(x.to(d), w.to(d), 1, 1, alpha=0.9, **dk)
            assert rel_err(got, want) < 1e-2, (list(kw), rel_err(got, want))
        # dgrad: the same filter rows traversed right to left
        dy = torch.randn(B, O, H, W, generator=g)
        got = _C.conv2d_dgrad(dy.to(d), w.to(d), (H, W), 1, 1)
        want = ops.conv2d_dgrad(dy, w, (H, W), 1, 1)
        assert rel_err(got, want) < 1e-2, rel_err(got, want)
        if not per and C % 32 == 0:
            # two-source K loop (decoder concatenation) through the same kernel
            x2 = torch.randn(B, 32, H, W, generator=g)
            w2 = torch.randn(O, C + 32, 3, 3, generator=g) / ((C + 32) * 9) ** 0.5
            got = _C.conv2d_forward(x.to(d), w2.to(d), 1, 1, x2=x2.to(d))
            want = ops.conv2d(torch.cat([x, x2], 1), w2, 1, 1)
            assert rel_err(got, want) < 1e-2, rel_err(got, want)
    finally:
        _C.conv_flags = old


@pytest.mark.parametrize("case", [
    # (B, C [dx channels], O [dy channels], H, W, k)
    (2, 64, 64, 127, 127, 3),          # the discriminator's odd feature maps: pixel tiles hang over both image edges
    (2, 128, 128, 64, 200, 3),         # 256-pixel tiles + row-tap kernel
    (1, 512, 512, 31, 31, 3),          # two channel tiles
    (3, 96, 32, 17, 40, 3),            # channel tail inside a 32-channel chunk
    (2, 256, 256, 130, 256, 3),        # CTA pairs
    (2, 64, 128, 20, 20, 1),           # 1x1 filter
])
def test_dgrad_with_activation_backward_epilogue(built_library, case):
    """msg_conv2d_dgrad_mask: g_pre = alpha * conv^T(dy, w) * mask(ref) * gain and its per-channel sums against the two-step
    reference (dgrad, then op_static/fused_act.py:31-40's backward).  The mask is exact (it is read from `ref`), so the
    tolerance is the convolution's; the bias gradient is compared relative to the sum of magnitudes."""
    from multi_stylegan_b200 import _C, _lib
    from tests import backend_oracle
    if not _C.tensor_core_path_available():
        pytest.skip("not an sm_100 device")
    B, C, O, H, W, k = case
    g = torch.Generator().manual_seed(sum(case))
    dy = torch.randn(B, O, H + 2 * (k // 2) - k + 1, W + 2 * (k // 2) - k + 1, generator=g)
    w = torch.randn(O, C, k, k, generator=g) / (O * k * k) ** 0.5
    ref = torch.randn(B, C, H, W, generator=g)
    d = dev()
    want, want_db = backend_oracle.conv2d_dgrad_act_bwd(dy, w, ref, k // 2, alpha=0.8, slope=0.2, gain=1.3)
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_TC
    try:
        r = _C.conv2d_dgrad_act_bwd(dy.to(d), w.to(d), ref.to(d), k // 2, alpha=0.8, slope=0.2, gain=1.3)
        assert r is not None, "shape expected to be eligible"
        got, got_db = r
        assert rel_err(got, want) < 1e-2, rel_err(got, want)
        scale = want.abs().sum((0, 2, 3)).max()
        assert (got_db.cpu() - want_db).abs().max() / scale < 1e-3
        # deterministic
        got2, got_db2 = _C.conv2d_dgrad_act_bwd(dy.to(d), w.to(d), ref.to(d), k // 2, alpha=0.8, slope=0.2, gain=1.3)
        assert torch.equal(got, got2) and torch.equal(got_db, got_db2)
        r = _C.conv2d_dgrad_act_bwd(dy.to(d), w.to(d), ref.to(d), k // 2, alpha=0.8, want_dbias=False)
        assert r[1] is None
    finally:
        _C.conv_flags = old
    _C.conv_flags = _lib.CONV_FORCE_SIMT
    try:
        assert _C.conv2d_dgrad_act_bwd(dy.to(d), w.to(d), ref.to(d), k // 2) is None      # caller falls back to two steps
    finally:
        _C.conv_flags = old


@pytest.mark.parametrize("case", [
    # (B, C, O, H, W, k, pad, per_sample): >= 256 output channels and 256-wide input-channel tiles -> CTA-pair wgrad
    (2, 256, 256, 40, 40, 3, 1, False),
    (1, 384, 384, 33, 17, 3, 1, False),       # channel counts that leave the last pair / last half tile partly empty
    (2, 512, 256, 20, 20, 1, 0, False),       # one tap (TG = 1)
    (2, 256, 256, 24, 24, 3, 1, True),        # one filter bank per sample
    (3, 256, 512, 9, 130, 2, 0, False),       # four taps: two full groups, no remainder
    (8, 512, 512, 64, 64, 3, 1, False),       # enough K for several splits (remainder tap group runs fewer of them)
])
def test_wgrad_cta_pair_kernel(built_library, case):
    from multi_stylegan_b200 import _C, _lib
    if not _C.tensor_core_path_available():
        pytest.skip("not an sm_100 device")
    B, C, O, H, W, k, pad, per = case
    g = torch.Generator().manual_seed(sum(case[:7]))
    x = torch.randn(B, C, H, W, generator=g)
    OH, OW = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    dy = torch.randn(B, O, OH, OW, generator=g)
    want = ops.conv2d_wgrad(dy, x, (k, k), 1, pad, per)
    d = dev()
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_TC
    try:
        got = _C.conv2d_wgrad(dy.to(d), x.to(d), (k, k), 1, pad, per)
        got2 = _C.conv2d_wgrad(dy.to(d), x.to(d), (k, k), 1, pad, per)
    finally:
        _C.conv_flags = old
    assert got.shape == want.shape
    assert rel_err(got, want) < 1e-2, rel_err(got, want)
    assert torch.equal(got, got2)


@pytest.mark.parametrize("case", [
    # (B, C, O, H, W, k, per_sample): few output tiles, long K loop -> the tap-split pass + fixed-order reduce with the epilogue
    (8, 512, 512, 4, 4, 3, False),
    (2, 256, 256, 8, 8, 3, True),
    (1, 96, 100, 16, 16, 3, False),         # channel tail in the last 32-channel chunk, N with a tail
    (2, 128, 64, 9, 7, 2, False),           # four taps, odd map
])
def test_conv_tap_split_path(built_library, case):
    from multi_stylegan_b200 import _C, _lib
    from tests import backend_oracle
    if not _C.tensor_core_path_available():
        pytest.skip("not an sm_100 device")
    B, C, O, H, W, k, per = case
    pad = k // 2
    g = torch.Generator().manual_seed(sum(case[:6]))
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn((B, O, C, k, k) if per else (O, C, k, k), generator=g) / (C * k * k) ** 0.5
    OH, OW = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    bias, add = torch.randn(O, generator=g), torch.randn(B, O, OH, OW, generator=g)
    noise, nw = torch.randn(B, 1, OH, OW, generator=g), torch.tensor([0.3])
    cs, s2 = torch.rand(B, O, generator=g) + 0.5, torch.randn(B, O, generator=g)
    d = dev()
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_TC
    try:
        for kw in (dict(), dict(bias=bias, act=True, gain=1.4, noise=noise, noise_w=nw), dict(add=add, gain=0.7),
                   dict(bias=bias, act=True, col_scale=cs, out2_scale=s2)):
            want = backend_oracle.conv2d_forward(x, w, 1, pad, alpha=0.9, **kw)
            dk = {a: (v.to(d) if isinstance(v, torch.Tensor) else v) for a, v in kw.items()}
            got = _C.conv2d_forward(x.to(d), w.to(d), 1, pad, alpha=0.9, **dk)
            if isinstance(want, tuple):
                assert rel_err(got[0], want[0]) < 1e-2 and rel_err(got[1], want[1]) < 1e-2, list(kw)
            else:
                assert rel_err(got, want) < 1e-2, (list(kw), rel_err(got, want))
            again = _C.conv2d_forward(x.to(d), w.to(d), 1, pad, alpha=0.9, **dk)
            assert torch.equal(again[0] if isinstance(again, tuple) else again, got[0] if isinstance(got, tuple) else got)
        dy = torch.randn(B, O, OH, OW, generator=g)
        acc = torch.randn(B, C, H, W, generator=g)
        got = _C.conv2d_dgrad(dy.to(d), w.to(d), (H, W), 1, pad, alpha=0.5, add=acc.to(d))
        want = ops.conv2d_dgrad(dy, w, (H, W), 1, pad) * 0.5 + acc
        assert rel_err(got, want) < 1e-2, rel_err(got, want)
    finally:
        _C.conv_flags = old
