"""GPU: fused_bias_act and upfirdn2d kernels, through the C-ABI, against the CPU oracle, the reference
fixtures, and size-independent properties at BASELINE sizes.  Tolerance 1e-4 relative (north star)."""
import pytest
import torch

from oracle import ops
from tests.conftest import load_golden, rel_err
from tests.test_host_logic import check_fir_autograd_case, check_lrelu_case

pytestmark = pytest.mark.gpu
TOL = 1e-4


def dev():
    return torch.device("cuda:0")


def test_library_loaded_and_counts_launches(built_library):
    from multi_stylegan_b200 import _C
    n0 = _C.launch_count()
    x = torch.randn(4, 8, 16, 16, device=dev())
    _C.fused_bias_act(x, torch.zeros(8, device=dev()), x.new_empty(0), 3, 0, 0.2, 1.0)
    assert _C.launch_count() == n0 + 1


@pytest.mark.parametrize("shape", [(2, 5, 6, 7), (3, 4), (2, 3, 40, 40), (2, 16, 64, 64), (1, 3, 33, 31), (0, 4, 8, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_fused_bias_act_all_modes(built_library, shape, dtype):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g, dtype=dtype)
    b = torch.randn(shape[1], generator=g, dtype=dtype)
    ref = torch.randn(shape, generator=g, dtype=dtype)
    e = torch.empty(0, dtype=dtype)
    for act, grad in [(3, 0), (3, 1), (3, 2), (1, 0), (1, 1), (1, 2)]:
        for bias in (b, e):
            r = ref if grad == 1 else e
            want = ops.fused_bias_act(x, bias, r, act, grad, 0.2, 1.5)
            got = _C.fused_bias_act(x.to(dev()), bias.to(dev()), r.to(dev()), act, grad, 0.2, 1.5)
            assert got.shape == want.shape and got.dtype == dtype
            if x.numel():
                assert rel_err(got, want) < 1e-6, (act, grad, bias.numel())


def test_fused_bias_act_non_contiguous_input(built_library):
    from multi_stylegan_b200 import _C
    x = torch.randn(2, 4, 8, 6).transpose(2, 3)       # [2,4,6,8] non-contiguous (not channels-last either)
    b = torch.randn(4)
    want = ops.fused_bias_act(x.contiguous(), b, torch.empty(0), 3, 0, 0.2, 1.0)
    got = _C.fused_bias_act(x.to(dev()), b.to(dev()), torch.empty(0, device=dev()), 3, 0, 0.2, 1.0)
    assert got.is_contiguous() and rel_err(got, want) < 1e-6
    # a channels-last view keeps its layout (no copy); values are identical by logical index
    xc = torch.randn(2, 6, 8, 4).permute(0, 3, 1, 2)
    want = ops.fused_bias_act(xc.contiguous(), b, torch.empty(0), 3, 0, 0.2, 1.0)
    got = _C.fused_bias_act(xc.to(dev()), b.to(dev()), torch.empty(0, device=dev()), 3, 0, 0.2, 1.0)
    assert got.is_contiguous(memory_format=torch.channels_last) and rel_err(got, want) < 1e-6


@pytest.mark.parametrize("shape", [(2, 5, 6, 7), (3, 4), (2, 3, 40, 40), (4, 32, 64, 64)])
def test_fused_bias_act_bwd_fused_reduction(built_library, shape):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(1)
    go = torch.randn(shape, generator=g)
    out = torch.randn(shape, generator=g)
    dx = ops.fused_bias_act(go, torch.empty(0), out, 3, 1, 0.2, 1.0)
    db = dx.sum([0] + list(range(2, dx.dim())))
    gdx, gdb = _C.fused_bias_act_bwd(go.to(dev()), out.to(dev()), 0.2, 1.0, shape[1])
    assert rel_err(gdx, dx) < 1e-6 and rel_err(gdb, db) < 1e-5
    # deterministic: bit-identical on a second run
    gdx2, gdb2 = _C.fused_bias_act_bwd(go.to(dev()), out.to(dev()), 0.2, 1.0, shape[1])
    assert torch.equal(gdb, gdb2)


def test_lrelu_reference_fixtures(built_library):
    for c in load_golden("ops.pt")["lrelu"]:
        check_lrelu_case(c, dev(), TOL)


def test_upfirdn2d_reference_fixtures(built_library):
    from multi_stylegan_b200 import _C
    for c in load_golden("ops.pt")["fir"]:
        y = _C.upfirdn2d(c["x"].to(dev()), c["k"].to(dev()), *c["cfg"])
        assert y.shape == c["y"].shape, c["cfg"]
        assert rel_err(y, c["y"]) < TOL, c["cfg"]
    for c in load_golden("ops.pt")["fir_autograd"]:
        check_fir_autograd_case(c, dev(), TOL)


CFGS = [  # (up, down, px0, px1, py0, py1, kh, kw)
    (1, 1, 2, 1, 2, 1, 4, 4), (1, 1, 1, 2, 1, 2, 4, 4), (1, 1, 2, 2, 2, 2, 4, 4), (1, 1, 1, 1, 1, 1, 4, 4),
    (2, 1, 2, 1, 2, 1, 4, 4), (1, 2, 1, 1, 1, 1, 4, 4), (1, 1, 1, 1, 1, 1, 3, 3), (2, 1, 1, 0, 1, 0, 2, 2),
    (1, 2, 0, 0, 0, 0, 2, 2), (1, 1, -1, 2, 0, -1, 4, 4), (3, 2, 3, 2, 1, 4, 5, 5), (2, 2, 0, 1, 2, 0, 4, 3),
]


@pytest.mark.parametrize("cfg", CFGS)
@pytest.mark.parametrize("hw", [(4, 4), (8, 8), (17, 23), (64, 64), (127, 127), (130, 260)])
def test_upfirdn2d_vs_oracle(built_library, cfg, hw):
    from multi_stylegan_b200 import _C
    up, down, px0, px1, py0, py1, kh, kw = cfg
    h, w = hw
    if h * up + py0 + py1 < kh or w * up + px0 + px1 < kw:
        pytest.skip("kernel larger than padded input")
    g = torch.Generator().manual_seed(h * 1000 + w)
    x = torch.randn(3, h, w, 1, generator=g)
    k = torch.randn(kh, kw, generator=g)
    want = ops.upfirdn2d(x, k, up, up, down, down, px0, px1, py0, py1)
    got = _C.upfirdn2d(x.to(dev()), k.to(dev()), up, up, down, down, px0, px1, py0, py1)
    assert got.shape == want.shape
    assert rel_err(got, want) < TOL


def test_upfirdn2d_minor_dim_and_fp64(built_library):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 9, 7, 3, generator=g, dtype=torch.float64)
    k = torch.randn(4, 4, generator=g, dtype=torch.float64)
    want = ops.upfirdn2d(x, k, 2, 2, 1, 1, 2, 1, 2, 1)
    got = _C.upfirdn2d(x.to(dev()), k.to(dev()), 2, 2, 1, 1, 2, 1, 2, 1)
    assert rel_err(got, want) < 1e-12
    assert _C.upfirdn2d(torch.zeros(0, 4, 4, 1, device=dev()), torch.ones(4, 4, device=dev()), 1, 1, 1, 1, 2, 1, 2, 1).shape == (0, 4, 4, 1)


def test_upfirdn2d_full_size_properties(built_library):
    """BASELINE config 2 sizes ([8,512,R,R]); the oracle is too slow there, so check properties:
    linearity, and <Fx, y> == <x, F^T y> with the adjoint configuration the autograd wrapper uses."""
    from multi_stylegan_b200.op_static import upfirdn2d
    torch.manual_seed(0)
    k = torch.tensor([1., 3., 3., 1.], device=dev())
    k = (k[None] * k[:, None]) / 64 * 4
    for (up, down, pad, R) in [(1, 1, (2, 1), 128), (2, 1, (2, 1), 64), (1, 1, (2, 2), 127)]:
        x = torch.randn(8, 512, R, R, device=dev(), requires_grad=True)
        y = upfirdn2d(x, k, up=up, down=down, pad=pad)
        gy = torch.randn_like(y)
        gx, = torch.autograd.grad(y, x, gy)
        lhs = (y.double() * gy.double()).sum()
        rhs = (x.double() * gx.double()).sum()
        assert abs((lhs - rhs) / lhs).item() < 1e-5
        x2 = torch.randn_like(x)
        y2 = upfirdn2d(x2, k, up=up, down=down, pad=pad)
        y12 = upfirdn2d(x.detach() + 2 * x2, k, up=up, down=down, pad=pad)
        assert rel_err(y12, y.detach() + 2 * y2) < 1e-5
        # DC gain: constant input -> interior equals sum(taps) (/up^2 for zero insertion)
        c = upfirdn2d(torch.ones(1, 1, R, R, device=dev()), k, up=up, down=down, pad=pad)
        assert abs(c[0, 0, 8, 8].item() - k.sum().item() / (up * up)) < 1e-5


def test_small_ops_vs_oracle(built_library):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(5)
    W = torch.randn(12, 8, 3, 3, generator=g)
    s = torch.randn(3, 8, generator=g)
    for demod in (True, False):
        want, wd = ops.modulate_weights(W, s, 0.1, demod)
        got, gd = _C.modulate_weights(W.to(dev()), s.to(dev()), 0.1, demod)
        assert rel_err(got, want) < 1e-5
        if demod:
            assert rel_err(gd, wd) < 1e-5
    x = torch.randn(3, 5, 16, 16, generator=g)
    for noise in (torch.randn(3, 1, 16, 16, generator=g), torch.randn(1, 1, 16, 16, generator=g), None):
        nw = torch.tensor([0.3])
        b = torch.randn(5, generator=g)
        want = ops.noise_bias_act(x, noise, nw, b, 0.2, 1.0)
        got = _C.noise_bias_act(x.to(dev()), None if noise is None else noise.to(dev()), nw.to(dev()), b.to(dev()), 0.2, 1.0)
        assert rel_err(got, want) < 1e-6
    import math
    th = []
    for a, sc in [(0.0, 1.0), (30.0, 1.2), (-100.0, 0.8)]:
        c, sn = math.cos(math.radians(a)) / sc, math.sin(math.radians(a)) / sc
        cx = cy = 7.5
        th.append([[c, -sn, cx - c * cx + sn * cy], [sn, c, cy - sn * cx - c * cy]])
    theta = torch.tensor(th)
    for mode in (0, 1):
        want = ops.affine_warp(x, theta, mode)
        got = _C.affine_warp(x.to(dev()), theta.to(dev()), mode)
        assert rel_err(got, want) < 1e-4


# ---- channels-last (NHWC) variants: same arithmetic, no layout copies ---------------------------------------
@pytest.mark.parametrize("shape", [(2, 8, 6, 7), (2, 16, 40, 40), (3, 132, 17, 9), (2, 512, 16, 16), (1, 4, 1, 1)])
def test_fused_bias_act_channels_last(built_library, shape):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(2)
    x = torch.randn(shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    ref = torch.randn(shape, generator=g)
    e = torch.empty(0)
    cl = lambda t: t.to(dev()).contiguous(memory_format=torch.channels_last)
    for act, grad in [(3, 0), (3, 1), (3, 2), (1, 0)]:
        for bias in (b, e):
            r = ref if grad == 1 else e
            want = ops.fused_bias_act(x, bias, r, act, grad, 0.2, 1.5)
            got = _C.fused_bias_act(cl(x), bias.to(dev()), cl(r) if r.numel() else r.to(dev()), act, grad, 0.2, 1.5)
            assert got.shape == want.shape
            if shape[2] * shape[3] > 1:
                assert got.is_contiguous(memory_format=torch.channels_last)
            assert rel_err(got, want) < 1e-6, (act, grad, bias.numel())
    dx = ops.fused_bias_act(x, e, ref, 3, 1, 0.2, 1.0)
    gdx, gdb = _C.fused_bias_act_bwd(cl(x), cl(ref), 0.2, 1.0, shape[1])
    assert rel_err(gdx, dx) < 1e-6 and rel_err(gdb, dx.sum([0, 2, 3])) < 1e-5
    gdx2, gdb2 = _C.fused_bias_act_bwd(cl(x), cl(ref), 0.2, 1.0, shape[1])
    assert torch.equal(gdb, gdb2)


@pytest.mark.parametrize("shape", [(2, 8, 6, 7), (3, 132, 17, 9), (2, 512, 32, 32)])
@pytest.mark.parametrize("shared_noise", [False, True])
def test_noise_bias_act_channels_last(built_library, shape, shared_noise):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(4)
    B, C, H, W = shape
    x = torch.randn(shape, generator=g)
    ref = torch.randn(shape, generator=g)
    noise = torch.randn(1 if shared_noise else B, 1, H, W, generator=g)
    nw = torch.tensor([0.37])
    b = torch.randn(C, generator=g)
    d = dev()
    for r in (None, ref):
        for nz in (noise, None):
            for bias in (b, None):
                want = ops.noise_bias_act_masked(x, r, nz, nw, bias, 0.2, 1.3)
                got = _C.noise_bias_act_cl(x.to(d), None if r is None else r.to(d), None if nz is None else nz.to(d),
                                           nw.to(d), None if bias is None else bias.to(d), 0.2, 1.3)
                assert rel_err(got, want) < 1e-6
    dx = ops.noise_bias_act_masked(x, ref, None, None, None, 0.2, 1.3)
    gdx, gdb, gdn = _C.noise_bias_act_cl_bwd(x.to(d), ref.to(d), noise.to(d), 0.2, 1.3)
    assert rel_err(gdx, dx) < 1e-6 and rel_err(gdb, dx.sum([0, 2, 3])) < 1e-5
    assert rel_err(gdn, (dx.sum(1, keepdim=True) * noise).sum().reshape(1)) < 1e-4
    _, gdb2, gdn2 = _C.noise_bias_act_cl_bwd(x.to(d), ref.to(d), noise.to(d), 0.2, 1.3)
    assert torch.equal(gdb, gdb2) and torch.equal(gdn, gdn2)


@pytest.mark.parametrize("cfg", CFGS)
@pytest.mark.parametrize("hwc", [(4, 4, 4), (8, 8, 8), (17, 23, 12), (33, 35, 132), (64, 64, 16), (127, 127, 8)])
def test_upfirdn2d_channels_last_vs_oracle(built_library, cfg, hwc):
    from multi_stylegan_b200 import _C
    up, down, px0, px1, py0, py1, kh, kw = cfg
    h, w, c = hwc
    if h * up + py0 + py1 < kh or w * up + px0 + px1 < kw:
        pytest.skip("kernel larger than padded input")
    g = torch.Generator().manual_seed(h * 1000 + w + c)
    x = torch.randn(2, h, w, c, generator=g)
    k = torch.randn(kh, kw, generator=g)
    want = ops.upfirdn2d(x, k, up, up, down, down, px0, px1, py0, py1)
    got = _C.upfirdn2d(x.to(dev()), k.to(dev()), up, up, down, down, px0, px1, py0, py1)
    assert got.shape == want.shape
    assert rel_err(got, want) < TOL


def test_upfirdn2d_channels_last_autograd_matches_planar(built_library):
    """The autograd wrapper on a channels-last activation == the same call on the NCHW tensor (fwd, grad, gradgrad)."""
    from multi_stylegan_b200.op_static import upfirdn2d
    torch.manual_seed(0)
    k = torch.tensor([1., 3., 3., 1.], device=dev())
    k = (k[None] * k[:, None]) / 64 * 4
    for (up, down, pad, R) in [(1, 1, (2, 1), 32), (2, 1, (2, 1), 16), (1, 2, (1, 1), 32), (1, 1, (2, 2), 31)]:
        x = torch.randn(2, 64, R, R, device=dev())
        outs = []
        for fmt in (torch.contiguous_format, torch.channels_last):
            xi = x.clone(memory_format=fmt).requires_grad_(True)
            y = upfirdn2d(xi, k, up=up, down=down, pad=pad)
            gy = torch.ones_like(y).requires_grad_(True)
            gx, = torch.autograd.grad(y, xi, gy * 0.5, create_graph=True)
            ggy, = torch.autograd.grad(gx.square().sum(), gy)
            outs.append((y, gx, ggy))
        assert outs[1][0].is_contiguous(memory_format=torch.channels_last)
        for a, b in zip(*outs):
            assert rel_err(b, a) < 1e-5


@pytest.mark.parametrize("shape", [(3, 12, 8, 3), (2, 64, 48, 1), (8, 512, 512, 3), (4, 3, 64, 1)])
@pytest.mark.parametrize("demod", [True, False])
def test_modulate_weights_backward_vs_autograd(built_library, shape, demod):
    """Fused backward of the weight modulation == autograd through the reference's arithmetic (:384-388)."""
    from multi_stylegan_b200 import _C
    from tests import backend_oracle
    B, O, C, k = shape
    g = torch.Generator().manual_seed(7)
    W = torch.randn(O, C, k, k, generator=g)
    s = torch.randn(B, C, generator=g)
    go = torch.randn(B, O, C, k, k, generator=g)
    scale = 0.05
    want_w, want_d = ops.modulate_weights(W, s, scale, demod)
    want_dW, want_ds = backend_oracle.modulate_weights_bwd(go, W, s, want_d, scale, demod)
    got_w, got_d = _C.modulate_weights(W.to(dev()), s.to(dev()), scale, demod)
    assert rel_err(got_w, want_w) < 1e-5
    got_dW, got_ds = _C.modulate_weights_bwd(go.to(dev()), W.to(dev()), s.to(dev()), got_d, scale, demod)
    assert rel_err(got_dW, want_dW) < 1e-4 and rel_err(got_ds, want_ds) < 1e-4


@pytest.mark.parametrize("shape", [(2, 8, 9, 7), (2, 132, 33, 35), (2, 512, 64, 64)])
@pytest.mark.parametrize("pad", [(2, 1), (1, 2), (2, 2)])
def test_blur_noise_bias_act_fused(built_library, shape, pad):
    """Blur + noise + bias + leaky ReLU in the FIR kernel's store == the three reference ops in sequence."""
    from multi_stylegan_b200 import _C
    from tests import backend_oracle
    B, C, H, W = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(shape, generator=g)
    k = torch.tensor([1., 3., 3., 1.])
    k = k[None] * k[:, None] / 16
    oh, ow = H + sum(pad) - 3, W + sum(pad) - 3
    bias = torch.randn(C, generator=g)
    nw = torch.tensor([0.3])
    d = dev()
    for noise in (torch.randn(B, 1, oh, ow, generator=g), torch.randn(1, 1, oh, ow, generator=g), None):
        p4 = (pad[0], pad[1], pad[0], pad[1])
        want = backend_oracle.blur_noise_bias_act(x, k, p4, noise, nw, bias, 0.2, 1.25)
        got = _C.blur_noise_bias_act(x.to(d), k.to(d), p4, None if noise is None else noise.to(d), nw.to(d), bias.to(d), 0.2, 1.25)
        assert got.shape == want.shape and rel_err(got, want) < TOL


def test_kernels_do_not_write_out_of_bounds(built_library):
    """compute-sanitizer is not available on the GPU pool, so the kernels are called through the raw C-ABI with the
    output (and the conv workspace) embedded in NaN-filled guard bands, on ragged sizes; the bands must stay untouched."""
    import ctypes
    from multi_stylegan_b200 import _lib
    L = _lib.lib()
    d = dev()
    G = 4096                                                    # guard floats on each side
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def guarded(n):
        buf = torch.full((n + 2 * G,), float("nan"), device=d)
        return buf, buf[G:G + n]

    def check(buf, n, what):
        torch.cuda.synchronize()
        assert torch.isnan(buf[:G]).all() and torch.isnan(buf[G + n:]).all(), what
        assert not torch.isnan(buf[G:G + n]).any(), what

    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    k = torch.rand(4, 4, device=d)
    for (B, H, W, C) in [(2, 33, 35, 12), (1, 7, 5, 132), (3, 64, 64, 8)]:
        x = torch.randn(B, H, W, C, device=d)
        for (up, down, p0, p1) in [(1, 1, 2, 1), (1, 1, 2, 2), (2, 1, 2, 1), (1, 2, 1, 1), (2, 2, 0, 1)]:
            oh = L.msg_upfirdn2d_out_size(H, up, down, p0, p1, 4)
            ow = L.msg_upfirdn2d_out_size(W, up, down, p0, p1, 4)
            n = B * oh * ow * C
            buf, out = guarded(n)
            rc = L.msg_upfirdn2d(ptr(out), ptr(x), ptr(k), B, H, W, C, 4, 4, up, up, down, down, p0, p1, p0, p1, 0, st)
            assert rc == 0
            check(buf, n, ("fir", B, H, W, C, up, down))
        # fused blur tail
        oh, ow = H + 3 - 3, W + 3 - 3
        n = B * oh * ow * C
        buf, out = guarded(n)
        noise = torch.randn(B, oh, ow, device=d)
        nw = torch.tensor([0.2], device=d)
        bias = torch.randn(C, device=d)
        rc = L.msg_upfirdn2d_bias_act(ptr(out), ptr(x), ptr(k), B, H, W, C, 4, 4, 2, 1, 2, 1, ptr(noise), ptr(nw), oh * ow,
                                      ptr(bias), 1, 0.2, 1.0, st)
        assert rc == 0
        check(buf, n, ("blur+act", B, H, W, C))
        # channel-inner activation forward / backward
        n = B * H * W * C
        buf, out = guarded(n)
        rc = L.msg_noise_bias_act_nhwc(ptr(out), ptr(x), None, None, None, ptr(bias), B * H * W, C, 1, 0.2, 1.0, st)
        assert rc == 0
        check(buf, n, ("nba", B, H, W, C))
    # conv forward / dgrad / wgrad: outputs and workspace in guard bands (tcgen05 engine, ragged shapes)
    for (B, C, O, H, W, kk, s, p, per) in [(2, 36, 40, 19, 23, 3, 1, 1, True), (2, 32, 48, 17, 18, 3, 2, 0, False),
                                           (1, 6, 20, 9, 9, 3, 1, 1, False), (3, 64, 132, 8, 8, 1, 1, 0, False)]:
        dsc = _lib.ConvDesc()
        dsc.B, dsc.C, dsc.H, dsc.W, dsc.O, dsc.kh, dsc.kw = B, C, H, W, O, kk, kk
        dsc.stride_h = dsc.stride_w = s
        dsc.pad_h = dsc.pad_w = p
        dsc.OH, dsc.OW = (H + 2 * p - kk) // s + 1, (W + 2 * p - kk) // s + 1
        dsc.w_batch_stride = O * C * kk * kk if per else 0
        dsc.layout = _lib.LAYOUT_NHWC
        x = torch.randn(B, H, W, C, device=d)
        w = torch.randn((B if per else 1) * O * C * kk * kk, device=d)
        dy = torch.randn(B, dsc.OH, dsc.OW, O, device=d)
        for which, n_out in ((0, B * dsc.OH * dsc.OW * O), (1, B * H * W * C), (2, w.numel())):
            wsn = L.msg_conv2d_workspace(ctypes.byref(dsc), which, _lib.CONV_FORCE_TC)
            wbuf = torch.full((wsn // 4 + 2 * G,), float("nan"), device=d)
            ws = wbuf[G:]
            buf, out = guarded(n_out)
            if which == 0:
                rc = L.msg_conv2d_forward(ptr(out), ptr(x), ptr(w), ctypes.byref(dsc), 1.0, ptr(ws), wsn, _lib.CONV_FORCE_TC, st)
            elif which == 1:
                rc = L.msg_conv2d_dgrad(ptr(out), ptr(dy), ptr(w), ctypes.byref(dsc), 1.0, ptr(ws), wsn, _lib.CONV_FORCE_TC, st)
            else:
                rc = L.msg_conv2d_wgrad(ptr(out), ptr(dy), ptr(x), ctypes.byref(dsc), 1.0, ptr(ws), wsn, _lib.CONV_FORCE_TC, st)
            assert rc == 0, (which, L.msg_last_error())
            check(buf, n_out, ("conv", which, B, C, O, H, W, kk, s))
            torch.cuda.synchronize()
            assert torch.isnan(wbuf[:G]).all() and torch.isnan(wbuf[G + wsn // 4:]).all(), ("conv workspace", which)


def test_reference_extension_module_shims(built_library):
    """`import fused_act_cuda` / `import upfirdn2d_cuda` (what the reference's op_static/fused_act.py:8 and upfirdn2d.py:8
    do) resolve to the shim modules once multi_stylegan_b200/shims is on sys.path; same signatures, same results as the
    oracle restatement of the reference kernels, including the calls the reference's autograd Functions make
    (fused_act.py:31-33,47-49,58; upfirdn2d.py:34-45,121-123)."""
    import importlib
    import sys
    from multi_stylegan_b200 import shims
    from oracle import ops
    sys.path.insert(0, shims.PATH)
    try:
        fused_act_cuda = importlib.import_module("fused_act_cuda")
        upfirdn2d_cuda = importlib.import_module("upfirdn2d_cuda")
    finally:
        sys.path.remove(shims.PATH)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 20, 9, 11, generator=g)
    b = torch.randn(20, generator=g)
    empty = x.new_empty(0)
    out = fused_act_cuda.fused_bias_act(x.cuda(), b.cuda(), empty.cuda(), 3, 0, 0.2, 1.0)                 # forward (:58)
    want = ops.fused_bias_act(x, b, empty, 3, 0, 0.2, 1.0)
    assert rel_err(out, want) < 1e-6
    gy = torch.randn(x.shape, generator=g)
    gi = fused_act_cuda.fused_bias_act(gy.cuda(), empty.cuda(), out, 3, 1, 0.2, 1.0)                      # backward (:31-33)
    assert rel_err(gi, ops.fused_bias_act(gy, empty, want, 3, 1, 0.2, 1.0)) < 1e-6
    gg = fused_act_cuda.fused_bias_act(gy.cuda(), b.cuda(), out, 3, 1, 0.2, 1.0)                          # double backward (:47-49)
    assert rel_err(gg, ops.fused_bias_act(gy, b, want, 3, 1, 0.2, 1.0)) < 1e-6
    k = torch.tensor([1., 3., 3., 1.])
    k = k[None] * k[:, None] / 64
    inp = torch.randn(6, 12, 14, 1, generator=g)
    for cfg in ((1, 1, 1, 1, 2, 1, 2, 1), (2, 2, 1, 1, 2, 1, 2, 1), (1, 1, 2, 2, 1, 1, 1, 1), (1, 1, 1, 1, 2, 2, 2, 2)):
        got = upfirdn2d_cuda.upfirdn2d(inp.cuda(), k.cuda(), *cfg)
        assert rel_err(got, ops.upfirdn2d(inp, k, *cfg)) < 1e-5, cfg


def test_round_2_kernels_do_not_write_out_of_bounds(built_library):
    """The same guard-band check for the kernels added in round 2 (small-M linears, non-local passes, MinibatchStdDev,
    demodulation backward, column sums), on ragged sizes, through the raw C-ABI."""
    import ctypes
    from multi_stylegan_b200 import _lib
    L = _lib.lib()
    d = dev()
    G = 4096
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    bufs = []

    def guarded(n, dtype=torch.float32):
        fill = float("nan") if dtype == torch.float32 else -7
        buf = torch.full((n + 2 * G,), fill, device=d, dtype=dtype)
        bufs.append((buf, n, dtype))
        return buf[G:G + n]

    def check(what):
        torch.cuda.synchronize()
        for buf, n, dtype in bufs:
            if dtype == torch.float32:
                assert torch.isnan(buf[:G]).all() and torch.isnan(buf[G + n:]).all(), what
                assert not torch.isnan(buf[G:G + n]).any(), what
            else:
                assert (buf[:G] == -7).all() and (buf[G + n:] == -7).all(), what
        bufs.clear()

    # mapping network: K = 20 (column slices of 3, 3, ..., 2 / empty), M = 19 rows (three forward clusters, two backward chunks)
    depth, M, K = 3, 19, 20
    Ws = [torch.randn(K, K, device=d) for _ in range(depth)]
    bs = [torch.randn(K, device=d) for _ in range(depth)]
    warr = (ctypes.c_void_p * depth)(*[w.data_ptr() for w in Ws])
    barr = (ctypes.c_void_p * depth)(*[b.data_ptr() for b in bs])
    z = torch.randn(M, K, device=d)
    acts, x0 = guarded(depth * M * K), guarded(M * K)
    assert L.msg_style_mapping_forward(ptr(acts), ptr(x0), ptr(z), warr, barr, depth, M, K, 0.3, 0.2, 1.0, 1e-8, st) == 0
    check("style_mapping_forward")
    dW, db, work = guarded(depth * K * K), guarded(depth * K), guarded(depth * M * K)
    gy = torch.randn(M, K, device=d)
    assert L.msg_style_mapping_backward(ptr(dW), ptr(db), ptr(gy), ptr(acts), ptr(x0), warr, barr, depth, M, K, 0.3, 0.2, 1.0,
                                        ptr(work), st) == 0
    check("style_mapping_backward")

    # grouped linears: N = 12 / 20 / 4, K = 8 / 12, M = 19
    specs = [(12, 8, 0), (20, 8, 0), (4, 12, 8)]                 # (N, K, in_off); input row = 20 floats
    items = (_lib.LinearItem * len(specs))()
    keep, out_off, w_off, b_off = [], 0, 0, 0
    for i, (N, Kk, off) in enumerate(specs):
        W, b = torch.randn(N, Kk, device=d), torch.randn(N, device=d)
        keep += [W, b]
        it = items[i]
        it.W, it.bias, it.N, it.K, it.in_off, it.out_off, it.w_off, it.b_off = W.data_ptr(), b.data_ptr(), N, Kk, off, out_off, w_off, b_off
        it.alpha, it.beta = 0.5, 2.0
        out_off, w_off, b_off = out_off + N, w_off + N * Kk, b_off + N
    slots = (_lib.LinearSlot * 2)()
    slots[0].in_off, slots[0].K, slots[0].first, slots[0].count = 0, 8, 0, 2
    slots[1].in_off, slots[1].K, slots[1].first, slots[1].count = 8, 12, 2, 1
    x = torch.randn(M, 20, device=d)
    out = guarded(M * out_off)
    assert L.msg_linear_group_forward(ptr(out), ptr(x), 20, items, 3, M, 20, 12, st) == 0
    check("linear_group_forward")
    gout = torch.randn(M * out_off, device=d)
    dW, db, din = guarded(w_off), guarded(b_off), guarded(M * 20)
    assert L.msg_linear_group_backward(ptr(dW), ptr(db), ptr(din), ptr(gout), ptr(x), 20, items, 3, slots, 2, M, 20, 12, st) == 0
    check("linear_group_backward")

    # non-local passes on an odd map
    B, H, W, cq, cv = 2, 9, 7, 8, 12
    qkv = torch.randn(B, H, W, 2 * cq + cv, device=d)
    theta, phi, gp = guarded(B * H * W * cq), guarded(B * (H // 2) * (W // 2) * cq), guarded(B * (H // 2) * (W // 2) * cv)
    idx = guarded(B * (H // 2) * (W // 2) * (cq + cv) // 4, torch.int32)
    assert L.msg_nl_split_pool(ptr(theta), ptr(phi), ptr(gp), ptr(idx), ptr(qkv), B, H, W, cq, cv, st) == 0
    check("nl_split_pool")
    dq = guarded(B * H * W * (2 * cq + cv))
    assert L.msg_nl_merge_unpool(ptr(dq), ptr(theta), ptr(phi), ptr(gp), ptr(idx), B, H, W, cq, cv, st) == 0
    check("nl_merge_unpool")
    for n in (12, 516, 1028, 2052):
        rows = 13
        s = guarded(rows * n)
        s.copy_(torch.randn(rows * n, device=d))
        assert L.msg_softmax_rows(ptr(s), rows, n, st) == 0
        dp = guarded(rows * n)
        dp.copy_(torch.randn(rows * n, device=d))
        assert L.msg_softmax_rows_bwd(ptr(dp), ptr(s), rows, n, st) == 0
        check(("softmax_rows", n))

    # MinibatchStdDev, groups = 3 of 2 samples, C = 5 (odd), HW = 35
    B, C, HW, groups = 6, 5, 35, 3
    x = torch.randn(B * HW * C, device=d)
    nws = L.msg_mbstd_workspace(groups)
    out, ws = guarded(B * HW * (C + 1)), guarded(nws // 4 + 64)
    assert L.msg_mbstd_forward(ptr(out), ptr(x), B, C, HW, groups, 1e-8, ptr(ws), nws, st) == 0
    ws.fill_(0)
    check("mbstd_forward")
    gx, ws = guarded(B * HW * C), guarded(nws // 4 + 64)
    go = torch.randn(B * HW * (C + 1), device=d)
    assert L.msg_mbstd_backward(ptr(gx), ptr(go), ptr(x), B, C, HW, groups, 1e-8, ptr(ws), nws, st) == 0
    ws.fill_(0)
    check("mbstd_backward")

    # demodulation backward and the deterministic reductions
    B, O, C, taps = 5, 12, 20, 9
    Wt, s_, gd, dd, wsq = (torch.randn(O * C * taps, device=d), torch.randn(B * C, device=d), torch.randn(B * O, device=d),
                           torch.rand(B * O, device=d), torch.rand(O * C, device=d))
    dW, ds = guarded(O * C * taps), guarded(B * C)
    assert L.msg_demod_factors_bwd(ptr(dW), ptr(ds), ptr(gd), ptr(dd), ptr(s_), ptr(wsq), ptr(Wt), B, O, C, taps, 0.1, st) == 0
    check("demod_factors_bwd")
    rows, C = 1000, 36
    x = torch.randn(rows * C, device=d)
    nws = L.msg_colsum_workspace(rows, C)
    out, ws = guarded(C), guarded(nws // 4 + 64)
    assert L.msg_colsum_nhwc(ptr(out), ptr(x), rows, C, 1.0, ptr(ws), nws, st) == 0
    ws.fill_(0)
    check("colsum_nhwc")
    out, ws = guarded(1), guarded(4096)
    assert L.msg_dot(ptr(out), ptr(x), ptr(x), rows * C, 1.0, ptr(ws), st) == 0
    ws.fill_(0)
    check("dot")
