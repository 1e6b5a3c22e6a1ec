"""GPU: the kernels of the shared-weight modulated convolution (include/msg_b200.h: msg_conv_epilogue.col_scale / y2,
msg_upfirdn2d_bias_act_mod, msg_styled_act_bwd, msg_demod_factors) and the fused layer Functions built on them
(multi_stylegan_b200/styled.py) against the CPU oracle — the reference's per-sample-weight arithmetic
(multi_stylegan_generator.py:379-411, oracle/model.py) on the same seeded inputs."""
import math

import pytest
import torch

from oracle import model as omodel, ops
from tests import backend_oracle
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def cl(t):
    return t.to(dev()).contiguous(memory_format=torch.channels_last)


# (B, C, O, H, W, k): TMA-store epilogue (O % 32 == 0 ...), CTA pairs (O = 256 with enough tiles), narrow / odd tiles
EPI_SHAPES = [
    (2, 32, 64, 32, 32, 3),
    (3, 64, 128, 16, 16, 3),
    (2, 64, 256, 64, 64, 3),      # BN = 256: CTA-pair kernel
    (2, 16, 16, 8, 8, 3),         # BN = 16: generic store path
    (2, 32, 48, 20, 20, 3),       # channel tail inside a 32-wide chunk
    (1, 64, 64, 4, 4, 3),         # generator start block
    (2, 32, 36, 12, 12, 1),       # N % 32 != 0, 1x1
]


@pytest.mark.parametrize("shape", EPI_SHAPES)
@pytest.mark.parametrize("engine", ["tc", "simt"])
@pytest.mark.parametrize("shared_scales", [False, True])
def test_conv_epilogue_col_scale_and_second_output(built_library, shape, engine, shared_scales):
    from multi_stylegan_b200 import _C, _lib
    if engine == "tc" and not _C.tensor_core_path_available():
        pytest.skip("not an sm_100 device")
    B, C, O, H, W, k = shape
    g = torch.Generator().manual_seed(B * 1000 + O)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(O, C, k, k, generator=g) / math.sqrt(C * k * k)
    nb = 1 if shared_scales else B
    d = torch.rand(nb, O, generator=g) + 0.5
    s2 = torch.randn(nb, O, generator=g)
    noise = torch.randn(B, 1, H, W, generator=g)
    nw = torch.tensor([0.3])
    bias = torch.randn(O, generator=g) * 0.2
    want, want2 = backend_oracle.conv2d_forward(x, w, 1, k // 2, alpha=0.7, bias=bias, noise=noise, noise_w=nw, act=True,
                                                slope=0.2, gain=1.3, col_scale=d, out2_scale=s2)
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_TC if engine == "tc" else _lib.CONV_FORCE_SIMT
    tol = 1e-2 if engine == "tc" else 1e-4
    try:
        got, got2 = _C.conv2d_forward(cl(x), w.to(dev()), 1, k // 2, alpha=0.7, bias=bias.to(dev()), noise=noise.to(dev()),
                                      noise_w=nw.to(dev()), act=True, slope=0.2, gain=1.3, col_scale=d.to(dev()),
                                      out2_scale=s2.to(dev()))
        assert _C.conv2d_last_engine() == ("tcgen05" if engine == "tc" else "simt")
        assert rel_err(got, want) < tol and rel_err(got2, want2) < tol, (rel_err(got, want), rel_err(got2, want2))
        # the second output is exactly the first times the scale (same registers, one extra multiply)
        assert torch.equal(got2, got * s2.to(dev()).reshape(-1, O, 1, 1).expand(B, O, 1, 1))
        # col_scale alone, linear epilogue
        only = _C.conv2d_forward(cl(x), w.to(dev()), 1, k // 2, alpha=0.7, col_scale=d.to(dev()))
        ref = backend_oracle.conv2d_forward(x, w, 1, k // 2, alpha=0.7, col_scale=d)
        assert rel_err(only, ref) < tol
    finally:
        _C.conv_flags = old


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 16, 16), (3, 64, 33, 20), (1, 512, 8, 8), (2, 128, 64, 64)])
@pytest.mark.parametrize("which", ["both", "g_out", "g_out2"])
def test_styled_act_bwd_matches_oracle(built_library, B, C, H, W, which):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(C + H)
    out = torch.randn(B, C, H, W, generator=g)
    g1 = torch.randn(B, C, H, W, generator=g) if which != "g_out2" else None
    g2 = torch.randn(B, C, H, W, generator=g) if which != "g_out" else None
    d = torch.rand(B, C, generator=g) + 0.5
    s2 = torch.randn(B, C, generator=g)
    noise = torch.randn(B, 1, H, W, generator=g)
    want_g, want_s = backend_oracle.styled_act_bwd(g1, g2, out, d, s2, noise, 0.2, 1.3)
    got_g, got_s = _C.styled_act_bwd(None if g1 is None else cl(g1), None if g2 is None else cl(g2), cl(out), d.to(dev()),
                                     s2.to(dev()), noise.to(dev()), 0.2, 1.3)
    assert rel_err(got_g, want_g) < 1e-5
    for k in range(4):
        assert rel_err(got_s[k], want_s[k]) < 1e-4, (k, rel_err(got_s[k], want_s[k]))
    # shared noise map, no demodulation factor
    want_g, want_s = backend_oracle.styled_act_bwd(g1, g2, out, None, s2, noise[:1], 0.2, 1.0)
    got_g, got_s = _C.styled_act_bwd(None if g1 is None else cl(g1), None if g2 is None else cl(g2), cl(out), None,
                                     s2.to(dev()), noise[:1].to(dev()), 0.2, 1.0)
    assert rel_err(got_g, want_g) < 1e-5
    for k in range(4):
        assert rel_err(got_s[k], want_s[k]) < 1e-4


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 16, 16), (2, 64, 64, 64), (3, 16, 9, 12)])
def test_blur_mod_matches_oracle(built_library, B, C, H, W):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(C * 7 + W)
    x = torch.randn(B, C, H, W, generator=g)
    k = torch.tensor([1., 3., 3., 1.])
    k = k[None] * k[:, None] / 16
    d = torch.rand(B, C, generator=g) + 0.5
    s2 = torch.randn(B, C, generator=g)
    noise = torch.randn(B, 1, H, W, generator=g)
    nw = torch.tensor([0.25])
    bias = torch.randn(C, generator=g) * 0.3
    pad = (2, 1, 2, 1)
    want, want2 = backend_oracle.blur_noise_bias_act_mod(x, k, pad, d, noise, nw, bias, 0.2, 1.1, s2)
    got, got2 = _C.blur_noise_bias_act_mod(cl(x), k.to(dev()), pad, d.to(dev()), noise.to(dev()), nw.to(dev()),
                                           bias.to(dev()), 0.2, 1.1, s2.to(dev()))
    assert rel_err(got, want) < 1e-5 and rel_err(got2, want2) < 1e-5
    got, none = _C.blur_noise_bias_act_mod(cl(x), k.to(dev()), pad, None, noise.to(dev()), nw.to(dev()), bias.to(dev()),
                                           0.2, 1.1, None)
    want, _ = backend_oracle.blur_noise_bias_act_mod(x, k, pad, None, noise, nw, bias, 0.2, 1.1, None)
    assert none is None and rel_err(got, want) < 1e-5


def test_demod_factors_match_reference_formula(built_library):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(5)
    W = torch.randn(96, 64, 3, 3, generator=g)
    s = torch.randn(4, 64, generator=g)
    scale = math.sqrt(2) / math.sqrt(64 * 9)
    # reference :384-388: rsqrt(sum_{c,kh,kw} (scale * W * s)^2 + 1e-8)
    wm = scale * W.unsqueeze(0) * s.view(4, 1, 64, 1, 1)
    want = torch.rsqrt(wm.pow(2).sum(dim=(2, 3, 4)) + 1e-8)
    d, wsq = _C.demod_factors(W.to(dev()), s.to(dev()), scale)
    assert rel_err(d, want) < 1e-5 and rel_err(wsq, W.pow(2).sum(dim=(2, 3))) < 1e-5


def _styled_pair(up, C, O, L, seed):
    from multi_stylegan_b200.multi_stylegan_generator import StyledConv2d
    torch.manual_seed(seed)
    a = StyledConv2d(C, O, (2, 2) if up else (3, 3), L, upsampling=up)
    with torch.no_grad():
        a.noise_injection.weight.fill_(0.2)
        a.activation.bias.normal_(0, 0.2)
    return a


@pytest.mark.parametrize("up", [False, True])
@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_fused_layer_matches_reference_layer(built_library, up, engine):
    """One StyledConv2d in the shared-weight form (style on the input, demodulation / noise / bias / leaky ReLU in the
    epilogue, second output) against oracle.model.styled_conv (the reference's per-sample-weight layer), forward and all
    first-order gradients."""
    from multi_stylegan_b200 import _C, _lib, styled
    if engine == "tc" and not _C.tensor_core_path_available():
        pytest.skip("not an sm_100 device")
    B, C, O, L, R = 2, 32, 64, 24, 16
    layer = _styled_pair(up, C, O, L, 3)
    sd = {"p." + n: v.detach().clone() for n, v in layer.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, C, R, R, generator=g)
    wlat = torch.randn(B, L, generator=g)
    Ro = 2 * R if up else R
    noise = torch.randn(B, 1, Ro, Ro, generator=g)
    s_next = torch.randn(B, O, generator=g)
    go = torch.randn(B, O, Ro, Ro, generator=g)
    go2 = torch.randn(B, O, Ro, Ro, generator=g)

    # oracle (CPU, fp32): reference arithmetic; second output = out * s_next
    xo = x.clone().requires_grad_(True)
    wo = wlat.clone().requires_grad_(True)
    so = s_next.clone().requires_grad_(True)
    po = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "blur" not in k}
    sdo = dict(sd)
    sdo.update(po)
    out_o, _ = omodel.styled_conv(sdo, "p", xo, wo, noise, up)
    loss_o = (out_o * go).sum() + (out_o * so.view(B, O, 1, 1) * go2).sum()
    names = sorted(po)
    grads_o = torch.autograd.grad(loss_o, [xo, wo, so] + [po[n] for n in names])

    layer = layer.to(dev())
    mc, act = layer.modulated_convolution, layer.activation
    xd = x.to(dev()).requires_grad_(True)
    wd = wlat.to(dev()).requires_grad_(True)
    sn = s_next.to(dev()).requires_grad_(True)
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_TC if engine == "tc" else _lib.CONV_FORCE_SIMT
    try:
        s = mc.modulation_mapping(wd)
        xs = xd * s.view(B, C, 1, 1)
        if up:
            out, out2 = styled.styled_up_conv(xs, mc.weight[0], s, mc.scale, True, mc.blur.kernel, mc.blur.padding,
                                              noise.to(dev()), layer.noise_injection.weight, act.bias, sn, mc.stride,
                                              mc.padding, act.negative_slope, act.scale)
        else:
            out, out2 = styled.styled_conv(xs, mc.weight[0], s, mc.scale, True, noise.to(dev()),
                                           layer.noise_injection.weight, act.bias, sn, mc.stride, mc.padding,
                                           act.negative_slope, act.scale)
        params = dict(layer.named_parameters())
        grads = torch.autograd.grad((out * go.to(dev())).sum() + (out2 * go2.to(dev())).sum(),
                                    [xd, wd, sn] + [params[n[2:]] for n in names])
    finally:
        _C.conv_flags = old
    tol = 1e-2 if engine == "tc" else 2e-4
    assert rel_err(out, out_o) < tol, rel_err(out, out_o)
    assert rel_err(out2, out_o * s_next.view(B, O, 1, 1)) < tol
    for name, got, want in zip(["x", "w", "s_next"] + names, grads, grads_o):
        if engine == "tc":
            # leaky-ReLU masks flip where a pre-activation is within the TF32 error of zero (see test_host_logic.GradCheck)
            err = ((got.detach().cpu().double() - want.double()).norm() / want.double().norm().clamp_min(1e-12)).item()
            assert err < 5e-2, (name, err)
        else:
            assert rel_err(got, want) < tol, (name, rel_err(got, want))


@pytest.mark.parametrize("shape", [(2, 32, 64, 32, 32, 3, 1, 1), (2, 64, 128, 16, 16, 1, 1, 0), (2, 32, 32, 64, 64, 3, 2, 0),
                                   (2, 48, 40, 32, 32, 2, 2, 0), (2, 6, 32, 64, 64, 3, 1, 1), (1, 256, 256, 64, 64, 3, 1, 1)])
@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_dgrad_accumulate_operand(built_library, shape, engine):
    """msg_conv2d_dgrad_acc: dx = alpha * conv^T(dy, w) + add, the sum formed in the kernel's epilogue (stride 1 and the
    phase-scattered stride-2 form; C = 6 takes the unaligned store path)."""
    from multi_stylegan_b200 import _C, _lib
    if engine == "tc" and not _C.tensor_core_path_available():
        pytest.skip("not an sm_100 device")
    B, C, O, H, W, k, s, p = shape
    g = torch.Generator().manual_seed(C + O)
    w = torch.randn(O, C, k, k, generator=g) / math.sqrt(C * k * k)
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    dy = torch.randn(B, O, OH, OW, generator=g)
    add = torch.randn(B, C, H, W, generator=g)
    want = ops.conv2d_dgrad(dy, w, (H, W), s, p) * 0.6 + add
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_TC if engine == "tc" else _lib.CONV_FORCE_SIMT
    try:
        got = _C.conv2d_dgrad(cl(dy), w.to(dev()), (H, W), s, p, alpha=0.6, add=cl(add))
    finally:
        _C.conv_flags = old
    assert rel_err(got, want) < (1e-2 if engine == "tc" else 1e-4), rel_err(got, want)


def test_demod_backward_colsum_and_dot_kernels(built_library):
    """msg_demod_factors_bwd / msg_colsum_nhwc / msg_dot against the tensor-op formulas autograd used before them."""
    from multi_stylegan_b200 import _C
    from tests import backend_oracle
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(4)
    for B, O, C, k in ((8, 512, 512, 3), (3, 12, 20, 1), (16, 64, 96, 2)):
        W = torch.randn(O, C, k, k, generator=g)
        s = torch.randn(B, C, generator=g)
        gd = torch.randn(B, O, generator=g)
        scale = 1.0 / (C * k * k) ** 0.5
        d, wsq = backend_oracle.demod_factors(W, s, scale)
        want_w, want_s = backend_oracle.demod_factors_bwd(gd, d, s, wsq, W, scale)
        got_w, got_s = _C.demod_factors_bwd(gd.to(dev), d.to(dev), s.to(dev), wsq.to(dev), W.to(dev), scale)
        assert rel_err(got_w, want_w) < 1e-5 and rel_err(got_s, want_s) < 1e-5
        only_s = _C.demod_factors_bwd(gd.to(dev), d.to(dev), s.to(dev), wsq.to(dev), W.to(dev), scale, need_w=False)
        assert only_s[0] is None and torch.equal(only_s[1], got_s)
    for shape in ((16, 128, 127, 127), (2, 4, 3, 5), (1, 36, 1, 1)):
        x = torch.randn(shape, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
        got = _C.colsum_cl(x, 0.5)
        want = x.double().sum((0, 2, 3)) * 0.5
        assert ((got.double() - want).abs().max() / want.abs().max().clamp_min(1e-6)) < 1e-5
        assert torch.equal(got, _C.colsum_cl(x, 0.5))
        y = torch.randn(shape, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
        got = _C.dot(x, y, 2.0)
        want = (x.double() * y.double()).sum() * 2.0
        assert abs(float(got) - float(want)) < 1e-4 * (x.double() * y.double()).abs().sum().item()
