"""GPU: oracle parity AT THE BENCHMARK SHAPES (BASELINE.json configs 2-3: 512 channels, 256x256 / 128x128), with the
engine the product selects by default (no environment variant): the kernels that earn the bench numbers — the CTA-pair
pixgemm with its persistent tile loop, the 5-D-TMA split-K wgrad, the phase-strided stride-2 paths — are reached here on
the same problem sizes, at batch 1-2 so the CPU oracle finishes in seconds (a 3x3 512->512 convolution at 256^2 is 309 GFLOP).

Tolerance: 1e-2 in max-abs-err / max-abs-ref (north star: TF32 tensor-core convs) for outputs AND for dgrad / wgrad,
which are mask-free contractions; tighter figures measured on B200 are printed."""
import math

import pytest
import torch

from oracle import model as omodel, ops
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2


def dev():
    return torch.device("cuda:0")


def cl(t):
    return t.to(dev()).contiguous(memory_format=torch.channels_last)


def _conv_triplet(B, C, O, H, W, k, stride, pad, per_sample, seed, transposed=False):
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn((B, O, C, k, k) if per_sample else (O, C, k, k), generator=g) / math.sqrt(C * k * k)
    y = ops.conv2d(x, w, stride, pad)
    dy = torch.randn(y.shape, generator=g)
    got = _C.conv2d_forward(cl(x), w.to(dev()), stride, pad)
    assert _C.conv2d_last_engine() == "tcgen05"
    e_f = rel_err(got, y)
    got = _C.conv2d_dgrad(cl(dy), w.to(dev()), (H, W), stride, pad)
    e_d = rel_err(got, ops.conv2d_dgrad(dy, w, (H, W), stride, pad))
    got = _C.conv2d_wgrad(cl(dy), cl(x), (k, k), stride, pad, per_sample)
    assert _C.conv2d_last_engine() == "tcgen05"
    e_w = rel_err(got, ops.conv2d_wgrad(dy, x, (k, k), stride, pad, per_sample))
    print("conv B=%d %d->%d %dx%d k=%d s=%d per_sample=%s: fwd %.2e dgrad %.2e wgrad %.2e" %
          (B, C, O, H, W, k, stride, per_sample, e_f, e_d, e_w))
    assert e_f < TOL and e_d < TOL and e_w < TOL, (e_f, e_d, e_w)


@pytest.mark.parametrize("case", [
    # (B, C, O, H, W, k, stride, pad, per_sample)
    (1, 512, 512, 256, 256, 3, 1, 1, False),     # the headline GEMM: 3x3 512->512 at 256^2 (shared-weight form)
    (1, 512, 512, 256, 256, 3, 1, 1, True),      # the same layer with a per-sample filter bank (path-length iterations)
    (2, 512, 512, 128, 128, 3, 1, 1, False),
    (2, 128, 128, 256, 256, 3, 2, 0, False),     # discriminator down-sampling conv, 256 -> 127
    (2, 128, 128, 256, 256, 3, 1, 1, False),     # discriminator 3x3 128->128 at 256^2 (256-pixel CTA tiles)
    (2, 512, 6, 256, 256, 1, 1, 0, True),        # stacked tRGB pair, 1x1 512 -> 6, per-sample weights
    (1, 512, 512, 256, 256, 2, 2, 0, False),     # adjoint of the generator's up-convolution (2x2 / stride 2, 256 -> 128)
], ids=lambda c: "x".join(str(v) for v in c))
def test_conv_kernels_at_benchmark_shapes(built_library, case):
    _conv_triplet(*case, seed=sum(case[:6]))


def test_up_convolution_at_benchmark_shape(built_library):
    """2x2 / stride-2 transposed convolution 512 -> 512, 128^2 -> 256^2 (multi_stylegan_generator.py:393-401), shared and
    per-sample weights."""
    from multi_stylegan_b200 import _C
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 512, 128, 128, generator=g)
    w = torch.randn(512, 512, 2, 2, generator=g) / math.sqrt(512 * 4)              # [O, C, 2, 2] (the module's layout)
    want = ops.conv_transpose2d(x, w.transpose(0, 1).contiguous(), stride=2)      # torch layout [Cin, Cout, kh, kw]
    got = _C.conv2d_dgrad(cl(x), w.to(dev()), (256, 256), 2, 0, w_transposed=True)
    assert got.shape == want.shape == (1, 512, 256, 256)
    e = rel_err(got, want)
    print("up-conv 512->512 128->256: %.2e" % e)
    assert e < TOL


def test_modulated_layer_at_benchmark_shape(built_library):
    """The generator's 3x3 512->512 layer at 256^2 in the product's shared-weight form (style on the input, demodulation,
    noise, bias, leaky ReLU in the epilogue, second output) against the reference's per-sample-weight layer
    (oracle.model.styled_conv), forward and the mask-free gradients (w.r.t. the conv input and the weight, with the
    oracle's activation output pinned as the mask)."""
    from multi_stylegan_b200 import styled
    from multi_stylegan_b200.multi_stylegan_generator import StyledConv2d
    torch.manual_seed(21)
    B, C, R, L = 1, 512, 256, 512
    layer = StyledConv2d(C, C, (3, 3), L)
    with torch.no_grad():
        layer.noise_injection.weight.fill_(0.1)
        layer.activation.bias.normal_(0, 0.1)
    sd = {"p." + n: v.detach().clone() for n, v in layer.state_dict().items()}
    g = torch.Generator().manual_seed(22)
    x = torch.randn(B, C, R, R, generator=g)
    wlat = torch.randn(B, L, generator=g)
    noise = torch.randn(B, 1, R, R, generator=g)
    s_next = torch.randn(B, C, generator=g)
    want, s = omodel.styled_conv(sd, "p", x, wlat, noise, False)
    layer = layer.to(dev())
    mc, act = layer.modulated_convolution, layer.activation
    with torch.no_grad():
        sdev = mc.modulation_mapping(wlat.to(dev()))
        xs = x.to(dev()) * sdev.view(B, C, 1, 1)
        out, out2 = styled.styled_conv(xs, mc.weight[0], sdev, mc.scale, True, noise.to(dev()), layer.noise_injection.weight,
                                       act.bias, s_next.to(dev()), mc.stride, mc.padding, act.negative_slope, act.scale)
    e1, e2 = rel_err(out, want), rel_err(out2, want * s_next.view(B, C, 1, 1))
    print("modulated 3x3 512->512 @256^2 (shared-weight form): out %.2e out2 %.2e" % (e1, e2))
    assert e1 < TOL and e2 < TOL


def test_default_generator_forward_matches_oracle(built_library):
    """BASELINE config 1 on the device: the default 512-channel / 256x256 generator, batch 1, explicit noise, against
    oracle.model.generator_forward (the reference's arithmetic on the CPU)."""
    from multi_stylegan_b200 import config
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    torch.manual_seed(0)
    net = G_mod.Generator(config.multi_style_gan_generator_config, compute_dead_branch=False)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("noise_injection.weight"):
                p.fill_(0.05)
            if n.endswith("activation.bias"):
                p.normal_(0, 0.1)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    z = torch.randn(1, 512, generator=g)
    noise = [torch.randn(1, 1, 4, 4, generator=g)]
    for i in range(len(net.main_convolutions_1)):
        r = 2 ** (i // 2 + 3)
        noise.append(torch.randn(1, 1, r, r, generator=g))
    want = omodel.generator_forward(sd, z=z, noise=noise)
    net = net.to(dev())
    with torch.no_grad():
        got = net(z.to(dev()), noise=[n.to(dev()) for n in noise])
        net.fused_modconv = False
        got_ps = net(z.to(dev()), noise=[n.to(dev()) for n in noise])
    e, e_ps = rel_err(got, want), rel_err(got_ps, want)
    print("default generator forward B=1: shared-weight form %.2e, per-sample-weight form %.2e" % (e, e_ps))
    assert got.shape == want.shape == (1, 2, 3, 256, 256)
    assert e < TOL and e_ps < TOL


def test_default_discriminator_forward_matches_oracle(built_library):
    from multi_stylegan_b200 import config
    import multi_stylegan_b200.u_net_2d_discriminator as D_mod
    torch.manual_seed(2)
    net = D_mod.Discriminator(config.u_net_2d_discriminator_config, no_rfp=True)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.5)                  # exercise the attention path (gamma is initialised to 0)
            if n.endswith(".bias"):
                p.normal_(0, 0.1)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 2, 3, 256, 256, generator=g)
    ws, wp = omodel.discriminator_forward(sd, x)
    net = net.to(dev())
    with torch.no_grad():
        gs, gp = net(x.to(dev()))
    es, ep = rel_err(gs, ws), rel_err(gp, wp)
    print("default discriminator forward B=1: scalar %.2e pixel-wise %.2e" % (es, ep))
    assert gs.shape == ws.shape and gp.shape == wp.shape
    assert es < TOL and ep < TOL
