"""GPU: the fused small-M linears (csrc/linear_ops.cu) against the module-by-module torch formulation they replace
(EqualizedLinear equalized_layer.py:210-254, PixelwiseNormalization :257-277, StyleMapping
multi_stylegan_generator.py:208-235, FusedLeakyReLU op_static/fused_act.py:76-85).  fp32 FMA on both sides: 1e-5."""
import pytest
import torch

from tests.conftest import rel_err

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("case", [(512, 8, 16), (512, 8, 8), (16, 2, 5), (64, 3, 37), (128, 1, 1)],
                         ids=lambda c: "K%d-depth%d-M%d" % c)
def test_style_mapping_one_launch(built_library, case):
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    K, depth, M = case
    torch.manual_seed(K + depth + M)
    net = G_mod.StyleMapping(latent_dimensions=K, depth=depth).to(dev())
    with torch.no_grad():
        for m in net.modules():
            if hasattr(m, "bias") and m.bias is not None:
                m.bias.normal_(0, 0.3)
    z = torch.randn(M, K, device=dev())
    gy = torch.randn(M, K, device=dev())
    assert net._fused_eligible(z)
    got = net(z)
    got.backward(gy)
    g_fused = [p.grad.clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    want = net.layers(z)                       # the module-by-module formulation
    want.backward(gy)
    g_ref = [p.grad.clone() for p in net.parameters()]
    assert rel_err(got, want) < 1e-5, rel_err(got, want)
    for (n, _), a, b in zip(net.named_parameters(), g_fused, g_ref):
        assert rel_err(a, b) < 1e-4, (n, rel_err(a, b))
    # bit-reproducible
    for p in net.parameters():
        p.grad = None
    again = net(z)
    again.backward(gy)
    assert torch.equal(again, got)
    for p, a in zip(net.parameters(), g_fused):
        assert torch.equal(p.grad, a)


@pytest.mark.parametrize("M", [8, 19])
def test_grouped_style_linears(built_library, M):
    from multi_stylegan_b200 import equalized_layer, linear
    torch.manual_seed(M)
    L, J = 64, 5
    dims = [(64, 0), (32, 1), (48, 1), (64, 3), (128, 4), (4, 4)]          # (out features, latent index); slot 2 unused
    mods = [equalized_layer.EqualizedLinear(L, n, bias=(i != 2)).to(dev()) for i, (n, _) in enumerate(dims)]
    with torch.no_grad():
        for m in mods:
            if m.bias is not None:
                m.bias.normal_(1.0, 0.2)
    latent = torch.randn(M, J, L, device=dev(), requires_grad=True)
    gouts = [torch.randn(M, n, device=dev()) for n, _ in dims]
    specs = [(m.weight, m.bias, j * L, m.scale, m.scale_bias) for m, (_, j) in zip(mods, dims)]
    outs = linear.style_linears(latent.reshape(M, -1), specs)
    assert all(o.is_contiguous() for o in outs)
    torch.autograd.backward(outs, gouts)
    got = [o.detach().clone() for o in outs]
    g_lat = latent.grad.clone()
    g_par = [(m.weight.grad.clone(), None if m.bias is None else m.bias.grad.clone()) for m in mods]
    latent.grad = None
    for m in mods:
        m.zero_grad(set_to_none=True)
    ref = [m(latent[:, j]) for m, (_, j) in zip(mods, dims)]
    torch.autograd.backward(ref, gouts)
    for a, b in zip(got, ref):
        assert rel_err(a, b) < 1e-5
    assert rel_err(g_lat, latent.grad) < 1e-4
    assert torch.all(g_lat[:, 2] == 0)                                     # a latent nobody reads gets a zero gradient
    for m, (gw, gb) in zip(mods, g_par):
        assert rel_err(gw, m.weight.grad) < 1e-4
        if gb is not None:
            assert rel_err(gb, m.bias.grad) < 1e-4


def test_generator_styles_in_one_launch_match_per_module_path(built_library):
    """The synthesis network with all style linears / the mapping network fused vs the same network with the fused
    small-M kernels switched off (both in the shared-weight conv form): forward and every parameter gradient."""
    from multi_stylegan_b200 import config, linear
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    torch.manual_seed(5)
    g_cfg, _ = config.scaled_configs(channel_div=16, g_stages=4)
    net = G_mod.Generator(g_cfg, compute_dead_branch=False).to(dev())
    z = [torch.randn(3, net.latent_dimensions, device=dev()) for _ in range(2)]
    torch.manual_seed(9)
    img = net(z, inject_index=3, randomize_noise=False)
    gi = torch.randn_like(img)
    img.backward(gi)
    g1 = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    net.zero_grad(set_to_none=True)
    import unittest.mock as um
    with um.patch.object(G_mod.StyleMapping, "_fused_eligible", lambda self, n: False), \
            um.patch.object(G_mod.Generator, "_all_styles", lambda self, latent: {}):
        ref = net(z, inject_index=3, randomize_noise=False)
        ref.backward(gi)
    g2 = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    assert rel_err(img, ref) < 5e-3, rel_err(img, ref)       # the convolutions round their operands to TF32
    assert set(g1) == set(g2)
    for n in g1:
        assert rel_err(g1[n], g2[n]) < 2e-2, (n, rel_err(g1[n], g2[n]))
