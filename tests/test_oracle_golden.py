"""CPU: pin the oracle (oracle/) against (a) hand-derived known answers, (b) fixtures dumped from the
unmodified reference (tests/golden, made by oracle/make_golden.py) and (c) the reference itself when it
is mounted.  The reference ships no tests of its own."""
import math

import pytest
import torch

from oracle import model, ops, ref_loader
from tests.conftest import load_golden, rel_err


# ---- (a) known answers derived by hand from fused_bias_act_kernel.cu:25-48 ---------------------------
def test_fused_bias_act_known_answers():
    x = torch.tensor([[[-1.0, 2.0], [0.5, -4.0]]])           # [1, 2 channels, 2]
    b = torch.tensor([0.5, -1.0])
    e = torch.empty(0)
    # act=3 grad=0: (x+b > 0 ? x+b : 0.2(x+b)) * 2
    y = ops.fused_bias_act(x, b, e, 3, 0, 0.2, 2.0)
    assert torch.allclose(y, torch.tensor([[[-0.2, 5.0], [-0.2, -2.0]]]))
    # act=3 grad=1: mask from ref, no bias
    ref = torch.tensor([[[1.0, -1.0], [0.0, 3.0]]])
    y = ops.fused_bias_act(x, e, ref, 3, 1, 0.2, 1.0)
    assert torch.allclose(y, torch.tensor([[[-1.0, 0.4], [0.1, -4.0]]]))
    # act=3 grad=2 and act=1 grad=2 are identically zero; act=1 grad 0/1 are linear
    assert ops.fused_bias_act(x, b, e, 3, 2, 0.2, 2.0).abs().max() == 0
    assert ops.fused_bias_act(x, b, e, 1, 2, 0.2, 2.0).abs().max() == 0
    assert torch.allclose(ops.fused_bias_act(x, b, e, 1, 0, 0.2, 3.0), (x + b.view(1, 2, 1)) * 3.0)
    # bias indexes dim 1 for 2-D inputs too (step_b = 1)
    x2 = torch.tensor([[1.0, -1.0], [-2.0, 2.0]])
    assert torch.allclose(ops.fused_bias_act(x2, b, e, 3, 0, 0.5, 1.0), torch.tensor([[1.5, -1.0], [-0.75, 1.0]]))


def test_upfirdn2d_known_answers():
    # identity tap
    x = torch.arange(12.0).view(1, 3, 4, 1)
    assert torch.equal(ops.upfirdn2d(x, torch.ones(1, 1), 1, 1, 1, 1, 0, 0, 0, 0), x)
    # true convolution: taps are applied flipped. x = delta -> output reproduces the taps un-flipped
    d = torch.zeros(1, 3, 3, 1)
    d[0, 0, 0, 0] = 1.0
    k = torch.tensor([[1.0, 2.0], [3.0, 4.0]])
    y = ops.upfirdn2d(d, k, 1, 1, 1, 1, 1, 0, 1, 0)[0, :, :, 0]
    assert torch.equal(y[:2, :2], k)
    # up=2 zero insertion with a 1x1 tap and the output-size formula (kernel.cu:167-168)
    y = ops.upfirdn2d(x, torch.ones(1, 1), 2, 2, 1, 1, 0, 0, 0, 0)
    assert y.shape == (1, 6, 8, 1) and torch.equal(y[0, ::2, ::2, 0], x[0, :, :, 0]) and y[0, 1::2].abs().max() == 0
    # down=2 keeps every second sample
    y = ops.upfirdn2d(x, torch.ones(1, 1), 1, 1, 2, 2, 0, 0, 0, 0)
    assert torch.equal(y[0, :, :, 0], x[0, ::2, ::2, 0])
    # negative pad crops
    y = ops.upfirdn2d(x, torch.ones(1, 1), 1, 1, 1, 1, -1, 0, 0, -1)
    assert torch.equal(y[0, :, :, 0], x[0, :2, 1:, 0])


# ---- (b) fixtures from the reference -----------------------------------------------------------------
def test_oracle_upfirdn2d_matches_reference_fixtures():
    g = load_golden("ops.pt")
    for case in g["fir"]:
        y = ops.upfirdn2d(case["x"], case["k"], *case["cfg"])
        assert y.shape == case["y"].shape, case["cfg"]
        assert rel_err(y, case["y"]) < 1e-6, case["cfg"]


def test_oracle_generator_matches_reference_fixture():
    g = load_golden("generator.pt")
    sd = g["state_dict"]
    img = model.generator_forward(sd, g["z"], g["noise"], g["inject_index"])
    assert rel_err(img, g["image"]) < 1e-6
    img_dead = model.generator_forward(sd, g["z"], g["noise"], g["inject_index"], dead_branch=True)
    assert torch.equal(img, img_dead)       # the second branch's main convs are unobservable (Q1)
    # fixed-noise buffers + single style
    n_main = 6
    noise = [sd["noises.noise_start"]] + [sd["noises.noise_%d" % i] for i in range(n_main)]
    assert rel_err(model.generator_forward(sd, g["z1"], noise), g["image_fixed"]) < 1e-6
    # latent layout
    lat = model.generator_latent(sd, g["z1"], None, n_main + 2)
    assert rel_err(lat, g["latent"]) < 1e-6


def test_oracle_generator_gradients_match_reference_fixture():
    g = load_golden("generator.pt")
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in g["state_dict"].items()}
    img = model.generator_forward(sd, g["z"], g["noise"], g["inject_index"])
    names = sorted(g["grads"])
    grads = torch.autograd.grad((img * g["direction"]).sum(), [sd[n] for n in names], allow_unused=True)
    for n, gr in zip(names, grads):
        assert gr is not None, n
        assert rel_err(gr, g["grads"][n]) < 2e-5, n
    # exactly the 36-style set of never-used parameters (main_convolutions_2.*) has no gradient
    assert all(n.startswith("main_convolutions_2.") for n in g["none_grads"]) and len(g["none_grads"]) == 18


def test_oracle_path_length_matches_reference_fixture():
    g = load_golden("generator.pt")
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in g["state_dict"].items()}
    latent = model.generator_latent(sd, g["z1"], None, 8)
    grad = model.path_length_grads(sd, latent, g["noise"], g["pl_noise"])
    assert rel_err(grad, g["pl_grad"]) < 2e-5
    penalty, pl, mean = model.path_length_penalty(grad, torch.zeros(1))
    assert rel_err(pl, g["pl_value"]) < 2e-5 and rel_err(penalty, g["pl_penalty"]) < 1e-4
    assert rel_err(mean, g["pl_mean"]) < 2e-5
    names = sorted(g["pl_param_grads"])
    grads = torch.autograd.grad(penalty, [sd[n] for n in names], allow_unused=True)
    for n, gr in zip(names, grads):
        ref = g["pl_param_grads"][n]
        if gr is None:
            assert ref.abs().max() == 0, n
        else:
            assert (gr - ref).abs().max() <= 1e-4 * max(ref.abs().max().item(), 1e-3), n


def test_oracle_discriminator_matches_reference_fixture():
    g = load_golden("discriminator.pt")
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in g["state_dict"].items()}
    scalar, pixel = model.discriminator_forward(sd, g["x"])
    assert rel_err(scalar, g["scalar"]) < 1e-6 and rel_err(pixel, g["pixel"]) < 1e-6
    names = sorted(g["grads"])
    grads = torch.autograd.grad((scalar * g["ds"]).sum() + (pixel * g["dp"]).sum(), [sd[n] for n in names])
    for n, gr in zip(names, grads):
        assert rel_err(gr, g["grads"][n]) < 5e-5, n
    r1 = model.r1_penalty(sd, g["x"])
    assert rel_err(r1, g["r1"]) < 1e-5
    names = sorted(g["r1_grads"])
    grads = torch.autograd.grad(r1, [sd[n] for n in names], allow_unused=True)
    for n, gr in zip(names, grads):
        ref = g["r1_grads"][n]
        if gr is None:
            assert ref.abs().max() == 0, n
        else:
            assert (gr - ref).abs().max() <= 1e-4 * max(ref.abs().max().item(), 1e-4), n


# ---- (c) live against the reference, when mounted ----------------------------------------------------
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")


@needs_ref
def test_oracle_upfirdn2d_live_against_reference_native():
    ref_loader.load()
    import sys
    native = sys.modules["multi_stylegan.op_static.upfirdn2d"].upfirdn2d_native
    gen = torch.Generator().manual_seed(0)
    for _ in range(25):
        up = int(torch.randint(1, 4, (1,), generator=gen))
        down = int(torch.randint(1, 4, (1,), generator=gen))
        kh, kw = (int(v) for v in torch.randint(1, 6, (2,), generator=gen))
        pads = [int(v) for v in torch.randint(-1, 4, (4,), generator=gen)]
        h, w = (int(v) for v in torch.randint(4, 12, (2,), generator=gen))
        minor = int(torch.randint(1, 3, (1,), generator=gen))
        x = torch.randn(2, h, w, minor, generator=gen)
        k = torch.randn(kh, kw, generator=gen)
        if h * up + pads[2] + pads[3] < kh or w * up + pads[0] + pads[1] < kw:
            continue
        a = ops.upfirdn2d(x, k, up, up, down, down, *pads)
        b = native(x, k, up, up, down, down, *pads)
        assert a.shape == b.shape and rel_err(a, b) < 1e-5


@needs_ref
def test_oracle_models_live_against_reference():
    ref = ref_loader.load()
    from oracle.make_golden import TINY_D, TINY_G, randomize
    torch.manual_seed(42)
    G = ref.generator.Generator(TINY_G)
    randomize(G, 3)
    z = torch.randn(2, 16)
    noise = [torch.randn(2, 1, 4, 4)] + [torch.randn(2, 1, 2 ** (i // 2 + 3), 2 ** (i // 2 + 3)) for i in range(6)]
    with torch.no_grad():
        assert rel_err(model.generator_forward(dict(G.state_dict()), z, noise), G(z, noise=noise)) < 1e-6
    D = ref.discriminator.Discriminator(TINY_D, no_rfp=True)
    randomize(D, 4)
    x = torch.rand(2, 2, 3, 32, 32)
    with torch.no_grad():
        s, p = D(x)
        s2, p2 = model.discriminator_forward(dict(D.state_dict()), x)
    assert rel_err(s2, s) < 1e-6 and rel_err(p2, p) < 1e-6
