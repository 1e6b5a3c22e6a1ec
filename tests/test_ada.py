"""Adaptive discriminator augmentation (reference adaptive_discriminator_augmentation.py:11-213): pipeline semantics with
injected draws, differentiability (the generator step back-propagates through it), and the p controller."""
import copy

import pytest
import torch

from tests.conftest import rel_err


def _draws(batch, **kw):
    import numpy as np
    d = dict(flip=[], rot90=[], rot90_angle=0., roll=[], roll_frac=(0., 0.), iso=[], iso_scale=np.zeros((0, 1)),
             rot_a=[], rot_a_angle=np.zeros(0), aniso=[], aniso_scale=np.zeros((0, 2)), rot_b=[], rot_b_angle=np.zeros(0))
    d.update(kw)
    return d


def _full_draws():
    import numpy as np
    return _draws(4, flip=[0, 2], rot90=[1], rot90_angle=90., roll=[0, 3], roll_frac=(0.1, -0.07), iso=[1, 2],
                  iso_scale=np.array([[1.02], [0.97]]), rot_a=[0], rot_a_angle=np.array([33.]), aniso=[3],
                  aniso_scale=np.array([[1.03, 0.96]]), rot_b=[2], rot_b_angle=np.array([-120.]))


def run_pipeline(dev):
    from multi_stylegan_b200.adaptive_discriminator_augmentation import AugmentationPipeline
    g = torch.Generator().manual_seed(0)
    x = torch.rand(4, 6, 32, 32, generator=g)
    pipe = AugmentationPipeline()
    # no gate fires: exact identity
    same = pipe(x.clone().to(dev), 0.5, _draws(4))
    assert torch.equal(same.cpu(), x)
    # flip only: exact
    fl = pipe(x.clone().to(dev), 0.5, _draws(4, flip=[1, 3]))
    want = x.clone()
    want[[1, 3]] = want[[1, 3]].flip(dims=(-1,))
    assert torch.equal(fl.cpu(), want)
    # integer roll: exact
    ro = pipe(x.clone().to(dev), 0.5, _draws(4, roll=[0], roll_frac=(0.125, -0.125)))
    want = x.clone()
    want[[0]] = torch.roll(want[[0]], shifts=(4, -4), dims=(-2, -1))
    assert torch.equal(ro.cpu(), want)
    # the reference assigns into a view of its input (:118-187): the caller's tensor holds the augmented batch afterwards
    xin = x.clone().to(dev)
    res = pipe(xin, 0.5, _full_draws())
    assert torch.equal(xin, res)
    # everything at once, differentiable w.r.t. the images
    xi = x.clone().to(dev).requires_grad_(True)
    out = pipe(xi * 1.0, 0.5, _full_draws())
    assert out.shape == x.shape and torch.isfinite(out).all()
    w = torch.rand(out.shape, generator=g).to(dev)
    gx, = torch.autograd.grad((out * w).sum(), xi)
    assert gx.shape == x.shape and gx.abs().sum() > 0
    return out.detach().cpu(), gx.cpu(), x, w.cpu()


def _stage_draws():
    """One dictionary per stage (each alone), then all stages together."""
    import numpy as np
    yield "rot90_+90", _draws(4, rot90=[0, 2], rot90_angle=90.)
    yield "rot90_-90", _draws(4, rot90=[1], rot90_angle=-90.)
    yield "rot90_180", _draws(4, rot90=[3], rot90_angle=180.)
    yield "iso", _draws(4, iso=[0, 3], iso_scale=np.array([[1.05], [0.93]]))
    yield "rot_a", _draws(4, rot_a=[1, 2], rot_a_angle=np.array([17.0, -143.0]))
    yield "aniso", _draws(4, aniso=[0, 1], aniso_scale=np.array([[1.04, 0.95], [0.97, 1.02]]))
    yield "rot_b", _draws(4, rot_b=[2], rot_b_angle=np.array([71.0]))
    yield "all", _full_draws()


def check_against_independent_oracle(dev, tol=1e-4):
    """The product pipeline (own inverse pixel maps + own warp kernel) against oracle/ada.py (kornia's route restated
    with F.affine_grid / F.grid_sample): outputs and the gradient w.r.t. the input images, stage by stage."""
    from multi_stylegan_b200.adaptive_discriminator_augmentation import AugmentationPipeline
    from oracle import ada as oada
    g = torch.Generator().manual_seed(3)
    x = torch.rand(4, 6, 32, 32, generator=g)
    w = torch.rand(4, 6, 32, 32, generator=g)
    pipe = AugmentationPipeline()
    for name, d in _stage_draws():
        xo = x.clone().requires_grad_(True)
        want = oada.augmentation_pipeline(xo, d)
        gwant, = torch.autograd.grad((want * w).sum(), xo)
        xi = x.clone().to(dev).requires_grad_(True)
        got = pipe(xi * 1.0, 0.5, d)
        ggot, = torch.autograd.grad((got * w.to(dev)).sum(), xi)
        assert rel_err(got, want) < tol, (name, rel_err(got, want))
        assert rel_err(ggot, gwant) < tol, (name, rel_err(ggot, gwant))


def test_pipeline_host_logic(oracle_backend):
    run_pipeline("cpu")


def test_pipeline_matches_independent_oracle_host_logic(oracle_backend):
    check_against_independent_oracle("cpu")


@pytest.mark.gpu
def test_pipeline_kernels_match_independent_oracle(built_library):
    """CUDA warp + adjoint kernels, all seven stages with injected draws, against oracle/ada.py."""
    run_pipeline("cuda:0")
    check_against_independent_oracle("cuda:0")


@pytest.mark.gpu
def test_affine_warp_adjoint(built_library):
    """<A x, y> == <x, A^T y> for the warp kernel and its backward kernel, both padding modes."""
    from multi_stylegan_b200 import _C
    import math
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 6, 40, 24, generator=g).cuda()
    y = torch.randn(3, 6, 40, 24, generator=g).cuda()
    th = []
    for a, sc in [(0.0, 1.0), (30.0, 1.2), (-100.0, 0.8)]:
        c, sn = math.cos(math.radians(a)) / sc, math.sin(math.radians(a)) / sc
        th.append([[c, -sn, 11.5 - c * 11.5 + sn * 19.5], [sn, c, 19.5 - sn * 11.5 - c * 19.5]])
    theta = torch.tensor(th).cuda()
    for mode in (0, 1):
        lhs = (_C.affine_warp(x, theta, mode).double() * y.double()).sum()
        rhs = (x.double() * _C.affine_warp_bwd(y, theta, mode).double()).sum()
        assert abs((lhs - rhs) / lhs).item() < 1e-5


def test_wrapper_p_controller_and_pairing(oracle_backend):
    """p moves by +-p_step every r_update fake calls, clamped to [0, p_max] (:80-95); forward_pair == two calls."""
    from multi_stylegan_b200.adaptive_discriminator_augmentation import AdaptiveDiscriminatorAugmentation

    class Stub(torch.nn.Module):
        def __init__(self, sign):
            super().__init__()
            self.sign = sign

        def forward(self, x, **kw):
            b = x.shape[0]
            return torch.full((b, 1), self.sign), torch.full((b, 1, 1, x.shape[-2], x.shape[-1]), self.sign)

    x = torch.rand(2, 2, 3, 16, 16)
    for sign, expect in ((1.0, 0.05 + 0.005), (-1.0, 0.05 - 0.005)):
        ada = AdaptiveDiscriminatorAugmentation(Stub(sign))
        for i in range(8):
            assert ada.p == 0.05
            ada(x.clone(), is_real=True)             # real calls do not count
            ada(x.clone(), is_real=False)
        assert abs(ada.p - expect) < 1e-12 and len(ada.r) == 0 and len(ada.r_history) == 1
    ada = AdaptiveDiscriminatorAugmentation(Stub(-1.0))
    ada.p = 0.002
    for i in range(8):
        ada(x.clone(), is_real=False)
    assert ada.p == 0.0                              # clamped
    assert ada(x, is_cut_mix=True)[0].shape == (2, 1)   # cut-mix path bypasses the augmentation (:64-65)
    # forward_pair falls back to two calls for a discriminator without forward_pair and records one r value
    ada = AdaptiveDiscriminatorAugmentation(Stub(1.0))
    (rs, rp), (fs, fp) = ada.forward_pair(x.clone(), x.clone())
    assert rs.shape == (2, 1) and fp.shape == (2, 1, 1, 16, 16) and len(ada.r) == 1
