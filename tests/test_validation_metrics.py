"""Validation metrics: the Frechet distance against outputs of the reference's own `_calc_fid` /
`_calc_fvd` (tests/golden/metrics.pt), the normalisation helpers against the reference's, and the FID / FVD / IS flows with
an injected feature network against the oracle's restatement of the reference's collection loops."""
import random

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import make_golden_metrics as mk
from oracle import metrics as ometrics
from oracle.make_golden import TINY_G, randomize
from tests.conftest import load_golden


def _check_frechet(device, rtol):
    from multi_stylegan_b200 import validation_metrics as vm
    golden = load_golden("metrics.pt")["frechet"]
    for name, (real, fake) in mk.cases().items():
        want = golden[name]
        assert abs(ometrics.frechet(real, fake) - want) <= 1e-6 * max(1.0, abs(want)), name      # oracle pinned to the reference
        got = vm.frechet_distance(torch.from_numpy(real).to(device), torch.from_numpy(fake).to(device))
        assert abs(got - want) <= rtol * max(1.0, abs(want)), (name, got, want)
        # streamed in ragged batches = all at once; a limit cuts the stream like the reference cuts its lists
        stats = vm.FrechetStatistics()
        for chunk in torch.from_numpy(real).to(device).split(7):
            stats.update(chunk, limit=real.shape[0] - 3)
        assert stats.n == real.shape[0] - 3
        whole = vm.frechet_distance(real[:-3], fake)
        assert abs(vm.frechet_distance(stats, torch.from_numpy(fake).to(device)) - whole) <= 1e-8 * max(1.0, abs(whole))
    big_mean = mk.cases()["full_rank"]
    shifted = (big_mean[0] + 1e6, big_mean[1] + 1e6)          # cancellation check: moments about a pilot mean
    assert abs(vm.frechet_distance(*[torch.from_numpy(a).to(device) for a in shifted]) - golden["full_rank"]) < 1e-4
    with pytest.raises(ValueError):
        vm.frechet_distance(real[:1], fake)


def test_frechet_distance_matches_the_reference():
    _check_frechet("cpu", 1e-7)


@pytest.mark.gpu
def test_frechet_distance_on_the_device():
    _check_frechet("cuda:0", 1e-7)


def test_normalisation_and_inception_score():
    from multi_stylegan_b200 import validation_metrics as vm
    g = load_golden("metrics.pt")["normalize"]
    assert torch.equal(vm.normalize_0_1_batch(g["input"].clone()), g["zero_one"])
    assert torch.equal(vm.normalize_m1_1_batch(g["input"].clone()), g["minus_one_one"])
    assert torch.allclose(ometrics.normalize_m1_1_batch(g["input"]), g["minus_one_one"], atol=1e-6)
    assert float(g["zero_one"].min()) == pytest.approx(1e-3)          # the reference clamps the normalised values from below
    p = torch.randn(50, 10, generator=torch.Generator().manual_seed(0)).softmax(dim=1)
    assert vm.inception_score(p) == pytest.approx(ometrics.inception_score(p.double()), rel=1e-9)
    assert vm.inception_score(torch.full((8, 4), 0.25)) == pytest.approx(1.0)
    onehot = torch.eye(4).repeat(3, 1) * (1 - 3e-6) + 1e-6
    assert vm.inception_score(onehot) == pytest.approx(4.0, rel=1e-3)


class _Features(nn.Module):
    """Stand-in for the pretrained networks: a fixed random projection of pooled pixels (frames or videos)."""

    def __init__(self, dims=12, classes=False):
        super().__init__()
        gen = torch.Generator().manual_seed(21)
        self.register_buffer("w", torch.randn(3 * 16, dims, generator=gen))
        self.classes = classes

    def forward(self, x):
        if x.dim() == 5:                                          # video [B, 3, T, H, W]: pool time too
            x = x.mean(dim=2)
        pooled = nn.functional.adaptive_avg_pool2d(x, 4).flatten(1)
        return pooled @ self.w


def _tiny_generator(dev):
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    torch.manual_seed(5)
    G = G_mod.Generator(TINY_G, compute_dead_branch=False)
    randomize(G, 3)
    return G.to(dev).eval()


def _check_flows(dev):
    from multi_stylegan_b200 import misc, validation_metrics as vm
    G = _tiny_generator(dev)
    net = _Features().to(dev)
    reals = [torch.rand(4, 2, 3, 32, 32, generator=torch.Generator().manual_seed(i)) for i in range(5)]
    samples, batch = 10, 4

    def seed():
        random.seed(1), np.random.seed(1), torch.manual_seed(1)

    def fakes():
        out = []
        with torch.no_grad():
            for _ in range(3):                                    # ceil(10 / 4)
                out.append(G(misc.get_noise(batch_size=batch, latent_dimension=16, p_mixed_noise=0.0, device=dev)).cpu())
        return out

    # FID: the product (device-side streaming statistics) against the reference's list / numpy / scipy flow
    fid = vm.FID(device=dev, batch_size=batch, data_samples=samples, no_rfp=True, network=net)
    seed()
    got = fid(G, reals)
    seed()
    real_act = ometrics.frame_activations(reals, (0, 1), lambda x: net(x.to(dev)), samples, True)
    # (the generator draws from the device's generator, the frame indices from the CPU's: regenerate in the same order)
    fake_batches = []
    fake_lists = [[], []]
    with torch.no_grad():
        for _ in range(3):
            images = G(misc.get_noise(batch_size=batch, latent_dimension=16, p_mixed_noise=0.0, device=dev)).cpu()
            acts = ometrics.frame_activations([images], (0, 1), lambda x: net(x.to(dev)), batch, False)
            for lst, a in zip(fake_lists, acts):
                lst.append(a)
    fake_act = [np.concatenate(lst)[:samples] for lst in fake_lists]
    want = tuple(ometrics.frechet(r, f) for r, f in zip(real_act, fake_act))
    assert len(got) == 2
    for g, w in zip(got, want):
        assert g == pytest.approx(w, rel=2e-3, abs=1e-6), (got, want)
    assert fid.real_statistics[0].n == samples                    # 3 batches of 4, cut at data_samples
    cached = fid.real_statistics
    seed()
    fid(G, [])                                                    # the real statistics are computed once
    assert fid.real_statistics is cached

    # FVD: whole sequences as videos, no frame draws
    fvd = vm.FVD(device=dev, batch_size=batch, data_samples=samples, no_rfp=True, network=net)
    seed()
    got = fvd(G, reals)
    seed()
    fk = fakes()

    def video_acts(batches, c, stop):
        lst = []
        for images in batches:
            v = torch.stack([images[:, c]] * 3, dim=1)
            lst.extend(net(ometrics.normalize_m1_1_batch(v).to(dev)).cpu().flatten(start_dim=1).unbind(0))
            if stop and len(lst) >= samples:
                break
        return torch.stack(lst[:samples]).double().numpy()
    want = tuple(ometrics.frechet(video_acts(reals, c, True), video_acts(fk, c, False)) for c in (0, 1))
    for g, w in zip(got, want):
        assert g == pytest.approx(w, rel=2e-3, abs=1e-6), (got, want)

    # IS: bright field only -> a single score; the preprocessing resizes to 299 x 299 and normalises per sample
    is_metric = vm.IS(device=dev, batch_size=batch, data_samples=samples, no_rfp=True, no_gfp=True, network=_Features(7).to(dev))
    seed()
    score = is_metric(G)
    assert isinstance(score, float) and 1.0 <= score <= 7.0
    x = vm.IS.preprocessing(torch.rand(2, 3, 1, 32, 32))
    assert x.shape == (2, 3, 299, 299) and float(x.max()) == pytest.approx(1.0) and float(x.min()) >= -1.0

    with pytest.raises(RuntimeError):                             # no weights are shipped or fetched
        vm.FID(device=dev, no_rfp=True)(G, reals)
    with pytest.raises(AttributeError):                           # the reference's broken combination stays an error
        vm.FID(device=dev, no_gfp=True, no_rfp=False, network=net)(G, reals)


def test_metric_flows_host_logic(oracle_backend):
    _check_flows("cpu")


@pytest.mark.gpu
def test_metric_flows_on_the_device(built_library):
    _check_flows(torch.device("cuda:0"))
