"""Validation metrics (second half): what multi_stylegan/validation_metrics.py computes around
its pretrained feature networks — the frame selection and normalisation that feed them, the Frechet distance of FID / FVD
(:192-219, :401-428) and the inception score (:128-146) — with the statistics kept on the device.

The pretrained networks themselves (torchvision's Inception v3 for IS, the pytorch-fid Inception for FID, the Kinetics I3D
for FVD) are downloads the reference fetches at construction (:43, :578, :391); nothing in this library ships or fetches
weights.  `IS` / `FID` / `FVD` therefore take the feature network as an argument (`network=`: any module mapping
[B, 3, H, W] — FVD: [B, 3, T, H, W] — to logits resp. activations) and raise if called without one.

Device-side statistics instead of the reference's per-batch `.cpu().unbind()` lists and a numpy covariance of a
[samples, 2048] matrix at the end: `FrechetStatistics` streams batches into a float64 sum and a float64 sum of outer
products (one [D, D] GEMM per batch), so 5 000 samples never exist as one array and nothing leaves the GPU until the
scalar.  tr sqrt(C_r C_f) is evaluated as the sum of the square roots of the eigenvalues of the symmetric matrix
C_r^(1/2) C_f C_r^(1/2) (two `eigh` calls; the same number as the real part of the trace of scipy's `sqrtm(C_r C_f)` that the
reference takes, including rank-deficient covariances; tests/test_validation_metrics.py against outputs of the reference's
own `_calc_fid`)."""
import math
from typing import Callable, Iterable, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import misc


# ---- normalisation fed to the feature networks (misc.py:216-235) ------------------------------------------------------
def normalize_0_1_batch(input: torch.Tensor) -> torch.Tensor:
    """misc.py:216-225: per-sample min-max normalisation of a 5-D batch — and the reference's clamp of the RESULT at 1e-3."""
    flat = input.reshape(input.shape[0], -1)
    lo = flat.min(dim=1)[0][:, None, None, None, None]
    hi = flat.max(dim=1)[0][:, None, None, None, None]
    return ((input - lo) / (hi - lo)).clamp(min=1e-03)


def normalize_m1_1_batch(input: torch.Tensor) -> torch.Tensor:
    """misc.py:228-235."""
    return 2. * normalize_0_1_batch(input) - 1.


# ---- statistics -------------------------------------------------------------------------------------------------------
class FrechetStatistics(object):
    """Streaming first and second moments of activation batches [b, D] in float64 on the batches' device.  The moments are
    taken about the first batch's mean, so the covariance does not suffer from cancellation when |mean| >> spread."""

    def __init__(self) -> None:
        self.n = 0
        self.shift: Optional[torch.Tensor] = None
        self.sum: Optional[torch.Tensor] = None
        self.outer: Optional[torch.Tensor] = None

    def update(self, activations: torch.Tensor, limit: Optional[int] = None) -> "FrechetStatistics":
        """Adds the rows of `activations` (only as many as keep the total at `limit`: the reference truncates its lists to
        `data_samples`, :296-301)."""
        a = activations.detach().flatten(start_dim=1).double()
        if limit is not None:
            a = a[:max(0, limit - self.n)]
        if a.shape[0] == 0:
            return self
        if self.sum is None:
            self.shift = a.mean(dim=0)
            self.sum = torch.zeros(a.shape[1], dtype=torch.float64, device=a.device)
            self.outer = torch.zeros(a.shape[1], a.shape[1], dtype=torch.float64, device=a.device)
        a = a - self.shift
        self.sum += a.sum(dim=0)
        self.outer.addmm_(a.t(), a)
        self.n += int(a.shape[0])
        return self

    def mean_cov(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """np.mean(axis=0), np.cov(rowvar=False) (:203-206): unbiased covariance."""
        if self.n < 2:
            raise ValueError("FrechetStatistics: at least two samples are needed for a covariance")
        m = self.sum / self.n
        cov = (self.outer - self.n * torch.outer(m, m)) / (self.n - 1)
        return self.shift + m, (cov + cov.t()) / 2

    @classmethod
    def of(cls, activations: Union[np.ndarray, torch.Tensor, "FrechetStatistics"]) -> "FrechetStatistics":
        if isinstance(activations, cls):
            return activations
        if isinstance(activations, np.ndarray):
            activations = torch.from_numpy(activations)
        return cls().update(activations)


def _sqrt_psd(c: torch.Tensor) -> torch.Tensor:
    w, v = torch.linalg.eigh(c)
    return (v * w.clamp(min=0).sqrt()) @ v.t()


def frechet_distance(real_activations, fake_activations) -> float:
    """FID._calc_fid / FVD._calc_fvd (:192-219 / :401-428): |mu_r - mu_f|^2 + tr C_r + tr C_f - 2 tr sqrt(C_r C_f).
    Arguments: activation arrays [samples, features] (numpy or torch) or FrechetStatistics."""
    real, fake = FrechetStatistics.of(real_activations), FrechetStatistics.of(fake_activations)
    mu_r, cov_r = real.mean_cov()
    mu_f, cov_f = fake.mean_cov()
    assert mu_r.shape == mu_f.shape and cov_r.shape == cov_f.shape
    cov_f = cov_f.to(cov_r.device)
    diff = mu_r - mu_f.to(mu_r.device)
    root_r = _sqrt_psd(cov_r)
    inner = root_r @ cov_f @ root_r
    tr_root = torch.linalg.eigvalsh((inner + inner.t()) / 2).clamp(min=0).sqrt().sum()
    return float(diff @ diff + torch.trace(cov_r) + torch.trace(cov_f) - 2 * tr_root)


def inception_score(probabilities: torch.Tensor) -> float:
    """:128-146: exp(mean_x KL(p(y|x) || p(y))) from the softmax outputs [samples, classes]."""
    p = probabilities.double()
    p_y = p.mean(dim=0, keepdim=True)
    return float(torch.sum(p * torch.log(p / p_y), dim=-1).mean().exp())


# ---- the metric objects -----------------------------------------------------------------------------------------------
def _latent_dimensions(generator) -> int:
    return generator.module.latent_dimensions if isinstance(generator, nn.DataParallel) else generator.latent_dimensions


def random_frame(images: torch.Tensor, channel: int) -> torch.Tensor:
    """:85-86, :248-249: one random time step of one channel for the whole batch as grey RGB, [B, 3, 1, H, W] (one
    torch.randint draw from the global CPU generator, like the reference)."""
    t = torch.randint(0, images.shape[2], (1,))
    return images[:, channel, t.to(images.device)].unsqueeze(dim=1).repeat_interleave(dim=1, repeats=3)


class _Metric(object):
    """Constructor arguments of the reference's IS / FID / FVD (`data_parallel` is accepted and ignored: one process per
    GPU) plus `network`, the pretrained feature extractor the reference downloads."""

    def __init__(self, device: Union[str, torch.device] = "cuda", data_parallel: bool = True, batch_size: int = 1,
                 data_samples: int = 5000, no_rfp: bool = False, no_gfp: bool = False,
                 network: Optional[nn.Module] = None) -> None:
        self.device, self.data_parallel, self.batch_size = device, data_parallel, batch_size
        self.data_samples, self.no_rfp, self.no_gfp = data_samples, no_rfp, no_gfp
        self.network = network

    def _channels(self) -> Tuple[int, ...]:
        if self.no_gfp and not self.no_rfp:
            # the reference reaches `self.activations_real_gfp` / a three-score return here without ever having created the
            # GFP lists (:185-188, :355-358): the combination does not work there either
            raise AttributeError("%s: no_gfp=True needs no_rfp=True (the reference fails on activations_real_gfp)"
                                 % type(self).__name__)
        return (0,) + (() if self.no_gfp else (1,)) + (() if self.no_rfp else (2,))

    def _net(self) -> nn.Module:
        if self.network is None:
            raise RuntimeError("%s needs its pretrained feature network (the reference downloads it; this library ships no "
                               "weights): pass network=<module>" % type(self).__name__)
        return self.network.to(self.device).eval()

    def _fakes(self, generator) -> Iterable[torch.Tensor]:
        generator.to(self.device)
        generator.eval()
        for _ in range(math.ceil(self.data_samples / self.batch_size)):
            noise_input = misc.get_noise(batch_size=self.batch_size, latent_dimension=_latent_dimensions(generator),
                                         p_mixed_noise=0.0, device=self.device)
            yield generator(input=noise_input)

    def _result(self, values):
        """The reference's return statements in their order (:152-156, :352-358): with GFP the first `return` wins and
        gives (bf, gfp) — also when an RFP score was computed; bright field alone otherwise."""
        if not self.no_gfp:
            return tuple(values[:2])
        return values[0]


class IS(_Metric):
    """Inception score of random generated frames per channel (:16-156)."""

    @staticmethod
    def preprocessing(input: torch.Tensor) -> torch.Tensor:
        """:44-53: bilinear, anti-aliased resize of [B, 3, 1, H, W] frames to 299 x 299, then [-1, 1] per sample.  (The
        reference calls kornia.resize(antialias=True); kornia is not in this image — torch's anti-aliased bilinear
        interpolation stands in, parity of this one step unpinned.)"""
        x = F.interpolate(input[:, :, 0], size=(299, 299), mode="bilinear", antialias=True, align_corners=False)
        return normalize_m1_1_batch(x[:, :, None])[:, :, 0]

    @torch.no_grad()
    def __call__(self, generator, **kwargs):
        net = self._net()
        channels = self._channels()
        probs = [[] for _ in channels]
        for fake_images in self._fakes(generator):
            for i, c in enumerate(channels):
                probs[i].append(net(self.preprocessing(random_frame(fake_images, c))).softmax(dim=1))
        return self._result([inception_score(torch.cat(p, dim=0)[:self.data_samples]) for p in probs])


class FID(_Metric):
    """Frechet inception distance between random real and generated frames per channel (:157-358); the real statistics
    are computed once and cached like the reference's `activations_real_*`."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.real_statistics = None

    def _features(self, net, images: torch.Tensor, channel: int) -> torch.Tensor:
        return net(normalize_m1_1_batch(random_frame(images, channel))[:, :, 0])

    @torch.no_grad()
    def __call__(self, generator, dataset):
        net = self._net()
        channels = self._channels()
        if self.real_statistics is None:
            stats = [FrechetStatistics() for _ in channels]
            for real_images in dataset:
                real_images = real_images.to(self.device)
                for s, c in zip(stats, channels):
                    s.update(self._features(net, real_images, c), self.data_samples)
                if stats[0].n >= self.data_samples:
                    break
            self.real_statistics = stats
        fake = [FrechetStatistics() for _ in channels]
        for fake_images in self._fakes(generator):
            for s, c in zip(fake, channels):
                s.update(self._features(net, fake_images, c), self.data_samples)
        return self._result([frechet_distance(r, f) for r, f in zip(self.real_statistics, fake)])


class FVD(FID):
    """Frechet video distance (:361-568): the whole sequence of a channel as grey RGB video through the I3D network."""

    def _features(self, net, images: torch.Tensor, channel: int) -> torch.Tensor:
        video = images[:, channel].unsqueeze(dim=1).repeat_interleave(dim=1, repeats=3)      # :462-463
        return net(normalize_m1_1_batch(video)).flatten(start_dim=1)
