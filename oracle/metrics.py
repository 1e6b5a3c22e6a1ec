"""Oracle for the validation metrics' arithmetic (multi_stylegan/validation_metrics.py).  TEST INFRASTRUCTURE ONLY.

Pinned: `frechet` against outputs of the reference's own `FID._calc_fid` / `FVD._calc_fvd`
(tests/golden/metrics.pt, oracle/make_golden_metrics.py).  The flows around it follow the cited lines; the pretrained
networks they feed are not available (parity of preprocessing against kornia unpinned)."""
from typing import Callable, List, Sequence

import numpy as np
import torch


def frechet(real_activations: np.ndarray, fake_activations: np.ndarray) -> float:
    """:192-219 with numpy / scipy as the reference has it (today's scipy.linalg.sqrtm returns the matrix alone)."""
    from scipy.linalg import sqrtm
    real_mu, fake_mu = np.mean(real_activations, axis=0), np.mean(fake_activations, axis=0)
    real_cov, fake_cov = np.cov(real_activations, rowvar=False), np.cov(fake_activations, rowvar=False)
    diff = real_mu - fake_mu
    cov_mean = sqrtm(real_cov @ fake_cov)
    if np.iscomplexobj(cov_mean):
        cov_mean = cov_mean.real
    return float(diff @ diff + np.trace(real_cov) + np.trace(fake_cov) - 2 * np.trace(cov_mean))


def inception_score(predictions: torch.Tensor) -> float:
    """:128-146 for one channel."""
    p_y = predictions.mean(dim=0, keepdim=True)
    kl = torch.sum(predictions * torch.log(predictions / p_y), dim=-1)
    return float(kl.mean().exp())


def normalize_m1_1_batch(x: torch.Tensor) -> torch.Tensor:
    """misc.py:216-235, per sample with explicit loops."""
    out = torch.empty_like(x)
    for b in range(x.shape[0]):
        lo, hi = x[b].min(), x[b].max()
        out[b] = 2.0 * ((x[b] - lo) / (hi - lo)).clamp(min=1e-03) - 1.0
    return out


def frame_activations(batches: Sequence[torch.Tensor], channels: Sequence[int], net: Callable, data_samples: int,
                      stop_when_full: bool) -> List[np.ndarray]:
    """The reference's collection loop of FID (:236-262 real, :277-301 fake): per batch one torch.randint per channel
    (all draws before the network calls), activations appended sample by sample, lists cut at data_samples."""
    lists = [[] for _ in channels]
    for images in batches:
        frames = []
        for c in channels:
            t = int(torch.randint(0, images.shape[2], (1,)))
            frames.append(torch.stack([images[:, c, t]] * 3, dim=1).unsqueeze(2))          # [B, 3, 1, H, W]
        for lst, f in zip(lists, frames):
            lst.extend(net(normalize_m1_1_batch(f)[:, :, 0]).cpu().unbind(dim=0))
        if stop_when_full and len(lists[0]) >= data_samples:
            break
    return [torch.stack(lst[:data_samples], dim=0).double().numpy() for lst in lists]
