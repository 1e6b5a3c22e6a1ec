"""Oracle for the EMA-generator sampling path.  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py): an element-wise restatement of what the reference's scripts compose from a generated sequence,
written with explicit loops / indexing so that it shares no tensor-op sequence with the product.

PARITY UNPINNED: the reference's scripts have no tests or golden outputs, cannot be imported (argparse at module level,
`.cuda()`, a trained checkpoint that is not in the repository); the restatement follows the cited lines.  The generator they
call IS pinned (oracle/model.py against tests/golden/generator.pt)."""
from typing import List, Tuple

import torch

from . import model


def sample_images(sequence: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """scripts/get_gan_samples.py:43-53 for ONE sample (the script's batch is 1): sequence [1, 2, T, H, W] ->
    bf, gfp of [T, 3, H, W]."""
    _, _, T, H, W = sequence.shape
    bf = torch.zeros(T, 3, H, W, dtype=sequence.dtype)
    gfp = torch.zeros(T, 3, H, W, dtype=sequence.dtype)
    for t in range(T):
        for c in range(3):
            bf[t, c] = sequence[0, 0, t]           # :46 repeat_interleave(3, dim=1), :51 permute(1, 0, 2, 3)
        gfp[t, 1] = sequence[0, 1, t]              # :47-49 only the green plane survives
    return bf, gfp


def interpolation_latents(anchors: torch.Tensor, frames_per_anchor: int, chunk: int) -> torch.Tensor:
    """scripts/gan_latent_space_interpolation.py:36-40: F.interpolate(mode="linear", align_corners=True) along the anchor
    axis = piecewise-linear through the anchors with the first and last output ON the first and last anchor."""
    A, D = anchors.shape
    n = frames_per_anchor * A
    out = torch.zeros(n, D, dtype=torch.float64)
    for i in range(n):
        pos = i * (A - 1) / (n - 1)
        lo = min(int(pos), A - 2) if A > 1 else 0
        w = pos - lo
        out[i] = anchors[lo].double() * (1.0 - w) + (anchors[lo + 1].double() * w if A > 1 else 0.0)
    return out.float().reshape(n // chunk, chunk, D)


def video_frames(samples: torch.Tensor) -> torch.Tensor:
    """scripts/gan_latent_space_interpolation.py:47-56: samples [N, 2, T, H, W] -> frames [N, 3, 2 H, T W]."""
    N, _, T, H, W = samples.shape
    video = torch.zeros(N, 3, 2 * H, T * W, dtype=samples.dtype)
    for t in range(T):
        for c in range(3):
            video[:, c, :H, t * W:(t + 1) * W] = samples[:, 0, t]      # bright field, grey, upper half
        video[:, 1, H:, t * W:(t + 1) * W] = samples[:, 1, t]          # GFP, green plane, lower half
    return video


def fixed_noise(sd: model.SD) -> List[torch.Tensor]:
    """multi_stylegan_generator.py:88-95,144-147: the registered noise buffers used when randomize_noise=False."""
    noise = [sd["noises.noise_start"]]
    i = 0
    while "noises.noise_%d" % i in sd:
        noise.append(sd["noises.noise_%d" % i])
        i += 1
    return noise
