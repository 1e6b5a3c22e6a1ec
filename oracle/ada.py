"""TEST INFRASTRUCTURE — CPU restatement of the reference's ADA augmentation pipeline
(multi_stylegan/adaptive_discriminator_augmentation.py:99-213) with torch's own ``F.affine_grid`` /
``F.grid_sample``.  Imported only by tests/ (never by the product).

It is deliberately independent of the product's implementation (which builds one inverse pixel-space map per sample
and samples it with its own CUDA kernel): here every stage follows the route of kornia 0.4.1, the reference's warp
library (requirements.txt:7) —

    kaf.apply_affine  ->  get_affine_matrix2d(translations, center, scale, angle)   [rotation by -angle]
                      ->  get_rotation_matrix2d(center, -angle, scale)              [2x3 forward matrix, pixels]
                      ->  warp_affine: homography -> normalise to [-1, 1] with (W-1, H-1) -> invert
                          -> F.affine_grid + F.grid_sample(bilinear, reflection, align_corners=True)
    kaf.rotate        ->  get_rotation_matrix2d(((W-1)/2, (H-1)/2), +angle, 1) -> warp_affine(bilinear, zeros)

PARITY UNPINNED: kornia 0.4.1 is neither vendored under /root/reference nor installed and there is no network, so
the matrix conventions are restated from its published source as remembered (SURVEY.md section 8c says the same).  One
quirk is reproduced on purpose: 0.4.1's get_rotation_matrix2d takes the translation column from the FIRST row of the
scaled rotation only (alpha = M00, beta = M01: t = ((1-alpha) x - beta y, beta x + (1-alpha) y)), which for anisotropic
scales does not keep the centre fixed.  The reference passes centre = 0.5 * (H, W) as (x, y) (:137-138).

All random draws are injected (same dictionary the product's `sample_draws` produces), in the reference's order."""
import math
from typing import Dict, List

import torch
import torch.nn.functional as F


def rotation_matrix2d(center: torch.Tensor, angle_deg: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """kornia 0.4.1 get_rotation_matrix2d: center [N,2] (x, y), angle [N] degrees, scale [N,2] -> [N,2,3]."""
    a = angle_deg.double() * math.pi / 180.0
    cos, sin = torch.cos(a), torch.sin(a)
    rot = torch.stack([torch.stack([cos, sin], -1), torch.stack([-sin, cos], -1)], -2)          # angle_to_rotation_matrix
    scaled = rot @ torch.diag_embed(scale.double())
    alpha, beta = scaled[:, 0, 0], scaled[:, 0, 1]
    x, y = center[:, 0].double(), center[:, 1].double()
    M = torch.zeros(a.shape[0], 2, 3, dtype=torch.float64)
    M[:, :, :2] = scaled
    M[:, 0, 2] = (1.0 - alpha) * x - beta * y
    M[:, 1, 2] = beta * x + (1.0 - alpha) * y
    return M


def warp_affine(src: torch.Tensor, M: torch.Tensor, padding_mode: str) -> torch.Tensor:
    """kornia 0.4.1 warp_affine with align_corners=True, bilinear: src [N,C,H,W], M [N,2,3] forward pixel map."""
    N, C, H, W = src.shape
    M3 = torch.zeros(N, 3, 3, dtype=torch.float64)
    M3[:, :2] = M
    M3[:, 2, 2] = 1.0
    norm = torch.tensor([[2.0 / (W - 1), 0.0, -1.0], [0.0, 2.0 / (H - 1), -1.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    dst_norm_trans_src_norm = norm @ M3 @ torch.inverse(norm)
    src_norm_trans_dst_norm = torch.inverse(dst_norm_trans_src_norm)
    grid = F.affine_grid(src_norm_trans_dst_norm[:, :2].to(src.dtype), [N, C, H, W], align_corners=True)
    return F.grid_sample(src, grid, mode="bilinear", padding_mode=padding_mode, align_corners=True)


def apply_affine(images: torch.Tensor, angle_deg, scale_xy) -> torch.Tensor:
    """kaf.apply_affine with zero translation / shear, centre 0.5 * (H, W), bilinear, reflection, align_corners=True."""
    N, _, H, W = images.shape
    center = torch.tensor([[0.5 * H, 0.5 * W]], dtype=torch.float64).repeat(N, 1)
    M = rotation_matrix2d(center, -torch.as_tensor(angle_deg, dtype=torch.float64).reshape(N),
                          torch.as_tensor(scale_xy, dtype=torch.float64).reshape(N, 2))
    return warp_affine(images, M, "reflection")


def rotate(images: torch.Tensor, angle_deg: float) -> torch.Tensor:
    """kaf.rotate: about the tensor centre ((W-1)/2, (H-1)/2), bilinear, zeros padding."""
    N, _, H, W = images.shape
    center = torch.tensor([[(W - 1) / 2.0, (H - 1) / 2.0]], dtype=torch.float64).repeat(N, 1)
    M = rotation_matrix2d(center, torch.full((N,), float(angle_deg), dtype=torch.float64), torch.ones(N, 2))
    return warp_affine(images, M, "zeros")


def augmentation_pipeline(images: torch.Tensor, d: Dict[str, object]) -> torch.Tensor:
    """reference :99-199 with the draws `d` injected; returns the augmented batch (the input is not modified)."""
    import numpy as np
    x = images.clone()
    H, W = x.shape[-2:]

    def put(idx: List[int], values: torch.Tensor) -> None:
        nonlocal x
        sel = torch.zeros(x.shape[0], dtype=torch.bool)
        sel[idx] = True
        full = torch.zeros_like(x)
        full[idx] = values
        x = torch.where(sel.view(-1, 1, 1, 1), full, x)                 # out-of-place images[idx] = values

    if d["flip"]:
        put(d["flip"], x[d["flip"]].flip(dims=(-1,)))                                        # :116-118
    if d["rot90"]:
        put(d["rot90"], rotate(x[d["rot90"]], float(d["rot90_angle"])))                      # :120-125
    if d["roll"]:
        shift = (int(H * d["roll_frac"][0]), int(W * d["roll_frac"][1]))                     # :210-213
        put(d["roll"], torch.roll(x[d["roll"]], shifts=shift, dims=(-2, -1)))
    if d["iso"]:
        s = np.asarray(d["iso_scale"], dtype=np.float64).reshape(-1, 1)
        put(d["iso"], apply_affine(x[d["iso"]], np.zeros(len(d["iso"])), np.repeat(s, 2, axis=1)))         # :131-147
    if d["rot_a"]:
        put(d["rot_a"], apply_affine(x[d["rot_a"]], np.asarray(d["rot_a_angle"]), np.ones((len(d["rot_a"]), 2))))
    if d["aniso"]:
        put(d["aniso"], apply_affine(x[d["aniso"]], np.zeros(len(d["aniso"])),
                                     np.asarray(d["aniso_scale"], dtype=np.float64).reshape(-1, 2)))       # :166-182
    if d["rot_b"]:
        put(d["rot_b"], apply_affine(x[d["rot_b"]], np.asarray(d["rot_b_angle"]), np.ones((len(d["rot_b"]), 2))))
    return x
