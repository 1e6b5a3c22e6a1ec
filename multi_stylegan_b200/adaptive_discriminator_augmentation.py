"""Adaptive discriminator augmentation (wrapper + pipeline) with the reference's interface
(multi_stylegan/adaptive_discriminator_augmentation.py:11-213).

Stage order, gates and parameter distributions follow the reference exactly, including its quirks: one
shared 90-degree rotation and one shared integer roll for all selected samples (:122, :210-211), lognormal
sigma (0.2 ln 2)^2 (:141, :176), centre = 0.5 * (H, W) passed as (x, y) (:137-138), four *sequential*
bilinear resamplings (each stage re-interpolates the previous stage's output).  The bilinear /
reflection / align_corners warps run in one CUDA kernel per stage (csrc/misc_ops.cu: affine_warp) with a
per-sample 2x3 matrix, identity for samples the gate skipped (exact copy).

kornia 0.4.1 (the reference's warp implementation, requirements.txt:7) is neither vendored nor installed,
so its matrix conventions are restated from its published source and are NOT pinned by any fixture
("parity unpinned").  All random draws can be injected (`draws=`) so the oracle and the
kernels consume identical parameters."""
import math
import random
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import _C


class _AffineWarpBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_output, theta, mode):
        ctx.save_for_backward(theta)
        ctx.mode = mode
        return _C.affine_warp_bwd(grad_output, theta, mode)

    @staticmethod
    def backward(ctx, gg):
        theta, = ctx.saved_tensors
        return _AffineWarp.apply(gg, theta, ctx.mode), None, None


class _AffineWarp(torch.autograd.Function):
    """Bilinear warp (csrc/misc_ops.cu), differentiable w.r.t. the image: the generator step back-propagates through
    the augmentation exactly as it does through the reference's kornia warps (:133,:152,:168,:187).  The warp is linear
    in the image, so its backward is the adjoint kernel and the backward of that is the warp again."""

    @staticmethod
    def forward(ctx, images, theta, mode):
        ctx.save_for_backward(theta)
        ctx.mode = mode
        return _C.affine_warp(images, theta, mode)

    @staticmethod
    def backward(ctx, grad_output):
        theta, = ctx.saved_tensors
        return _AffineWarpBackward.apply(grad_output, theta, ctx.mode), None, None


def affine_warp(images: torch.Tensor, theta: torch.Tensor, mode: int = 0) -> torch.Tensor:
    return _AffineWarp.apply(images, theta, mode)


def _select(gate: torch.Tensor) -> List[int]:
    return [i for i, v in enumerate(gate.tolist()) if v]


def sample_draws(batch: int, p: float) -> Dict[str, object]:
    """Consumes the host RNGs in the reference's order (:116-199)."""
    d: Dict[str, object] = {}
    d["flip"] = _select(torch.rand(batch) <= p)
    d["rot90"] = _select(torch.rand(batch) <= p)
    d["rot90_angle"] = random.choice([0., -90., 90., 180.]) if d["rot90"] else 0.
    d["roll"] = _select(torch.rand(batch) <= p)
    d["roll_frac"] = (random.uniform(-0.125, 0.125), random.uniform(-0.125, 0.125)) if d["roll"] else (0., 0.)
    d["iso"] = _select(torch.rand(batch) <= p)
    d["iso_scale"] = np.random.lognormal(mean=0, sigma=(0.2 * math.log(2)) ** 2, size=(len(d["iso"]), 1)) \
        if d["iso"] else np.zeros((0, 1))
    q = 1 - math.sqrt(1 - p)
    d["rot_a"] = _select(torch.rand(batch) <= q)
    d["rot_a_angle"] = np.random.uniform(low=-180, high=180, size=len(d["rot_a"])) if d["rot_a"] else np.zeros(0)
    d["aniso"] = _select(torch.rand(batch) <= p)
    d["aniso_scale"] = np.random.lognormal(mean=0, sigma=(0.2 * math.log(2)) ** 2, size=(len(d["aniso"]), 2)) \
        if d["aniso"] else np.zeros((0, 2))
    d["rot_b"] = _select(torch.rand(batch) <= q)
    d["rot_b_angle"] = np.random.uniform(low=-180, high=180, size=len(d["rot_b"])) if d["rot_b"] else np.zeros(0)
    return d


def affine_inverse_theta(batch: int, idx: List[int], angle_deg, scale_xy, center: Tuple[float, float]) -> torch.Tensor:
    """Per-sample output->input pixel map [B,2,3] of kornia's apply_affine (rotation by -angle about
    `center`, then per-axis scale; forward matrix inverted for sampling); identity outside `idx`."""
    theta = torch.zeros(batch, 2, 3, dtype=torch.float64)
    theta[:, 0, 0] = 1.0
    theta[:, 1, 1] = 1.0
    cx, cy = center
    for j, i in enumerate(idx):
        a = math.radians(-float(angle_deg[j]))
        sx, sy = float(scale_xy[j][0]), float(scale_xy[j][1])
        cos, sin = math.cos(a), math.sin(a)
        m = np.array([[cos * sx, sin * sy, 0.0], [-sin * sx, cos * sy, 0.0], [0.0, 0.0, 1.0]])
        m[0, 2] = cx - (m[0, 0] * cx + m[0, 1] * cy)
        m[1, 2] = cy - (m[1, 0] * cx + m[1, 1] * cy)
        theta[i] = torch.from_numpy(np.linalg.inv(m)[:2])
    return theta.float()


class AugmentationPipeline(nn.Module):
    def forward(self, images: torch.Tensor, p: float, draws: Optional[Dict[str, object]] = None) -> torch.Tensor:
        """images [B, C, H, W]; mutated in place for the index-type stages like the reference (:118,:124,:129)."""
        B, _, H, W = images.shape
        d = sample_draws(B, p) if draws is None else draws
        if d["flip"]:
            images[d["flip"]] = images[d["flip"]].flip(dims=(-1,))
        if d["rot90"]:
            ang = float(d["rot90_angle"])
            th = affine_inverse_theta(len(d["rot90"]), list(range(len(d["rot90"]))), [-ang] * len(d["rot90"]),
                                      [(1., 1.)] * len(d["rot90"]), ((W - 1) / 2, (H - 1) / 2))
            images[d["rot90"]] = affine_warp(images[d["rot90"]], th.to(images.device), mode=1)
        if d["roll"]:
            shift = (int(H * d["roll_frac"][0]), int(W * d["roll_frac"][1]))
            images[d["roll"]] = torch.roll(images[d["roll"]], shifts=shift, dims=(-2, -1))
        centre = (0.5 * H, 0.5 * W)
        stages = [(d["iso"], np.zeros(len(d["iso"])), np.repeat(np.asarray(d["iso_scale"]).reshape(-1, 1), 2, axis=1)),
                  (d["rot_a"], d["rot_a_angle"], np.ones((len(d["rot_a"]), 2))),
                  (d["aniso"], np.zeros(len(d["aniso"])), np.asarray(d["aniso_scale"]).reshape(-1, 2)),
                  (d["rot_b"], d["rot_b_angle"], np.ones((len(d["rot_b"]), 2)))]
        for idx, angles, scales in stages:
            if idx:
                theta = affine_inverse_theta(B, idx, angles, scales, centre).to(images.device)
                images = affine_warp(images, theta, mode=0)
        return images


class AdaptiveDiscriminatorAugmentation(nn.Module):
    graph_capturable = False         # the pipeline draws its per-call decisions on the host (ModelWrapper.cuda_graphs)

    def __init__(self, discriminator: nn.Module, r_target: float = 0.6, p_step: float = 5e-03, r_update: int = 8,
                 p_max: float = 0.8, process_group=None) -> None:
        super().__init__()
        self.discriminator = discriminator
        self.r_target, self.p_step, self.r_update, self.p_max = r_target, p_step, r_update, p_max
        self.r: List[torch.Tensor] = []
        self.p = 0.05
        self.r_history: List[float] = []
        self.augmentation_pipeline = AugmentationPipeline()
        self.process_group = process_group

    @torch.no_grad()
    def _calc_r(self, prediction_scalar: torch.Tensor, prediction_pixel_wise: torch.Tensor) -> torch.Tensor:
        """Overfitting heuristic (:51-52), kept on the device; all-reduced across ranks when sharded."""
        r = 0.5 * torch.mean(torch.sign(prediction_scalar)) \
            + 0.5 * torch.mean(torch.sign(prediction_pixel_wise.mean(dim=(-1, -2))))
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(r, group=self.process_group)
            r = r / dist.get_world_size(self.process_group)
        return r

    def _update_p(self) -> None:
        if len(self.r) >= self.r_update:
            r = float(torch.stack(self.r).mean().item())          # the only host sync: once per r_update calls
            self.p = self.p + self.p_step if r > self.r_target else self.p - self.p_step
            self.p = min(max(self.p, 0.), self.p_max)
            self.r = []
            self.r_history.append(r)

    def forward_pair(self, real: torch.Tensor, fake: torch.Tensor, draws_real=None, draws_fake=None):
        """forward(real, is_real=True) and forward(fake, is_real=False) with one batched discriminator pass: both halves
        are augmented separately (real first, the order in which the reference consumes the host RNGs), `p` cannot change
        between the two calls of a step (it only moves after a fake call), so the results are those of the two calls."""
        pair = getattr(self.discriminator, "forward_pair", None)
        if pair is None or real.shape != fake.shape:
            return self.forward(real, is_real=True, draws=draws_real), self.forward(fake, is_real=False, draws=draws_fake)
        shape = real.shape
        a = self.augmentation_pipeline(real.flatten(start_dim=1, end_dim=2), self.p, draws_real).reshape(shape)
        b = self.augmentation_pipeline(fake.flatten(start_dim=1, end_dim=2), self.p, draws_fake).reshape(shape)
        out_real, out_fake = pair(a, b)
        self.r.append(self._calc_r(out_fake[0].detach(), out_fake[1].detach()))
        self._update_p()
        return out_real, out_fake

    def forward(self, images: torch.Tensor, is_real: bool = False, is_cut_mix: bool = False,
                draws: Optional[Dict[str, object]] = None):
        if is_cut_mix:
            return self.discriminator(images)
        original_shape = images.shape
        flat = images.flatten(start_dim=1, end_dim=2)
        flat = self.augmentation_pipeline(flat, self.p, draws)
        prediction_scalar, prediction_pixel_wise = self.discriminator(flat.reshape(original_shape))
        if not is_real:
            self.r.append(self._calc_r(prediction_scalar.detach(), prediction_pixel_wise.detach()))
        self._update_p()
        return prediction_scalar, prediction_pixel_wise
