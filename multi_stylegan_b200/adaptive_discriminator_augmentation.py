"""Adaptive discriminator augmentation (wrapper + pipeline) with the reference's interface
(multi_stylegan/adaptive_discriminator_augmentation.py:11-213).

Stage order, gates and parameter distributions follow the reference exactly, including its quirks: one
shared 90-degree rotation and one shared integer roll for all selected samples (:122, :210-211), lognormal
sigma (0.2 ln 2)^2 (:141, :176), centre = 0.5 * (H, W) passed as (x, y) (:137-138), four *sequential*
bilinear resamplings (each stage re-interpolates the previous stage's output), and the in-place update of the
caller's batch (:118-187 assign into a view of the input, so the train step's later uses of the same real /
fake tensors see the augmented images).

Execution is split in two: the host draws the per-call random decisions in the reference's order
(`sample_draws`) and packs them into ONE small tensor (`build_plan`: gate masks, the shared roll, five per-sample
2x3 inverse maps); the device applies that plan with shape-static kernels (`apply_plan`: select-flip, warp,
select-roll, four warps; samples a gate skipped get identity maps = exact copies).  Because the device side has no
data-dependent control flow, a captured CUDA graph replays it with a fresh plan copied into the same buffer
(`graph_capturable`, ModelWrapper.cuda_graphs).

kornia 0.4.1 (the reference's warp implementation, requirements.txt:7) is neither vendored nor installed,
so its matrix conventions are restated from its published source and are NOT pinned by any fixture
("parity unpinned"); tests compare this pipeline with an independent restatement built on
F.affine_grid / F.grid_sample (oracle/ada.py).  All random draws can be injected (`draws=`)."""
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
        # kornia 0.4.1 get_rotation_matrix2d: the translation column is built from the first row only
        # (alpha = m00, beta = m01); identical to "keep the centre fixed" for isotropic scales
        m[0, 2] = (1.0 - m[0, 0]) * cx - m[0, 1] * cy
        m[1, 2] = m[0, 1] * cx + (1.0 - m[0, 0]) * cy
        theta[i] = torch.from_numpy(np.linalg.inv(m)[:2])
    return theta.float()


def plan_size(batch: int) -> int:
    return 32 * batch + 2


def build_plan(d: Dict[str, object], batch: int, height: int, width: int) -> torch.Tensor:
    """Pack one call's draws into a float32 host tensor: [flip mask B | roll mask B | roll shift (y, x) | 5 x B x 6 maps]
    (maps: the shared 90-degree rotation, then iso-scale, rotation, aniso-scale, rotation; identity where not selected)."""
    B, H, W = batch, height, width
    plan = torch.zeros(plan_size(B), dtype=torch.float32)
    plan[torch.as_tensor(d["flip"], dtype=torch.long)] = 1.0
    plan[B + torch.as_tensor(d["roll"], dtype=torch.long)] = 1.0
    if d["roll"]:
        plan[2 * B] = float(int(H * d["roll_frac"][0]))
        plan[2 * B + 1] = float(int(W * d["roll_frac"][1]))
    n90 = len(d["rot90"])
    ang = float(d["rot90_angle"])
    # kaf.rotate(+angle) about the tensor centre == the affine map with angle -(-angle)
    thetas = [affine_inverse_theta(B, list(d["rot90"]), [-ang] * n90, [(1., 1.)] * n90, ((W - 1) / 2, (H - 1) / 2))]
    centre = (0.5 * H, 0.5 * W)
    stages = [(d["iso"], np.zeros(len(d["iso"])), np.repeat(np.asarray(d["iso_scale"]).reshape(-1, 1), 2, axis=1)),
              (d["rot_a"], d["rot_a_angle"], np.ones((len(d["rot_a"]), 2))),
              (d["aniso"], np.zeros(len(d["aniso"])), np.asarray(d["aniso_scale"]).reshape(-1, 2)),
              (d["rot_b"], d["rot_b_angle"], np.ones((len(d["rot_b"]), 2)))]
    for idx, angles, scales in stages:
        thetas.append(affine_inverse_theta(B, list(idx), angles, scales, centre))
    plan[2 * B + 2:] = torch.stack(thetas).reshape(-1)
    return plan


def apply_plan(images: torch.Tensor, plan: torch.Tensor) -> torch.Tensor:
    """The seven stages of the reference (:116-199) driven by a device-resident plan; no host decision, no sync."""
    B, _, H, W = images.shape
    dev = images.device
    flip = plan[:B].view(B, 1, 1, 1) > 0.5
    roll = plan[B:2 * B].view(B, 1, 1, 1) > 0.5
    shift = plan[2 * B:2 * B + 2].round().to(torch.long)
    theta = plan[2 * B + 2:].view(5, B, 2, 3)
    x = torch.where(flip, images.flip(dims=(-1,)), images)
    x = affine_warp(x, theta[0], mode=1)
    iy = torch.remainder(torch.arange(H, device=dev) - shift[0], H)            # torch.roll(x, s)[i] = x[(i - s) mod n]
    ix = torch.remainder(torch.arange(W, device=dev) - shift[1], W)
    x = torch.where(roll, x.index_select(2, iy).index_select(3, ix), x)
    for k in range(1, 5):
        x = affine_warp(x, theta[k], mode=0)
    return x


class AugmentationPipeline(nn.Module):
    def forward(self, images: torch.Tensor, p: float, draws: Optional[Dict[str, object]] = None,
                plan: Optional[torch.Tensor] = None) -> torch.Tensor:
        """images [B, C, H, W].  Returns the augmented batch and — like the reference, whose stages assign into a view
        of the caller's tensor — also leaves it in `images` when that tensor does not take part in autograd."""
        B, _, H, W = images.shape
        if plan is None:
            d = sample_draws(B, p) if draws is None else draws
            plan = build_plan(d, B, H, W).to(images.device, non_blocking=True)
        out = apply_plan(images, plan)
        if not images.requires_grad and not torch.is_inference(images):
            with torch.no_grad():
                images.copy_(out)
        return out


class AdaptiveDiscriminatorAugmentation(nn.Module):
    """Wraps a discriminator (:11-96).  `p` moves by +-p_step whenever r_update fake batches have been seen, from the mean
    of their overfitting heuristic r.  r is accumulated on the device and read back once per r_update calls (the
    reference calls .item() every time); with several ranks the accumulated value is averaged over the ranks, which is
    what the reference's DataParallel gather computes on the global batch.

    CUDA graphs (ModelWrapper.cuda_graphs): while a capture is open every pipeline call allocates a plan slot — a static
    device tensor the captured kernels read — and `refresh_plans` fills all slots of that capture with fresh draws before
    each replay; `after_replay` does the host side of the p controller."""
    graph_capturable = True

    def __init__(self, discriminator: nn.Module, r_target: float = 0.6, p_step: float = 5e-03, r_update: int = 8,
                 p_max: float = 0.8, process_group=None) -> None:
        super().__init__()
        self.discriminator = discriminator
        self.r_target, self.p_step, self.r_update, self.p_max = r_target, p_step, r_update, p_max
        self.p = 0.05
        self.r_history: List[float] = []
        self.augmentation_pipeline = AugmentationPipeline()
        self.process_group = process_group
        self._r_sum: Optional[torch.Tensor] = None      # device accumulator of r over the fake calls since the last update
        self._r_count = 0
        self._capture: Optional[dict] = None
        self._plan_event = None

    @property
    def r(self) -> List[float]:
        """Number-of-pending-values view kept for code that inspects `len(ada.r)` like the reference's list."""
        return [float("nan")] * self._r_count

    @torch.no_grad()
    def _calc_r(self, prediction_scalar: torch.Tensor, prediction_pixel_wise: torch.Tensor) -> torch.Tensor:
        """Overfitting heuristic (:51-52), kept on the device."""
        return 0.5 * torch.mean(torch.sign(prediction_scalar)) \
            + 0.5 * torch.mean(torch.sign(prediction_pixel_wise.mean(dim=(-1, -2))))

    @torch.no_grad()
    def _record_r(self, prediction_scalar: torch.Tensor, prediction_pixel_wise: torch.Tensor) -> None:
        r = self._calc_r(prediction_scalar.detach(), prediction_pixel_wise.detach()).reshape(1).float()
        if self._r_sum is None or self._r_sum.device != r.device:
            if self._capture is not None:
                raise RuntimeError("ADA: the first fake batch must be seen eagerly before an iteration is captured")
            self._r_sum = torch.zeros(1, dtype=torch.float32, device=r.device)
        self._r_sum += r
        if self._capture is not None:
            self._capture["fake_calls"] += 1             # the host half happens in after_replay
        else:
            self._r_count += 1

    def _update_p(self) -> None:
        if self._r_count >= self.r_update and self._capture is None:
            from . import dist as mdist
            total = self._r_sum.clone()
            if mdist.world_size(self.process_group) > 1:
                mdist.all_reduce_mean_(total, self.process_group)
            r = float(total.item()) / self._r_count           # the only host sync: once per r_update fake calls
            self.p = self.p + self.p_step if r > self.r_target else self.p - self.p_step
            self.p = min(max(self.p, 0.), self.p_max)
            self._r_sum.zero_()
            self._r_count = 0
            self.r_history.append(r)

    # ---- CUDA-graph support ---------------------------------------------------------------------------
    def begin_plan_capture(self, max_batch: int, device, max_calls: int = 8) -> None:
        """Call BEFORE the capture opens: the plan tensors are allocated here, outside the graph's private memory pool
        (they are written by eager copies between replays)."""
        bank = [torch.zeros(plan_size(max_batch), dtype=torch.float32, device=device) for _ in range(max_calls)]
        self._capture = {"slots": [], "fake_calls": 0, "bank": bank}

    def end_plan_capture(self) -> dict:
        cap, self._capture = self._capture, None
        cap.pop("bank", None)
        return cap

    def refresh_plans(self, cap: dict) -> None:
        """Fresh draws (reference order, current p) for every pipeline call of a captured iteration, copied into the
        static plan tensors on the current stream; call right before replaying the graphs of `cap`."""
        if not cap["slots"]:
            return
        if self._plan_event is not None:
            self._plan_event.synchronize()       # the previous upload has left the pinned staging buffers
        for slot in cap["slots"]:
            B, H, W = slot["shape"]
            if slot["host"] is None:
                slot["host"] = torch.empty(plan_size(B), dtype=torch.float32).pin_memory()
            slot["host"].copy_(build_plan(sample_draws(B, self.p), B, H, W))
            slot["device"].copy_(slot["host"], non_blocking=True)
        if self._plan_event is None:
            self._plan_event = torch.cuda.Event()
        self._plan_event.record()

    def after_replay(self, cap: dict) -> None:
        self._r_count += cap["fake_calls"]
        self._update_p()

    def _augment(self, flat: torch.Tensor, draws) -> torch.Tensor:
        if self._capture is None:
            return self.augmentation_pipeline(flat, self.p, draws)
        B, _, H, W = flat.shape
        # (the pinned staging buffer is allocated by the first refresh: no host allocation while a capture is open)
        bank = self._capture["bank"]
        if not bank or bank[-1].numel() < plan_size(B) or bank[-1].device != flat.device:
            raise RuntimeError("ADA: no pre-allocated plan tensor for a batch of %d on %s" % (B, flat.device))
        slot = {"shape": (B, H, W), "device": bank.pop()[:plan_size(B)], "host": None}
        self._capture["slots"].append(slot)
        return self.augmentation_pipeline(flat, self.p, plan=slot["device"])

    def forward_pair(self, real: torch.Tensor, fake: torch.Tensor, draws_real=None, draws_fake=None):
        """forward(real, is_real=True) and forward(fake, is_real=False) with one batched discriminator pass: both halves
        are augmented separately (real first, the order in which the reference consumes the host RNGs), `p` cannot change
        between the two calls of a step (it only moves after a fake call), so the results are those of the two calls."""
        pair = getattr(self.discriminator, "forward_pair", None)
        if pair is None or real.shape != fake.shape:
            return self.forward(real, is_real=True, draws=draws_real), self.forward(fake, is_real=False, draws=draws_fake)
        shape = real.shape
        a = self._augment(real.flatten(start_dim=1, end_dim=2), draws_real).reshape(shape)
        b = self._augment(fake.flatten(start_dim=1, end_dim=2), draws_fake).reshape(shape)
        out_real, out_fake = pair(a, b)
        self._record_r(out_fake[0], out_fake[1])
        self._update_p()
        return out_real, out_fake

    def forward(self, images: torch.Tensor, is_real: bool = False, is_cut_mix: bool = False,
                draws: Optional[Dict[str, object]] = None):
        if is_cut_mix:
            return self.discriminator(images)
        original_shape = images.shape
        flat = self._augment(images.flatten(start_dim=1, end_dim=2), draws)
        prediction_scalar, prediction_pixel_wise = self.discriminator(flat.reshape(original_shape))
        if not is_real:
            self._record_r(prediction_scalar, prediction_pixel_wise)
        self._update_p()
        return prediction_scalar, prediction_pixel_wise
