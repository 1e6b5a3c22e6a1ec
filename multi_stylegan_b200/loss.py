"""Losses and lazy regularisers of the train step (multi_stylegan/loss.py:97-195, 283-317, 353-395).
All are a handful of scalar-sized torch ops; the heavy lifting is the double backward through the
custom kernels (R1 through D, path length through G)."""
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import autograd


def _weighted(value: torch.Tensor, weight: Optional[torch.Tensor]) -> torch.Tensor:
    if weight is None:
        return value
    return value * weight.view(1, 1, 1, weight.shape[-2], weight.shape[-1]).to(value.device)


class NonSaturatingLogisticGeneratorLoss(nn.Module):
    def forward(self, prediction_fake: torch.Tensor, weight: torch.Tensor = None) -> torch.Tensor:
        return torch.mean(_weighted(F.softplus(-prediction_fake), weight))


class NonSaturatingLogisticDiscriminatorLoss(nn.Module):
    def forward(self, prediction_real: torch.Tensor, prediction_fake: torch.Tensor,
                weight: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        return (torch.mean(_weighted(F.softplus(-prediction_real), weight)),
                torch.mean(_weighted(F.softplus(prediction_fake), weight)))


class NonSaturatingLogisticDiscriminatorLossCutMix(nn.Module):
    def forward(self, prediction: torch.Tensor, label: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return (torch.mean(F.softplus(-prediction) * label),
                torch.mean(F.softplus(prediction) * (-label + 1.)))


class R1Regularization(nn.Module):
    """0.5 * mean_b ||d(sum scalar + sum pixel-wise)/d image||^2 — loss.py:302-317."""

    def forward(self, prediction_real: torch.Tensor, image_real: torch.Tensor,
                prediction_real_pixel_wise: Optional[torch.Tensor] = None) -> torch.Tensor:
        outputs = prediction_real.sum() if prediction_real_pixel_wise is None \
            else (prediction_real.sum(), prediction_real_pixel_wise.sum())
        grad_real, = autograd.grad(outputs=outputs, inputs=image_real, create_graph=True)
        return 0.5 * grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


class PathLengthRegularization(nn.Module):
    """(l - running mean)^2 with l = mean_b sqrt(mean_k sum_d g^2 + 1e-8) collapsed to a scalar before
    the EMA (loss.py:378-395); `mean_path_length` is a plain attribute exactly as in the reference."""

    def __init__(self, decay: float = 0.01) -> None:
        super().__init__()
        self.decay = decay
        self.mean_path_length = torch.zeros(1, dtype=torch.float)

    def forward(self, grad: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        mean = self.mean_path_length.detach().to(grad.device)
        path_lengths = torch.sqrt(grad.pow(2).sum(2).mean(1) + 1e-08).mean()
        mean = mean + self.decay * (path_lengths.mean() - mean)
        penalty = torch.mean((path_lengths - mean) ** 2)        # the gradient flows through the updated mean (:392-394)
        # The reference keeps the attribute attached to the graph until the next call detaches it (:385).  The values
        # are the same with the attribute detached right away, and the generator's whole double-backward graph — and
        # with it the parameters' AccumulateGrad nodes, which are bound to the stream they were created on and would
        # break a later CUDA-graph capture — is released as soon as the penalty's backward has run.
        self.mean_path_length = mean.detach()
        return penalty, path_lengths


class TopK(nn.Module):
    """Top-k filtering of the fake predictions in the generator step — loss.py:398-444."""

    def __init__(self, starting_iteration: int, final_iteration: int) -> None:
        super().__init__()
        self.starting_iteration, self.final_iteration, self.iterations = starting_iteration, final_iteration, 0

    def calc_v(self) -> float:
        self.iterations += 1
        if self.iterations <= self.starting_iteration:
            return 1.
        if self.iterations >= self.final_iteration:
            return 0.5
        return 0.5 * (1. - float(self.iterations - self.starting_iteration)
                      / float(self.final_iteration - self.starting_iteration)) + 0.5

    def forward(self, input: torch.Tensor):
        v = self.calc_v()
        flat = input.view(-1)
        return torch.topk(flat, k=max(1, int(flat.shape[0] * v)))
