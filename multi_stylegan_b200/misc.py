"""EMA, mixed-noise sampler and the with-replacement frame permutation of multi_stylegan/misc.py:183-252."""
import random
from typing import List, Union

import numpy as np
import torch
import torch.nn as nn


@torch.no_grad()
def exponential_moving_average(model_ema: nn.Module, model_train: nn.Module, decay: float = 0.999) -> None:
    """p_ema <- decay * p_ema + (1 - decay) * p over named parameters (buffers are not averaged),
    as ONE multi-tensor pass (lerp) instead of the reference's per-tensor Python loop."""
    assert type(model_ema) is type(model_train), "EMA can only be performed on networks of the same type!"
    ema = dict(model_ema.named_parameters())
    train = dict(model_train.named_parameters())
    keys = list(ema.keys())
    dst = [ema[k].data for k in keys]
    src = [train[k].data for k in keys]
    # one pass: p_ema + (1 - decay) * (p - p_ema)  ==  decay * p_ema + (1 - decay) * p
    torch._foreach_lerp_(dst, src, 1 - decay)


def random_permutation(n: int) -> torch.Tensor:
    permutation = torch.from_numpy(np.random.choice(range(n), size=n))
    if torch.equal(permutation, torch.arange(n)):
        permutation = torch.arange(start=n - 1, end=-1, step=-1)
    return permutation


def get_noise(batch_size: int, latent_dimension: int, p_mixed_noise: float = 0.9,
              device: Union[str, torch.device] = "cuda") -> Union[torch.Tensor, List[torch.Tensor]]:
    if (p_mixed_noise > 0) and (random.random() < p_mixed_noise):
        return list(torch.randn(2, batch_size, latent_dimension, dtype=torch.float32, device=device).unbind(0))
    return torch.randn(batch_size, latent_dimension, dtype=torch.float32, device=device)
