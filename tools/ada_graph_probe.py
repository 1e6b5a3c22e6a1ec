"""Diagnostic: eager vs CUDA-graph replay of the train step with ADA — logs every plan that is drawn."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from multi_stylegan_b200 import adaptive_discriminator_augmentation as A
from multi_stylegan_b200.model_wrapper import ModelWrapper
from tests.test_train_step import _hp, build

dev = torch.device("cuda:0")
hp = _hp()
hp["p_mixed_noise"] = 0.5
log = []
orig = A.build_plan


def logged(d, B, H, W):
    p = orig(d, B, H, W)
    log.append(round(float(p.double().abs().sum()), 4))
    return p


A.build_plan = logged
for graphed in (False, True):
    G, D = build(dev)
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"], fused=True, capturable=True)
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"], fused=True, capturable=True)
    ada = A.AdaptiveDiscriminatorAugmentation(D, r_update=2, p_step=0.1)
    ada.p = 0.5
    mw = ModelWrapper(G, ada, opt_g, opt_d, hyperparameters=hp, device=dev, cuda_graphs=True)
    mw._d_params = lambda: list(D.parameters())
    random.seed(7), np.random.seed(7), torch.manual_seed(7)
    gen = torch.Generator().manual_seed(11)
    for it in range(int(os.environ.get("ITERS", "8"))):
        if not graphed:
            mw._graphs.clear()
        log.clear()
        real = torch.rand(4, 2, 3, 32, 32, generator=gen).to(dev)
        p0 = ada.p
        out = mw.train_step(real)
        torch.cuda.synchronize()
        print("graphed=%s it=%d p=%.2f->%.2f replays=%d plans=%s real=%.5f fake=%.5f g=%.5f rsum=%s cnt=%d" % (
            graphed, it, p0, ada.p, mw.graph_replays, log, float(out["loss_discriminator_real"]),
            float(out["loss_discriminator_fake"]), float(out["loss_generator"]),
            None if ada._r_sum is None else round(float(ada._r_sum), 4), ada._r_count))
