"""Oracle restatement of one iteration of the reference trainer (model_wrapper.py:253-451) on reference-named
parameter dicts, CPU fp32.  Epoch-0 behaviour by default (no CutMix, no wrong-order fakes, no top-k, no trap weights);
`late` switches the late-epoch branches on with every random draw injected:
    late = {"perm": frame permutation (:268-273), "n_wrong": number of wrong-order reals appended to the fakes,
            "cut_mix": (binary map of the augmentation step, binary map of the consistency step) (:331-376,
                        u_net_2d_discriminator.py:384-448), "top_k_v": fraction kept in the generator step (:392-401,
                        loss.py:420-444), "trap": pixel-wise loss weight map or None (:262-263,:404-405)}
TEST INFRASTRUCTURE ONLY."""
import math
from typing import Dict, List

import torch

from oracle import model as om


def _clip_(params: List[torch.Tensor], grads: List[torch.Tensor], max_norm: float = 5.0):
    """torch.nn.utils.clip_grad_norm_ (model_wrapper.py:296,323,410,438) on explicit grads."""
    total = torch.sqrt(sum(g.pow(2).sum() for g in grads if g is not None))
    coef = (max_norm / (total + 1e-6)).clamp(max=1.0)
    return [None if g is None else g * coef for g in grads]


class OracleTrainer:
    def __init__(self, sd_g: Dict[str, torch.Tensor], sd_d: Dict[str, torch.Tensor], g_groups, lr_d: float, betas, hp,
                 dead_branch: bool = False):
        self.dead_branch = dead_branch
        self.sd_g = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.startswith("noises.")
                                                 and not k.endswith(".kernel")) for k, v in sd_g.items()}
        self.sd_d = {k: v.clone().requires_grad_(not k.endswith(".kernel")) for k, v in sd_d.items()}
        self.g_names = [k for k, v in self.sd_g.items() if v.requires_grad]
        self.d_names = [k for k, v in self.sd_d.items() if v.requires_grad]
        # same parameter groups as Generator.get_parameters (lr_style for the mapping network)
        self.opt_g = torch.optim.Adam([{"params": [self.sd_g[k] for k in self.g_names if not k.startswith("style_mapping.")], "lr": g_groups[0]},
                                       {"params": [self.sd_g[k] for k in self.g_names if k.startswith("style_mapping.")], "lr": g_groups[1]}],
                                      betas=betas)
        self.opt_d = torch.optim.Adam([self.sd_d[k] for k in self.d_names], lr=lr_d, betas=betas)
        self.ema = {k: v.detach().clone() for k, v in self.sd_g.items()}
        self.hp = hp
        self.mean_pl = torch.zeros(1)
        self.iteration = 0

    def _apply(self, sd, names, loss, opt):
        params = [sd[k] for k in names]
        grads = torch.autograd.grad(loss, params, allow_unused=True)
        grads = _clip_(params, list(grads))
        for p, g in zip(params, grads):
            p.grad = g
        opt.step()
        for p in params:
            p.grad = None

    def step(self, real, z_d, z_g, z_pl, noise_d, noise_g, noise_pl, inject, pl_noise, late=None):
        import torch.nn.functional as F
        hp, out = self.hp, {}
        late = late or {}
        trap = late.get("trap")

        def weighted(v):                                     # loss.py:116-131,146-170 with `weight`
            return v if trap is None else v * trap.view(1, 1, 1, trap.shape[-2], trap.shape[-1])
        self.iteration += 1
        with torch.no_grad():
            fake = om.generator_forward({k: v.detach() for k, v in self.sd_g.items()}, z_d, noise_d, inject,
                                        dead_branch=self.dead_branch)
            if late.get("n_wrong"):                          # model_wrapper.py:268-273
                fake = torch.cat([fake, real[:late["n_wrong"]][:, :, late["perm"]]], dim=0)
        rs, rp = om.discriminator_forward(self.sd_d, real)
        fs, fp = om.discriminator_forward(self.sd_d, fake)
        lr_, lf_ = om.ns_discriminator_loss(rs, fs)
        lrp, lfp = weighted(F.softplus(-rp)).mean(), weighted(F.softplus(fp)).mean()
        out.update(loss_discriminator_real=lr_.detach(), loss_discriminator_fake=lf_.detach(),
                   loss_discriminator_real_pixel_wise=lrp.detach(), loss_discriminator_fake_pixel_wise=lfp.detach())
        self._apply(self.sd_d, self.d_names, lr_ + lf_ + lrp + lfp, self.opt_d)
        if self.iteration % hp["lazy_discriminator_regularization"] == 0:
            r1 = om.r1_penalty(self.sd_d, real)
            out["loss_discriminator_regularization"] = r1.detach()
            self._apply(self.sd_d, self.d_names, hp["w_discriminator_regularization_r1"] * r1, self.opt_d)
        if late.get("cut_mix") is not None:                  # model_wrapper.py:331-376
            w = hp["w_discriminator_regularization"]
            map_a, map_b = late["cut_mix"]
            n = real.shape[0]
            mixed = real * map_a + fake[:n] * (1.0 - map_a)                              # u_net_2d_discriminator.py:395-401
            _, pred = om.discriminator_forward(self.sd_d, mixed)
            cm = (F.softplus(-pred) * map_a).mean() + (F.softplus(pred) * (1.0 - map_a)).mean()   # loss.py:189-195
            out["loss_cut_mix_augmentation"] = cm.detach()
            self._apply(self.sd_d, self.d_names, w * cm, self.opt_d)
            mixed = real * map_b + fake[:n] * (1.0 - map_b)                              # :416-425 (predictions of the D step)
            target = rp.detach() * map_b + fp.detach()[:n] * (1.0 - map_b)
            _, pred = om.discriminator_forward(self.sd_d, mixed)
            reg = F.mse_loss(pred, target)
            out["loss_cut_mix_regularization"] = reg.detach()
            self._apply(self.sd_d, self.d_names, w * reg, self.opt_d)
        fake = om.generator_forward(self.sd_g, z_g, noise_g, inject, dead_branch=self.dead_branch)
        fs, fp = om.discriminator_forward(self.sd_d, fake)
        if late.get("top_k_v") is not None:                  # model_wrapper.py:392-401, loss.py:436-443
            flat = fs.view(-1)
            fs, idx = torch.topk(flat, k=max(1, int(flat.shape[0] * late["top_k_v"])))
            fp = fp[idx]
        lg, lgp = om.ns_generator_loss(fs), weighted(F.softplus(-fp)).mean()
        out.update(loss_generator=lg.detach(), loss_generator_pixel_wise=lgp.detach())
        self._apply(self.sd_g, self.g_names, lg + lgp, self.opt_g)
        if self.iteration % hp["lazy_generator_regularization"] == 0:
            n_lat = sum(1 for k in self.sd_g if k.startswith("main_convolutions_1.") and k.endswith("modulated_convolution.weight")) + 2
            latent = om.generator_latent(self.sd_g, z_pl, inject, n_lat)
            image = om.generator_forward(self.sd_g, latent=latent, noise=noise_pl)
            scale = math.sqrt(image.shape[2] * image.shape[3] * image.shape[4])
            g = torch.autograd.grad((image * (pl_noise / scale)).sum(), latent, create_graph=True)[0]
            pen, pl, self.mean_pl = om.path_length_penalty(g, self.mean_pl.detach())
            out.update(path_length=pl.detach(), loss_path_length_regularization=pen.detach())
            self._apply(self.sd_g, self.g_names, hp["w_generator_regularization"] * pen, self.opt_g)
        with torch.no_grad():      # misc.py:195-199 — parameters only, buffers are not averaged
            for k in self.g_names:
                self.ema[k].mul_(0.999).add_(self.sd_g[k].detach(), alpha=0.001)
        return out
