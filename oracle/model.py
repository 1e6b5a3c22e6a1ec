"""Oracle for the generator, discriminator and regularisers: a *functional* CPU restatement driven by a
reference-named state_dict.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines it follows (paths under multi_stylegan/)."""
import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

from . import ops

SD = Dict[str, torch.Tensor]
SQRT2 = math.sqrt(2.0)


# ---- layers ----------------------------------------------------------------------------------------
def eq_linear(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """equalized_layer.py:225-254: weight*sqrt(2)/sqrt(in), bias*sqrt(2)/sqrt(out)."""
    w = sd[p + ".weight"]
    b = sd.get(p + ".bias")
    return F.linear(x, w * (SQRT2 / math.sqrt(w.shape[1])), None if b is None else b * (SQRT2 / math.sqrt(w.shape[0])))


def eq_conv2d(sd: SD, p: str, x: torch.Tensor, stride=1, padding=0) -> torch.Tensor:
    """equalized_layer.py:42-44,70-73."""
    w = sd[p + ".weight"]
    b = sd.get(p + ".bias")
    scale = SQRT2 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3])
    return F.conv2d(x, w * scale, None if b is None else b * (SQRT2 / math.sqrt(w.shape[0])), stride=stride,
                    padding=padding)


def lrelu_bias(x: torch.Tensor, b: torch.Tensor, slope: float = 0.2, scale: float = 1.0) -> torch.Tensor:
    """FusedLeakyReLU module: op_static/fused_act.py:76-85 (module gain 1.0) over kernel.cu:25-48."""
    return F.leaky_relu(x + b.view([1, -1] + [1] * (x.dim() - 2)), slope) * scale


def fir(x: torch.Tensor, kernel: torch.Tensor, up: int = 1, down: int = 1, pad: Tuple[int, int] = (0, 0)):
    """op_static/upfirdn2d.py:148-153 on NCHW (viewed as [B*C,H,W,1], :102)."""
    B, C, H, W = x.shape
    y = ops.upfirdn2d(x.reshape(B * C, H, W, 1), kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    return y.view(B, C, y.shape[1], y.shape[2])


def blur_kernel_2d(taps: Sequence[int] = (1, 3, 3, 1), gain: float = 1.0) -> torch.Tensor:
    """multi_stylegan_generator.py:619-632 (and :601-602 for the gain)."""
    k = torch.tensor(list(taps), dtype=torch.float32)
    k = k[None, :] * k[:, None]
    return k / k.sum() * gain


# ---- generator -------------------------------------------------------------------------------------
def modulated_conv(sd: SD, p: str, x: torch.Tensor, style: torch.Tensor, demodulate: bool, upsampling: bool):
    """multi_stylegan_generator.py:365-414.  Returns (out, modulated_style); `style` is the latent slice
    when the layer owns a modulation linear, else the already-modulated style of its twin."""
    W = sd[p + ".weight"]                                  # [1, O, C, kh, kw]
    _, O, C, kh, kw = W.shape
    B = x.shape[0]
    if (p + ".modulation_mapping.weight") in sd:
        s = eq_linear(sd, p + ".modulation_mapping", style).view(B, 1, C, 1, 1)      # :380
    else:
        s = style                                                                     # :382
    w = (SQRT2 / math.sqrt(C * kh * kw)) * W * s                                      # :335,:384
    if demodulate:
        w = w * torch.rsqrt((w ** 2).sum(dim=[2, 3, 4]) + 1e-8).view(B, O, 1, 1, 1)   # :386-388
    if upsampling:
        y = ops.conv_transpose2d(x, w.transpose(1, 2), stride=2, padding=0)           # :393-401
        k = blur_kernel_2d(gain=4.0)                                                  # :325,:601-602
        pf = (4 - 2) + (kh - 1)                                                       # :615-616
        y = fir(y, k.to(y), pad=((pf + 1) // 2, pf // 2))                       # :403
    else:
        y = ops.conv2d(x, w, stride=1, padding=(kh // 2, kw // 2))                    # :406-411
    return y, s


def styled_conv(sd: SD, p: str, x, style, noise, upsampling: bool):
    """multi_stylegan_generator.py:452-469 (+ :281-292 noise, fused_act.py:84-85 activation)."""
    y, s = modulated_conv(sd, p + ".modulated_convolution", x, style, True, upsampling)
    if noise is None:
        noise = torch.randn(y.shape[0], 1, y.shape[2], y.shape[3], dtype=torch.float32, device=y.device)
    y = y + sd[p + ".noise_injection.weight"] * noise
    return lrelu_bias(y, sd[p + ".activation.bias"]), s


def output_block(sd: SD, p: str, x, style, skip=None):
    """multi_stylegan_generator.py:504-526; Upsample taps without gain, pad (2,1) (:545-551)."""
    y, s = modulated_conv(sd, p + ".modulated_convolution", x, style, False, False)
    y = y + sd[p + ".bias"]
    if skip is not None:
        y = y + fir(skip, blur_kernel_2d().to(skip), up=2, pad=(2, 1))
    return y, s


def style_mapping(sd: SD, z: torch.Tensor) -> torch.Tensor:
    """multi_stylegan_generator.py:222-235, equalized_layer.py:276."""
    x = z / torch.sqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    i = 1
    while "style_mapping.layers.%d.weight" % i in sd:
        x = eq_linear(sd, "style_mapping.layers.%d" % i, x)
        x = lrelu_bias(x, sd["style_mapping.layers.%d.bias" % (i + 1)])
        i += 2
    return x


def generator_latent(sd: SD, z: Union[torch.Tensor, List[torch.Tensor]], inject_index: Optional[int],
                     n_latent: int) -> torch.Tensor:
    """multi_stylegan_generator.py:134-138,152-160."""
    if isinstance(z, list):
        s0, s1 = style_mapping(sd, z[0]), style_mapping(sd, z[1])
        return torch.cat([s0.unsqueeze(1).repeat(1, inject_index, 1),
                          s1.unsqueeze(1).repeat(1, n_latent - inject_index, 1)], dim=1)
    return style_mapping(sd, z).unsqueeze(1).repeat(1, n_latent, 1)


def generator_forward(sd: SD, z=None, noise: Optional[List[torch.Tensor]] = None,
                      inject_index: Optional[int] = None, latent: Optional[torch.Tensor] = None,
                      dead_branch: bool = False) -> torch.Tensor:
    """multi_stylegan_generator.py:114-191.  `noise` = [noise_start] + one map per main conv (:148-150).
    The second branch's main convolutions never reach the image (:184,187,189) and are skipped unless
    dead_branch=True (then they are evaluated and discarded, as the reference does)."""
    n_main = 0
    while "main_convolutions_1.%d.modulated_convolution.weight" % n_main in sd:
        n_main += 1
    if latent is None:
        latent = generator_latent(sd, z, inject_index, n_main + 2)
    B = latent.shape[0]
    noise = [None] * (n_main + 1) if noise is None else noise
    x1 = sd["constant_input_1.input"].repeat_interleave(B, dim=0)                      # :174
    x2 = sd["constant_input_2.input"].repeat_interleave(B, dim=0)                      # :175
    x1, s = styled_conv(sd, "starting_convolution_1", x1, latent[:, 0], noise[0], False)   # :176
    x2, _ = styled_conv(sd, "starting_convolution_2", x2, s, noise[0], False)          # :177
    skip1, s = output_block(sd, "starting_output_block_1", x1, latent[:, 1])           # :178
    skip2, _ = output_block(sd, "starting_output_block_2", x2, s)                      # :179
    for i in range(n_main // 2):                                                       # :181-189
        x1, s = styled_conv(sd, "main_convolutions_1.%d" % (2 * i), x1, latent[:, 2 * i + 1], noise[1 + 2 * i], True)
        if dead_branch:
            x2, _ = styled_conv(sd, "main_convolutions_2.%d" % (2 * i), x2, s, noise[1 + 2 * i], True)
        x1, s = styled_conv(sd, "main_convolutions_1.%d" % (2 * i + 1), x1, latent[:, 2 * i + 2], noise[2 + 2 * i], False)
        if dead_branch:
            x2, _ = styled_conv(sd, "main_convolutions_2.%d" % (2 * i + 1), x2, s, noise[2 + 2 * i], False)
        skip1, s = output_block(sd, "output_blocks_1.%d" % i, x1, latent[:, 2 * i + 3], skip1)
        skip2, _ = output_block(sd, "output_blocks_2.%d" % i, x1, s, skip2)            # x1, not x2 (:189)
    return torch.stack([skip1, skip2], dim=1)                                          # :191


def path_length_grads(sd: SD, latent: torch.Tensor, noise, pl_noise: torch.Tensor) -> torch.Tensor:
    """multi_stylegan_generator.py:193-200 with the random direction passed in (already divided by
    sqrt(frames*H*W))."""
    image = generator_forward(sd, latent=latent, noise=noise)
    return torch.autograd.grad((image * pl_noise).sum(), latent, create_graph=True, retain_graph=True)[0]


def path_length_penalty(grad: torch.Tensor, mean_path_length: torch.Tensor, decay: float = 0.01):
    """loss.py:388-395.  Returns (penalty, path_length, new_mean)."""
    pl = torch.sqrt(grad.pow(2).sum(2).mean(1) + 1e-8).mean()
    new_mean = mean_path_length + decay * (pl.mean() - mean_path_length)
    return torch.mean((pl - new_mean) ** 2), pl, new_mean


# ---- discriminator ---------------------------------------------------------------------------------
def minibatch_stddev(x: torch.Tensor, alpha: float = 1e-8) -> torch.Tensor:
    """u_net_2d_discriminator.py:212-217."""
    c = x - x.mean(dim=0, keepdim=True)
    s = torch.sqrt((c ** 2).mean(dim=0).clamp(min=alpha)).mean().view(1, 1, 1)
    return torch.cat([x, s.repeat(x.shape[0], 1, x.shape[2], x.shape[3])], dim=1)


def resnet_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """u_net_2d_discriminator.py:174-186."""
    w0 = sd[p + ".main_mapping.0.weight"]
    h = minibatch_stddev(x) if w0.shape[1] == x.shape[1] + 1 else x
    h = lrelu_bias(eq_conv2d(sd, p + ".main_mapping.0", h, 1, 1), sd[p + ".main_mapping.1.bias"])
    h = lrelu_bias(eq_conv2d(sd, p + ".main_mapping.2", h, 1, 1), sd[p + ".main_mapping.3.bias"])
    r = eq_conv2d(sd, p + ".residual_mapping", x) if (p + ".residual_mapping.weight") in sd else x
    return (h + r) / SQRT2


def non_local_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """u_net_2d_discriminator.py:359-381."""
    B, _, H, W = x.shape
    theta = eq_conv2d(sd, p + ".theta", x).flatten(2)
    phi = F.max_pool2d(eq_conv2d(sd, p + ".phi", x), 2, 2).flatten(2)
    g = F.max_pool2d(eq_conv2d(sd, p + ".g", x), 2, 2).flatten(2)
    beta = F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1)
    o = eq_conv2d(sd, p + ".o", torch.bmm(g, beta.transpose(1, 2)).view(B, -1, H, W))
    r = eq_conv2d(sd, p + ".residual_mapping", x) if (p + ".residual_mapping.weight") in sd else x
    return (sd[p + ".gamma"] * o + r) / SQRT2


def _block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return non_local_block(sd, p, x) if (p + ".theta.weight") in sd else resnet_block(sd, p, x)


def discriminator_forward(sd: SD, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """u_net_2d_discriminator.py:99-140: [B, 2, 3, H, W] -> ([B,1], [B,1,1,H,W])."""
    x = images.flatten(1, 2)                                                           # :124
    n_enc = 0
    while any(k.startswith("encoder_blocks.%d." % n_enc) for k in sd):
        n_enc += 1
    k1 = blur_kernel_2d()
    feats = []
    for i in range(n_enc):                                                             # :127-131
        x = _block(sd, "encoder_blocks.%d" % i, x)
        if i != n_enc - 1:
            feats.append(x)
            x = eq_conv2d(sd, "downscale_convolutions.%d.0" % i, x, stride=2, padding=0)   # :59-62
            x = fir(x, k1.to(x), pad=(2, 2))                                         # Blur() :62,:306-307
    h = x.mean(dim=(2, 3))                                                             # :66-67
    h = lrelu_bias(eq_linear(sd, "classification_head.2", h), sd["classification_head.3.bias"])
    scalar = eq_linear(sd, "classification_head.4", h)                                 # :70
    for j, skip in enumerate(reversed(feats)):                                         # :135-137
        u = fir(x, k1.to(x), up=2, pad=(2, 1))                                   # Upsample() :87,:236-242
        u = eq_conv2d(sd, "transposed_convolutions.%d.1" % j, u)
        x = _block(sd, "decoder_blocks.%d" % j, torch.cat([u, skip], dim=1))
    x = lrelu_bias(x, sd["final_mapping.0.bias"])                                      # :94-97
    pixel = eq_conv2d(sd, "final_mapping.1", x).unsqueeze(2)                           # :139
    return scalar, pixel


# ---- losses ----------------------------------------------------------------------------------------
def ns_discriminator_loss(real: torch.Tensor, fake: torch.Tensor):
    """loss.py:166-170."""
    return F.softplus(-real).mean(), F.softplus(fake).mean()


def ns_generator_loss(fake: torch.Tensor) -> torch.Tensor:
    """loss.py:128-129."""
    return F.softplus(-fake).mean()


def r1_penalty(sd: SD, real: torch.Tensor) -> torch.Tensor:
    """loss.py:311-316 with the D call of model_wrapper.py:312-315."""
    real = real.detach().requires_grad_(True)
    scalar, pixel = discriminator_forward(sd, real)
    g, = torch.autograd.grad((scalar.sum(), pixel.sum()), real, create_graph=True)
    return 0.5 * g.pow(2).view(g.shape[0], -1).sum(1).mean()
