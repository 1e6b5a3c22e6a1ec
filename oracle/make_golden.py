"""Generate tests/golden/*.pt from the UNMODIFIED reference (CPU).  TEST INFRASTRUCTURE ONLY.

Run in the authoring container (needs /root/reference):  python -m oracle.make_golden
The fixtures are small seeded input/output/gradient sets of the reference's own modules; they travel
to the GPU box, where /root/reference does not exist."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

TINY_G = dict(channels=(16, 16, 16, 16), channel_factor=1, latent_dimensions=16, depth_style_mapping=2,
              starting_resolution=(4, 4))
TINY_D = dict(encoder_channels=((3, 8), (8, 16), (16, 24), (24, 48), (48, 64)),
              decoder_channels=((64, 48), (48, 24), (24, 16), (16, 8)), fft=False)


def randomize(module: torch.nn.Module, seed: int) -> None:
    """Reference init leaves biases / noise weights / gamma at 0, which would hide epilogue bugs."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("noise_injection.weight"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            elif name.endswith("gamma"):
                p.fill_(0.7)
            elif name.endswith("bias") and "modulation_mapping" not in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)


def fir_cases():
    # (in_h, in_w, kernel taps, gain, up, down, pad_x0, pad_x1, pad_y0, pad_y1)
    k4 = [1, 3, 3, 1]
    return [
        (8, 8, k4, 4.0, 1, 1, 2, 1, 2, 1),      # G blur fwd
        (8, 8, k4, 4.0, 1, 1, 1, 2, 1, 2),      # G blur bwd
        (7, 7, k4, 1.0, 1, 1, 2, 2, 2, 2),      # D blur fwd 127->128 analogue
        (8, 8, k4, 1.0, 1, 1, 1, 1, 1, 1),      # D blur bwd
        (6, 5, k4, 1.0, 2, 1, 2, 1, 2, 1),      # upsample fwd
        (12, 10, k4, 1.0, 1, 2, 1, 1, 1, 1),    # upsample bwd
        (9, 6, [1, 2, 1], 1.0, 1, 1, 1, 1, 1, 1),   # 3x3 taps
        (9, 6, k4, 1.0, 1, 1, -1, 2, 0, -1),    # negative pads crop
        (5, 7, [1, 1], 1.0, 2, 1, 1, 0, 1, 0),  # 2x2 taps up2
        (10, 12, [1, 1], 1.0, 1, 2, 0, 0, 0, 0),    # 2x2 taps down2
        (33, 70, k4, 1.0, 1, 1, 2, 1, 2, 1),    # spans several tiles
        (16, 16, [1, 4, 6, 4, 1], 1.0, 3, 2, 3, 2, 1, 4),   # outside the reference's CUDA modes
    ]


def golden_ops(ref):
    native = sys.modules["multi_stylegan.op_static.upfirdn2d"].upfirdn2d_native
    out = {"fir": []}
    g = torch.Generator().manual_seed(11)
    for (h, w, taps, gain, up, down, px0, px1, py0, py1) in fir_cases():
        k = torch.tensor(taps, dtype=torch.float32)
        k = k[None, :] * k[:, None]
        k = k / k.sum() * gain
        k = k + 0.01 * torch.randn(k.shape, generator=g)      # asymmetric taps: catches a missing flip
        x = torch.randn(3, h, w, 1, generator=g)
        y = native(x, k, up, up, down, down, px0, px1, py0, py1)
        out["fir"].append(dict(x=x, k=k, cfg=(up, up, down, down, px0, px1, py0, py1), y=y))
    # the reference's autograd wrappers (fwd, grad, grad-of-grad) through the stand-in modules
    up_fn = ref.op_static.upfirdn2d
    out["fir_autograd"] = []
    for (up, down, pad, gain) in [(1, 1, (2, 1), 4.0), (2, 1, (2, 1), 1.0), (1, 1, (2, 2), 1.0), (1, 2, (1, 1), 1.0)]:
        k = torch.tensor([1., 3., 3., 1.])
        k = k[None, :] * k[:, None]
        k = k / k.sum() * gain
        x = torch.randn(2, 3, 9 if (pad == (2, 2)) else 8, 8, generator=g).requires_grad_(True)
        y = up_fn(x, k, up=up, down=down, pad=pad)
        gy = torch.randn(y.shape, generator=g).requires_grad_(True)
        gx, = torch.autograd.grad(y, x, gy, create_graph=True)
        v = torch.randn(gx.shape, generator=g)
        ggy, = torch.autograd.grad(gx, gy, v)
        out["fir_autograd"].append(dict(x=x.detach(), k=k, up=up, down=down, pad=pad, y=y.detach(), gy=gy.detach(),
                                        gx=gx.detach(), v=v, ggy=ggy))
    # FusedLeakyReLU module (gain 1.0) and fused_leaky_relu (gain sqrt 2): fwd, grads, grad-of-grad
    out["lrelu"] = []
    for shape, scale in [((2, 5, 6, 7), 1.0), ((3, 4), 2 ** 0.5), ((2, 3, 40, 40), 1.0)]:
        x = torch.randn(shape, generator=g).requires_grad_(True)
        b = (torch.randn(shape[1], generator=g) * 0.5).requires_grad_(True)
        y = ref.op_static.fused_leaky_relu(x, b, 0.2, scale)
        gy = torch.randn(y.shape, generator=g).requires_grad_(True)
        gx, gb = torch.autograd.grad(y, (x, b), gy, create_graph=True)
        v = torch.randn(gx.shape, generator=g)
        vb = torch.randn(gb.shape, generator=g)
        ggy, = torch.autograd.grad((gx, gb), gy, (v, vb))
        out["lrelu"].append(dict(x=x.detach(), b=b.detach(), scale=scale, y=y.detach(), gy=gy.detach(),
                                 gx=gx.detach(), gb=gb.detach(), v=v, vb=vb, ggy=ggy))
    return out


def golden_block(ref):
    """One dual-style block pair (3x3 and the 2x2 up variant): forward and first-order grads."""
    G = ref.generator
    out = []
    gen = torch.Generator().manual_seed(5)
    for up in (False, True):
        torch.manual_seed(3 + int(up))
        a = G.StyledConv2d(8, 12, (2, 2) if up else (3, 3), 16, upsampling=up)
        b = G.StyledConv2d(8, 12, (2, 2) if up else (3, 3), 16, upsampling=up, modulation_mapping=False)
        randomize(a, 1)
        randomize(b, 2)
        x1 = torch.randn(2, 8, 8, 8, generator=gen).requires_grad_(True)
        x2 = torch.randn(2, 8, 8, 8, generator=gen).requires_grad_(True)
        w = torch.randn(2, 16, generator=gen).requires_grad_(True)
        r = 16 if up else 8
        noise = torch.randn(2, 1, r, r, generator=gen)
        y1, s = a(x1, w, noise=noise)
        y2 = b(x2, s, noise=noise)
        g1 = torch.randn(y1.shape, generator=gen)
        g2 = torch.randn(y2.shape, generator=gen)
        params = list(a.parameters()) + list(b.parameters())
        grads = torch.autograd.grad((y1 * g1).sum() + (y2 * g2).sum(), [x1, x2, w] + params)
        out.append(dict(up=up, sd_a={k: v.clone() for k, v in a.state_dict().items()},
                        sd_b={k: v.clone() for k, v in b.state_dict().items()},
                        x1=x1.detach(), x2=x2.detach(), w=w.detach(), noise=noise, y1=y1.detach(), y2=y2.detach(),
                        s=s.detach(), g1=g1, g2=g2, gx1=grads[0], gx2=grads[1], gw=grads[2],
                        gparams_a={n: g for (n, _), g in zip(a.named_parameters(), grads[3:])},
                        gparams_b={n: g for (n, _), g in
                                   zip(b.named_parameters(), grads[3 + len(list(a.parameters())):])}))
    return out


def golden_generator(ref):
    torch.manual_seed(0)
    G = ref.generator.Generator(TINY_G)
    randomize(G, 7)
    gen = torch.Generator().manual_seed(1)
    z = [torch.randn(2, 16, generator=gen), torch.randn(2, 16, generator=gen)]
    noise = [torch.randn(2, 1, 4, 4, generator=gen)] + \
            [torch.randn(2, 1, 2 ** (i // 2 + 3), 2 ** (i // 2 + 3), generator=gen) for i in range(6)]
    image = G(z, noise=noise, inject_index=3)
    direction = torch.randn(image.shape, generator=gen)
    G.zero_grad()
    (image * direction).sum().backward()
    grads = {n: p.grad.clone() for n, p in G.named_parameters() if p.grad is not None}
    none_grads = [n for n, p in G.named_parameters() if p.grad is None]
    # single-style, fixed-noise buffers
    z1 = torch.randn(2, 16, generator=gen)
    with torch.no_grad():
        image_fixed = G(z1, randomize_noise=False)
        image_lat, latent = G(z1, noise=noise, return_main_style_vectors=True)
    # path-length: reference draws its direction inside forward (:195); seed torch's global RNG
    torch.manual_seed(123)
    pl_grad = G(z1, noise=noise, return_path_length_grads=True)
    torch.manual_seed(123)
    pl_noise = torch.randn(image.shape) / (3 * 32 * 32) ** 0.5
    plr = ref.loss.PathLengthRegularization()
    penalty, pl = plr(pl_grad)
    G.zero_grad()
    penalty.backward()
    pl_param_grads = {n: p.grad.clone() for n, p in G.named_parameters() if p.grad is not None}
    return dict(config=TINY_G, state_dict={k: v.clone() for k, v in G.state_dict().items()}, z=z, noise=noise,
                inject_index=3, image=image.detach(), direction=direction, grads=grads, none_grads=none_grads,
                z1=z1, image_fixed=image_fixed, latent=latent, image_lat=image_lat, pl_grad=pl_grad.detach(),
                pl_noise=pl_noise, pl_penalty=penalty.detach(), pl_value=pl.detach(),
                pl_mean=plr.mean_path_length.detach().clone(), pl_param_grads=pl_param_grads)


def golden_discriminator(ref):
    torch.manual_seed(0)
    D = ref.discriminator.Discriminator(TINY_D, no_rfp=True)
    randomize(D, 9)
    gen = torch.Generator().manual_seed(2)
    x = torch.rand(3, 2, 3, 32, 32, generator=gen)
    scalar, pixel = D(x, is_real=True, is_cut_mix=False)
    ds = torch.randn(scalar.shape, generator=gen)
    dp = torch.randn(pixel.shape, generator=gen)
    D.zero_grad()
    ((scalar * ds).sum() + (pixel * dp).sum()).backward()
    grads = {n: p.grad.clone() for n, p in D.named_parameters()}
    # R1 (loss.py:311-316) — double backward through every op of D
    xr = x.clone().requires_grad_(True)
    s2, p2 = D(xr, is_real=False, is_cut_mix=True)
    r1 = ref.loss.R1Regularization()(s2, xr, p2)
    D.zero_grad()
    r1.backward()
    r1_grads = {n: p.grad.clone() for n, p in D.named_parameters() if p.grad is not None}
    return dict(config=TINY_D, state_dict={k: v.clone() for k, v in D.state_dict().items()}, x=x,
                scalar=scalar.detach(), pixel=pixel.detach(), ds=ds, dp=dp, grads=grads, r1=r1.detach(),
                r1_grads=r1_grads)


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    torch.save(golden_ops(ref), os.path.join(OUT, "ops.pt"))
    torch.save(golden_block(ref), os.path.join(OUT, "dual_style_block.pt"))
    torch.save(golden_generator(ref), os.path.join(OUT, "generator.pt"))
    torch.save(golden_discriminator(ref), os.path.join(OUT, "discriminator.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
