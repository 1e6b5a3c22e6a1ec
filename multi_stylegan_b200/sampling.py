"""EMA-generator sampling path: what scripts/get_gan_samples.py:30-60 and
scripts/gan_latent_space_interpolation.py:28-62 do with a trained `generator_ema`, on the library's kernels.

    generator = load_generator_ema("checkpoint_100.pt")            # the published checkpoints load unchanged (README.md:104-111)
    for bf, gfp in generate_samples(generator, samples=100):       # [T, 3, H, W] each, as the script hands them to save_image
        ...
    video = latent_space_interpolation(generator)                  # [frames, 3, 2 H, T W]

Inference only: `torch.no_grad()`, the generator's unobservable second branch is skipped (`compute_dead_branch=False`),
the fused shared-weight synthesis runs (multi_stylegan_generator.Generator._synthesis_fused).  There is no CPU fallback: on
a box without the CUDA library the generator's ops raise.

    python -m multi_stylegan_b200.sampling --load_checkpoint checkpoint_100.pt --samples 100 [--batch_size 8] [--out DIR]
    python -m multi_stylegan_b200.sampling --load_checkpoint checkpoint_100.pt --interpolation [--out DIR]
"""
import os
from typing import Dict, Iterator, Optional, Tuple, Union

import torch
import torch.nn.functional as F

from . import config as default_config
from . import misc


def _strip_data_parallel(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The reference trains and saves through nn.DataParallel (train_multi_stylegan.py:67-70): its state dicts carry a
    `module.` prefix that its scripts put back by wrapping the generator again (get_gan_samples.py:33-34)."""
    if state and all(k.startswith("module.") for k in state):
        return {k[len("module."):]: v for k, v in state.items()}
    return dict(state)


def load_generator_ema(checkpoint: Union[str, Dict], config: Optional[Dict] = None, device: Union[str, torch.device] = "cuda",
                       key: str = "generator_ema", compute_dead_branch: bool = False):
    """get_gan_samples.py:32-36: build the generator, load `checkpoint[key]` strictly, eval mode.  `checkpoint` is a path,
    the loaded checkpoint dict, or the state dict itself."""
    from .multi_stylegan_generator import Generator
    if isinstance(checkpoint, (str, os.PathLike)):
        checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=False)
    state = checkpoint[key] if key in checkpoint and isinstance(checkpoint[key], dict) else checkpoint
    generator = Generator(config or default_config.multi_style_gan_generator_config, compute_dead_branch=compute_dead_branch)
    generator.load_state_dict(_strip_data_parallel(state), strict=True)
    return generator.to(device).eval()


def sequence_to_images(sequence: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """get_gan_samples.py:43-53 for a whole batch: [B, 2, T, H, W] -> bright-field and GFP sequences [B, T, 3, H, W]; the
    bright-field channel is repeated to grey RGB, the GFP channel fills green only."""
    bf = sequence[:, 0:1].repeat_interleave(3, dim=1)
    gfp = sequence[:, 1:2].repeat_interleave(3, dim=1)
    gfp[:, 0] = 0.0
    gfp[:, 2] = 0.0
    return bf.permute(0, 2, 1, 3, 4), gfp.permute(0, 2, 1, 3, 4)


@torch.no_grad()
def generate_samples(generator, samples: int = 100, batch_size: int = 1,
                     device: Optional[Union[str, torch.device]] = None) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    """get_gan_samples.py:38-53: yields `samples` pairs (bf, gfp) of [T, 3, H, W] tensors.  Latents come from
    misc.get_noise(p_mixed_noise=0.0) and the per-layer noise maps are fresh draws, like the script's; with batch_size 1 the
    generator consumes the device's random stream in the script's order, larger batches draw the same distributions in
    batch-sized blocks (throughput: bench.py's `ema_sampling`)."""
    if device is None:
        device = next(generator.parameters()).device
    done = 0
    while done < samples:
        b = min(batch_size, samples - done)
        noise_input = misc.get_noise(batch_size=b, device=device, p_mixed_noise=0.0,
                                     latent_dimension=generator.latent_dimensions)
        bf, gfp = sequence_to_images(generator(noise_input))
        for i in range(b):
            yield bf[i], gfp[i]
        done += b


def save_samples(generator, samples: int = 100, batch_size: int = 1, out_dir: str = ".") -> int:
    """get_gan_samples.py:54-60: sample_bf_{i}.png / sample_gfp_{i}.png, the T frames side by side without padding."""
    import torchvision
    os.makedirs(out_dir, exist_ok=True)
    n = 0
    for n, (bf, gfp) in enumerate(generate_samples(generator, samples, batch_size), 1):
        torchvision.utils.save_image(tensor=bf, fp=os.path.join(out_dir, "sample_bf_{}.png".format(n - 1)), nrow=bf.shape[0],
                                     padding=0)
        torchvision.utils.save_image(tensor=gfp, fp=os.path.join(out_dir, "sample_gfp_{}.png".format(n - 1)),
                                     nrow=gfp.shape[0], padding=0)
    return n


def interpolation_latents(anchors: torch.Tensor, frames_per_anchor: int = 100, chunk: int = 32) -> torch.Tensor:
    """gan_latent_space_interpolation.py:36-40: linear interpolation (align_corners) through the anchor latents
    [A, latent] to A * frames_per_anchor latents, cut into chunks of `chunk` -> [A * frames / chunk, chunk, latent]."""
    a, d = anchors.shape
    z = F.interpolate(anchors.permute(1, 0).unsqueeze(dim=1), size=(frames_per_anchor * a), mode="linear",
                      align_corners=True).squeeze(dim=1).permute(1, 0)
    return z.reshape(z.shape[0] // chunk, chunk, d)


def compose_video(samples: torch.Tensor) -> torch.Tensor:
    """gan_latent_space_interpolation.py:47-56: [N, 2, T, H, W] -> frames [N, 3, 2 H, T W]; the T time steps side by side,
    bright field (grey) above GFP (green)."""
    def strip(channel: torch.Tensor) -> torch.Tensor:
        x = channel.permute(0, 2, 3, 1)                                              # [N, H, W, T]
        x = torch.cat([x[..., t] for t in range(x.shape[-1])], dim=-1)               # [N, H, T W]
        return x.unsqueeze(dim=1).repeat_interleave(repeats=3, dim=1)
    bf, gfp = strip(samples[:, 0]), strip(samples[:, 1])
    gfp[:, 0] = 0.0
    gfp[:, 2] = 0.0
    return torch.cat([bf, gfp], dim=2)


@torch.no_grad()
def latent_space_interpolation(generator, anchors: Optional[torch.Tensor] = None, n_anchors: int = 16,
                               frames_per_anchor: int = 100, chunk: int = 32) -> torch.Tensor:
    """gan_latent_space_interpolation.py:35-56: frames of a walk through latent space with the generator's FIXED noise
    buffers (`randomize_noise=False`, multi_stylegan_generator.py:88-95).  Returns the frames on the CPU like the script."""
    device = next(generator.parameters()).device
    if anchors is None:
        anchors = torch.randn(n_anchors, generator.latent_dimensions, dtype=torch.float32, device=device)
    z = interpolation_latents(anchors.to(device), frames_per_anchor, chunk)
    samples = [generator(input=z[i], randomize_noise=False).cpu() for i in range(z.shape[0])]
    return compose_video(torch.cat(samples, dim=0))


def save_video_frames(video: torch.Tensor, out_dir: str = "video_gan") -> int:
    """gan_latent_space_interpolation.py:57-60 (the ffmpeg call of :62 is left to the user)."""
    import torchvision
    os.makedirs(out_dir, exist_ok=True)
    for index in range(video.shape[0]):
        torchvision.utils.save_image(tensor=video[index][None], fp=os.path.join(out_dir, "frame_{}.png".format(str(index).zfill(5))))
    return int(video.shape[0])


def main(argv=None) -> None:
    from argparse import ArgumentParser
    parser = ArgumentParser(description="EMA-generator sampling (scripts/get_gan_samples.py, gan_latent_space_interpolation.py)")
    parser.add_argument("--cuda_devices", default="0", type=str)
    parser.add_argument("--samples", default=100, type=int)
    parser.add_argument("--load_checkpoint", default="checkpoint_100.pt", type=str)
    parser.add_argument("--batch_size", default=1, type=int)
    parser.add_argument("--interpolation", action="store_true")
    parser.add_argument("--out", default=None, type=str)
    args = parser.parse_args(argv)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.cuda_devices)
    generator = load_generator_ema(args.load_checkpoint, device="cuda")
    if args.interpolation:
        n = save_video_frames(latent_space_interpolation(generator), args.out or "video_gan")
        print("wrote %d frames" % n)
    else:
        n = save_samples(generator, args.samples, args.batch_size, args.out or ".")
        print("wrote %d samples" % n)


if __name__ == "__main__":
    main()
