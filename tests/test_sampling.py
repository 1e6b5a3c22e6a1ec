"""EMA-generator sampling path (scripts/get_gan_samples.py, scripts/gan_latent_space_interpolation.py):
checkpoint loading, the image / video composition and the generated sequences against the oracle."""
import os

import pytest
import torch

from oracle import model as omodel
from oracle import sampling as osampling
from oracle.make_golden import TINY_G, randomize
from tests.conftest import rel_err


def _tiny_generator():
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    torch.manual_seed(5)
    G = G_mod.Generator(TINY_G, compute_dead_branch=True)
    randomize(G, 3)
    return G


def test_composition_matches_oracle():
    from multi_stylegan_b200 import sampling
    gen = torch.Generator().manual_seed(0)
    seq = torch.randn(3, 2, 3, 8, 8, generator=gen)
    bf, gfp = sampling.sequence_to_images(seq.clone())
    for i in range(seq.shape[0]):
        want_bf, want_gfp = osampling.sample_images(seq[i:i + 1])
        assert torch.equal(bf[i], want_bf) and torch.equal(gfp[i], want_gfp)
    assert torch.equal(sampling.compose_video(seq.clone()), osampling.video_frames(seq))
    anchors = torch.randn(4, 16, generator=gen)
    got = sampling.interpolation_latents(anchors, frames_per_anchor=8, chunk=16)
    want = osampling.interpolation_latents(anchors, 8, 16)
    assert got.shape == want.shape == (2, 16, 16)
    assert torch.allclose(got, want, atol=1e-6)
    assert torch.equal(got[0, 0], anchors[0]) and torch.allclose(got[-1, -1], anchors[-1], atol=1e-6)


def test_checkpoint_loading_accepts_the_reference_layouts(tmp_path):
    """The reference saves DataParallel state dicts (`module.` prefix) under "generator_ema" (get_gan_samples.py:33-34)."""
    from multi_stylegan_b200 import sampling
    G = _tiny_generator()
    sd = G.state_dict()
    path = os.path.join(str(tmp_path), "checkpoint_100.pt")
    torch.save({"generator_ema": {"module." + k: v for k, v in sd.items()}, "generator": {}}, path)
    for source in (path, {"generator_ema": sd}, sd):
        loaded = sampling.load_generator_ema(source, config=TINY_G, device="cpu")
        assert not loaded.training and not loaded.compute_dead_branch
        for k, v in loaded.state_dict().items():
            assert torch.equal(v, sd[k]), k
    with pytest.raises(RuntimeError):
        sampling.load_generator_ema({"generator_ema": {k: v for k, v in sd.items() if "noises" not in k}}, config=TINY_G,
                                    device="cpu")


def _check_sampling(dev, tol):
    from multi_stylegan_b200 import sampling
    G = _tiny_generator()
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    G = sampling.load_generator_ema(sd, config=TINY_G, device=dev)
    # latent walk with the fixed noise buffers: every frame against the oracle generator
    anchors = torch.randn(4, 16, generator=torch.Generator().manual_seed(1))
    video = sampling.latent_space_interpolation(G, anchors=anchors, frames_per_anchor=4, chunk=8)
    z = osampling.interpolation_latents(anchors, 4, 8)
    want = torch.cat([omodel.generator_forward(sd, z[i], noise=osampling.fixed_noise(sd)) for i in range(z.shape[0])])
    want_video = osampling.video_frames(want)
    assert video.shape == want_video.shape == (16, 3, 64, 96) and video.device.type == "cpu"
    assert rel_err(video, want_video) < tol
    # fresh-noise samples: shapes, channel layout, and the script's random stream at batch size 1
    torch.manual_seed(9)
    pairs = list(sampling.generate_samples(G, samples=3, batch_size=2))
    assert len(pairs) == 3
    for bf, gfp in pairs:
        assert bf.shape == gfp.shape == (3, 3, 32, 32)
        assert torch.equal(bf[:, 0], bf[:, 1]) and torch.equal(bf[:, 0], bf[:, 2])
        assert float(gfp[:, 0].abs().max()) == 0.0 and float(gfp[:, 2].abs().max()) == 0.0 and float(gfp[:, 1].abs().max()) > 0
    torch.manual_seed(9)
    a = next(sampling.generate_samples(G, samples=1, batch_size=1))
    torch.manual_seed(9)
    with torch.no_grad():
        seq = G(torch.randn(1, 16, device=dev))                      # get_gan_samples.py:40-42 with p_mixed_noise = 0
    want_bf, want_gfp = osampling.sample_images(seq.cpu())
    assert torch.equal(a[0].cpu(), want_bf) and torch.equal(a[1].cpu(), want_gfp)


def test_sampling_host_logic_matches_oracle(oracle_backend):
    _check_sampling("cpu", 1e-4)


@pytest.mark.gpu
def test_sampling_matches_oracle_on_the_device(built_library, tmp_path):
    from multi_stylegan_b200 import sampling
    _check_sampling(torch.device("cuda:0"), 1e-2)
    G = sampling.load_generator_ema(_tiny_generator().state_dict(), config=TINY_G, device="cuda:0")
    assert sampling.save_samples(G, samples=2, batch_size=2, out_dir=str(tmp_path)) == 2
    assert sorted(os.listdir(str(tmp_path))) == ["sample_bf_0.png", "sample_bf_1.png", "sample_gfp_0.png", "sample_gfp_1.png"]
