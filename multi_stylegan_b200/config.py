"""Shapes and hyper-parameters of the default model (values of multi_stylegan/config.py:6-57)."""
import math
from typing import Any, Dict

u_net_2d_discriminator_config: Dict[str, Any] = {
    "encoder_channels": ((3, 128), (128, 256), (256, 384), (384, 768), (768, 1024)),
    "decoder_channels": ((1024, 768), (768, 384), (384, 256), (256, 128)),
    "fft": False,
}

multi_style_gan_generator_config: Dict[str, Any] = {
    "channels": (512, 512, 512, 512, 512, 512, 512),
    "channel_factor": 1,
    "latent_dimensions": 512,
    "depth_style_mapping": 8,
    "starting_resolution": (4, 4),
}

generation_hyperparameters: Dict[str, Any] = {
    "p_mixed_noise": 0.9,
    "lazy_generator_regularization": 16,
    "w_generator_regularization": math.log(2) / ((256 ** 2) * (math.log(256) - math.log(2))),
    "lazy_discriminator_regularization": 16,
    "w_discriminator_regularization_r1": 10.0,
    "w_discriminator_regularization": 4.0,
    "batch_factor_wrong_order": 1. / 4.,
    "batch_size_shrink_path_length_regularization": 2. / 4.,
    "betas": (0.0, 0.999),
    "top_k_start": 1. / 4.,
    "top_k_finish": 3. / 4.,
    "wrong_order_start": 3. / 4.,
    "trap_weight": 1. / 4.,
}


def scaled_configs(channel_div: int = 1, g_stages: int = 7):
    """Smaller models of the same topology for tests (channels divided, fewer generator stages)."""
    g = dict(multi_style_gan_generator_config)
    g["channels"] = tuple(512 // channel_div for _ in range(g_stages))
    g["latent_dimensions"] = 512 // channel_div
    d = dict(u_net_2d_discriminator_config)
    d["encoder_channels"] = tuple((a // channel_div if a > 3 else a, b // channel_div)
                                  for a, b in u_net_2d_discriminator_config["encoder_channels"])
    d["decoder_channels"] = tuple((a // channel_div, b // channel_div)
                                  for a, b in u_net_2d_discriminator_config["decoder_channels"])
    return g, d
