"""Twin-branch ("dual-style") StyleGAN2 generator with the reference's module tree, constructor
arguments, forward signature and state_dict names (multi_stylegan/multi_stylegan_generator.py:15-641),
so reference checkpoints load unchanged — with every convolution, FIR and activation on the sm_100a
kernels of this package.

Behaviours reproduced on purpose: FusedLeakyReLU gain 1.0; `Upsample` without the
x4 gain; 2x2/stride-2 transposed up-convs followed by a 4x4 blur with pad (2,1); constant inputs of
ones; the second branch's main convolutions never reach the image (`output_blocks_2` is fed branch-1
features, reference :189).  Because that branch is unobservable, `compute_dead_branch=False` skips it
(identical outputs and gradients; with noise=None it also skips the dead layers' randn draws).
"""
import math
from typing import Any, Dict, Iterable, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
from torch import autograd

from . import _C, conv, equalized_layer, linear, styled
from .op_static import FusedLeakyReLU, upfirdn2d
from .op_static.fused_act import noise_bias_leaky_relu
from .op_static.upfirdn2d import blur_noise_bias_leaky_relu


# The generator has two arithmetically equivalent execution forms (see styled.py, _mode.py): the shared-weight form with
# hand-written first-order backwards (default; used by no_grad forwards and the generator step) and the reference's
# per-sample-weight form, every piece of which is differentiable to any order (path-length regularisation).
from ._mode import higher_order_gradients  # noqa: E402,F401  (re-exported)
from . import _mode  # noqa: E402


def _fir_kernel(taps: List[int]) -> torch.Tensor:
    k = torch.tensor(taps, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


class Upsample(nn.Module):
    """x2 zero-insertion + FIR, taps normalised to sum 1 (no factor^2 gain) — reference :529-575."""

    def __init__(self, blur_kernel: List[int] = [1, 3, 3, 1], factor: int = 2) -> None:
        super().__init__()
        self.factor = factor
        kernel = _fir_kernel(blur_kernel)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.padding = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return upfirdn2d(input, self.kernel, up=self.factor, pad=self.padding)


class Blur(nn.Module):
    """FIR low-pass; taps x sampling_factor^2 when sampling_factor > 1 — reference :578-641."""

    def __init__(self, kernel: List[int], sampling_factor: int = 1, sampling_factor_padding: int = 2,
                 kernel_size: int = 3) -> None:
        super().__init__()
        p = (len(kernel) - sampling_factor_padding) + (kernel_size - 1)
        self.padding = ((p + 1) // 2, p // 2)
        k = _fir_kernel(kernel)
        if sampling_factor > 1:
            k = k * (sampling_factor ** 2)
        self.register_buffer("kernel", k)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return upfirdn2d(input, self.kernel, pad=self.padding)


class StyleMapping(nn.Module):
    def __init__(self, latent_dimensions: int = 512, depth: int = 8) -> None:
        super().__init__()
        layers: List[nn.Module] = [equalized_layer.PixelwiseNormalization()]
        for _ in range(depth):
            layers += [equalized_layer.EqualizedLinear(latent_dimensions, latent_dimensions, bias=False),
                       FusedLeakyReLU(latent_dimensions)]
        self.layers = nn.Sequential(*layers)

    def forward(self, noise: torch.Tensor) -> torch.Tensor:
        if self._fused_eligible(noise):
            # the whole network (pixel norm + depth x [linear -> bias + leaky ReLU]) as one launch (csrc/linear_ops.cu)
            mods = list(self.layers)
            lin0, act0 = mods[1], mods[2]
            layers = [(mods[i].weight, mods[i + 1].bias) for i in range(1, len(mods), 2)]
            return linear.style_mapping(noise, layers, lin0.scale, act0.negative_slope, act0.scale, mods[0].alpha)
        return self.layers(noise)

    def _fused_eligible(self, noise: torch.Tensor) -> bool:
        if not (noise.is_cuda and noise.dim() == 2 and noise.dtype == torch.float32) or _mode.higher_order():
            return False
        if noise.requires_grad and torch.is_grad_enabled():
            return False                       # the fused backward does not produce the gradient w.r.t. the noise
        mods = list(self.layers)
        depth = (len(mods) - 1) // 2
        if depth < 1 or not _C.style_mapping_supported(depth, noise.shape[1]):
            return False
        acts = mods[2::2]
        return all(a.negative_slope == acts[0].negative_slope and a.scale == acts[0].scale for a in acts)


class ConstantInput(nn.Module):
    def __init__(self, channel: int, size: Tuple[int, int] = (4, 4)) -> None:
        super().__init__()
        self.input = nn.Parameter(torch.ones(1, channel, size[0], size[1]))

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return self.input.repeat_interleave(dim=0, repeats=input.shape[0])


class NoiseInjection(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1, dtype=torch.float32))

    def forward(self, input: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if noise is None:
            noise = torch.randn(input.shape[0], 1, input.shape[2], input.shape[3], device=input.device,
                                dtype=torch.float32)
        return input + self.weight * noise


class _ModulateWeights(autograd.Function):
    """w_mod[b,o,c,:,:] = scale * W[o,c,:,:] * s[b,c] (* rsqrt(sum_{c,kh,kw}(.)^2 + 1e-8)) — reference :384-388 as one
    kernel, with a fused first-order backward.  When the backward itself is being recorded (create_graph=True: the
    path-length regulariser, :193-200) it is evaluated with differentiable tensor ops instead, so gradients of any
    order stay available."""

    @staticmethod
    def forward(ctx, W, s, scale, demodulate):
        from . import _C
        w_mod, d = _C.modulate_weights(W, s, scale, demodulate)
        ctx.save_for_backward(W, s, d if d is not None else W.new_empty(0))
        ctx.scale, ctx.demodulate = scale, demodulate
        return w_mod

    @staticmethod
    def backward(ctx, g):
        from . import _C
        W, s, d = ctx.saved_tensors
        if not torch.is_grad_enabled():
            dW, ds = _C.modulate_weights_bwd(g, W, s, d if ctx.demodulate else None, ctx.scale, ctx.demodulate)
            return dW, ds, None, None
        sb = s.view(s.shape[0], 1, s.shape[1], 1, 1)
        u = ctx.scale * W.unsqueeze(0) * sb
        if ctx.demodulate:
            dm = torch.rsqrt((u * u).sum(dim=[2, 3, 4]) + 1e-08).view(s.shape[0], W.shape[0], 1, 1, 1)
            dd = (g * u).sum(dim=[2, 3, 4]).view(s.shape[0], W.shape[0], 1, 1, 1)
            du = dm * (g - dm * dm * dd * u)
        else:
            du = g
        dW = ctx.scale * (du * sb).sum(dim=0)
        ds = ctx.scale * (du * W.unsqueeze(0)).sum(dim=[1, 3, 4])
        return dW, ds, None, None


class ModulatedConv2d(nn.Module):
    """Weight-modulated / demodulated per-sample convolution — reference :295-414.

    The per-sample filter banks scale*W*s (*demod) are built with the reference's own arithmetic and
    handed to the per-sample conv kernels as the GEMM B operand; upsampling layers are the 2x2/stride-2
    transposed convolution (= the conv's dgrad kernel) followed by the x4 blur."""

    def __init__(self, in_channels: int, out_channels: int, style_dimension: int,
                 kernel_size: Union[int, Tuple[int, int]] = (3, 3), demodulate: bool = True, upsampling: bool = True,
                 blur_kernel: List[int] = [1, 3, 3, 1], modulation_mapping: bool = True) -> None:
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.demodulate, self.upsampling = demodulate, upsampling
        self.kernel_size = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
        self.blur = Blur(kernel=blur_kernel, sampling_factor=2, sampling_factor_padding=2,
                         kernel_size=self.kernel_size[0]) if upsampling else None
        if upsampling:
            self.padding, self.stride = (0, 0), (2, 2)
        else:
            self.padding, self.stride = (self.kernel_size[0] // 2, self.kernel_size[1] // 2), (1, 1)
        self.scale = math.sqrt(2) / math.sqrt(in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight = nn.Parameter(torch.randn(1, out_channels, in_channels, *self.kernel_size))
        self.modulation_mapping = equalized_layer.EqualizedLinear(style_dimension, in_channels, bias=True) \
            if modulation_mapping else None
        if modulation_mapping:
            self.modulation_mapping.bias.data.fill_(1.0)

    def extra_repr(self) -> str:
        return "{}, {}, kernel_size={}, stride={}, padding={}, upsampling={}".format(
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.upsampling)

    def modulated_weight(self, style: torch.Tensor, batch_size: int, modulated_style: Optional[torch.Tensor] = None):
        """Per-sample filter banks [B, O, C, kh, kw] (reference :379-388) and the modulated style [B,1,C,1,1].
        `modulated_style` [B, C]: the output of this layer's style linear when the caller has already computed it
        (all style linears of the network in one launch, linear.style_linears)."""
        if modulated_style is not None:
            modulated_style = modulated_style.reshape(batch_size, 1, self.in_channels, 1, 1)
        elif self.modulation_mapping is not None:
            modulated_style = self.modulation_mapping(style).view(batch_size, 1, self.in_channels, 1, 1)
        else:
            modulated_style = style
        weight = _ModulateWeights.apply(self.weight[0], modulated_style.reshape(batch_size, self.in_channels),
                                        self.scale, self.demodulate)             # [B, O, C, kh, kw]
        return weight, modulated_style

    def forward(self, input: torch.Tensor, style: torch.Tensor, modulated_style: Optional[torch.Tensor] = None):
        batch_size, features, height, width = input.shape
        assert features == self.in_channels, \
            "Expect input feature shape of {} but get {}.".format(self.in_channels, features)
        weight, modulated_style = self.modulated_weight(style, batch_size, modulated_style)
        if self.upsampling:
            output = conv.conv_transpose2d(input, weight, stride=self.stride, padding=self.padding,
                                           weight_is_conv_layout=True)
            output = self.blur(output)
        else:
            output = conv.conv2d(input, weight, stride=self.stride, padding=self.padding)
        if self.modulation_mapping is not None:
            return output, modulated_style
        return output


class StyledConv2d(nn.Module):
    """Modulated conv -> noise injection -> bias + leaky ReLU; returns (out, style) when it owns the
    style linear, so the paired block of the second branch can reuse the style — reference :417-469."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[int, Tuple[int, int]],
                 style_dimension: int, demodulate: bool = True, upsampling: bool = False,
                 blur_kernel: List[int] = [1, 3, 3, 1], modulation_mapping: bool = True) -> None:
        super().__init__()
        self.modulation_mapping = modulation_mapping
        self.modulated_convolution = ModulatedConv2d(in_channels=in_channels, out_channels=out_channels,
                                                     kernel_size=kernel_size, style_dimension=style_dimension,
                                                     demodulate=demodulate, upsampling=upsampling,
                                                     blur_kernel=blur_kernel, modulation_mapping=modulation_mapping)
        self.noise_injection = NoiseInjection()
        self.activation = FusedLeakyReLU(out_channels)

    def forward(self, input: torch.Tensor, style: torch.Tensor, noise: torch.Tensor = None):
        mc = self.modulated_convolution
        if not mc.upsampling and mc.out_channels % 4 == 0:
            # conv -> noise (:289-292) -> bias + leaky ReLU (fused_act.py:58): one kernel, epilogue on the accumulators
            batch_size, features, height, width = input.shape
            assert features == mc.in_channels, \
                "Expect input feature shape of {} but get {}.".format(mc.in_channels, features)
            weight, style = mc.modulated_weight(style, batch_size)
            if noise is None:
                noise = torch.randn(batch_size, 1, height, width, device=input.device, dtype=torch.float32)
            output = conv.conv2d_bias_act(input, weight, bias=self.activation.bias, noise=noise,
                                          noise_w=self.noise_injection.weight, stride=mc.stride, padding=mc.padding,
                                          negative_slope=self.activation.negative_slope, gain=self.activation.scale)
            if self.modulation_mapping:
                return output, style
            return output
        if mc.upsampling and mc.out_channels % 4 == 0:
            # transposed conv (:393-401) -> [blur (:403) + noise + bias + leaky ReLU] as one FIR pass
            batch_size = input.shape[0]
            weight, style = mc.modulated_weight(style, batch_size)
            output = conv.conv_transpose2d(input, weight, stride=mc.stride, padding=mc.padding, weight_is_conv_layout=True)
            kh, kw = mc.blur.kernel.shape
            oh = output.shape[2] + sum(mc.blur.padding) - kh + 1
            ow = output.shape[3] + sum(mc.blur.padding) - kw + 1
            if noise is None:
                noise = torch.randn(batch_size, 1, oh, ow, device=input.device, dtype=torch.float32)
            output = blur_noise_bias_leaky_relu(output, mc.blur.kernel, mc.blur.padding, noise,
                                                self.noise_injection.weight, self.activation.bias,
                                                self.activation.negative_slope, self.activation.scale)
            if self.modulation_mapping:
                return output, style
            return output
        if self.modulation_mapping:
            output, style = self.modulated_convolution(input, style)
        else:
            output = self.modulated_convolution(input, style)
        if output.shape[1] % 4 == 0:
            # noise injection (:289-292) + bias + leaky ReLU (fused_act.py:58) in one pass over the activation
            if noise is None:
                noise = torch.randn(output.shape[0], 1, output.shape[2], output.shape[3], device=output.device,
                                    dtype=torch.float32)
            output = noise_bias_leaky_relu(output, noise, self.noise_injection.weight, self.activation.bias,
                                           self.activation.negative_slope, self.activation.scale)
        else:
            output = self.activation(self.noise_injection(output, noise=noise))
        if self.modulation_mapping:
            return output, style
        return output


class OutputBlock(nn.Module):
    """tRGB: 1x1 modulated conv without demodulation + scalar bias + upsampled skip — reference :472-526."""

    def __init__(self, in_channels: int, style_dimension: int, out_channels: int = 1, upsampling: bool = False,
                 blur_kernel: List[int] = [1, 3, 3, 1], modulation_mapping: bool = True) -> None:
        super().__init__()
        self.modulation_mapping = modulation_mapping
        self.upsampling = Upsample(blur_kernel=blur_kernel, factor=2) if upsampling else nn.Identity()
        self.modulated_convolution = ModulatedConv2d(in_channels=in_channels, out_channels=out_channels,
                                                     style_dimension=style_dimension, kernel_size=(1, 1),
                                                     upsampling=False, demodulate=False,
                                                     modulation_mapping=modulation_mapping)
        self.bias = nn.Parameter(torch.zeros(1, 1, 1, 1, dtype=torch.float32))

    def forward(self, input: torch.Tensor, style: torch.Tensor, skip: torch.Tensor = None,
                modulated_style: Optional[torch.Tensor] = None):
        if self.modulation_mapping:
            output, style = self.modulated_convolution(input, style, modulated_style)
        else:
            output = self.modulated_convolution(input, style)
        output = output + self.bias
        if skip is not None:
            output = output + self.upsampling(skip)
        if self.modulation_mapping:
            return output, style
        return output


class Generator(nn.Module):
    def __init__(self, config: Dict[str, Any], compute_dead_branch: bool = True) -> None:
        super().__init__()
        channels: Tuple[int, ...] = config["channels"]
        f = config["channel_factor"]
        ch = [int(c // f) for c in channels]
        self.out_channels = 3
        self.latent_dimensions: int = config["latent_dimensions"]
        self.starting_resolution: Tuple[int, int] = config["starting_resolution"]
        self.compute_dead_branch = compute_dead_branch
        self.fused_modconv = True        # False: always the per-sample-weight formulation (tests, A/B measurements)
        L = self.latent_dimensions
        self.style_mapping = StyleMapping(latent_dimensions=L, depth=config["depth_style_mapping"])
        self.constant_input_1 = ConstantInput(channel=ch[0], size=self.starting_resolution)
        self.constant_input_2 = ConstantInput(channel=ch[0], size=self.starting_resolution)
        self.starting_convolution_1 = StyledConv2d(ch[0], ch[0], (3, 3), L, upsampling=False, demodulate=True)
        self.starting_convolution_2 = StyledConv2d(ch[0], ch[0], (3, 3), L, upsampling=False, demodulate=True,
                                                   modulation_mapping=False)
        self.starting_output_block_1 = OutputBlock(ch[0], L, self.out_channels, upsampling=False)
        self.starting_output_block_2 = OutputBlock(ch[0], L, self.out_channels, upsampling=False,
                                                   modulation_mapping=False)
        self.main_convolutions_1 = nn.ModuleList()
        self.output_blocks_1 = nn.ModuleList()
        self.main_convolutions_2 = nn.ModuleList()
        self.output_blocks_2 = nn.ModuleList()
        for i in range(len(ch) - 1):
            self.main_convolutions_1.append(StyledConv2d(ch[i], ch[i + 1], (2, 2), L, upsampling=True))
            self.main_convolutions_1.append(StyledConv2d(ch[i + 1], ch[i + 1], (3, 3), L, upsampling=False))
            self.output_blocks_1.append(OutputBlock(ch[i + 1], L, self.out_channels, upsampling=True))
            self.main_convolutions_2.append(StyledConv2d(ch[i], ch[i + 1], (2, 2), L, upsampling=True,
                                                         modulation_mapping=False))
            self.main_convolutions_2.append(StyledConv2d(ch[i + 1], ch[i + 1], (3, 3), L, upsampling=False,
                                                         modulation_mapping=False))
            self.output_blocks_2.append(OutputBlock(ch[i + 1], L, self.out_channels, upsampling=True,
                                                    modulation_mapping=False))
        self.noises = nn.Module()
        r0 = self.starting_resolution
        self.noises.register_buffer("noise_start", torch.randn(1, 1, r0[0], r0[1]))
        for i in range(len(ch) - 1):
            self.noises.register_buffer("noise_{}".format(2 * i), torch.randn(1, 1, 2 ** (i + 3), 2 ** (i + 3)))
            self.noises.register_buffer("noise_{}".format(2 * i + 1), torch.randn(1, 1, 2 ** (i + 3), 2 ** (i + 3)))

    def get_parameters(self, lr_main: float = 1e-03, lr_style: float = 1e-05) -> Iterable:
        """Parameter groups of the reference (:97-112): everything at lr_main, the mapping network at lr_style."""
        main = [self.constant_input_1, self.starting_convolution_1, self.starting_output_block_1,
                self.main_convolutions_1, self.output_blocks_1, self.constant_input_2,
                self.starting_convolution_2, self.starting_output_block_2, self.main_convolutions_2,
                self.output_blocks_2]
        groups = [{"params": m.parameters(), "lr": lr_main} for m in main]
        groups.append({"params": self.style_mapping.parameters(), "lr": lr_style})
        return groups

    @property
    def n_latent(self) -> int:
        return len(self.main_convolutions_1) + 2

    def _latent(self, input, input_is_latent: bool, inject_index) -> torch.Tensor:
        n_latent = self.n_latent
        if not input_is_latent:
            if isinstance(input, list):
                if len(input) > 1 and all(z.shape == input[0].shape and z.dim() == 2 for z in input):
                    styles = list(self.style_mapping(torch.cat(input, dim=0)).split(input[0].shape[0], dim=0))   # row-wise network
                else:
                    styles = [self.style_mapping(z) for z in input]
                if torch.is_tensor(inject_index):
                    # crossover index held on the device (CUDA-graph replay): the same [B, n_latent, L] tensor as below
                    first = torch.arange(n_latent, device=styles[0].device).view(1, n_latent, 1) < inject_index
                    return torch.where(first, styles[0].unsqueeze(1), styles[1].unsqueeze(1))
                if inject_index is None:
                    inject_index = np.random.randint(1, n_latent - 1)
                return torch.cat((styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                  styles[1].unsqueeze(1).repeat(1, n_latent - inject_index, 1)), dim=1)
            return self.style_mapping(input).unsqueeze(1).repeat(1, n_latent, 1)
        if input.ndim < 3:
            return input.unsqueeze(1).repeat(1, n_latent, 1)
        if input.shape[1] != n_latent:
            assert input.shape[1] == 0
            return input.repeat(1, n_latent, 1)
        return input

    @staticmethod
    def _paired_output(block_1: "OutputBlock", block_2: "OutputBlock", features: torch.Tensor, w: torch.Tensor,
                       skip_12: torch.Tensor, modulated_style: Optional[torch.Tensor] = None):
        """output_blocks_1[i](x, w, skip_1) and output_blocks_2[i](x, style, skip_2) (reference :188-189, :509-526) as ONE
        1x1 modulated convolution with the two 3-channel filter banks stacked (both read the same 512-channel
        features) and ONE skip upsampling on the stacked [B, 6, H, W] image; returns the stacked result and the style."""
        batch = features.shape[0]
        weight_1, style = block_1.modulated_convolution.modulated_weight(w, batch, modulated_style)
        weight_2, _ = block_2.modulated_convolution.modulated_weight(style, batch)
        mc = block_1.modulated_convolution
        output = conv.conv2d(features, torch.cat([weight_1, weight_2], dim=1), stride=mc.stride, padding=mc.padding)
        c = weight_1.shape[1]
        bias = torch.cat([block_1.bias.expand(1, c, 1, 1), block_2.bias.expand(1, c, 1, 1)], dim=1)
        output = output + bias
        if skip_12 is not None:
            output = output + block_1.upsampling(skip_12)
        return output, style

    def forward(self, input: Union[List[torch.Tensor], torch.Tensor], return_main_style_vectors: bool = False,
                noise: Optional[List[torch.Tensor]] = None, randomize_noise: bool = True,
                inject_index: Optional[int] = None, input_is_latent: bool = False,
                return_path_length_grads: bool = False, path_length_noise: Optional[torch.Tensor] = None):
        n_main = len(self.main_convolutions_1)
        if noise is None:
            if randomize_noise:
                noise_start, noise = None, [None] * n_main
            else:
                noise_start = self.noises.noise_start
                noise = [getattr(self.noises, "noise_{}".format(i)) for i in range(n_main)]
        else:
            noise_start, noise = noise[0], noise[1:]
        latent = self._latent(input, input_is_latent, inject_index)
        dead = self.compute_dead_branch
        if return_path_length_grads:
            with higher_order_gradients():
                return self._forward_tail(latent, noise_start, noise, dead, True, path_length_noise, False)
        return self._forward_tail(latent, noise_start, noise, dead, False, None, return_main_style_vectors)

    def _fused_eligible(self, latent: torch.Tensor) -> bool:
        """Shared-weight fast path: live branch only, channel counts the channels-last kernels take."""
        if self.compute_dead_branch or _mode.higher_order() or not self.fused_modconv:
            return False
        convs = [self.starting_convolution_1, self.starting_convolution_2] + list(self.main_convolutions_1)
        return all(c.modulated_convolution.out_channels % 4 == 0 and c.modulated_convolution.in_channels % 4 == 0
                   and c.modulated_convolution.demodulate for c in convs)

    def _synthesis_fused(self, latent: torch.Tensor, noise_start, noise) -> torch.Tensor:
        """The live part of the network (branch 1 + the starting block of branch 2, reference :176-189) in the
        shared-weight form: each layer's epilogue writes its activation and the same activation multiplied by the
        NEXT layer's style, which is that layer's GEMM operand; styles and demodulation factors are [B, C] vectors."""
        B = latent.shape[0]
        dev = latent.device

        styles = self._all_styles(latent)

        def style_of(block, w: torch.Tensor) -> torch.Tensor:
            mc = block.modulated_convolution
            s = styles.get(id(mc))
            return mc.modulation_mapping(w) if s is None else s                            # [B, C_in]

        def draw(noise_map, h, w):
            return torch.randn(B, 1, h, w, device=dev, dtype=torch.float32) if noise_map is None else noise_map

        def run(block: "StyledConv2d", xs, s, noise_map, s_next):
            mc, act = block.modulated_convolution, block.activation
            if mc.upsampling:
                kh, kw = mc.blur.kernel.shape
                oh = (xs.shape[2] - 1) * mc.stride[0] + mc.kernel_size[0] + sum(mc.blur.padding) - kh + 1
                ow = (xs.shape[3] - 1) * mc.stride[1] + mc.kernel_size[1] + sum(mc.blur.padding) - kw + 1
                return styled.styled_up_conv(xs, mc.weight[0], s, mc.scale, mc.demodulate, mc.blur.kernel, mc.blur.padding,
                                             draw(noise_map, oh, ow), block.noise_injection.weight, act.bias, s_next,
                                             mc.stride, mc.padding, act.negative_slope, act.scale)
            return styled.styled_conv(xs, mc.weight[0], s, mc.scale, mc.demodulate, draw(noise_map, xs.shape[2], xs.shape[3]),
                                      block.noise_injection.weight, act.bias, s_next, mc.stride, mc.padding,
                                      act.negative_slope, act.scale)

        mains = self.main_convolutions_1
        n_main = len(mains)
        s0 = style_of(self.starting_convolution_1, latent[:, 0])
        s_next = style_of(mains[0], latent[:, 1]) if n_main else None
        sb = s0.view(B, -1, 1, 1)
        out_1, xs = run(self.starting_convolution_1, self.constant_input_1.input * sb, s0, noise_start, s_next)
        out_2, _ = run(self.starting_convolution_2, self.constant_input_2.input * sb, s0, noise_start, None)
        skip_1, style = self.starting_output_block_1(out_1, latent[:, 1],
                                                     modulated_style=styles.get(id(self.starting_output_block_1.modulated_convolution)))
        skip_2 = self.starting_output_block_2(out_2, style)
        skip_12 = torch.cat([skip_1, skip_2], dim=1)
        for i in range(n_main // 2):
            s_up = s_next
            s_cv = style_of(mains[2 * i + 1], latent[:, 2 * i + 2])
            _, xs = run(mains[2 * i], xs, s_up, noise[2 * i], s_cv)
            s_next = style_of(mains[2 * i + 2], latent[:, 2 * i + 3]) if 2 * i + 2 < n_main else None
            out_1, xs = run(mains[2 * i + 1], xs, s_cv, noise[2 * i + 1], s_next)
            skip_12, _ = self._paired_output(self.output_blocks_1[i], self.output_blocks_2[i], out_1,
                                             latent[:, 2 * i + 3], skip_12,
                                             styles.get(id(self.output_blocks_1[i].modulated_convolution)))
        return skip_12.reshape(B, 2, self.out_channels, skip_12.shape[2], skip_12.shape[3])

    def _all_styles(self, latent: torch.Tensor) -> Dict[int, torch.Tensor]:
        """Every style linear of the live network (`modulation_mapping` of the first branch's convolutions and tRGB
        blocks, reference :355-361) in ONE launch: they are independent linears on slices of the [B, n_latent * L]
        latent (linear.style_linears).  Keys: id() of the ModulatedConv2d.  Empty when the fused form does not apply."""
        mains = self.main_convolutions_1
        if not latent.is_cuda or latent.dim() != 3 or latent.dtype != torch.float32 or latent.shape[2] % 4 != 0:
            return {}
        L = latent.shape[2]
        users = [(self.starting_convolution_1.modulated_convolution, 0), (self.starting_output_block_1.modulated_convolution, 1)]
        users += [(m.modulated_convolution, j + 1) for j, m in enumerate(mains)]
        users += [(ob.modulated_convolution, 2 * i + 3) for i, ob in enumerate(self.output_blocks_1)]
        users = [(mc, j) for mc, j in users if mc.modulation_mapping is not None and j < latent.shape[1]]
        users.sort(key=lambda u: u[1])                                       # items reading the same latent are adjacent
        if not users or len(users) > 48:
            return {}
        specs = [(mc.modulation_mapping.weight, mc.modulation_mapping.bias, j * L, mc.modulation_mapping.scale,
                  mc.modulation_mapping.scale_bias) for mc, j in users]
        outs = linear.style_linears(latent.reshape(latent.shape[0], -1), specs)
        return {id(mc): o for (mc, _), o in zip(users, outs)}

    def _forward_tail(self, latent, noise_start, noise, dead, return_path_length_grads, path_length_noise,
                      return_main_style_vectors):
        n_main = len(self.main_convolutions_1)
        if self._fused_eligible(latent):
            image = self._synthesis_fused(latent, noise_start, noise)
            if return_main_style_vectors:
                return image, latent
            return image

        out_1 = self.constant_input_1(latent)
        out_2 = self.constant_input_2(latent)
        out_1, style = self.starting_convolution_1(out_1, latent[:, 0], noise=noise_start)
        out_2 = self.starting_convolution_2(out_2, style, noise=noise_start)
        skip_1, style = self.starting_output_block_1(out_1, latent[:, 1])
        skip_2 = self.starting_output_block_2(out_2, style)
        paired = not dead            # both tRGB blocks of a level read branch-1 features (:188-189): evaluate them jointly
        skip_12 = torch.cat([skip_1, skip_2], dim=1) if paired else None
        for i in range(n_main // 2):
            out_1, style = self.main_convolutions_1[2 * i](out_1, latent[:, 2 * i + 1], noise=noise[2 * i])
            if dead:
                out_2 = self.main_convolutions_2[2 * i](out_2, style, noise=noise[2 * i])
            out_1, style = self.main_convolutions_1[2 * i + 1](out_1, latent[:, 2 * i + 2], noise=noise[2 * i + 1])
            if dead:
                out_2 = self.main_convolutions_2[2 * i + 1](out_2, style, noise=noise[2 * i + 1])
            if paired:
                skip_12, style = self._paired_output(self.output_blocks_1[i], self.output_blocks_2[i], out_1,
                                                     latent[:, 2 * i + 3], skip_12)
            else:
                skip_1, style = self.output_blocks_1[i](out_1, latent[:, 2 * i + 3], skip=skip_1)
                skip_2 = self.output_blocks_2[i](out_1, style, skip=skip_2)     # branch-1 features (reference :189)
        if paired:
            image = skip_12.reshape(skip_12.shape[0], 2, self.out_channels, skip_12.shape[2], skip_12.shape[3])
        else:
            image = torch.stack([skip_1, skip_2], dim=1)
        if return_path_length_grads:
            # reference :195-196; `path_length_noise` (an extension) injects the N(0,1) draw for parity tests
            pl_noise = (torch.randn(image.shape, device=image.device, dtype=torch.float32, requires_grad=True)
                        if path_length_noise is None else path_length_noise) \
                / math.sqrt(image.shape[2] * image.shape[3] * image.shape[4])
            return autograd.grad(outputs=(image * pl_noise).sum(), inputs=latent, create_graph=True,
                                 retain_graph=True, only_inputs=True)[0]
        if return_main_style_vectors:
            return image, latent
        return image
