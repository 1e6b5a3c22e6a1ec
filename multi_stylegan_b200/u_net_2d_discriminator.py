"""U-Net discriminator (encoder/decoder of residual + non-local blocks, scalar and per-pixel heads) with
the reference's module tree and state_dict names (multi_stylegan/u_net_2d_discriminator.py:14-381).
Convolutions, blurs, upsampling and activations run on this package's sm_100a kernels, and so does the non-local
block in its first-order form (attention.py); the composite torch.bmm / softmax / max_pool2d formulation of the
reference (:370-380) remains for double backward (R1) and for the CPU.
The FFT input option relied on torch.rfft, which no longer exists; `fft: True` raises."""
import math
from typing import Any, Dict, List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _C, _mode, attention, conv, equalized_layer
from .multi_stylegan_generator import _fir_kernel
from .op_static import FusedLeakyReLU, upfirdn2d


class _tf32_matmul(object):
    """Scoped torch.backends.cuda.matmul.allow_tf32 = True (forward and the backward nodes recorded inside keep
    the setting of the time they run, so the train step also enables it around backward; see model_wrapper)."""

    def __enter__(self):
        self.old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.old
        return False


class Upsample(nn.Module):
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
    def __init__(self, kernel: List[int] = [1, 3, 3, 1], sampling_factor: int = 1, sampling_factor_padding: int = 2,
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


class _MbStdFused(torch.autograd.Function):
    """MinibatchStdDev on the library's kernels (first-order; msg_mbstd_forward / _backward)."""

    @staticmethod
    def forward(ctx, x, groups, alpha):
        ctx.save_for_backward(x)
        ctx.cfg = (groups, alpha)
        return _C.mbstd_forward(x, groups, alpha)

    @staticmethod
    def backward(ctx, gout):
        if torch.is_grad_enabled():
            raise RuntimeError(_mode.NO_DOUBLE_BACKWARD)
        x, = ctx.saved_tensors
        return _C.mbstd_backward(gout, x, *ctx.cfg), None, None


class MinibatchStdDev(nn.Module):
    """One extra channel holding the mean over (c,h,w) of the per-position batch std — reference :189-217."""

    def __init__(self, alpha: float = 1e-8) -> None:
        super().__init__()
        self.alpha = alpha
        self.groups = 1      # > 1: the batch is `groups` independent sub-batches (Discriminator.forward_pair)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if (input.is_cuda and input.dtype == torch.float32 and input.dim() == 4 and not _mode.higher_order()
                and input.shape[0] % self.groups == 0):
            return _MbStdFused.apply(input, self.groups, self.alpha)       # three launches instead of ~10 ATen kernels
        if self.groups > 1:
            G, Bg = self.groups, input.shape[0] // self.groups
            x = input.reshape(G, Bg, *input.shape[1:])
            centred = x - x.mean(dim=1, keepdim=True)
            std = torch.sqrt((centred ** 2).mean(dim=1).clamp(min=self.alpha)).mean(dim=(1, 2, 3))      # [G]
            plane = std.view(G, 1, 1, 1, 1).expand(G, Bg, 1, input.shape[2], input.shape[3])
            return torch.cat((input, plane.reshape(input.shape[0], 1, input.shape[2], input.shape[3])), 1)
        centred = input - input.mean(dim=0, keepdim=True)
        std = torch.sqrt((centred ** 2).mean(dim=0).clamp(min=self.alpha)).mean().view(1, 1, 1)
        return torch.cat((input, std.repeat(input.shape[0], 1, input.shape[2], input.shape[3])), 1)


class ResNetBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, mini_batch_std_dev: bool = False) -> None:
        super().__init__()
        self.mini_batch_std_dev = MinibatchStdDev() if mini_batch_std_dev else nn.Identity()
        self.main_mapping = nn.Sequential(
            equalized_layer.EqualizedConv2d(in_channels + 1 if mini_batch_std_dev else in_channels, out_channels,
                                            kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False),
            FusedLeakyReLU(out_channels),
            equalized_layer.EqualizedConv2d(out_channels, out_channels, kernel_size=(3, 3), stride=(1, 1),
                                            padding=(1, 1), bias=False),
            FusedLeakyReLU(out_channels))
        self.residual_mapping = equalized_layer.EqualizedConv2d(
            in_channels, out_channels, kernel_size=1, stride=1, padding=0, bias=False) \
            if in_channels != out_channels else nn.Identity()

    def forward(self, input: torch.Tensor, input_2: torch.Tensor = None) -> torch.Tensor:
        """`input_2` (an extension used by the U-Net decoder): the block is applied to the channel concatenation
        [input | input_2] (reference :137) without materialising it — the two convolutions that consume the block input
        read both tensors through the K loop of the implicit GEMM, and the backward yields the two input gradients as
        separate dense tensors instead of strided slices of one."""
        c1, a1, c2, a2 = self.main_mapping
        res = self.residual_mapping
        fusable = (c1.weight.shape[0] % 4 == 0 and c1.bias is None and c2.bias is None
                   and isinstance(res, equalized_layer.EqualizedConv2d) and res.bias is None)
        if input_2 is not None and not (fusable and isinstance(self.mini_batch_std_dev, nn.Identity)
                                        and _C.cat2_supported(input, input_2, c1.stride)):
            input, input_2 = torch.cat([input, input_2], dim=1), None
        if not fusable:
            output = self.main_mapping(self.mini_batch_std_dev(input))
            return (output + self.residual_mapping(input)) / math.sqrt(2)
        # (lrelu(conv(lrelu(conv(x)))) + conv1x1(x)) / sqrt(2) in three kernels: both activations and the residual
        # join run in the conv epilogues (reference :174-186)
        j = 1.0 / math.sqrt(2)
        if (isinstance(self.mini_batch_std_dev, nn.Identity) and not _mode.higher_order()
                and tuple(c1.weight.shape[-2:]) == (3, 3) and tuple(c2.weight.shape[-2:]) == (3, 3)
                and tuple(res.weight.shape[-2:]) == (1, 1) and c1.stride == (1, 1) and c1.padding == (1, 1)
                and c2.stride == (1, 1) and c2.padding == (1, 1) and res.stride == (1, 1) and res.padding == (0, 0)
                and a1.negative_slope == a2.negative_slope):
            # first-order form: one Function for the block, input gradients summed in the dgrad epilogue (conv.ResBlockFused)
            return conv.res_block(input, input_2, c1.weight, a1.bias, c2.weight, a2.bias, res.weight, c1.scale, c2.scale,
                                  res.scale * j, a1.negative_slope, a1.scale, a2.scale * j)
        x = self.mini_batch_std_dev(input)
        h = conv.conv2d_bias_act(x, c1.weight, bias=a1.bias, stride=c1.stride, padding=c1.padding,
                                 negative_slope=a1.negative_slope, gain=a1.scale, alpha=c1.scale, x2=input_2)
        # the join's 1/sqrt(2) is folded into the second activation's gain and the residual convolution's alpha, so
        # neither the forward nor the backward spends a pass on the scaling
        h = conv.conv2d_bias_act(h, c2.weight, bias=a2.bias, stride=c2.stride, padding=c2.padding,
                                 negative_slope=a2.negative_slope, gain=a2.scale * j, alpha=c2.scale)
        return conv.conv2d_add_scale(input, res.weight, h, stride=res.stride, padding=res.padding,
                                     gain=1.0, alpha=res.scale * j, x2=input_2)


class NonLocalBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        def c1(i, o):
            return equalized_layer.EqualizedConv2d(i, o, kernel_size=(1, 1), padding=(0, 0), bias=False)
        self.theta = c1(in_channels, out_channels // 8)
        self.phi = c1(in_channels, out_channels // 8)
        self.g = c1(in_channels, out_channels // 2)
        self.o = c1(out_channels // 2, out_channels)
        self.residual_mapping = c1(in_channels, out_channels) if in_channels != out_channels else nn.Identity()
        self.register_parameter(name="gamma", param=nn.Parameter(torch.tensor(0.)))

    def forward(self, input: torch.Tensor, input_2: torch.Tensor = None) -> torch.Tensor:
        """`input_2` (U-Net decoder): the block is applied to the channel concatenation [input | input_2] (reference :137)."""
        res = self.residual_mapping
        cq, cv = self.theta.weight.shape[0], self.g.weight.shape[0]
        if (not _mode.higher_order() and isinstance(res, equalized_layer.EqualizedConv2d) and res.bias is None
                and attention.eligible(input, cq, cv)):
            if input_2 is not None and not _C.cat2_supported(input, input_2, 1):
                input, input_2 = torch.cat([input, input_2], dim=1), None
            return attention.non_local_block(input, input_2, self.theta.weight, self.phi.weight, self.g.weight,
                                             self.o.weight, res.weight, self.gamma, self.theta.scale, self.o.scale,
                                             res.scale / math.sqrt(2))
        if input_2 is not None:
            input = torch.cat([input, input_2], dim=1)
        batch_size, _, height, width = input.shape

        def rows(t: torch.Tensor) -> torch.Tensor:
            """[B, C, h, w] -> [B, h*w, C]: a free view of a channels-last activation (its memory order)."""
            return t.permute(0, 2, 3, 1).reshape(t.shape[0], t.shape[2] * t.shape[3], t.shape[1])
        theta = rows(self.theta(input))                                                   # [B, HW, C/8]   = theta^T
        phi = rows(F.max_pool2d(self.phi(input), kernel_size=(2, 2), stride=(2, 2)))      # [B, HW/4, C/8] = phi^T
        g = rows(F.max_pool2d(self.g(input), kernel_size=(2, 2), stride=(2, 2)))          # [B, HW/4, C/2] = g^T
        # The 4096 x 1024 attention stays on the library bmm like the reference (:370-380), written on the transposed
        # operands so that no layout copy is needed: beta = softmax(theta^T phi), (g beta^T)^T = beta g^T, and
        # [B, HW, C/2] is already the channels-last image of the result.  The reference's pinned PyTorch 1.8.1 runs CUDA
        # matmuls with TF32 enabled by default; request the same here instead of the fp32 CUDA-core sgemm newer
        # PyTorch versions fall back to.
        with _tf32_matmul():
            beta = F.softmax(torch.bmm(theta, phi.transpose(1, 2)), -1)                   # [B, HW, HW/4]
            attended = torch.bmm(beta, g)                                                 # [B, HW, C/2]
        attended = attended.view(batch_size, height, width, -1).permute(0, 3, 1, 2)       # [B, C/2, H, W] channels-last
        output = self.o(attended)
        res = self.residual_mapping
        if isinstance(res, equalized_layer.EqualizedConv2d) and res.bias is None:
            # (gamma * o + conv1x1(x)) / sqrt(2) with the join in the residual conv's epilogue (:381)
            j = 1.0 / math.sqrt(2)
            return conv.conv2d_add_scale(input, res.weight, (self.gamma * j) * output, stride=res.stride,
                                         padding=res.padding, gain=1.0, alpha=res.scale * j)
        return (self.gamma * output + self.residual_mapping(input)) / math.sqrt(2)


class Discriminator(nn.Module):
    def __init__(self, config: Dict[str, Any], no_rfp: bool = False, no_gfp: bool = False) -> None:
        super().__init__()
        enc: Tuple[Tuple[int, int], ...] = config["encoder_channels"]
        dec: Tuple[Tuple[int, int], ...] = config["decoder_channels"]
        self.fft: bool = config["fft"]
        if self.fft:
            raise NotImplementedError("fft input relied on torch.rfft (removed from PyTorch); config 'fft' must be False")
        self.encoder_blocks = nn.ModuleList()
        for index, (c_in, c_out) in enumerate(enc):
            if index == 0:
                self.encoder_blocks.append(ResNetBlock(3 if no_gfp else (6 if no_rfp else 9), c_out))
            elif index == 2:
                self.encoder_blocks.append(NonLocalBlock(c_in, c_out))
            else:
                self.encoder_blocks.append(ResNetBlock(c_in, c_out, mini_batch_std_dev=index >= (len(enc) - 2)))
        self.downscale_convolutions = nn.ModuleList(
            [nn.Sequential(equalized_layer.EqualizedConv2d(c[1], c[1], kernel_size=(3, 3), stride=(2, 2),
                                                           padding=(0, 0)), Blur()) for c in enc[:-1]])
        self.classification_head = nn.Sequential(
            nn.AdaptiveAvgPool2d(output_size=(1, 1)),
            nn.Flatten(start_dim=1),
            equalized_layer.EqualizedLinear(enc[-1][-1], 128, bias=False),
            FusedLeakyReLU(channel=128),
            equalized_layer.EqualizedLinear(128, 1, bias=False))
        self.decoder_blocks = nn.ModuleList()
        for index, (c_in, c_out) in enumerate(dec):
            self.decoder_blocks.append(NonLocalBlock(c_in, c_out) if index == 1 else ResNetBlock(c_in, c_out))
        self.transposed_convolutions = nn.ModuleList()
        for current, past, d in zip(reversed(enc[1:]), reversed(enc[:-1]), dec):
            self.transposed_convolutions.append(nn.Sequential(
                Upsample(),
                equalized_layer.EqualizedConv2d(current[-1], d[0] - past[-1], kernel_size=(1, 1), stride=(1, 1),
                                                padding=(0, 0), bias=False)))
        self.final_mapping = nn.Sequential(
            FusedLeakyReLU(channel=dec[-1][-1]),
            equalized_layer.EqualizedConv2d(dec[-1][-1], 1, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                                            bias=False))

    def forward(self, input: torch.Tensor, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        """input [B, channels, frames, H, W] -> (scalar [B,1], pixel-wise [B,1,1,H,W]); extra kwargs
        (is_real / is_cut_mix from the train step) are accepted and ignored like the reference (:99)."""
        x = input.flatten(start_dim=1, end_dim=2)
        features = []
        last = len(self.encoder_blocks) - 1
        for index, block in enumerate(self.encoder_blocks):
            x = block(x)
            if index != last:
                down, blur = self.downscale_convolutions[index][0], self.downscale_convolutions[index][1]
                if not _mode.higher_order() and x.dtype == torch.float32:
                    # strided conv + bias and the skip tap as one Function: bias in the conv epilogue, the skip gradient
                    # summed in the conv's dgrad epilogue (conv.DownscaleTapFused)
                    y, skip = conv.downscale_with_tap(x, down.weight, down.bias, down.stride, down.padding, down.scale,
                                                      down.scale_bias)
                    features.append(skip)
                    x = blur(y)
                else:
                    features.append(x)
                    x = blur(down(x))
        classification = self.classification_head(x)
        for block, up, skip in zip(self.decoder_blocks, self.transposed_convolutions, reversed(features)):
            # reference: conv1x1(Upsample(x)) (:87,:135-137).  A 1x1 convolution (per pixel, across channels) and the
            # FIR upsampling (per channel, across pixels) commute exactly, so the convolution runs first, on a quarter
            # of the pixels, and the upsampling on the smaller channel count.
            upsample, conv1x1 = up[0], up[1]
            up_x = upsample(conv1x1(x))
            x = block(up_x, skip)                      # both block types read the concatenation [up_x | skip] in place
        return classification, self.final_mapping(x).unsqueeze(dim=2)


    def forward_pair(self, first: torch.Tensor, second: torch.Tensor):
        """(D(first), D(second)) as ONE batched pass.  Every operation of the network is per-sample except
        MinibatchStdDev, which is evaluated per half, so the results equal two separate calls (the train step's
        real / fake passes, model_wrapper.py:279-283) with half the launches and better-filled small layers."""
        assert first.shape == second.shape, "forward_pair needs two batches of the same shape"
        n = first.shape[0]
        stds = [m for m in self.modules() if isinstance(m, MinibatchStdDev)]
        for m in stds:
            m.groups = 2
        try:
            classification, pixel_wise = self.forward(torch.cat([first, second], dim=0))
        finally:
            for m in stds:
                m.groups = 1
        return (classification[:n], pixel_wise[:n]), (classification[n:], pixel_wise[n:])


# ---- CutMix helpers (u_net_2d_discriminator.py:384-448) ------------------------------------------------
def _generate_binary_cut_mix_map(height: int, width: int, device="cpu") -> torch.Tensor:
    import random
    binary_map = torch.zeros(1, 1, 1, height, width, dtype=torch.float, device=device)
    ch = int(torch.randint(int(0.1 * height), int(0.9 * height), size=(1,)))
    cw = int(torch.randint(int(0.1 * width), int(0.9 * width), size=(1,)))
    if random.random() > 0.5:
        binary_map[..., ch:, cw:] = 1.0
    else:
        binary_map[..., :ch, :cw] = 1.0
    if random.random() > 0.5:
        binary_map = -binary_map + 1.
    return binary_map


def generate_cut_mix_augmentation_data(image_real: torch.Tensor, image_fake: torch.Tensor):
    image_fake = image_fake[:image_real.shape[0]]
    target = _generate_binary_cut_mix_map(image_real.shape[-2], image_fake.shape[-1], image_real.device)
    return image_real * target + image_fake * (-target + 1.), target


def generate_cut_mix_transformation_data(image_real, image_fake, prediction_real, prediction_fake):
    image_fake = image_fake[:image_real.shape[0]]
    prediction_fake = prediction_fake[:image_real.shape[0]]
    m = _generate_binary_cut_mix_map(image_real.shape[-2], image_fake.shape[-1], image_real.device)
    return image_real * m + image_fake * (-m + 1.), prediction_real * m + prediction_fake * (-m + 1.)
