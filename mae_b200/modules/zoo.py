"""Registered conv stacks and flows of the reference's shipped configurations, built from
``networks.py`` with the reference's module names, so that ``MAE.from_params(json)`` works on the reference's config files
and its checkpoints load with ``strict=True``.  Registry keys as in the reference:

  EncoderCore  'resnet_binary_28x28'   encoder_cores/resnet.py:13-51     'resnet_color_32x32'  encoder_cores/resnet.py:54-111
  Decoder      'resnet_binary_28x28'   image_decoders/resnet.py:12-55    'resnet_color_32x32'  image_decoders/resnet.py:58-121
               'pixelcnn_binary_28x28' image_decoders/pixelcnn.py:22-163
               'pixelcnn++_color_32x32' / '_64x64'  image_decoders/pixelcnnpp.py:54-170
  EncoderCore  'densenet_color_32x32' / '_64x64'    encoder_cores/densenet.py:13-112
  Flow         'af_made'               flows/af/made.py:12-123           'af2d_made'  flows/af/made2d.py:12-131

Everything here is ordinary PyTorch; the loss / evaluation arithmetic behind these stacks is the CUDA path
(``BinaryImageDecoder`` / ``ColorImageDecoder`` / ``GaussianEncoder`` in this package)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .decoder import BinaryImageDecoder, ColorImageDecoder
from .encoder import EncoderCore
from .flow import Flow
from .networks import (MADE, MADE2d, Conv2dWeightNorm, ConvTranspose2dWeightNorm, DenseNet, DeResNet, LinearWeightNorm,
                       PixelCNN, PixelCNNPP, ReShape, ResNet, _call, _run, cached_weights)

__all__ = ["ResnetEncoderCoreBinaryImage28x28", "ResnetEncoderCoreColorImage32x32", "ResnetDecoderBinaryImage28x28",
           "ResnetDecoderColorImage32x32", "PixelCNNDecoderBinaryImage28x28", "AFMADE", "DenseNetEncoderCoreColorImage32x32",
           "DenseNetEncoderCoreColorImage64x64", "PixelCNNPPDecoderColorImage32x32", "PixelCNNPPDecoderColorImage64x64",
           "AF2dMADE"]


def _unwrap(m):
    return m.module if isinstance(m, nn.DataParallel) else m


def _replicate(m, ngpu):
    return nn.DataParallel(m, device_ids=list(range(ngpu))) if ngpu > 1 else m


# ------------------------------------------------------------------------------------------------ encoder cores
class ResnetEncoderCoreBinaryImage28x28(EncoderCore):
    """28x28x1 -> (mu, logvar) [B, nz]: three stride-2 residual blocks, a 4x4 valid convolution to 512 units, a
    weight-normalised linear head whose output is chunked (the strided views the hot path consumes without a copy)."""

    def __init__(self, nz):
        super().__init__()
        self.nz, self.nc = nz, 1
        hidden_units = 512
        self.main = nn.Sequential(
            ResNet(self.nc, [32, 64, 64], [2, 2, 2]),
            Conv2dWeightNorm(64, hidden_units, 4, 1, 0),
            nn.ELU(),
        )
        self.linear = LinearWeightNorm(hidden_units, 2 * nz)

    def _features(self, x, init, init_scale):
        h = _run(self.main, x, init, init_scale)
        return _run([self.linear], h.view(h.size(0), h.size(1)), init, init_scale)

    def forward(self, input):
        return self._features(input, False, 1.0).chunk(2, 1)

    def initialize(self, x, init_scale=1.0):
        return self._features(x, True, init_scale).chunk(2, 1)

    def output_size(self) -> Tuple:
        return self.nz,

    @classmethod
    def from_params(cls, params: Dict) -> "ResnetEncoderCoreBinaryImage28x28":
        return cls(**params)


class ResnetEncoderCoreColorImage32x32(EncoderCore):
    """32x32x3 -> (mu, hardtanh(logvar, -7, 7)) [B, z_channels, 8, 8]."""

    def __init__(self, z_channels):
        super().__init__()
        self.z_channels, self.H, self.W, self.nc = z_channels, 8, 8, 3
        self.main = nn.Sequential(
            ResNet(self.nc, [48, 48], [1, 1]),
            Conv2dWeightNorm(48, 96, 3, 2, 1),
            nn.ELU(),
            ResNet(96, [96, 96], [1, 1]),
            Conv2dWeightNorm(96, 96, 3, 2, 1),
            nn.ELU(),
            ResNet(96, [96, 96, 96], [1, 1, 1]),
            Conv2dWeightNorm(96, 48, 1, 1),
            nn.ELU(),
            Conv2dWeightNorm(48, 2 * z_channels, 1, 1),
        )

    def _split(self, out):
        mu, logvar = out.chunk(2, 1)
        return mu, F.hardtanh(logvar, min_val=-7., max_val=7.)

    def forward(self, input):
        return self._split(self.main(input))

    def initialize(self, x, init_scale=1.0):
        return self._split(_run(self.main, x, True, init_scale))

    def output_size(self) -> Tuple:
        return self.z_channels, self.H, self.W

    @classmethod
    def from_params(cls, params: Dict) -> "ResnetEncoderCoreColorImage32x32":
        return cls(**params)


# ----------------------------------------------------------------------------------------------------- decoders
class ResnetDecoderBinaryImage28x28(BinaryImageDecoder):
    """z [B, nz, 1, 1] -> Bernoulli probabilities [B, 1, 28, 28]."""

    def __init__(self, nz, ngpu=1):
        super().__init__(nz, ngpu=ngpu)
        self.nc = 1
        self.core = _replicate(nn.Sequential(
            ConvTranspose2dWeightNorm(nz, 64, 4, 1, 0),
            nn.ELU(),
            DeResNet(64, [64, 32, self.nc], [2, 2, 2], [0, 1, 1]),
            nn.Sigmoid(),
        ), ngpu)

    def forward(self, z):
        return self.core(z)

    def initialize(self, x, z, init_scale=1.0):
        return _run(_unwrap(self.core), z, True, init_scale)

    def output_size(self) -> Tuple:
        return 1, 28, 28

    @classmethod
    def from_params(cls, params: Dict) -> "ResnetDecoderBinaryImage28x28":
        return cls(**params)


class ResnetDecoderColorImage32x32(ColorImageDecoder):
    """z [B, z_channels, 8, 8] -> raw mixture parameters [B, 10*nmix, 32, 32] (the likelihood kernel's input)."""

    def __init__(self, z_channels, nmix, ngpu=1):
        super().__init__(nmix, ngpu=ngpu)
        self.z_channels, self.H, self.W = z_channels, 8, 8
        self.core = _replicate(nn.Sequential(
            ConvTranspose2dWeightNorm(z_channels, 96, 1, 1, 0),
            nn.ELU(),
            DeResNet(96, [96, 96, 96], [1, 1, 1], [0, 0, 0]),
            ConvTranspose2dWeightNorm(96, 96, 3, 2, 1, 1),
            nn.ELU(),
            DeResNet(96, [96, 96], [1, 1], [0, 0]),
            ConvTranspose2dWeightNorm(96, 48, 3, 2, 1, 1),
            nn.ELU(),
            DeResNet(48, [48, 48], [1, 1], [0, 0]),
            Conv2dWeightNorm(48, (self.nc * 3 + 1) * self.nmix, 1),
        ), ngpu)

    def z_shape(self) -> Tuple:
        return self.z_channels, self.H, self.W

    def output_size(self) -> Tuple:
        return self.nc, 32, 32

    def forward(self, x, z):
        return self.core(z)

    def head_parts(self):
        """image_decoders/resnet.py:85: the stack ends in Conv2dWeightNorm(48, 10*nmix, 1) on the last DeResNet's output."""
        if isinstance(self.core, nn.DataParallel):
            return None
        core = self.core
        return (lambda x, z: core[:-1](z)), core[-1], False

    def initialize(self, x, z, init_scale=1.0):
        return _run(_unwrap(self.core), z, True, init_scale)

    @classmethod
    def from_params(cls, params: Dict) -> "ResnetDecoderColorImage32x32":
        return cls(**params)


class PixelCNNDecoderBinaryImage28x28(BinaryImageDecoder):
    """Autoregressive Bernoulli decoder: z is mapped to 4 conditioning feature maps that are concatenated to the image;
    a PixelCNN whose first (type-A) layer masks only the image channel predicts every pixel from the pixels before it
    and from z (van den Oord et al. 2016; Chen et al. 2016)."""

    _KERNELS = {'small': [7, 7, 7, 5, 5, 3, 3], 'large': [7, 7, 7, 7, 7, 5, 5, 5, 5, 3, 3, 3, 3]}

    def __init__(self, nz, mode, ngpu=1):
        super().__init__(nz, ngpu=ngpu)
        if mode not in self._KERNELS:
            raise ValueError('unknown mode: %s' % mode)
        kernels = self._KERNELS[mode]
        self.nc, self.fm_latent = 1, 4
        self.img_latent = 28 * 28 * self.fm_latent
        hidden_channels = 64
        self.z_transform = _replicate(nn.Sequential(
            LinearWeightNorm(nz, self.img_latent),
            ReShape((self.fm_latent, 28, 28)),
            nn.ELU(),
        ), ngpu)
        self.core = _replicate(nn.Sequential(
            PixelCNN(self.nc + self.fm_latent, hidden_channels, len(kernels), kernels, self.nc),
            Conv2dWeightNorm(hidden_channels, hidden_channels, 1),
            nn.ELU(),
            Conv2dWeightNorm(hidden_channels, self.nc, 1),
            nn.Sigmoid(),
        ), ngpu)

    def z_shape(self) -> Tuple:
        return self.nz,

    def output_size(self) -> Tuple:
        return self.nc, 28, 28

    def decode(self, z, random_sample):
        """Ancestral sampling, one pixel per network evaluation (image_decoders/pixelcnn.py:72-104)."""
        side = 28
        cond = self.z_transform(z)
        canvas = torch.cat([cond.new_zeros(z.size(0), self.nc, side, side), cond], dim=1)
        with cached_weights():
            for i in range(side):
                for j in range(side):
                    probs = self.core(canvas)[:, :, i, j]
                    canvas[:, :self.nc, i, j] = torch.bernoulli(probs) if random_sample else probs.ge(0.5).float()
            return canvas[:, :self.nc], self.core(canvas)

    def likelihood_input(self, x, z):
        """Bernoulli probabilities [B, ns, 784] of every pixel given the TRUE previous pixels and z [B, ns, nz]: what the
        likelihood kernel reads (image_decoders/pixelcnn.py:119-134 up to the BCE, which is the CUDA kernel's)."""
        b, ns = z.size(0), z.size(1)
        cond = self.z_transform(z.reshape(b * ns, *z.size()[2:]))
        img = x.unsqueeze(1).expand(b, ns, *x.size()[1:]).reshape(b * ns, *x.size()[1:])
        return self.core(torch.cat([img, cond], dim=1)).view(b, ns, -1)

    def initialize(self, x, z, init_scale=1.0):
        cond = _run(_unwrap(self.z_transform), z, True, init_scale)
        return _run(_unwrap(self.core), torch.cat([x, cond], dim=1), True, init_scale)

    @classmethod
    def from_params(cls, params: Dict) -> "PixelCNNDecoderBinaryImage28x28":
        return cls(**params)


# -------------------------------------------------------------------------------------------------------- flows
class AFMADEBlock(nn.Module):
    """One autoregressive affine layer.  ``backward`` (the cheap direction: one MADE evaluation) maps y -> x = mu(y) +
    y * exp(0.5 logvar(y)); ``forward`` inverts it coordinate by coordinate (flows/af/made.py:12-55)."""

    def __init__(self, input_size, num_hiddens, hidden_size, order, bias=True, var=True):
        super().__init__()
        if input_size <= 0:
            raise ValueError('input size (%s) should be positive' % input_size)
        self.order = order
        self.mu = MADE(input_size, num_hiddens, hidden_size, order, bias=bias)
        self.logvar = MADE(input_size, num_hiddens, hidden_size, order, bias=bias) if var else None
        self._input_size = input_size

    def _affine(self, y, init, init_scale):
        mu = _run([self.mu], y, init, init_scale)
        logstd = _run([self.logvar], y, init, init_scale) * 0.5 if self.logvar is not None else torch.zeros_like(mu)
        return mu, logstd

    def backward(self, y):
        mu, logstd = self._affine(y, False, 1.0)
        return mu + y * logstd.exp(), logstd.sum(dim=1)

    def initialize(self, y, init_scale=1.0):
        mu, logstd = self._affine(y, True, init_scale)
        return mu + y * logstd.exp(), logstd.sum(dim=1)

    def forward(self, x):
        eps = 1e-12
        y = torch.zeros_like(x)
        logstd = torch.zeros_like(x)
        coords = range(self._input_size) if self.order == 'A' else reversed(range(self._input_size))
        with cached_weights():
            for i in coords:       # coordinate i of (mu, logstd) only depends on the coordinates already solved
                mu, logstd = self._affine(y, False, 1.0)
                y[:, i] = ((x - mu) / (logstd.exp() + eps))[:, i]
        return y, logstd.sum(dim=1)


class AFMADEDualBlock(nn.Module):
    """Two autoregressive layers with opposite orderings ('A' then 'B')."""

    def __init__(self, input_size, num_hiddens, hidden_size, bias=True, var=True):
        super().__init__()
        self.fwd = AFMADEBlock(input_size, num_hiddens, hidden_size, 'A', bias=bias, var=var)
        self.bwd = AFMADEBlock(input_size, num_hiddens, hidden_size, 'B', bias=bias, var=var)

    def forward(self, x):
        x, ld_f = self.fwd.forward(x)
        y, ld_b = self.bwd.forward(x)
        return y, ld_f + ld_b

    def backward(self, y):
        y, ld_b = self.bwd.backward(y)
        x, ld_f = self.fwd.backward(y)
        return x, ld_f + ld_b

    def initialize(self, y, init_scale=1.0):
        y, ld_b = self.bwd.initialize(y, init_scale=init_scale)
        x, ld_f = self.fwd.initialize(y, init_scale=init_scale)
        return x, ld_f + ld_b


class AFMADE(Flow):
    """Prior flow of the binary-image configuration: ``backward`` (z -> standard-normal space, used by
    ``log_probability_prior`` on [B*K, nz]) costs one MADE pass per block; ``forward`` (sampling) is sequential."""

    def __init__(self, input_size, num_blocks, num_hiddens=1, hidden_size=None, bias=True, var=True):
        super().__init__()
        self._input_size = input_size
        self.num_blocks, self.num_hiddens = num_blocks, num_hiddens
        self.hidden_size = input_size * 10 if hidden_size is None else hidden_size
        self.blocks = nn.ModuleList([AFMADEDualBlock(input_size, num_hiddens, self.hidden_size, bias=bias, var=var)
                                     for _ in range(num_blocks)])

    def input_size(self) -> Tuple:
        return self._input_size,

    def forward(self, x: torch.Tensor, init=False, init_scale=1.0):
        if init:
            raise AssertionError('AF does not support initialization via forward.')
        logdet = x.new_zeros(x.size(0))
        for block in reversed(self.blocks):
            x, ld = block.forward(x)
            logdet = logdet + ld
        return x, logdet

    def backward(self, y: torch.Tensor, init=False, init_scale=1.0):
        logdet = y.new_zeros(y.size(0))
        for block in self.blocks:
            y, ld = block.initialize(y, init_scale=init_scale) if init else block.backward(y)
            logdet = logdet + ld
        return y, logdet

    @classmethod
    def from_params(cls, params: Dict) -> "AFMADE":
        return cls(**params)


# ------------------------------------------------------------------- colour configuration (color_image32x32.json)
class DenseNetEncoderCoreColorImage32x32(EncoderCore):
    """32x32x3 -> (mu, hardtanh(logvar, -7, 7)) [B, z_channels, 8, 8] through three dense stages."""

    def __init__(self, z_channels):
        super().__init__()
        self.z_channels, self.H, self.W, self.nc = z_channels, 8, 8, 3
        self.main = nn.Sequential(*self._stages(z_channels))

    def _stages(self, z_channels):
        return [
            DenseNet(self.nc, 15, 3),                    # 48 x 32 x 32
            Conv2dWeightNorm(48, 48, 3, 2, 1), nn.ELU(),  # 48 x 16 x 16
            DenseNet(48, 16, 3),                         # 96 x 16 x 16
            Conv2dWeightNorm(96, 96, 3, 2, 1), nn.ELU(),  # 96 x 8 x 8
            DenseNet(96, 16, 6),                         # 192 x 8 x 8
            Conv2dWeightNorm(192, 96, 1, 1), nn.ELU(),
            Conv2dWeightNorm(96, 2 * z_channels, 1, 1),
        ]

    def _split(self, out):
        mu, logvar = out.chunk(2, 1)
        return mu, F.hardtanh(logvar, min_val=-7., max_val=7.)

    def forward(self, input):
        return self._split(self.main(input))

    def initialize(self, x, init_scale=1.0):
        return self._split(_run(self.main, x, True, init_scale))

    def output_size(self) -> Tuple:
        return self.z_channels, self.H, self.W

    @classmethod
    def from_params(cls, params: Dict):
        return cls(**params)


class DenseNetEncoderCoreColorImage64x64(DenseNetEncoderCoreColorImage32x32):
    """64x64x3 variant: one more stride-2 stage."""

    def _stages(self, z_channels):
        return [
            DenseNet(self.nc, 15, 3),                      # 48 x 64 x 64
            Conv2dWeightNorm(48, 48, 3, 2, 1), nn.ELU(),    # 48 x 32 x 32
            DenseNet(48, 16, 3),                           # 96 x 32 x 32
            Conv2dWeightNorm(96, 96, 3, 2, 1), nn.ELU(),    # 96 x 16 x 16
            DenseNet(96, 16, 3),                           # 144 x 16 x 16
            Conv2dWeightNorm(144, 144, 3, 2, 1), nn.ELU(),  # 144 x 8 x 8
            DenseNet(144, 16, 3),                          # 192 x 8 x 8
            Conv2dWeightNorm(192, 96, 1, 1), nn.ELU(),
            Conv2dWeightNorm(96, 2 * z_channels, 1, 1),
        ]


class _PixelCNNPPCore(nn.Module):
    """z [B, zc, 8, 8] -> conditioning map h at image resolution (transposed convolutions), PixelCNN++ trunk over
    (x, h), two 1x1 convolutions to the 10*nmix raw mixture parameters the likelihood kernel reads."""

    def __init__(self, nc, z_channels, h_channels, hidden_channels, levels, num_resnets, nmix, dropout=0., activation='concat_elu'):
        super().__init__()
        from collections import OrderedDict
        layers, c = OrderedDict(), z_channels
        for i in range(levels - 1):
            layers['level%d' % i] = ConvTranspose2dWeightNorm(c, c // 2, 3, 2, 1, 1)
            layers['elu%d' % i] = nn.ELU()
            c //= 2
        layers['h_level'] = Conv2dWeightNorm(c, h_channels, 1)
        self.z_transform = nn.Sequential(layers)
        self.core = PixelCNNPP(levels, nc, hidden_channels, num_resnets, h_channels, dropout=dropout, activation=activation)
        self.output = nn.Sequential(
            nn.ELU(),
            Conv2dWeightNorm(hidden_channels, hidden_channels, 1),
            nn.ELU(),
            Conv2dWeightNorm(hidden_channels, (nc * 3 + 1) * nmix, 1),
        )

    def _pass(self, x, z, init, s):
        h = _run(self.z_transform, z, init, s)
        return _run(self.output, _call(self.core, init, s, x, h=h), init, s)

    def forward(self, x, z):
        return self._pass(x, z, False, 1.0)

    def trunk(self, x, z):
        """Everything up to the input of the last ELU -> 1x1 convolution pair (image_decoders/pixelcnnpp.py:28-34)."""
        h = self.z_transform(z)
        return self.output[1](self.output[0](self.core(x, h=h)))

    def initialize(self, x, z, init_scale=1.0):
        return self._pass(x, z, True, init_scale)


class PixelCNNPPDecoderColorImage32x32(ColorImageDecoder):
    """PixelCNN++ decoder of the CIFAR configuration: raw [B, 10*nmix, 32, 32] from (x, z)."""

    _GEOMETRY = dict(image_size=32, levels=3, hidden_channels=96, num_resnets=6)

    def __init__(self, z_channels, h_channels, nmix, dropout=0., activation='elu', ngpu=1):
        super().__init__(nmix, ngpu=ngpu)
        g = self._GEOMETRY
        self.z_channels, self.H, self.W, self.image_size = z_channels, 8, 8, g['image_size']
        self.core = _replicate(_PixelCNNPPCore(self.nc, z_channels, h_channels, g['hidden_channels'], g['levels'],
                                               g['num_resnets'], nmix, dropout=dropout, activation=activation), ngpu)

    def z_shape(self) -> Tuple:
        return self.z_channels, self.H, self.W

    def output_size(self) -> Tuple:
        return self.nc, self.image_size, self.image_size

    def forward(self, x, z):
        return self.core(x, z)

    def head_parts(self):
        if isinstance(self.core, nn.DataParallel):
            return None
        return self.core.trunk, self.core.output[3], True

    def decode(self, z, random_sample):
        """Ancestral sampling, one pixel per network evaluation (image_decoders/pixelcnnpp.py:101-123)."""
        side = self.image_size
        x = z.new_zeros(z.size(0), self.nc, side, side)
        with cached_weights():
            for i in range(side):
                for j in range(side):
                    x[:, :, i, j] = self.sample_pixels(*self.execute(z, x), random_sample)[:, :, i, j]
        return x, None

    def initialize(self, x, z, init_scale=1.0):
        return _unwrap(self.core).initialize(x, z, init_scale=init_scale)

    @classmethod
    def from_params(cls, params: Dict):
        return cls(**params)


class PixelCNNPPDecoderColorImage64x64(PixelCNNPPDecoderColorImage32x32):
    _GEOMETRY = dict(image_size=64, levels=4, hidden_channels=64, num_resnets=4)


class AF2dMADEBlock(nn.Module):
    """Autoregressive affine layer over [C, H, W] in raster order (flows/af/made2d.py:12-62); see AFMADEBlock."""

    def __init__(self, in_channels, kernel_size, num_hiddens, hidden_channels, hidden_kernels, order, bias=True, var=True):
        super().__init__()
        self.order = order
        args = (in_channels, kernel_size, num_hiddens, hidden_channels, hidden_kernels, order)
        self.mu = MADE2d(*args, bias=bias)
        self.logvar = MADE2d(*args, bias=bias) if var else None

    def _affine(self, y, init, init_scale):
        mu = _run([self.mu], y, init, init_scale)
        logstd = _run([self.logvar], y, init, init_scale) * 0.5 if self.logvar is not None else torch.zeros_like(mu)
        return mu, logstd

    def backward(self, y):
        mu, logstd = self._affine(y, False, 1.0)
        return mu + y * logstd.exp(), logstd.flatten(1).sum(dim=1)

    def initialize(self, y, init_scale=1.0):
        mu, logstd = self._affine(y, True, init_scale)
        return mu + y * logstd.exp(), logstd.flatten(1).sum(dim=1)

    def forward(self, x):
        eps = 1e-12
        y, logstd = torch.zeros_like(x), torch.zeros_like(x)
        rows, cols = list(range(x.size(2))), list(range(x.size(3)))
        if self.order == 'B':
            rows, cols = rows[::-1], cols[::-1]
        with cached_weights():
            for i in rows:
                for j in cols:  # position (i, j) of (mu, logstd) only depends on the positions already solved
                    mu, logstd = self._affine(y, False, 1.0)
                    y[:, :, i, j] = ((x - mu) / (logstd.exp() + eps))[:, :, i, j]
        return y, logstd.flatten(1).sum(dim=1)


class AF2dMADEDualBlock(AFMADEDualBlock):
    def __init__(self, in_channels, kernel_size, num_hiddens, hidden_channels, hidden_kernels, bias=True, var=True):
        nn.Module.__init__(self)
        args = (in_channels, kernel_size, num_hiddens, hidden_channels, hidden_kernels)
        self.fwd = AF2dMADEBlock(*args, 'A', bias=bias, var=var)
        self.bwd = AF2dMADEBlock(*args, 'B', bias=bias, var=var)


class AF2dMADE(AFMADE):
    """Prior flow of the colour configurations over z [C, 8, 8] (flows/af/made2d.py:88-131)."""

    def __init__(self, in_channels, height, width, num_blocks, kernel_size, num_hiddens=1, hidden_channels=None,
                 hidden_kernels=None, bias=True, var=True):
        Flow.__init__(self)
        self._input_shape = (in_channels, height, width)
        hidden_channels = in_channels * 8 if hidden_channels is None else hidden_channels
        hidden_kernels = [kernel_size] * num_hiddens if hidden_kernels is None else hidden_kernels
        self.blocks = nn.ModuleList([AF2dMADEDualBlock(in_channels, kernel_size, num_hiddens, hidden_channels, hidden_kernels,
                                                       bias=bias, var=var) for _ in range(num_blocks)])

    def input_size(self) -> Tuple:
        return self._input_shape


ResnetEncoderCoreBinaryImage28x28.register('resnet_binary_28x28')
ResnetEncoderCoreColorImage32x32.register('resnet_color_32x32')
ResnetDecoderBinaryImage28x28.register('resnet_binary_28x28')
ResnetDecoderColorImage32x32.register('resnet_color_32x32')
PixelCNNDecoderBinaryImage28x28.register('pixelcnn_binary_28x28')
AFMADE.register('af_made')
DenseNetEncoderCoreColorImage32x32.register('densenet_color_32x32')
DenseNetEncoderCoreColorImage64x64.register('densenet_color_64x64')
PixelCNNPPDecoderColorImage32x32.register('pixelcnn++_color_32x32')
PixelCNNPPDecoderColorImage64x64.register('pixelcnn++_color_64x64')
AF2dMADE.register('af2d_made')
