"""PyTorch building blocks of the conv stacks around the hot path: weight-normalised layers with
data-dependent initialisation, residual stacks, masked layers, PixelCNN and MADE.

These are the callers' side of the boundary: plain ``torch.nn`` (cuDNN / cuBLAS), no kernels of ours.  They are written
from the published layer definitions with the SAME module / parameter names as the reference so that its ``model.pt``
files load with ``strict=True``:

  reference                                               state_dict keys kept
  networks/weight_norm.py:8-128                           ``<name>.linear|conv|deconv.{weight_g, weight_v, bias}``
  networks/resnet.py:23-154                               ``main.<i>.{conv1, conv2, downsample}`` / ``{deconv1, deconv2, downsample}``
  networks/masked.py:15-212                               ``{weight_v, weight_g (log scale), bias, mask}``
  networks/auto_regressives/pixelcnn.py:8-105             ``main.<i>.main.<j>``, ``direct_connects.<i>.main.<j>``
  networks/auto_regressives/made.py:8-121                 ``input_layer|top_layer, direct_connect, hidden_layers.<i>, output_layer``
  networks/dense.py:13-54                                 ``main.block<i>.main.<j>``
  networks/auto_regressives/pixelcnnpp.py:21-338          ``up_layers, down_layers, nins1, nins2, up_hs, down_hs``

Differences in mechanism (not in function): weight normalisation is evaluated explicitly in ``forward`` instead of through
``torch.nn.utils.weight_norm`` hooks, and masks are applied functionally (``weight_v * mask``) instead of overwriting
``weight_v.data`` on every call (masked entries are zero from construction on and receive a zero gradient, so they stay
zero; outputs and state dicts are identical).  ``cached_weights()`` memoises the effective weights across
the sequential evaluations of the autoregressive directions.
"""
from __future__ import annotations

import contextlib
import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["LinearWeightNorm", "Conv2dWeightNorm", "ConvTranspose2dWeightNorm", "ResNet", "DeResNet", "MaskedLinear",
           "MaskedConv2d", "PixelCNN", "MADE", "ReShape", "DenseNet", "PixelCNNPP", "MADE2d", "cached_weights"]

_INIT_STD = 0.05
_STD_FLOOR = 1e-6

_weight_cache = None     # id(layer) -> effective weight, only inside cached_weights() with autograd off


@contextlib.contextmanager
def cached_weights():
    """Inside this context (and only while autograd is off) every weight-normalised / masked layer evaluates its effective
    weight once and reuses it: the sequential directions of the autoregressive models (``Flow.forward``: D evaluations of
    the same MADE, pixel-by-pixel ``decode``: H*W evaluations of the same PixelCNN) otherwise re-normalise and re-mask
    every layer's weight on every call (masked.py:102-106).  SURVEY.md §8f N4."""
    global _weight_cache
    outer = _weight_cache
    if outer is None:
        _weight_cache = {}
    try:
        yield
    finally:
        _weight_cache = outer


def _memo(layer, compute):
    if _weight_cache is None or torch.is_grad_enabled():
        return compute()
    w = _weight_cache.get(id(layer))
    if w is None:
        w = _weight_cache[id(layer)] = compute()
    return w


def _rownorm(v: torch.Tensor, dim: int) -> torch.Tensor:
    """L2 norm over every dimension except ``dim`` (kept, so it broadcasts against ``v``)."""
    dims = [i for i in range(v.dim()) if i != dim]
    return torch.linalg.vector_norm(v, dim=dims, keepdim=True)     # zero sub-gradient at 0 (fully masked rows)


def _channel_stats(out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """mean and unbiased std of every output channel (dim 1) over batch and positions."""
    flat = out.transpose(0, 1).reshape(out.size(1), -1)
    return flat.mean(dim=1), flat.std(dim=1)


class _NormedWeight(nn.Module):
    """Holds (weight_g, weight_v, bias) under the names ``torch.nn.utils.weight_norm`` would create and rebuilds
    ``w = g * v / ||v||`` (norm over everything but ``norm_dim``) on demand."""

    def __init__(self, shape: Sequence[int], norm_dim: int, n_out: int, bias: bool):
        super().__init__()
        self.norm_dim = norm_dim
        v = torch.empty(*shape).normal_(mean=0.0, std=_INIT_STD)
        self.bias = nn.Parameter(torch.zeros(n_out)) if bias else None
        self.weight_g = nn.Parameter(_rownorm(v, norm_dim).clone())
        self.weight_v = nn.Parameter(v)

    def weight(self) -> torch.Tensor:
        return _memo(self, lambda: self.weight_v * (self.weight_g / _rownorm(self.weight_v, self.norm_dim)))

    def rescale(self, mean: torch.Tensor, std: torch.Tensor, init_scale: float) -> None:
        """Data-dependent init (Salimans & Kingma 2016): make the layer's output zero-mean with std ``init_scale``."""
        gain = init_scale / (std + _STD_FLOOR)
        g_shape = [1] * self.weight_g.dim()
        g_shape[self.norm_dim] = -1
        self.weight_g.mul_(gain.view(g_shape))
        if self.bias is not None:
            self.bias.sub_(mean).mul_(gain)


class LinearWeightNorm(nn.Module):
    """Weight-normalised affine map; parameters live in ``self.linear`` (weight_norm.py:8-43)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.linear = _NormedWeight((out_features, in_features), 0, out_features, bias)

    def forward(self, input):
        return F.linear(input, self.linear.weight(), self.linear.bias)

    def initialize(self, x, init_scale=1.0):
        with torch.no_grad():
            out = self(x)
            self.linear.rescale(out.mean(dim=0), out.std(dim=0), init_scale)
            return self(x)

    def extra_repr(self):
        return 'in_features={}, out_features={}, bias={}'.format(self.in_features, self.out_features, self.linear.bias is not None)


class Conv2dWeightNorm(nn.Module):
    """Weight-normalised convolution; parameters live in ``self.conv`` (weight_norm.py:46-86)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        k = nn.modules.utils._pair(kernel_size)
        self.conv = _NormedWeight((out_channels, in_channels // groups, *k), 0, out_channels, bias)
        self.geometry = dict(stride=stride, padding=padding, dilation=dilation, groups=groups)

    def _conv(self, input):
        return F.conv2d(input, self.conv.weight(), self.conv.bias, **self.geometry)

    def forward(self, input):
        return self._conv(input)

    def initialize(self, x, init_scale=1.0):
        with torch.no_grad():
            mean, std = _channel_stats(self(x))
            self.conv.rescale(mean, std, init_scale)
            return self(x)


class ConvTranspose2dWeightNorm(nn.Module):
    """Weight-normalised transposed convolution, normalised per OUTPUT channel (dim 1 of the weight); parameters live in
    ``self.deconv`` (weight_norm.py:89-128)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, groups=1, bias=True,
                 dilation=1):
        super().__init__()
        k = nn.modules.utils._pair(kernel_size)
        self.deconv = _NormedWeight((in_channels, out_channels // groups, *k), 1, out_channels, bias)
        self.geometry = dict(stride=stride, padding=padding, output_padding=output_padding, groups=groups, dilation=dilation)

    def _deconv(self, input):
        return F.conv_transpose2d(input, self.deconv.weight(), self.deconv.bias, **self.geometry)

    def forward(self, input):
        return self._deconv(input)

    def initialize(self, x, init_scale=1.0):
        with torch.no_grad():
            mean, std = _channel_stats(self(x))
            self.deconv.rescale(mean, std, init_scale)
            return self(x)


def _run(layers, x, init: bool, init_scale: float):
    """Feed x through ``layers``; with ``init`` every layer that has a data-dependent ``initialize`` runs it."""
    for layer in layers:
        x = layer.initialize(x, init_scale=init_scale) if init and hasattr(layer, 'initialize') else layer(x)
    return x


# ------------------------------------------------------------------------------------------- residual stacks
class ResNetBlock(nn.Module):
    """ELU(conv2(ELU(conv1(x))) + shortcut(x)), 3x3 weight-normalised convolutions (resnet.py:23-61)."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = Conv2dWeightNorm(inplanes, planes, 3, stride=stride, padding=1)
        self.activation = nn.ELU()
        self.conv2 = Conv2dWeightNorm(planes, planes, 3, padding=1)
        self.downsample = (Conv2dWeightNorm(inplanes, planes, 1, stride=stride)
                           if (stride != 1 or inplanes != planes) else None)
        self.stride = stride

    def _pass(self, x, init, init_scale):
        shortcut = x if self.downsample is None else _run([self.downsample], x, init, init_scale)
        h = self.activation(_run([self.conv1], x, init, init_scale))
        h = _run([self.conv2], h, init, init_scale)
        return self.activation(h + shortcut)

    def forward(self, x):
        return self._pass(x, False, 1.0)

    def initialize(self, x, init_scale=1.0):
        return self._pass(x, True, init_scale)


class DeResNetBlock(nn.Module):
    """Pre-activated transposed block: a = ELU(x); deconv2(ELU(deconv1(a))) + shortcut(a)  (resnet.py:64-107)."""

    def __init__(self, inplanes, planes, stride=1, output_padding=0):
        super().__init__()
        self.deconv1 = ConvTranspose2dWeightNorm(inplanes, planes, 3, stride=stride, padding=1, output_padding=output_padding)
        self.activation = nn.ELU()
        self.deconv2 = ConvTranspose2dWeightNorm(planes, planes, 3, padding=1)
        self.downsample = (ConvTranspose2dWeightNorm(inplanes, planes, 1, stride=stride, output_padding=output_padding)
                           if (stride != 1 or inplanes != planes) else None)
        self.stride = stride

    def _pass(self, x, init, init_scale):
        a = self.activation(x)
        shortcut = a if self.downsample is None else _run([self.downsample], a, init, init_scale)
        h = self.activation(_run([self.deconv1], a, init, init_scale))
        return _run([self.deconv2], h, init, init_scale) + shortcut

    def forward(self, x):
        return self._pass(x, False, 1.0)

    def initialize(self, x, init_scale=1.0):
        return self._pass(x, True, init_scale)


class ResNet(nn.Module):
    def __init__(self, inplanes, planes: List[int], strides: List[int]):
        super().__init__()
        if len(planes) != len(strides):
            raise ValueError('planes and strides must have the same length')
        blocks, cin = [], inplanes
        for cout, stride in zip(planes, strides):
            blocks.append(ResNetBlock(cin, cout, stride=stride))
            cin = cout
        self.main = nn.Sequential(*blocks)

    def forward(self, x):
        return self.main(x)

    def initialize(self, x, init_scale=1.0):
        return _run(self.main, x, True, init_scale)


class DeResNet(nn.Module):
    def __init__(self, inplanes, planes: List[int], strides: List[int], output_paddings: List[int]):
        super().__init__()
        if not (len(planes) == len(strides) == len(output_paddings)):
            raise ValueError('planes, strides and output_paddings must have the same length')
        blocks, cin = [], inplanes
        for cout, stride, op in zip(planes, strides, output_paddings):
            blocks.append(DeResNetBlock(cin, cout, stride=stride, output_padding=op))
            cin = cout
        self.main = nn.Sequential(*blocks)

    def forward(self, x):
        return self.main(x)

    def initialize(self, x, init_scale=1.0):
        return _run(self.main, x, True, init_scale)


# ----------------------------------------------------------------------------------------------- masked layers
class _MaskedWeight(nn.Module):
    """Common part of the masked layers (masked.py:15-212): ``w = (v*mask) * exp(g) / (||v*mask|| + 1e-8)`` with a
    LOG-scale gain ``weight_g`` and a constant 0/1 ``mask`` buffer."""

    def _setup(self, shape, mask: torch.Tensor, n_out: int, bias: bool):
        self.register_buffer('mask', mask.to(torch.float32))
        v = torch.empty(*shape).normal_(mean=0.0, std=_INIT_STD) * self.mask
        g_shape = (n_out,) + (1,) * (len(shape) - 1)
        self.weight_v = nn.Parameter(v)
        self.weight_g = nn.Parameter((_rownorm(v, 0) + 1e-8).log().view(g_shape))
        if bias:
            self.bias = nn.Parameter(torch.zeros(n_out))
        else:
            self.register_parameter('bias', None)

    def masked_weight(self) -> torch.Tensor:
        def compute():
            v = self.weight_v * self.mask
            return v * (self.weight_g.exp() / (_rownorm(v, 0) + 1e-8))
        return _memo(self, compute)

    def _rescale(self, mean, std, init_scale):
        gain = init_scale / (std + _STD_FLOOR)
        self.weight_g.add_(gain.log().view(self.weight_g.shape))
        if self.bias is not None:
            self.bias.sub_(mean).mul_(gain)


def made_degrees(n_units: int, total_units: int, kind: str) -> torch.Tensor:
    """Degree ("max unit") of each unit of a MADE layer (masked.py:44-60): inputs 1..n; hidden units spread evenly over
    1..total_units (ceil((i+1) * total/n)); outputs 0..n-1 capped at total_units."""
    idx = torch.arange(n_units, dtype=torch.float64)
    if kind == 'input':
        return (idx + 1).to(torch.int64)
    if kind == 'output':
        return torch.clamp(idx, max=float(total_units)).to(torch.int64)
    # float32 product as numpy's np.ceil((i + 1) * (float(total)/n)) evaluates it in double: do the same in double
    return torch.ceil((idx + 1) * (float(total_units) / n_units)).to(torch.int64)


class MaskedLinear(_MaskedWeight):
    """MADE layer: output unit i sees input unit j iff degree_out[i] >= degree_in[j]; order 'B' reverses both axes."""

    def __init__(self, in_features, out_features, mask_type, total_units, max_units=None, bias=True):
        super().__init__()
        layer_type, order = mask_type
        if layer_type not in ('input-hidden', 'hidden-hidden', 'hidden-output', 'input-output') or order not in ('A', 'B'):
            raise ValueError('unknown mask type %s' % (mask_type,))
        self.in_features, self.out_features = in_features, out_features
        self.layer_type, self.order = layer_type, order
        if layer_type.startswith('input'):
            deg_in = made_degrees(in_features, total_units, 'input')
        else:
            if max_units is None or len(max_units) != in_features:
                raise ValueError('hidden layers need the degrees of their %d inputs' % in_features)
            deg_in = torch.as_tensor(max_units, dtype=torch.int64)
        if layer_type.endswith('output'):
            if out_features <= total_units:
                raise ValueError('an output layer needs more than total_units units')
            deg_out = made_degrees(out_features, total_units, 'output')
        else:
            deg_out = made_degrees(out_features, total_units, 'hidden')
        self.max_units = deg_out.numpy()
        mask = (deg_out.view(-1, 1) >= deg_in.view(1, -1)).to(torch.float32)
        if order == 'B':
            mask = mask.flip(0, 1)
        self._setup((out_features, in_features), mask.contiguous(), out_features, bias)

    def forward(self, input):
        return F.linear(input, self.masked_weight(), self.bias)

    def initialize(self, x, init_scale=1.0):
        with torch.no_grad():
            out = self(x)
            std = out.std(dim=0)
            std = std + (std <= 0).to(std.dtype)      # constant units (fully masked rows) keep their scale
            self._rescale(out.mean(dim=0), std, init_scale)
            return self(x)

    def extra_repr(self):
        return 'in_features={}, out_features={}, bias={}, type={}, order={}'.format(
            self.in_features, self.out_features, self.bias is not None, self.layer_type, self.order)


class MaskedConv2d(_MaskedWeight):
    """PixelCNN convolution: of the first ``masked_channels`` input channels only the taps strictly before the centre in
    raster order are visible (type 'B': the centre too); the remaining input channels (conditioning) are unmasked;
    order 'B' flips the raster order (masked.py:113-212)."""

    def __init__(self, in_channels, out_channels, kernel_size, mask_type='A', order='A', masked_channels=None, stride=1,
                 padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        if mask_type not in ('A', 'B') or order not in ('A', 'B'):
            raise ValueError('mask_type and order must be "A" or "B"')
        if in_channels % groups or out_channels % groups:
            raise ValueError('in_channels and out_channels must be divisible by groups')
        pair = nn.modules.utils._pair
        self.in_channels, self.out_channels = in_channels, out_channels
        self.mask_type, self.order = mask_type, order
        self.masked_channels = in_channels if masked_channels is None else masked_channels
        self.kernel_size, self.stride, self.padding, self.dilation = pair(kernel_size), pair(stride), pair(padding), pair(dilation)
        self.groups = groups
        kh, kw = self.kernel_size
        shape = (out_channels, in_channels // groups, kh, kw)
        mask = torch.ones(shape)
        first_hidden_col = kw // 2 + (1 if mask_type == 'B' else 0)
        mask[:, :self.masked_channels, kh // 2, first_hidden_col:] = 0
        mask[:, :self.masked_channels, kh // 2 + 1:, :] = 0
        if order == 'B':
            mask = mask.flip(2, 3)
        self._setup(shape, mask.contiguous(), out_channels, bias)

    def forward(self, input):
        return F.conv2d(input, self.masked_weight(), self.bias, self.stride, self.padding, self.dilation, self.groups)

    def initialize(self, x, init_scale=1.0):
        with torch.no_grad():
            mean, std = _channel_stats(self(x))
            self._rescale(mean, std, init_scale)
            return self(x)


# ---------------------------------------------------------------------------------------------------- PixelCNN
class PixelCNNBlock(nn.Module):
    """Bottleneck residual block: 1x1 -> masked kxk (type B) -> 1x1 at half width (pixelcnn.py:8-33)."""

    def __init__(self, in_channels, kernel_size):
        super().__init__()
        self.mask_type = 'B'
        mid = in_channels // 2
        self.main = nn.Sequential(
            Conv2dWeightNorm(in_channels, mid, 1),
            nn.ELU(),
            MaskedConv2d(mid, mid, kernel_size, mask_type='B', padding=kernel_size // 2),
            nn.ELU(),
            Conv2dWeightNorm(mid, in_channels, 1),
        )
        self.activation = nn.ELU()

    def forward(self, input):
        return self.activation(self.main(input) + input)

    def initialize(self, x, init_scale=1.0):
        return self.activation(_run(self.main, x, True, init_scale) + x)


class MaskABlock(nn.Module):
    """First layer: type-A masked convolution over [image | conditioning] channels (pixelcnn.py:36-53)."""

    def __init__(self, in_channels, out_channels, kernel_size, masked_channels):
        super().__init__()
        self.mask_type = 'A'
        self.main = nn.Sequential(
            MaskedConv2d(in_channels, out_channels, kernel_size, mask_type='A', masked_channels=masked_channels,
                         padding=kernel_size // 2),
            nn.ELU(),
        )

    def forward(self, input):
        return self.main(input)

    def initialize(self, x, init_scale=1.0):
        return _run(self.main, x, True, init_scale)


class PixelCNN(nn.Module):
    """Stack of masked blocks with long skip connections: the output of block i re-enters, through its own bottleneck
    block ``direct_connects[i]``, in front of block i+3 (and the third-last output after the last block)
    (pixelcnn.py:56-105)."""

    def __init__(self, in_channels, out_channels, num_blocks, kernel_sizes, masked_channels):
        super().__init__()
        if num_blocks != len(kernel_sizes):
            raise ValueError('one kernel size per block')
        blocks = [MaskABlock(in_channels, out_channels, kernel_sizes[0], masked_channels)]
        blocks += [PixelCNNBlock(out_channels, k) for k in kernel_sizes[1:]]
        self.main = nn.ModuleList(blocks)
        self.direct_connects = nn.ModuleList([PixelCNNBlock(out_channels, kernel_sizes[i]) for i in range(1, num_blocks - 1)])

    def _pass(self, x, init, init_scale):
        pending = []                                   # outputs waiting for their skip connection (lag 3)
        for i, block in enumerate(self.main):
            if i >= 3:
                x = x + _run([self.direct_connects[i - 3]], pending.pop(0), init, init_scale)
            x = _run([block], x, init, init_scale)
            pending.append(x)
        if len(pending) != 3:
            raise RuntimeError('PixelCNN needs at least 3 blocks')
        return x + _run([self.direct_connects[-1]], pending.pop(0), init, init_scale)

    def forward(self, input):
        return self._pass(input, False, 1.0)

    def initialize(self, x, init_scale=1.0):
        return self._pass(x, True, init_scale)


# -------------------------------------------------------------------------------------------------------- MADE
class MADE(nn.Module):
    """Masked autoencoder for distribution estimation (Germain et al. 2015) with a masked direct input->output
    connection; ReLU hidden units (auto_regressives/made.py:8-60).  Output i depends on inputs < i (order 'A')."""

    def __init__(self, input_size, num_hiddens, hidden_size, order, bias=True):
        super().__init__()
        if num_hiddens < 1:
            raise ValueError('MADE needs at least one hidden layer')
        self.input_size, self.num_hiddens, self.hidden_size = input_size, num_hiddens, hidden_size
        self.activation = nn.ReLU()
        total = input_size - 1
        self.input_layer = MaskedLinear(input_size, hidden_size, ('input-hidden', order), total, bias=bias)
        self.direct_connect = MaskedLinear(input_size, input_size, ('input-output', order), total, bias=bias)
        degrees = self.input_layer.max_units
        hidden = []
        for _ in range(1, num_hiddens):
            layer = MaskedLinear(hidden_size, hidden_size, ('hidden-hidden', order), total, max_units=degrees, bias=bias)
            degrees = layer.max_units
            hidden.append(layer)
        self.hidden_layers = nn.ModuleList(hidden)
        self.output_layer = MaskedLinear(hidden_size, input_size, ('hidden-output', order), total, max_units=degrees, bias=bias)

    def _pass(self, x, init, init_scale):
        h = self.activation(_run([self.input_layer], x, init, init_scale))
        for layer in self.hidden_layers:
            h = self.activation(_run([layer], h, init, init_scale))
        out = _run([self.output_layer], h, init, init_scale)
        return out + _run([self.direct_connect], x, init, init_scale)

    def forward(self, x):
        return self._pass(x, False, 1.0)

    def initialize(self, x, init_scale=1.0):
        return self._pass(x, True, init_scale)


class ReShape(nn.Module):
    def __init__(self, out_shape):
        super().__init__()
        self.out_shape = tuple(out_shape)

    def forward(self, input):
        return input.view(input.size(0), *self.out_shape)


# ---------------------------------------------------------------------------------------------------- DenseNet
class DenseNetBlock(nn.Module):
    """x -> cat[x, conv3x3(ELU(conv1x1(x)))] with a 4*growth bottleneck (dense.py:13-32)."""

    def __init__(self, inplanes, growth_rate):
        super().__init__()
        self.main = nn.Sequential(
            Conv2dWeightNorm(inplanes, 4 * growth_rate, 1),
            nn.ELU(),
            Conv2dWeightNorm(4 * growth_rate, growth_rate, 3, padding=1),
        )

    def forward(self, x):
        return torch.cat([x, self.main(x)], dim=1)

    def initialize(self, x, init_scale=1.0):
        return torch.cat([x, _run(self.main, x, True, init_scale)], dim=1)


class DenseNet(nn.Module):
    """``steps`` dense blocks, each followed by an ELU; channels grow by ``growth_rate`` per step (dense.py:35-54)."""

    def __init__(self, inplanes, growth_rate, steps):
        super().__init__()
        from collections import OrderedDict
        layers = OrderedDict()
        for step in range(steps):
            layers['block%d' % step] = DenseNetBlock(inplanes + step * growth_rate, growth_rate)
            layers['activation%d' % step] = nn.ELU()
        self.main = nn.Sequential(layers)

    def forward(self, x):
        return self.main(x)

    def initialize(self, x, init_scale=1.0):
        return _run(self.main, x, True, init_scale)


# ------------------------------------------------------------------------------------------ PixelCNN++ layers
class DownShiftConv2d(Conv2dWeightNorm):
    """Convolution that only sees rows at or above the output row: pad kh-1 rows on top, (kw-1)/2 columns on both sides
    (masked.py:215-227; Salimans et al. 2017)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=(1, 1), bias=True):
        kh, kw = kernel_size
        super().__init__(in_channels, out_channels, (kh, kw), stride=tuple(stride), bias=bias)
        self.shift_padding = ((kw - 1) // 2, (kw - 1) // 2, kh - 1, 0)     # left, right, top, bottom

    def forward(self, input):
        return self._conv(F.pad(input, self.shift_padding))


class DownRightShiftConv2d(DownShiftConv2d):
    """... and only columns at or left of the output column: pad kw-1 columns on the left (masked.py:230-237)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=(1, 1), bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, bias=bias)
        kh, kw = kernel_size
        self.shift_padding = (kw - 1, 0, kh - 1, 0)


class DownShiftConvTranspose2d(ConvTranspose2dWeightNorm):
    """Transposed counterpart: full transposed convolution, then drop the last kh-1 rows and (kw-1)/2 columns on both
    sides (masked.py:240-255)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=(1, 1), bias=True):
        kh, kw = kernel_size
        super().__init__(in_channels, out_channels, (kh, kw), stride=tuple(stride),
                         output_padding=(stride[0] - 1, stride[1] - 1), bias=bias)
        self.crop = (kh - 1, (kw - 1) // 2, (kw - 1) // 2)       # rows at the bottom, columns left, columns right

    def forward(self, input):
        out = self._deconv(input)
        bottom, left, right = self.crop
        return out[:, :, :out.size(2) - bottom, left:out.size(3) - right]


class DownRightShiftConvTranspose2d(DownShiftConvTranspose2d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=(1, 1), bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, bias=bias)
        kh, kw = kernel_size
        self.crop = (kh - 1, 0, kw - 1)


def concat_elu(x, dim=1):
    """ELU of [x, -x] (Shang et al. 2016 with ELU): doubles the channels."""
    return F.elu(torch.cat([x, -x], dim=dim))


def _call(layer, init, init_scale, *args, **kwargs):
    return layer.initialize(*args, init_scale=init_scale, **kwargs) if init else layer(*args, **kwargs)


class GatedResnetBlock(nn.Module):
    """One residual step of both PixelCNN++ streams (pixelcnnpp.py:21-103): the vertical stream x1 (down-shifted
    convolutions) and the horizontal stream x2 (down-right-shifted, fed by x1 through a 1x1 ``nin``); gated output
    ``a * sigmoid(b)``; optional conditioning h enters both gates through two 3x3 convolutions."""

    def __init__(self, in_channels, h_channels=0, dropout=0.0, activation=concat_elu):
        super().__init__()
        gain = 2 if activation is concat_elu else 1
        c = in_channels
        self.activation = activation
        self.down_conv1 = DownShiftConv2d(gain * c, c, (2, 3))
        self.down_conv2 = DownShiftConv2d(gain * c, c, (2, 3))
        self.down_conv3 = DownShiftConv2d(gain * c, 2 * c, (2, 3))
        self.down_right_conv1 = DownRightShiftConv2d(gain * c, c, (2, 2))
        self.down_right_conv2 = DownRightShiftConv2d(gain * c, c, (2, 2))
        self.nin = Conv2dWeightNorm(gain * c, c, (1, 1))
        self.down_right_conv3 = DownRightShiftConv2d(gain * c, 2 * c, (2, 2))
        self.h_conv1 = Conv2dWeightNorm(gain * h_channels, c, (3, 3), padding=1) if h_channels else None
        self.h_conv2 = Conv2dWeightNorm(gain * c, 2 * c, (3, 3), padding=1) if h_channels else None
        self.dropout = nn.Dropout(p=dropout)

    def _pass(self, x1, x2, h, init, s):
        act = self.activation
        hc = 0
        if h is not None:
            hc = _call(self.h_conv1, init, s, act(h))
            hc = _call(self.h_conv2, init, 0.1 * s, act(hc))

        def gate(t, skip):
            a, b = (t + hc).chunk(2, 1)
            return a * torch.sigmoid(b) + skip

        c1 = _call(self.down_conv1, init, s, act(x1))
        c1 = _call(self.down_conv2, init, s, act(c1))
        c1 = gate(_call(self.down_conv3, init, 0.1 * s, self.dropout(act(c1))), x1)
        aux = _call(self.nin, init, s, act(c1))
        c2 = _call(self.down_right_conv1, init, s, act(x2))
        c2 = _call(self.down_right_conv2, init, s, act(c2))
        c2 = gate(_call(self.down_right_conv3, init, 0.1 * s, self.dropout(act(c2 + aux))), x2)
        return c1, c2

    def forward(self, x1, x2, h=None):
        return self._pass(x1, x2, h, False, 1.0)

    def initialize(self, x1, x2, h=None, init_scale=1.0):
        return self._pass(x1, x2, h, True, init_scale)


def _shift_down(t):
    return F.pad(t, (0, 0, 1, 0))[:, :, :-1, :]


def _shift_right(t):
    return F.pad(t, (1, 0, 0, 0))[:, :, :, :-1]


class TopShiftBlock(nn.Module):
    """Entry of the two streams: shifted convolutions of the image so that neither stream sees the current pixel
    (pixelcnnpp.py:106-135)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.down_conv1 = DownShiftConv2d(in_channels, out_channels, (2, 3))
        self.down_conv2 = DownShiftConv2d(in_channels, out_channels, (1, 3))
        self.down_right_conv = DownRightShiftConv2d(in_channels, out_channels, (2, 1))

    def _pass(self, x, init, s):
        x1 = _shift_down(_call(self.down_conv1, init, s, x))
        x2 = _shift_down(_call(self.down_conv2, init, s, x)) + _shift_right(_call(self.down_right_conv, init, s, x))
        return x1, x2

    def forward(self, input):
        return self._pass(input, False, 1.0)

    def initialize(self, x, init_scale=1.0):
        return self._pass(x, True, init_scale)


class _StreamPair(nn.Module):
    """Applies one layer per stream (strided down-sampling or transposed up-sampling of both streams)."""
    first: str
    second: str

    def _pass(self, x1, x2, init, s):
        return _call(getattr(self, self.first), init, s, x1), _call(getattr(self, self.second), init, s, x2)

    def forward(self, x1, x2, h=None):
        return self._pass(x1, x2, False, 1.0)

    def initialize(self, x1, x2, h=None, init_scale=1.0):
        return self._pass(x1, x2, True, init_scale)


class DownSamplingBlock(_StreamPair):
    first, second = 'down_conv', 'down_right_conv'

    def __init__(self, num_filters):
        super().__init__()
        self.down_conv = DownShiftConv2d(num_filters, num_filters, (2, 3), stride=(2, 2))
        self.down_right_conv = DownRightShiftConv2d(num_filters, num_filters, (2, 2), stride=(2, 2))


class UpSamplingBlock(_StreamPair):
    first, second = 'down_deconv', 'down_right_deconv'

    def __init__(self, num_filters):
        super().__init__()
        self.down_deconv = DownShiftConvTranspose2d(num_filters, num_filters, (2, 3), stride=(2, 2))
        self.down_right_deconv = DownRightShiftConvTranspose2d(num_filters, num_filters, (2, 2), stride=(2, 2))


class NINBlock(nn.Module):
    """x + conv1x1(act(residual)): merges a stored up-pass activation into the down pass."""

    def __init__(self, num_filters, activation=concat_elu):
        super().__init__()
        self.activation = activation
        self.nin = Conv2dWeightNorm((2 if activation is concat_elu else 1) * num_filters, num_filters, (1, 1))

    def forward(self, x, residual):
        return x + self.nin(self.activation(residual))

    def initialize(self, x, residual, init_scale=1.0):
        return x + self.nin.initialize(self.activation(residual), init_scale=init_scale)


class Identity(nn.Module):
    def forward(self, input):
        return input

    def initialize(self, x, init_scale=1.0):
        return x


class DownSampling(nn.Module):
    """conditioning h: stride-2 convolution doubling the channels"""

    def __init__(self, num_filters):
        super().__init__()
        self.conv = Conv2dWeightNorm(num_filters, num_filters * 2, (3, 3), stride=(2, 2), padding=1)

    def forward(self, x):
        return self.conv(x)

    def initialize(self, x, init_scale=1.0):
        return self.conv.initialize(x, init_scale=init_scale)


class UpSampling(nn.Module):
    """conditioning h: stride-2 transposed convolution halving the channels"""

    def __init__(self, num_filters):
        super().__init__()
        self.deconv = ConvTranspose2dWeightNorm(num_filters, num_filters // 2, (3, 3), stride=(2, 2), padding=1, output_padding=1)

    def forward(self, x):
        return self.deconv(x)

    def initialize(self, x, init_scale=1.0):
        return self.deconv.initialize(x, init_scale=init_scale)


class PixelCNNPP(nn.Module):
    """PixelCNN++ trunk (Salimans et al. 2017; pixelcnnpp.py:205-338): an "up" pass of ``levels`` resolutions with
    ``num_resnets`` gated blocks each (down-sampling in between), then a mirrored "down" pass that pops the stored
    up-pass activations through 1x1 ``nins``; the conditioning map h is down/up-sampled alongside.  Returns the
    horizontal stream [B, out_channels, H, W]."""

    def __init__(self, levels, in_channels, out_channels, num_resnets, h_channels=0, dropout=0.0, activation='concat_elu'):
        super().__init__()
        if activation not in ('elu', 'concat_elu'):
            raise ValueError('unknown activation %s' % activation)
        act = F.elu if activation == 'elu' else concat_elu
        hc = h_channels
        up_layers, up_hs, nins1, nins2 = [], [], [], []
        for level in range(levels):
            if level == 0:
                up_layers.append(TopShiftBlock(in_channels, out_channels))
                up_hs.append(Identity())
            else:
                up_layers.append(DownSamplingBlock(out_channels))
                nins1.append(NINBlock(out_channels, act))
                nins2.append(NINBlock(out_channels, act))
                up_hs.append(DownSampling(hc) if hc else Identity())
                hc *= 2
            for _ in range(num_resnets):
                up_layers.append(GatedResnetBlock(out_channels, hc, dropout, act))
                nins1.append(NINBlock(out_channels, act))
                nins2.append(NINBlock(out_channels, act))
                up_hs.append(Identity())
        down_layers, down_hs = [], []
        for level in range(levels):
            for _ in range(num_resnets):
                down_layers.append(GatedResnetBlock(out_channels, hc, dropout, act))
                down_hs.append(Identity())
            if level < levels - 1:
                down_layers.append(UpSamplingBlock(out_channels))
                down_hs.append(UpSampling(hc) if hc else Identity())
                hc //= 2
        if not (len(nins1) == len(down_layers) and len(up_hs) == len(up_layers) and hc == h_channels):
            raise RuntimeError('PixelCNN++ layer bookkeeping is inconsistent')
        self.up_layers = nn.ModuleList(up_layers)
        self.down_layers = nn.ModuleList(down_layers)
        self.nins1 = nn.ModuleList(nins1)
        self.nins2 = nn.ModuleList(nins2)
        self.up_hs = nn.ModuleList(up_hs)
        self.down_hs = nn.ModuleList(down_hs)

    def _pass(self, x, h, init, s):
        stack = []
        x1 = x2 = None
        for i, (layer, h_layer) in enumerate(zip(self.up_layers, self.up_hs)):
            if i == 0:
                x1, x2 = _call(layer, init, s, x)
            else:
                x1, x2 = _call(layer, init, s, x1, x2, h=h)
                stack.append((x1, x2))
            h = _call(h_layer, init, s, h)
        for i, (layer, h_layer, nin1, nin2) in enumerate(zip(self.down_layers, self.down_hs, self.nins1, self.nins2)):
            u1, u2 = stack.pop()
            if i == 0:
                x1, x2 = u1, u2
            else:
                x1 = _call(nin1, init, s, x1, u1)
                x2 = _call(nin2, init, s, x2, u2)
            x1, x2 = _call(layer, init, s, x1, x2, h)
            h = _call(h_layer, init, s, h)
        if stack:
            raise RuntimeError('unbalanced up / down pass')
        return x2

    def forward(self, input, h=None):
        return self._pass(input, h, False, 1.0)

    def initialize(self, x, h=None, init_scale=1.0):
        return self._pass(x, h, True, init_scale)


# ------------------------------------------------------------------------------------------------------ MADE2d
class MADE2d(nn.Module):
    """Convolutional MADE over [C, H, W] in raster order (made.py:63-121): type-A masked convolution in, residual type-B
    masked hidden layers, type-B output layer, plus a type-A masked direct connection; ReLU units."""

    def __init__(self, in_channels, kernel_size, num_hiddens, hidden_channels, hidden_kernels, order, bias=True):
        super().__init__()
        if num_hiddens < 1 or num_hiddens != len(hidden_kernels):
            raise ValueError('one kernel per hidden layer')
        pair = nn.modules.utils._pair
        pad = lambda k: (pair(k)[0] // 2, pair(k)[1] // 2)    # noqa: E731
        self.activation = nn.ReLU()
        self.top_layer = MaskedConv2d(in_channels, hidden_channels, kernel_size, mask_type='A', order=order,
                                      padding=pad(kernel_size), bias=bias)
        self.direct_connect = MaskedConv2d(in_channels, in_channels, kernel_size, mask_type='A', order=order,
                                           padding=pad(kernel_size), bias=bias)
        self.hidden_layers = nn.ModuleList([
            MaskedConv2d(hidden_channels, hidden_channels, pair(k), mask_type='B', order=order, padding=pad(k), bias=bias)
            for k in hidden_kernels[:num_hiddens - 1]])
        k_out = hidden_kernels[-1]
        self.output_layer = MaskedConv2d(hidden_channels, in_channels, pair(k_out), mask_type='B', order=order,
                                         padding=pad(k_out), bias=bias)

    def _pass(self, x, init, s):
        h = self.activation(_call(self.top_layer, init, s, x))
        for layer in self.hidden_layers:
            h = self.activation(_call(layer, init, s, h) + h)
        return _call(self.output_layer, init, s, h) + _call(self.direct_connect, init, s, x)

    def forward(self, x):
        return self._pass(x, False, 1.0)

    def initialize(self, x, init_scale=1.0):
        return self._pass(x, True, init_scale)
