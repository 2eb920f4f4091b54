"""Tiny stand-ins for the conv stacks, built on mae_b200.modules' interfaces with the same parameter names as
the stubs in tests/golden/make_golden.py (which sit on the REFERENCE's interfaces), so golden state dicts load."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from mae_b200.modules import BinaryImageDecoder, ColorImageDecoder, EncoderCore, Flow


class LinearCore(EncoderCore):
    def __init__(self, nin, nz, clamp):
        super().__init__()
        self.nz, self.clamp = nz, clamp
        self.linear = nn.Linear(nin, 2 * nz)

    def forward(self, x):
        mu, logvar = self.linear(x.view(x.size(0), -1)).chunk(2, 1)   # mu is a strided view, like the reference cores
        if self.clamp:
            logvar = F.hardtanh(logvar, min_val=-7., max_val=7.)
        return mu, logvar

    def output_size(self):
        return self.nz,

    @classmethod
    def from_params(cls, params):
        return cls(**params)


class AffineFlow(Flow):
    def __init__(self, nz):
        super().__init__()
        self.nz = nz
        self.a = nn.Parameter(torch.zeros(nz))
        self.b = nn.Parameter(torch.zeros(nz))

    def input_size(self):
        return self.nz,

    def forward(self, x, init=False, init_scale=1.0):
        return (x - self.b) * torch.exp(-self.a), -self.a.sum().expand(x.size(0))

    def backward(self, y, init=False, init_scale=1.0):
        return y * torch.exp(self.a) + self.b, self.a.sum().expand(y.size(0))

    @classmethod
    def from_params(cls, params):
        return cls(**params)


class BinaryLinearDecoder(BinaryImageDecoder):
    def __init__(self, nz, npix):
        super().__init__(nz)
        self.linear = nn.Linear(nz, npix)

    def z_shape(self):
        return self.nz,

    def forward(self, z):
        return torch.sigmoid(self.linear(z))

    @classmethod
    def from_params(cls, params):
        return cls(**params)


class ColorLinearDecoder(ColorImageDecoder):
    def __init__(self, nz, nmix, h, w):
        super().__init__(nmix)
        self.nz, self.h, self.w = nz, h, w
        self.linear = nn.Linear(nz, 10 * nmix * h * w)

    def z_shape(self):
        return self.nz,

    def forward(self, x, z):
        return self.linear(z).view(z.size(0), 10 * self.nmix, self.h, self.w)

    @classmethod
    def from_params(cls, params):
        return cls(**params)


LinearCore.register('test_linear_core')
AffineFlow.register('test_affine_flow')
BinaryLinearDecoder.register('test_binary_linear')
ColorLinearDecoder.register('test_color_linear')
