"""Drop-in for ``mae.modules``: the loss / evaluation hot path on the CUDA kernels (same names, signatures and registries)
plus the PyTorch conv stacks and flows of the reference's binary-image and ResNet colour configurations."""
from .flow import Flow
from .encoder import Encoder, EncoderCore, GaussianEncoder
from .decoder import Decoder, BinaryImageDecoder, ColorImageDecoder
from . import networks
from .zoo import (AF2dMADE, AFMADE, DenseNetEncoderCoreColorImage32x32, DenseNetEncoderCoreColorImage64x64,
                  PixelCNNDecoderBinaryImage28x28, PixelCNNPPDecoderColorImage32x32, PixelCNNPPDecoderColorImage64x64,
                  ResnetDecoderBinaryImage28x28, ResnetDecoderColorImage32x32, ResnetEncoderCoreBinaryImage28x28,
                  ResnetEncoderCoreColorImage32x32)
from .mae import MAE
from .utils import exponentialMovingAverage, logavgexp, logsumexp, norm, sample_from_discretized_mix_logistic

__all__ = ["MAE", "Encoder", "EncoderCore", "GaussianEncoder", "Flow", "Decoder", "BinaryImageDecoder",
           "ColorImageDecoder", "exponentialMovingAverage", "logsumexp", "logavgexp", "norm",
           "sample_from_discretized_mix_logistic", "networks", "AFMADE",
           "PixelCNNDecoderBinaryImage28x28", "ResnetDecoderBinaryImage28x28", "ResnetDecoderColorImage32x32",
           "ResnetEncoderCoreBinaryImage28x28", "ResnetEncoderCoreColorImage32x32", "AF2dMADE",
           "DenseNetEncoderCoreColorImage32x32", "DenseNetEncoderCoreColorImage64x64", "PixelCNNPPDecoderColorImage32x32",
           "PixelCNNPPDecoderColorImage64x64"]
