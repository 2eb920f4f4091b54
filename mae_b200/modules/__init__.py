"""Drop-in for the hot-path part of ``mae.modules`` (same names, signatures and registries)."""
from .flow import Flow
from .encoder import Encoder, EncoderCore, GaussianEncoder
from .decoder import Decoder, BinaryImageDecoder, ColorImageDecoder
from .mae import MAE
from .utils import exponentialMovingAverage, logsumexp

__all__ = ["MAE", "Encoder", "EncoderCore", "GaussianEncoder", "Flow", "Decoder", "BinaryImageDecoder",
           "ColorImageDecoder", "exponentialMovingAverage", "logsumexp"]
