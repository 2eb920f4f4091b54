"""mae_b200 — B200-native (sm_100a) loss-and-evaluation hot path of XuezheMax/mae.

* ``mae_b200.ops``      autograd operators over the C ABI of ``libmae_b200.so`` (include/mae_b200.h)
* ``mae_b200.modules``  drop-in for ``mae.modules`` (MAE, GaussianEncoder, Binary/ColorImageDecoder, registries)
* ``mae_b200.synth``    deterministic synthetic encoder/decoder outputs for tests and benchmarks
"""
__version__ = "0.1.0"
