"""wefax_b200 — B200-native implementation of the WEFAX file-decoding hot path.

``wefax_b200.wefax.Demodulator`` is the drop-in for the reference's
``wefax.Demodulator``; ``wefax_b200.decoder.Decoder`` is the batched API under it;
both call hand-written sm_100a kernels through ``libwefax_b200.so`` (C-ABI in
``include/wefax_b200.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

__all__ = ["Demodulator", "Decoder", "BatchResult"]


def __getattr__(name):
    if name == "Demodulator":
        from .wefax import Demodulator
        return Demodulator
    if name in ("Decoder", "BatchResult"):
        from . import decoder
        return getattr(decoder, name)
    raise AttributeError(name)
