"""Start / stop tone test of audio packets on the GPU.

Mirrors the reference's ``DataPacket.contain_start_tone`` / ``contain_stop_tone``
(``data_packet.py:345-385``) and the counting part of the live decoder's state machine
(``wefax_live.py:175-200``): a start (stop) tone is *found* once consecutive packets that
contain it add up to 4 s.  The spectra and scipy's ``find_peaks`` run on the device
(``csrc/tones.cu`` through ``wefax_tone_scan``); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .config import Config


def tone_settings(config: Config | None = None) -> N.ToneSettings:
    """``config.json`` ``tones_settings`` with the reference's key names (data_packet.py:24-33);
    missing keys fall back to the reference's shipped values."""
    s = {}
    if config is not None:
        s = config.settings.get("tones_settings", {})
    else:
        try:
            s = Config().settings.get("tones_settings", {})
        except FileNotFoundError:
            s = {}
    g = s.get
    return N.ToneSettings(float(g("start_tone_peaks_minimum_distance", 250)),
                          float(g("stop_tone_peaks_minimum_distance", 380)),
                          float(g("peaks_minimum_height", 0.05)), float(g("peaks_minimum_prominence", 0.2)),
                          float(g("peaks_minimum_frequency", 800)), float(g("peaks_maximum_frequency", 3200)),
                          int(g("peaks_minimum_amount", 4)), int(g("peaks_maximum_amount", 6)))


def scan_tones(decoder, pcm, sample_rate: int, packet_seconds: float = 1.0, settings: N.ToneSettings | None = None):
    """``(start, stop, n_start_peaks, n_stop_peaks)`` for every consecutive packet of ``pcm``.

    ``decoder`` is a :class:`wefax_b200.decoder.Decoder`; ``pcm`` is int16, ``(n,)`` mono or
    ``(n, 2)`` stereo, a host numpy array or a CUDA torch tensor.  Packet ``k`` covers frames
    ``[k*P, (k+1)*P)`` with ``P = int(sample_rate * packet_seconds)`` (the live decoder's
    ``AUDIO_PACKET_DURATION`` packets); a trailing partial packet is ignored.
    """
    settings = settings or tone_settings()
    plen = int(sample_rate * packet_seconds)
    if plen < 4:
        raise ValueError("packet too short")
    flags = 0
    if type(pcm).__module__.startswith("torch"):
        import torch
        if pcm.dtype != torch.int16 or not pcm.is_cuda:
            raise TypeError("device PCM must be a CUDA int16 tensor")
        pcm = pcm.contiguous()
        n, ch = int(pcm.shape[0]), (1 if pcm.dim() == 1 else int(pcm.shape[1]))
        ptr = pcm.data_ptr()
        flags |= N.F_PCM_ON_DEVICE
    else:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        n, ch = int(pcm.shape[0]), (1 if pcm.ndim == 1 else int(pcm.shape[1]))
        ptr = pcm.ctypes.data
    npk = n // plen
    start = np.zeros(npk, dtype=np.uint8)
    stop = np.zeros(npk, dtype=np.uint8)
    ns = np.zeros(npk, dtype=np.int32)
    nt = np.zeros(npk, dtype=np.int32)
    if npk:
        rc = decoder._lib.wefax_tone_scan(decoder._h, C.c_void_p(ptr), n, ch, int(sample_rate), plen, flags,
                                          C.byref(settings), start.ctypes.data, stop.ctypes.data,
                                          ns.ctypes.data, nt.ctypes.data)
        decoder._check(rc)
    return start.astype(bool), stop.astype(bool), ns, nt


def find_transmissions(start: np.ndarray, stop: np.ndarray, packet_seconds: float = 1.0, hold_seconds: float = 4.0):
    """The counting logic of ``wefax_live.py:175-200`` run over a whole recording's packet flags.

    A start tone is found at the first packet where ``hold_seconds`` worth of consecutive
    start-tone packets have been seen; after that, a stop tone is found likewise (the live
    decoder additionally waits for a phasing pulse before it counts stop packets; a file
    scan has no such gate).  Returns a list of ``(start_found_packet, stop_found_packet)``
    pairs; ``stop_found_packet`` is ``None`` when the recording ends first.
    """
    out = []
    run = 0
    k = 0
    n = len(start)
    while k < n:
        run = run + 1 if start[k] else 0
        if run * packet_seconds >= hold_seconds:
            begin = k
            run = 0
            end = None
            k += 1
            while k < n:
                run = run + 1 if stop[k] else 0
                if run * packet_seconds >= hold_seconds:
                    end = k
                    break
                k += 1
            out.append((begin, end))
            run = 0
        k += 1
    return out
