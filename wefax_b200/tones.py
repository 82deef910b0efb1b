"""Start / stop tone test and per-packet sync pulse of audio packets on the GPU.

Mirrors the reference's ``DataPacket.contain_start_tone`` / ``contain_stop_tone``
(``data_packet.py:345-385``), ``DataPacket.find_sync_pulse`` (``data_packet.py:301-343``) and the
live decoder's state machine (``wefax_live.py:175-200``): a start tone is *found* once consecutive
packets that contain it add up to 4 s, THEN the first packet whose sync pulse is found starts the
picture at its last pulse, THEN a stop tone is found like the start tone.  Spectra, scipy's
``find_peaks``, the packets' own notch / Hilbert / percentile stretch and the template search run on
the device (``wefax_tone_scan`` / ``wefax_sync_pulse_scan``); there is no CPU fallback.
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


def sync_pulse_settings(config: Config | None = None) -> N.SyncPulseSettings:
    """``config.json`` ``sync_pulse_settings`` + ``notch_filter_settings`` (data_packet.py:20-46); missing keys fall
    back to the reference's shipped values."""
    s, nf = {}, {}
    try:
        cfg = config if config is not None else Config()
        s = cfg.settings.get("sync_pulse_settings", {})
        nf = cfg.settings.get("notch_filter_settings", {})
    except FileNotFoundError:
        pass
    return N.SyncPulseSettings(float(s.get("peaks_minimum_height", 0.5)), float(s.get("peaks_minimum_prominence", 0.2)),
                               float(s.get("peaks_minimum_frequency", 1400)), float(s.get("peaks_maximum_frequency", 1600)),
                               float(nf.get("notch_filter_frequency", 2600)), float(nf.get("notch_filter_quality_factor", 1)))


def scan_sync_pulses(decoder, pcm, sample_rate: int, packet_seconds: float = 1.0,
                     settings: N.SyncPulseSettings | None = None, want_samples: bool = False) -> dict:
    """``DataPacket.find_sync_pulse()`` (data_packet.py:301-343) for every consecutive packet of ``pcm`` (same
    packets as :func:`scan_tones`).  Returns ``pulse_found``, ``frequency_peak_found`` (bool arrays),
    ``n_fft_peaks``, ``peaks_samples`` (list of lists) and, if asked for, the packets' grey levels ``samples``."""
    settings = settings or sync_pulse_settings()
    plen = int(sample_rate * packet_seconds)
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
    npk = n // plen if plen > 0 else 0
    found = np.zeros(npk, dtype=np.uint8)
    freq = np.zeros(npk, dtype=np.uint8)
    nfft = np.zeros(npk, dtype=np.int32)
    npul = np.zeros(npk, dtype=np.int32)
    last = np.full(npk, -1, dtype=np.int32)
    pulses = np.full((npk, N.MAX_PULSES), -1, dtype=np.int32)
    samples = np.zeros((npk, plen), dtype=np.uint8) if want_samples else None
    if npk:
        rc = decoder._lib.wefax_sync_pulse_scan(decoder._h, C.c_void_p(ptr), n, ch, int(sample_rate), plen, flags,
                                                C.byref(settings), found.ctypes.data, freq.ctypes.data, nfft.ctypes.data,
                                                npul.ctypes.data, last.ctypes.data, pulses.ctypes.data,
                                                samples.ctypes.data if want_samples else None)
        decoder._check(rc)
    return dict(pulse_found=found.astype(bool), frequency_peak_found=freq.astype(bool), n_fft_peaks=nfft,
                n_pulses=npul, last_pulse=last,
                peaks_samples=[[int(v) for v in pulses[k, : min(int(npul[k]), N.MAX_PULSES)]] for k in range(npk)],
                samples=samples)


def state_machine(start: np.ndarray, stop: np.ndarray, pulse_found: np.ndarray, last_pulse: np.ndarray, packet_frames: int,
                  packet_seconds: float = 1.0, hold_seconds: float = 4.0):
    """``wefax_live.py:175-200`` over a whole recording's packet results: start tone for ``hold_seconds`` of
    consecutive packets, THEN (from that same packet on) the first packet whose sync pulse is found starts the
    picture at ``packet * packet_frames + peaks_samples[-1]`` (wefax_live.py:191), THEN (from that same packet on)
    the stop tone for ``hold_seconds`` ends it.  The live decoder ends its session there; a file scan goes on to
    look for the next transmission.  Returns ``[(start_packet, image_start_sample, stop_packet), ...]`` with None
    where the recording ends first."""
    out = []
    start_found = phasing_found = False
    n_start = n_stop = 0
    cur = None
    for k in range(len(start)):
        if not start_found:
            n_start = n_start + 1 if start[k] else 0
            if n_start * packet_seconds >= hold_seconds:
                start_found = True
                cur = [k, None, None]
        if start_found and not phasing_found:
            if pulse_found[k]:
                phasing_found = True
                cur[1] = int(k * packet_frames + int(last_pulse[k]))
        if start_found and phasing_found:
            n_stop = n_stop + 1 if stop[k] else 0
            if n_stop * packet_seconds >= hold_seconds:
                cur[2] = k
                out.append(tuple(cur))
                start_found = phasing_found = False
                n_start = n_stop = 0
                cur = None
    if cur is not None:
        out.append(tuple(cur))
    return out


def scan_recording(decoder, pcm, sample_rate: int, packet_seconds: float = 1.0):
    """Tone scan + sync pulse scan + state machine of one recording: where the live decoder would start and stop
    each picture.  ``[(start_packet, image_start_sample, stop_packet), ...]``."""
    start, stop, _, _ = scan_tones(decoder, pcm, sample_rate, packet_seconds)
    sp = scan_sync_pulses(decoder, pcm, sample_rate, packet_seconds)
    return state_machine(start, stop, sp["pulse_found"], sp["last_pulse"], int(sample_rate * packet_seconds), packet_seconds)


def find_transmissions(start: np.ndarray, stop: np.ndarray, packet_seconds: float = 1.0, hold_seconds: float = 4.0):
    """The counting logic of ``wefax_live.py:175-200`` run over a whole recording's packet flags.

    A start tone is found at the first packet where ``hold_seconds`` worth of consecutive
    start-tone packets have been seen; after that, a stop tone is found likewise.  This is the
    tone-only view; :func:`state_machine` adds the phasing gate the live decoder has
    (wefax_live.py:187-192).  Returns a list of ``(start_found_packet, stop_found_packet)``
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
