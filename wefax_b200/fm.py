"""EXTENSION: FM-discriminator decode with exact line length and IOC pixel columns.

The reference never implemented what its README describes (README.md:85-101: black 1500 Hz, white
2300 Hz, start tone 300 / 675 Hz, phasing lines, stop tone 450 Hz); it slope-detects and rasters
``int(samples per line)`` columns.  This module is the decoder the README describes, on the GPU
(``csrc/fm.cu`` through ``wefax_decode_fm``): band-pass, analytic signal, instantaneous frequency,
phasing-pulse line start, ``round(pi * IOC)`` pixels per line.  It is off the reference-exact path
and has no reference parity; see ``oracle/fm_oracle.py`` and ``tests/test_gpu_fm.py``.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _native as N


@dataclass
class FmResult:
    image: np.ndarray          # (rows, width) uint8, 0 = black (1500 Hz), 255 = white (2300 Hz)
    line_start: int            # sample (at 11025 Hz) of the first image row
    width: int                 # round(pi * ioc)
    search_from: int
    image_end: int
    grey: np.ndarray | None = None     # per-sample grey before clipping (float32), when asked for


def image_span(start_flags, stop_flags, sample_rate_out: int = N.TARGET_RATE, packet_seconds: float = 1.0):
    """Where the picture lies according to the tone scan: from the end of the first run of start-tone
    packets (>= 3 in a row) to the beginning of the first later run of stop-tone packets (>= 3)."""
    n = len(start_flags)
    k, run = 0, 0
    begin = 0
    while k < n:
        run = run + 1 if start_flags[k] else 0
        if run >= 3:
            while k + 1 < n and start_flags[k + 1]:
                k += 1
            begin = k + 1
            break
        k += 1
    end, run = n, 0
    for j in range(begin, n):
        run = run + 1 if stop_flags[j] else 0
        if run >= 3:
            end = j - 2
            break
    per = int(round(sample_rate_out * packet_seconds))
    return begin * per, end * per


def decode_fm(decoder, pcm, sample_rate: int, lpm: float = 120, ioc: int = 576, search_from: int | None = None,
              image_end: int | None = None, fold_lines: int = 20, black_hz: float = 1500.0, white_hz: float = 2300.0,
              band=(1200.0, 2600.0), fir_taps: int = 63, want_grey: bool = False) -> FmResult:
    """Decode one recording (int16 numpy, mono ``(n,)`` or stereo ``(n, 2)``).

    ``search_from`` / ``image_end`` (samples at 11025 Hz) default to what the start / stop tone scan
    finds (:func:`image_span`)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    n_in, ch = int(pcm.shape[0]), (1 if pcm.ndim == 1 else int(pcm.shape[1]))
    n_out = n_in if sample_rate == N.TARGET_RATE else N.resampled_length(n_in, sample_rate)
    if search_from is None or image_end is None:
        from .tones import scan_tones
        start, stop, _, _ = scan_tones(decoder, pcm, sample_rate)
        s0, s1 = image_span(start, stop)
        search_from = s0 if search_from is None else search_from
        image_end = min(s1, n_out) if image_end is None else image_end
    ls = 60.0 / lpm * N.TARGET_RATE
    width = int(round(math.pi * ioc))
    rows_max = max(0, int(math.floor((image_end - search_from) / ls)))
    img = np.zeros(max(1, rows_max * width), dtype=np.uint8)
    grey = np.zeros(n_out, dtype=np.float32) if want_grey else None
    rows, w, start_s = C.c_int32(0), C.c_int32(0), C.c_int64(0)
    desc = N.BatchDesc(1, n_in, ch, int(sample_rate), 0.0, 0.0, 0)
    prm = N.FmParams(float(lpm), int(ioc), float(black_hz), float(white_hz), float(band[0]), float(band[1]),
                     int(fir_taps), int(search_from), int(fold_lines), int(image_end))
    out = N.FmOut(grey.ctypes.data if want_grey else None, img.ctypes.data, img.size,
                  C.addressof(rows), C.addressof(w), C.addressof(start_s))
    decoder._check(decoder._lib.wefax_decode_fm(decoder._h, C.byref(desc), C.c_void_p(pcm.ctypes.data),
                                                C.byref(prm), C.byref(out)))
    return FmResult(img[: rows.value * width].reshape(rows.value, width), int(start_s.value), width,
                    int(search_from), int(image_end), grey)
