"""Multi-threaded 8-bit greyscale PNG writer for the large rasters the decoder
produces (a 60-min recording is a 5512 x 28800 image, 159 MB).

The reference saves with ``PIL.Image.save`` (wefax.py:408), a single zlib stream
compressed on one core.  A PNG's IDAT data is one zlib stream, but that stream
may be assembled from independently compressed pieces (the pigz construction):
every band of rows is deflated raw (``wbits=-15``) on its own thread — zlib
releases the GIL — and ended with a full flush so it is byte aligned; the pieces
are concatenated behind a zlib header and closed with the Adler-32 of the whole
filtered image.  Any PNG reader decodes the result to exactly the same pixels.
"""
from __future__ import annotations

import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_PNG_MAGIC = b"\x89PNG\r\n\x1a\n"
_ADLER_MOD = 65521


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def _adler32_combine(a1: int, a2: int, len2: int) -> int:
    """Adler-32 of the concatenation of two blocks from their Adler-32s (zlib's adler32_combine)."""
    rem = len2 % _ADLER_MOD
    s1 = a1 & 0xFFFF
    s2 = (rem * s1) % _ADLER_MOD
    s1 += (a2 & 0xFFFF) + _ADLER_MOD - 1
    s2 += ((a1 >> 16) & 0xFFFF) + ((a2 >> 16) & 0xFFFF) + _ADLER_MOD - rem
    if s1 >= _ADLER_MOD:
        s1 -= _ADLER_MOD
    if s1 >= _ADLER_MOD:
        s1 -= _ADLER_MOD
    if s2 >= (_ADLER_MOD << 1):
        s2 -= (_ADLER_MOD << 1)
    if s2 >= _ADLER_MOD:
        s2 -= _ADLER_MOD
    return (s2 << 16) | s1


def _deflate_band(args):
    rows, level, last = args
    # PNG scanlines: one filter-type byte (0 = None) in front of every row
    h, w = rows.shape
    raw = np.empty((h, w + 1), dtype=np.uint8)
    raw[:, 0] = 0
    raw[:, 1:] = rows
    buf = raw.tobytes()
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = comp.compress(buf)
    body += comp.flush(zlib.Z_FINISH if last else zlib.Z_FULL_FLUSH)
    return body, zlib.adler32(buf) & 0xFFFFFFFF, len(buf)


def write_png_gray8(path: str, image: np.ndarray, threads: int | None = None, level: int = 6,
                    band_bytes: int = 4 << 20) -> None:
    """Write a 2-D uint8 array as an 8-bit greyscale PNG, compressing row bands in parallel."""
    img = np.ascontiguousarray(image, dtype=np.uint8)
    if img.ndim != 2:
        raise ValueError("expected a 2-D uint8 array")
    h, w = img.shape
    if h == 0 or w == 0:
        raise ValueError("cannot write an empty image")
    rows_per_band = max(1, band_bytes // (w + 1))
    starts = list(range(0, h, rows_per_band))
    jobs = [(img[s: s + rows_per_band], level, i == len(starts) - 1) for i, s in enumerate(starts)]
    threads = threads or min(len(jobs), os.cpu_count() or 1)
    if threads > 1 and len(jobs) > 1:
        with ThreadPoolExecutor(threads) as pool:
            parts = list(pool.map(_deflate_band, jobs))
    else:
        parts = [_deflate_band(j) for j in jobs]
    adler = 1
    for _, a, n in parts:
        adler = _adler32_combine(adler, a, n)
    with open(path, "wb") as fh:
        fh.write(_PNG_MAGIC)
        fh.write(_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)))
        # one zlib stream (header, the raw deflate pieces, Adler-32), split over IDAT chunks
        fh.write(_chunk(b"IDAT", b"\x78\x9c" + parts[0][0]))
        for body, _, _ in parts[1:]:
            fh.write(_chunk(b"IDAT", body))
        fh.write(_chunk(b"IDAT", struct.pack(">I", adler)))
        fh.write(_chunk(b"IEND", b""))
