"""Minimal RIFF/WAVE reader for the decode path (the reference calls
``scipy.io.wavfile.read``, wefax.py:343,349).

Returns ``(sample_rate, data)`` with scipy's conventions: ``data`` is ``(n,)``
for mono and ``(n, channels)`` otherwise, in the dtype scipy delivers (uint8,
int16, int32 — 24-bit PCM left-justified —, float32, float64).  16-bit PCM is
what the GPU ingest kernel eats directly; the drop-in class converts the other
formats to float32 with the reference's arithmetic (wefax.py:360-373) first.
"""
from __future__ import annotations

import struct

import numpy as np

WAVE_FORMAT_PCM = 0x0001
WAVE_FORMAT_IEEE_FLOAT = 0x0003
WAVE_FORMAT_EXTENSIBLE = 0xFFFE


def read_header(path: str) -> dict:
    """Parse the chunks up to ``data``; returns format fields + data offset/size."""
    with open(path, "rb") as fh:
        riff = fh.read(12)
        if len(riff) < 12 or riff[:4] not in (b"RIFF", b"RF64") or riff[8:12] != b"WAVE":
            raise ValueError(f"File format {riff[:4]!r} not understood. Only 'RIFF' and 'RIFX' supported.")
        fmt = None
        while True:
            head = fh.read(8)
            if len(head) < 8:
                raise ValueError("No data chunk!")
            cid, size = head[:4], struct.unpack("<I", head[4:])[0]
            if cid == b"fmt ":
                raw = fh.read(size + (size & 1))
                tag, ch, rate, _brate, align, bits = struct.unpack("<HHIIHH", raw[:16])
                if tag == WAVE_FORMAT_EXTENSIBLE and size >= 26:
                    tag = struct.unpack("<H", raw[24:26])[0]
                fmt = dict(format_tag=tag, channels=ch, sample_rate=rate, block_align=align, bits=bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError("No fmt chunk before data")
                offset = fh.tell()
                fh.seek(0, 2)
                avail = fh.tell() - offset
                fmt.update(data_offset=offset, data_bytes=min(size, avail))
                return fmt
            else:
                fh.seek(size + (size & 1), 1)


def read(path: str):
    """``scipy.io.wavfile.read``: ``(sample_rate, data)`` with scipy's dtypes — uint8 (8-bit PCM), int16, int32
    (32-bit PCM, and 24-bit PCM left-justified in int32 as scipy >= 1.6 does), float32 / float64 (IEEE float)."""
    h = read_header(path)
    tag, bits = h["format_tag"], h["bits"]
    frame = h["block_align"] or h["channels"] * ((bits + 7) // 8)
    n = h["data_bytes"] // frame
    count = n * h["channels"]
    if tag == WAVE_FORMAT_PCM and bits in (8, 16, 32):
        dtype = {8: np.uint8, 16: np.dtype("<i2"), 32: np.dtype("<i4")}[bits]
        data = np.fromfile(path, dtype=dtype, count=count, offset=h["data_offset"])
    elif tag == WAVE_FORMAT_PCM and bits == 24:
        raw = np.fromfile(path, dtype=np.uint8, count=count * 3, offset=h["data_offset"]).reshape(-1, 3)
        data = np.zeros((raw.shape[0], 4), dtype=np.uint8)
        data[:, 1:] = raw                                   # little endian: the 24 bits become the top of an int32
        data = data.view("<i4").reshape(-1)
    elif tag == WAVE_FORMAT_IEEE_FLOAT and bits in (32, 64):
        data = np.fromfile(path, dtype=np.dtype("<f4") if bits == 32 else np.dtype("<f8"), count=count,
                           offset=h["data_offset"])
    else:
        raise ValueError(f"Unsupported bit depth: the WAV file has {bits}-bit data of format tag {tag}.")
    if h["channels"] > 1:
        data = data.reshape(n, h["channels"])
    return h["sample_rate"], data
