"""Minimal RIFF/WAVE reader for the decode path (the reference calls
``scipy.io.wavfile.read``, wefax.py:343,349).

Returns ``(sample_rate, data)`` with scipy's conventions: ``data`` is ``(n,)``
for mono and ``(n, channels)`` otherwise, in the stored integer type (uint8 for
8-bit, int16 for 16-bit PCM).  Only the formats the GPU ingest kernel consumes
are accepted; anything else raises ``ValueError`` instead of being converted
silently.
"""
from __future__ import annotations

import struct

import numpy as np

WAVE_FORMAT_PCM = 0x0001
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
    """``scipy.io.wavfile.read`` for 8/16-bit PCM."""
    h = read_header(path)
    if h["format_tag"] != WAVE_FORMAT_PCM or h["bits"] not in (8, 16):
        raise ValueError(f"unsupported WAV sample format (tag {h['format_tag']}, {h['bits']} bit): "
                         "the GPU ingest path takes 8/16-bit PCM")
    dtype = np.uint8 if h["bits"] == 8 else np.dtype("<i2")
    frame = h["channels"] * (h["bits"] // 8)
    n = h["data_bytes"] // frame
    data = np.fromfile(path, dtype=dtype, count=n * h["channels"], offset=h["data_offset"])
    if h["channels"] > 1:
        data = data.reshape(n, h["channels"])
    return h["sample_rate"], data
