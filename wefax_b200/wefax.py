"""Drop-in for the reference's file decoder class (``wefax.py:18-408``).

Same call surface as ``wefax.Demodulator`` — constructor arguments, error
strings, ``update_lines_per_minute``, ``process``, ``file_info``,
``save_output_image``, the post-``process()`` attributes and the progress
messages pushed on ``websocket_stack`` — but ``process()`` runs the whole DSP
chain on a B200 through ``libwefax_b200.so``.  Differences a caller can see are
listed in INTEGRATION.md (no 7 s of ``time.sleep``, ``sys.stdout`` is not
redirected process-wide, arrays are float32/uint8 numpy arrays).

    from wefax_b200.wefax import Demodulator
    d = Demodulator("recording.wav", lines_per_minute=120, tcp_stream=False, quiet=True)
    d.process()
    d.save_output_image("out.png")
"""
from __future__ import annotations

import os
import sys
import threading

import numpy as np

from . import _native as N
from . import wavio
from .config import Config
from .decoder import Decoder

# One native context per (host thread, device): main.py runs several conversions at once, each on its own
# request thread (main.py:53,294-309), and a context serves one thread at a time.  No lock: concurrent
# process() calls run concurrently, each on its own stream.  The context of a thread dies with the thread.
_thread_state = threading.local()


def _thread_decoder(device: int) -> Decoder:
    pool = getattr(_thread_state, "decoders", None)
    if pool is None:
        pool = _thread_state.decoders = {}
    if device not in pool:
        pool[device] = Decoder(device)
    return pool[device]


class Demodulator:
    def __init__(self, filepath: str,
                 lines_per_minute: int = 120,
                 quiet: bool = False,
                 tcp_stream: bool = True,
                 device: int = 0):
        # wefax.py:24-28 (exact messages; tests/tests.py:71-73 compares the second)
        if not os.path.exists(filepath):
            raise Exception(f"INVALID FILE: file at path: {filepath} does not exist")
        if filepath.split('.')[-1] != 'wav':
            raise Exception("INVALID FILETYPE: only .wav files are supported at this moment")

        self.filepath = filepath
        self.filename = self.filepath.split('/')[-1]
        self.lines_per_minute = lines_per_minute
        self.time_for_one_frame = 1 / (self.lines_per_minute / 60)  # in s
        self.quiet = quiet
        self.stream = tcp_stream
        self.device = device
        self.websocket_stack = []
        if not self.quiet:
            print("#" * 10 + ' ' * 5 + str(self.filename).ljust(20) + ' ' * 5 + "#" * 10)

    def update_lines_per_minute(self, lpm):
        self.lines_per_minute = lpm
        self.time_for_one_frame = 1 / (lpm / 60)  # in s

    # ------------------------------------------------------------------ process
    def process(self):
        """wefax.py:46-93 on the GPU; results land in the same attributes."""
        settings = Config().settings["notch_filter_settings"]
        notch_freq = int(settings["notch_filter_frequency"])          # wefax.py:63
        notch_q = settings["notch_filter_quality_factor"]             # wefax.py:64

        sample_rate, pcm = wavio.read(self.filepath)                   # wefax.py:349
        stereo = pcm.ndim == 2
        if stereo:
            self._log("\033[0;33mWARNING: two channels audio detected. Program will try to merge audio to one "
                      "channel\033[0m")
            parts = pcm.shape[0]
            for q in range(0, parts, 1000):                             # wefax.py:364-370: every 1000th frame ...
                self._progress("merging channels", (q + 1) / parts * 100)
            if parts and (parts - 1) % 1000 != 0:                       # ... and the last one
                self._progress("merging channels", parts / parts * 100)
            pcm = self._stereo_merge(pcm)
        elif pcm.dtype == np.uint8:
            pcm = pcm.astype(np.int16)                                  # same numeric values as the stored uint8
        elif pcm.dtype != np.int16:
            pcm = pcm.astype(np.float32)                                # int32 / 24-bit / float WAVs (wefax.py:349)
        self.sample_rate = sample_rate
        self.length = pcm.shape[0] / sample_rate                        # wefax.py:357
        resample = sample_rate != N.TARGET_RATE
        if resample:
            self._log("\033[0;33mWARNING: audio sample rate is not 11025 samples per second. Program will try to "
                      "resample audio\033[0m")
            self._progress("resampling audio", 0)

        res = _thread_decoder(self.device).decode(
            pcm, sample_rate, lpm=self.lines_per_minute, notch_freq=notch_freq, notch_q=notch_q,
            want=("audio", "demodulated", "digitalized", "raster"))

        if resample:
            self._progress("resampling audio", 100)
            self.sample_rate = N.TARGET_RATE                            # wefax.py:392-393
            self.length = res.n_out / self.sample_rate
        self.audio_data = res.audio[0]                                  # wefax.py:72
        self._progress("demodulating signal", 0)
        self.demodulated_data = res.demodulated[0]                      # wefax.py:74
        self._progress("demodulating signal", 100)
        self._progress("digitalizing signal", 0)
        status = int(res.status[0])
        if status & N.REC_NAN:
            raise res.error(0)                                          # int(nan) in wefax.py:216
        self.low, self.high = (float(v) for v in res.low_high[0])
        self.digitalized_data = res.digitalized[0]                      # wefax.py:76
        self._progress("digitalizing signal", 99)
        self._progress("digitalizing signal", 100)

        # wefax.py:238-258: one message per opened peak, at the sample that opened it
        consts = N.line_constants(self.lines_per_minute)
        peaks = res.peaks[0]
        n = res.n_out
        for k in range(1, len(peaks)):
            opened_at = peaks[k - 1] + consts["mindistance"] + 1
            self._progress("finding sync pulse", (opened_at / n) * 100)
        if len(peaks) == N.MAX_PEAKS:
            self._progress("finding sync pulse", 100)
        self.peaks = peaks
        if status & N.REC_NO_GROUPS:
            raise res.error(0)                                          # max([]) in wefax.py:294
        self.phasing_signals = res.phasing_signals[0]
        self.start_frame = self.phasing_signals[-1] if self.phasing_signals else 0   # wefax.py:80
        if status & N.REC_NO_LINES:
            raise res.error(0)                                          # putpixel in wefax.py:304

        from PIL import Image
        img = res.image(0)
        h = img.shape[0] // 4
        for py in range(0, h, 50):                                      # wefax.py:307-313
            self._progress("converting signal to image", (py + 1) / h * 100)
        self._progress("converting signal to image", 100)
        self.output_image = Image.fromarray(img, mode="L") if img.size else Image.new("L", (consts["width"], 0))

        if self.stream:                                                 # wefax.py:87-90
            self.__send_websocket_packet({"data_type": "message", "message_content": "convert_end"})

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _stereo_merge(pcm: np.ndarray) -> np.ndarray:
        """First two channels merged as the reference does, ``np.divide(np.add(L, R), 2)`` in the stored dtype
        (wefax.py:372: the sum of two integer scalars WRAPS).  int16 goes to the GPU as interleaved L/R and is
        merged by the ingest kernel; uint8 as a wrapped sum the kernel halves; every other sample format is merged
        here with numpy's own arithmetic and handed over as mono float32."""
        if pcm.dtype == np.int16:
            return np.ascontiguousarray(pcm[:, :2])
        if pcm.dtype == np.uint8:
            s = np.add(pcm[:, 0], pcm[:, 1])                            # wraps at 256
            out = np.zeros((pcm.shape[0], 2), dtype=np.int16)
            out[:, 0] = s
            return out
        with np.errstate(over="ignore"):
            s = np.add(pcm[:, 0], pcm[:, 1])                            # int32 wraps; floats add in their own type
        return np.divide(s, 2).astype(np.float32)

    def _log(self, text: str) -> None:
        if not self.quiet:
            print(text, file=sys.__stdout__)

    def _progress(self, title: str, percentage) -> None:
        if self.stream:
            self.__send_websocket_packet({"data_type": "progress_bar", "progress_title": title,
                                          "percentage": percentage})

    def __send_websocket_packet(self, message: dict):
        self.websocket_stack.append(message)

    # ------------------------------------------------------------------ the rest of the surface
    def file_info(self):
        """wefax.py:342-346 (``channels`` is ``len(data.shape)``: 1 for mono, 2 otherwise)."""
        h = wavio.read_header(self.filepath)
        frames = h["data_bytes"] // (h["channels"] * (h["bits"] // 8))
        channels = 1 if h["channels"] == 1 else 2
        return {"filename": self.filename, "channels": channels, "sample_rate": h["sample_rate"],
                "length": frames / h["sample_rate"]}

    def show_output_image(self):
        from matplotlib import pyplot as plt   # debug helper, as in wefax.py:403-405
        plt.imshow(self.output_image, cmap='gray')
        plt.show()

    def save_output_image(self, filepath: str):
        """wefax.py:407-408.  Large greyscale PNGs go through the multi-threaded writer
        (same pixels, ~10x faster than one zlib stream on one core); anything else through Pillow."""
        img = self.output_image
        if filepath.lower().endswith(".png") and img.mode == "L" and img.width * img.height >= (1 << 20):
            from . import pngio
            pngio.write_png_gray8(filepath, np.asarray(img))
        else:
            img.save(filepath)


if __name__ == "__main__":
    filename = str(sys.argv[1])
    lpm = int(sys.argv[2])
    output = str(sys.argv[3])
    demodulator = Demodulator(filename, lines_per_minute=lpm, tcp_stream=False, quiet=False)
    for key, value in demodulator.file_info().items():
        print(key, ":", value)
    demodulator.process()
    demodulator.save_output_image(output)
