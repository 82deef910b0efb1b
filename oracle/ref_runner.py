"""TEST / BENCH INFRASTRUCTURE — runs the UNMODIFIED reference decoder in-process.

Usable where ``/root/reference`` is mounted (the build container) or where
``oracle/make_ref.py`` staged its byte-for-byte copy under ``oracle/_ref/`` (git-ignored;
it travels to the GPU box).  It is used by ``tests/golden/make_golden.py`` to produce the
golden vectors that pin ``oracle/wefax_oracle.py``, by the container-only tests, and by
``bench.py`` to time the reference on the host cores (``--impl reference``, ``cpu_baseline``).
Nothing on the product path may import this.

Recipe: stub matplotlib (``wefax.py:6,9,10`` import it for debug
plots only), chdir to the reference root (``config.py:11`` opens
``config/config.json`` CWD-relative), patch ``wefax.time.sleep`` (the reference
sleeps 7.0 s per ``process()``: ``wefax.py:59,73,75,77,193,202``).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np

def _find_reference_root() -> str:
    """The read-only mount in the build container, else the unmodified copy oracle/make_ref.py staged under
    oracle/_ref/ (what travels to the GPU box; it holds the file decoder only, not data_packet.py)."""
    env = os.environ.get("WEFAX_REFERENCE_ROOT")
    if env:
        return env
    staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    for cand in ("/root/reference", staged):
        if os.path.isfile(os.path.join(cand, "wefax.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "wefax.py"))


def _stub_matplotlib() -> None:
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_wefax_stub", False):
        return  # a real matplotlib is importable; leave it alone
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker",
                 "matplotlib.animation", "matplotlib.cm"):
        mod = types.ModuleType(name)
        mod._wefax_stub = True
        sys.modules.setdefault(name, mod)
    sys.modules["matplotlib.ticker"].FormatStrFormatter = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].ticker = sys.modules["matplotlib.ticker"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]


@contextlib.contextmanager
def _in_reference_root():
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(cwd)


def import_reference():
    """Import the reference's ``wefax`` module (cached) with sleeps patched out."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    if "wefax" in sys.modules and getattr(sys.modules["wefax"], "_is_reference", False):
        return sys.modules["wefax"]
    try:
        import matplotlib  # noqa: F401
    except Exception:
        _stub_matplotlib()
    with _in_reference_root():
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            import wefax  # the reference module, not ours
        finally:
            sys.path.remove(REFERENCE_ROOT)
    wefax.time.sleep = lambda s: None
    wefax._is_reference = True
    return wefax


def run_reference(wav_path: str, lpm: int = 120) -> dict:
    """``Demodulator(wav, lpm).process()`` of the reference; returns every attribute.

    If ``process()`` raises (e.g. the ``ValueError`` from ``max([])`` at
    ``wefax.py:294``), the exception is returned under ``"error"`` with whatever
    attributes were set before it.
    """
    wefax = import_reference()
    wav_path = os.path.abspath(wav_path)
    out: dict = {"error": None}
    with _in_reference_root():
        d = wefax.Demodulator(wav_path, lines_per_minute=lpm, tcp_stream=True, quiet=True)
        try:
            d.process()
        except Exception as exc:  # the caller compares error type + message
            out["error"] = (type(exc).__name__, str(exc))
        finally:
            if sys.stdout is not sys.__stdout__:
                try:
                    sys.stdout.close()
                except Exception:
                    pass
            sys.stdout = sys.__stdout__
    for name in ("sample_rate", "length", "start_frame"):
        if hasattr(d, name):
            out[name] = getattr(d, name)
    if hasattr(d, "audio_data"):
        out["audio_data"] = np.asarray(d.audio_data, dtype=np.float64)
    if hasattr(d, "demodulated_data"):
        out["demodulated_data"] = np.asarray(d.demodulated_data, dtype=np.float64)
    if hasattr(d, "digitalized_data"):
        out["digitalized_data"] = np.asarray(d.digitalized_data, dtype=np.int64)
    if hasattr(d, "phasing_signals"):
        out["phasing_signals"] = [int(v) for v in d.phasing_signals]
    if hasattr(d, "output_image"):
        out["output_image"] = np.asarray(d.output_image)
    out["progress_titles"] = [m.get("progress_title", m.get("message_content"))
                              for m in d.websocket_stack]
    return out


def time_reference_process(wav_path: str, lpm: int = 120) -> tuple:
    """Wall-clock seconds of the reference's own ``Demodulator(wav, lpm).process()`` (sleeps patched out; the
    reference sleeps a fixed 7.0 s per call, wefax.py:59,73,75,77,193,202) and the number of samples it
    decoded.  An exception the reference itself raises (wefax.py:294) still counts: the work up to it ran."""
    import time
    wefax = import_reference()
    wav_path = os.path.abspath(wav_path)
    with _in_reference_root():
        d = wefax.Demodulator(wav_path, lines_per_minute=lpm, tcp_stream=False, quiet=True)
        t0 = time.perf_counter()
        try:
            d.process()
        except (ValueError, IndexError):
            pass
        finally:
            dt = time.perf_counter() - t0
            if sys.stdout is not sys.__stdout__:
                try:
                    sys.stdout.close()
                except Exception:
                    pass
            sys.stdout = sys.__stdout__
    n = len(d.digitalized_data) if hasattr(d, "digitalized_data") else 0
    return dt, n


def import_reference_data_packet():
    """The reference's ``data_packet`` module (live-path packet DSP, tone detection)."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    if "data_packet" in sys.modules and getattr(sys.modules["data_packet"], "_is_reference", False):
        return sys.modules["data_packet"]
    try:
        import matplotlib  # noqa: F401
    except Exception:
        _stub_matplotlib()
    with _in_reference_root():
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            import data_packet
        finally:
            sys.path.remove(REFERENCE_ROOT)
    data_packet._is_reference = True
    return data_packet


def run_reference_tones(samples: np.ndarray, sample_rate: int, lpm: int = 120) -> tuple:
    """``(contain_start_tone, contain_stop_tone)`` of the unmodified reference for one packet
    (data_packet.py:345-385)."""
    dp = import_reference_data_packet()
    with _in_reference_root():
        packet = dp.DataPacket(sample_rate, np.asarray(samples), lpm, "/tmp/", len(samples) / sample_rate, 0)
        return bool(packet.contain_start_tone()), bool(packet.contain_stop_tone())


def run_reference_sync_pulse(samples: np.ndarray, sample_rate: int, lpm: int = 120) -> dict:
    """``DataPacket.find_sync_pulse()`` of the unmodified reference for one packet (data_packet.py:301-343) and
    the packet's processed samples (data_packet.py:408-419)."""
    dp = import_reference_data_packet()
    with _in_reference_root():
        packet = dp.DataPacket(sample_rate, np.asarray(samples), lpm, "/tmp/", len(samples) / sample_rate, 0)
        info = packet.find_sync_pulse()
    return dict(pulse_found=bool(info["pulse_found"]), frequency_peak_found=bool(info["frequency_peak_found"]),
                samples_peak_found=bool(info["samples_peak_found"]),
                peaks_samples=[int(p) for p in info["peaks_samples"]],
                n_fft_peaks=int(len(info["peaks_fft"][0])),
                samples=np.asarray(packet.samples, dtype=np.int64))
