"""CPU tests of the FM-extension oracle (oracle/fm_oracle.py) and its host logic.
No reference parity exists for this mode; the oracle is pinned to scipy for its filter / Hilbert pieces
and to the ground truth of the synthetic generator for the whole definition."""
import numpy as np

from oracle import fm_oracle as F
from wefax_b200 import synth

SR = 11025


def test_fir_filtfilt_is_scipy_filtfilt_with_padlen_9():
    import scipy.signal
    rng = np.random.default_rng(2)
    x = rng.normal(size=4000) * 1000
    h = F.bandpass_taps(1200, 2600, SR, 63)
    ref = scipy.signal.filtfilt(h, [1.0], x, padlen=9)
    assert np.allclose(F.fir_filtfilt(h, x), ref, rtol=0, atol=1e-9 * np.abs(ref).max())
    assert np.allclose(F.hilbert_imag(x), scipy.signal.hilbert(x).imag, rtol=0, atol=1e-9 * np.abs(x).max())
    assert np.allclose(F.hilbert_imag(x[:-1]), scipy.signal.hilbert(x[:-1]).imag, rtol=0, atol=1e-9 * np.abs(x).max())


def test_bandpass_taps_pass_the_carrier_and_stop_the_rest():
    h = F.bandpass_taps(1200, 2600, SR, 63)
    assert h.shape == (63,) and np.allclose(h, h[::-1])
    f = np.fft.rfftfreq(8192, 1 / SR)
    H2 = np.abs(np.fft.rfft(h, 8192)) ** 2            # applied forwards and backwards
    band = (f > 1500) & (f < 2300)
    assert H2[band].min() > 0.9 and H2[band].max() < 1.1
    assert H2[(f < 700) | (f > 3100)].max() < 1e-4


def test_fm_oracle_recovers_the_transmitted_picture():
    pcm, truth = synth.synth_recording(60.0, seed=7, block=64, return_grey=True)
    d = F.decode(pcm, 120, 576, search_from=5 * SR, fold_lines=20, image_end=45 * SR)
    assert abs(d["line_start"] - 5 * SR) <= 2
    timg = F.image(truth, 5 * SR, 120, 576, 45 * SR)
    rows = min(timg.shape[0], d["image"].shape[0])
    err = np.abs(d["image"][:rows].astype(int) - timg[:rows].astype(int))
    assert err.mean() < 4.0 and (err <= 16).mean() > 0.95
    # a frequency offset shifts every grey level by offset / 800 Hz: the mean error shows it
    pcm_off = synth.synth_recording(60.0, seed=7, block=64, carrier_offset_hz=40.0)
    d2 = F.decode(pcm_off, 120, 576, search_from=5 * SR, fold_lines=20, image_end=45 * SR)
    rows = min(timg.shape[0], d2["image"].shape[0])
    shift = (d2["image"][30:rows].astype(int) - timg[30:rows].astype(int)).mean()
    assert 8 < shift < 16            # 40 / 800 * 255 = 12.75 levels, less what clips at white


def test_image_span_from_tone_flags():
    from wefax_b200.fm import image_span
    start = np.zeros(60, dtype=bool)
    stop = np.zeros(60, dtype=bool)
    start[:5] = True
    start[45:50] = True        # the stop tone also passes the start-tone test
    stop[45:50] = True
    assert image_span(start, stop) == (5 * SR, 45 * SR)
    assert image_span(np.zeros(10, dtype=bool), np.zeros(10, dtype=bool)) == (0, 10 * SR)
