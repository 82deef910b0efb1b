"""The oracle (oracle/wefax_oracle.py) against the golden vectors made by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from conftest import golden_full_names, load_digests, load_golden_full
from oracle import wefax_oracle as O
from wefax_b200 import synth

FLOAT_TOL = 1e-9          # max|d| / max|ref| for the float64 stages


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", golden_full_names())
def test_oracle_matches_reference_full(name):
    g = load_golden_full(name)
    o = O.decode(g["pcm"], g["sample_rate_in"], g["lpm"])
    for key in ("audio_data", "demodulated_data"):
        err = np.abs(o[key] - g[key]).max() / np.abs(g[key]).max()
        assert err < FLOAT_TOL, (key, err)
    assert np.array_equal(o["digitalized_data"], g["digitalized_data"].astype(np.int64))
    if g["error"] is None:
        assert o["error"] is None
        assert list(o["phasing_signals"]) == [int(v) for v in g["phasing_signals"]]
        assert o["start_frame"] == g["start_frame"]
        assert np.array_equal(o["output_image"], g["output_image"])
    else:
        assert list(o["error"]) == list(g["error"])


@pytest.mark.parametrize("name", sorted(load_digests()["cases"]))
def test_oracle_matches_reference_digest(name):
    c = load_digests()["cases"][name]
    pcm = synth.synth_recording(**c["synth"])
    assert _sha(pcm) == c["pcm_sha256"], "synthetic generator drifted from the committed vectors"
    sr = c["synth"].get("sample_rate", 11025)
    o = O.decode(pcm, sr, c["lpm"])
    assert o["audio_data"].shape[0] == c["n_out"]
    idx = np.linspace(0, c["n_out"] - 1, 4096).astype(np.int64)
    for key, skey, mkey in (("audio_data", "audio_sample", "audio_absmax"),
                            ("demodulated_data", "demod_sample", "demod_absmax")):
        err = np.abs(o[key][idx] - np.asarray(c[skey])).max() / c[mkey]
        assert err < FLOAT_TOL, (key, err)
    assert _sha(o["digitalized_data"].astype(np.uint8)) == c["digitalized_sha256"]
    if c["error"] is None:
        assert o["error"] is None
        assert list(o["phasing_signals"]) == c["phasing_signals"]
        assert o["start_frame"] == c["start_frame"]
        assert list(o["output_image"].shape) == c["image_shape"]
        assert _sha(o["output_image"]) == c["image_sha256"]
    else:
        assert list(o["error"]) == list(c["error"])


def test_line_constants_table():
    # SURVEY.md §8 per-LPM table (values probed from the reference's expressions)
    table = {60: (11025, 55, 11, 121, 8820), 90: (7350, 36, 7, 79, 5880),
             100: (6615, 33, 6, 72, 5292), 120: (5512, 27, 5, 59, 4410),
             180: (3675, 18, 3, 39, 2940), 240: (2756, 13, 2, 28, 2205)}
    for lpm, (w, n1, n0, L, mind) in table.items():
        c = O.line_constants(lpm)
        assert (c["width"], c["n1"], c["n0"], c["template_len"], c["mindistance"]) == (w, n1, n0, L, mind)


def test_notch_coefficients_known_answer():
    b, a = O.notch_coefficients(2600, 1, 11025)
    assert np.allclose(b, [0.52227657, -0.09289188, 0.52227657], atol=5e-9)
    assert np.allclose(a, [1.0, -0.09289188, 0.04455315], atol=5e-9)


def test_scipy_pillow_crosscheck():
    """The restated stages against the installed third-party routines themselves."""
    import scipy.signal
    from PIL import Image
    rng = np.random.default_rng(0)
    x = rng.normal(size=5001) * 1000
    b, a = O.notch_coefficients(2600, 1, 11025)
    bs, as_ = scipy.signal.iirnotch(2600, 1, 11025)
    assert np.array_equal(b, bs) and np.array_equal(a, as_)
    assert np.allclose(O.filtfilt(b, a, x), scipy.signal.filtfilt(bs, as_, x), rtol=0, atol=1e-9)
    assert np.allclose(O.hilbert(x), scipy.signal.hilbert(x), rtol=0, atol=1e-9)
    assert np.array_equal(O.medfilt5(np.abs(x)), scipy.signal.medfilt(np.abs(x), 5))
    for num in (2500, 2501, 7000, 7001, 5001):
        assert np.allclose(O.resample(x, num), scipy.signal.resample(x, num), rtol=0, atol=1e-9)
    for q in (0.5, 99.5, 50.0):
        assert O.percentile_linear(np.abs(x), q) == np.percentile(np.abs(x), q)
    img = rng.integers(0, 256, size=(37, 53), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img, mode="L").resize((53, 4 * 37)))
    assert np.array_equal(O.resize_rows_x4(img), ref)
    for h in (1, 2, 3, 4, 5):
        img = rng.integers(0, 256, size=(h, 7), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img, mode="L").resize((7, 4 * h)))
        assert np.array_equal(O.resize_rows_x4(img), ref)


# ---- N2: start / stop tone test (oracle/tones_oracle.py vs the reference's DataPacket) ----------

def _load_tones():
    import json
    import os
    from conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "tones.json")) as fh:
        return json.load(fh)


def test_tone_find_peaks_matches_scipy():
    import scipy.signal
    from oracle import tones_oracle as T
    rng = np.random.default_rng(11)
    for trial in range(20):
        n = int(rng.integers(50, 3000))
        x = np.abs(rng.normal(size=n)) * (rng.random(n) < 0.3)
        x = np.round(x / (x.max() + 1e-4), 2)          # many ties and plateaus
        for distance in (1, 7, 250, 380):
            ref = scipy.signal.find_peaks(x, distance=distance, height=0.05, prominence=0.2)[0]
            assert np.array_equal(T.find_peaks(x, distance, 0.05, 0.2), ref), (trial, distance)


@pytest.mark.parametrize("name", sorted(_load_tones()["synthetic"]))
def test_tone_oracle_matches_reference_synthetic(name):
    from oracle import tones_oracle as T
    c = _load_tones()["synthetic"][name]
    pcm = synth.synth_recording(**c["synth"])
    assert _sha(pcm) == c["pcm_sha256"]
    start, stop = T.scan(pcm, 11025)
    assert start.tolist() == c["start"]
    assert stop.tolist() == c["stop"]


def test_tone_oracle_known_answers_of_the_survey():
    """peaks of the start tone sit 300 Hz apart around 1900 Hz, of the stop tone
    450 Hz apart; a synthetic IOC576 start tone reproduces the 300 Hz spacing."""
    from oracle import tones_oracle as T
    pcm = synth.synth_recording(30.0, seed=3)
    freq, amp = T.fourier_transform(pcm[:11025], 11025)
    pk = T.find_peaks(amp, 250, 0.05, 0.2)
    assert 4 <= len(pk) <= 6
    assert np.all(np.abs(np.diff(freq[pk]) - 600.0) < 2.0) or np.all(np.abs(np.diff(freq[pk]) - 300.0) < 2.0)


@pytest.mark.parametrize("name", sorted(_load_tones()["fixtures"]))
def test_tone_oracle_matches_reference_fixtures(name):
    """The five shipped 1-s packets (their PCM travels inside the full_fixture_*.npz vectors)."""
    from oracle import tones_oracle as T
    c = _load_tones()["fixtures"][name]
    g = load_golden_full("fixture_" + name[:-len(".wav")])
    assert g["sample_rate_in"] == c["sample_rate"]
    assert T.contain_start_tone(g["pcm"], c["sample_rate"]) == c["start"]
    assert T.contain_stop_tone(g["pcm"], c["sample_rate"]) == c["stop"]


# ---- N2, second half: the per-packet sync pulse (oracle/tones_oracle.py vs DataPacket.find_sync_pulse) ----------

def _check_sync(info, samples, want):
    import hashlib
    assert info["pulse_found"] == want["pulse_found"]
    assert info["frequency_peak_found"] == want["frequency_peak_found"]
    assert info["samples_peak_found"] == want["samples_peak_found"]
    assert info["n_fft_peaks"] == want["n_fft_peaks"]
    assert info["peaks_samples"] == want["peaks_samples"]
    assert hashlib.sha256(np.asarray(samples).astype(np.uint8).tobytes()).hexdigest() == want["samples_sha256"]


@pytest.mark.parametrize("name", sorted(_load_tones()["fixtures"]))
def test_sync_pulse_oracle_matches_reference_fixtures(name):
    """find_sync_pulse() and the packet's processed samples (data_packet.py:301-343, 408-465) on the five shipped
    packets, 11025 Hz and 48 kHz."""
    from oracle import tones_oracle as T
    c = _load_tones()["fixtures"][name]
    g = load_golden_full("fixture_" + name[:-len(".wav")])
    sr = c["sample_rate"]
    _check_sync(T.find_sync_pulse(g["pcm"], sr), T.process_samples(g["pcm"], sr), c["sync_pulse"])


@pytest.mark.parametrize("name", sorted(_load_tones()["synthetic"]))
def test_sync_pulse_oracle_matches_reference_synthetic(name):
    from oracle import tones_oracle as T
    c = _load_tones()["synthetic"][name]
    pcm = synth.synth_recording(**c["synth"])
    for k, want in enumerate(c["sync_pulse"]):
        seg = pcm[k * 11025:(k + 1) * 11025]
        _check_sync(T.find_sync_pulse(seg, 11025), T.process_samples(seg, 11025), want)


def test_state_machine_gates_the_picture_on_the_sync_pulse():
    """wefax_live.py:175-200 on a synthetic transmission: the start tone is found after 4 s, the picture starts at
    the last pulse of the first packet whose pulse is found, the stop tone ends it."""
    from oracle import tones_oracle as T
    pcm = synth.synth_recording(30.0, lpm=120, ioc=576, seed=3)
    (start_packet, image_start, stop_packet), = T.state_machine(pcm, 11025)
    assert start_packet == 3 and stop_packet == 21
    first = next(k for k in range(start_packet, 30) if T.find_sync_pulse(pcm[k * 11025:(k + 1) * 11025], 11025)["pulse_found"])
    assert image_start == first * 11025 + T.find_sync_pulse(pcm[first * 11025:(first + 1) * 11025], 11025)["peaks_samples"][-1]
