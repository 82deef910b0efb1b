"""TEST INFRASTRUCTURE — CPU restatement of the reference's start/stop-tone test.

SURVEY.md §8(f) row N2: tone detection exists only in the live path
(``/root/reference/data_packet.py:345-406``): a 1-s packet of raw samples, its
one-sided amplitude spectrum normalised by its maximum, ``scipy.signal.find_peaks``
with ``distance`` (250 bins for the start tone, 380 for the stop tone), ``height``
0.05 and ``prominence`` 0.2 (``config/config.json`` ``tones_settings``), and the
packet "contains the tone" iff it has 4..6 peaks, all between 800 and 3200 Hz.

``find_peaks`` lives in scipy (pinned 1.10.0 in ``requirements.txt``, not under
``/root/reference``): its published algorithm (local maxima with plateau midpoints,
height filter, priority-ordered distance suppression, prominence with unbounded
window) is restated here in numpy / plain Python.

PARITY PIN: ``tests/golden/tones.json`` holds the answers of the unmodified
reference's ``DataPacket.contain_start_tone / contain_stop_tone`` on the five shipped
fixtures and on seeded synthetic packets (``tests/golden/make_golden.py``), and
``tests/test_oracle_golden.py`` checks this file against them and against
``scipy.signal.find_peaks`` itself.
"""
from __future__ import annotations

import numpy as np

#: config/config.json "tones_settings" of the reference
DEFAULT_TONE_SETTINGS = dict(start_distance=250, stop_distance=380, height=0.05, prominence=0.2,
                             min_frequency=800, max_frequency=3200, min_amount=4, max_amount=6)


def fourier_transform(raw_samples: np.ndarray, sample_rate: int):
    """data_packet.py:387-406: one-sided amplitude spectrum, ``abs(fft[:N//2] / (N//2))`` divided by
    ``max + 0.0001``; bin k sits at ``k / (N / sample_rate)`` Hz."""
    fft = np.fft.fft(np.asarray(raw_samples))
    n = fft.shape[0]
    freq = np.arange(n) / (n / sample_rate)
    half = n // 2
    amp = np.abs(fft[:half] / half)
    return freq[:half], amp / (amp.max() + 0.0001)


def local_maxima_1d(x: np.ndarray):
    """scipy ``_local_maxima_1d``: strict rise before, strict fall after; plateaus give their midpoint."""
    peaks = []
    n = x.shape[0]
    i, i_max = 1, n - 1
    while i < i_max:
        if x[i - 1] < x[i]:
            ahead = i + 1
            while ahead < i_max and x[ahead] == x[i]:
                ahead += 1
            if x[ahead] < x[i]:
                peaks.append((i + ahead - 1) // 2)
                i = ahead
        i += 1
    return np.asarray(peaks, dtype=np.int64)


def select_by_peak_distance(peaks: np.ndarray, priority: np.ndarray, distance: float) -> np.ndarray:
    """scipy ``_select_by_peak_distance``: visit peaks from the highest priority down; a kept peak
    removes every not yet visited peak closer than ``ceil(distance)``."""
    n = peaks.shape[0]
    dist = int(np.ceil(distance))
    keep = np.ones(n, dtype=bool)
    order = np.argsort(priority)
    for i in range(n - 1, -1, -1):
        j = order[i]
        if not keep[j]:
            continue
        k = j - 1
        while k >= 0 and peaks[j] - peaks[k] < dist:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < n and peaks[k] - peaks[j] < dist:
            keep[k] = False
            k += 1
    return keep


def peak_prominences(x: np.ndarray, peaks: np.ndarray) -> np.ndarray:
    """scipy ``_peak_prominences`` with ``wlen=None``: walk left / right while the signal stays
    <= the peak, remember the minima; prominence = peak - max(left minimum, right minimum)."""
    out = np.empty(peaks.shape[0], dtype=np.float64)
    n = x.shape[0]
    for idx, p in enumerate(peaks):
        h = x[p]
        i, left_min = p, h
        while i >= 0 and x[i] <= h:
            if x[i] < left_min:
                left_min = x[i]
            i -= 1
        i, right_min = p, h
        while i <= n - 1 and x[i] <= h:
            if x[i] < right_min:
                right_min = x[i]
            i += 1
        out[idx] = h - max(left_min, right_min)
    return out


def find_peaks(x: np.ndarray, distance: float, height: float, prominence: float) -> np.ndarray:
    """``scipy.signal.find_peaks(x, distance=, height=, prominence=)[0]`` (the order of the filters
    is scipy's: height, distance, prominence)."""
    x = np.asarray(x, dtype=np.float64)
    peaks = local_maxima_1d(x)
    peaks = peaks[x[peaks] >= height]
    peaks = peaks[select_by_peak_distance(peaks, x[peaks], distance)]
    prom = peak_prominences(x, peaks)
    return peaks[prom >= prominence]


def contain_tone(raw_samples: np.ndarray, sample_rate: int, distance: float, settings: dict | None = None) -> bool:
    """data_packet.py:365-385."""
    s = dict(DEFAULT_TONE_SETTINGS, **(settings or {}))
    freq, amp = fourier_transform(raw_samples, sample_rate)
    peaks = find_peaks(amp, distance, s["height"], s["prominence"])
    in_range = all(s["min_frequency"] <= f <= s["max_frequency"] for f in freq[peaks])
    return bool(in_range and s["min_amount"] <= len(peaks) <= s["max_amount"])


def contain_start_tone(raw_samples, sample_rate, settings=None) -> bool:
    """data_packet.py:345-353."""
    s = dict(DEFAULT_TONE_SETTINGS, **(settings or {}))
    return contain_tone(raw_samples, sample_rate, s["start_distance"], s)


def contain_stop_tone(raw_samples, sample_rate, settings=None) -> bool:
    """data_packet.py:355-363."""
    s = dict(DEFAULT_TONE_SETTINGS, **(settings or {}))
    return contain_tone(raw_samples, sample_rate, s["stop_distance"], s)


# --------------------------------------------------------------------------- #
# per-packet sync pulse (phasing gate of the live decoder)                      #
# --------------------------------------------------------------------------- #
#: config/config.json "sync_pulse_settings" of the reference (peaks_minimum_distance is read but never used)
DEFAULT_SYNC_SETTINGS = dict(height=0.5, prominence=0.2, min_frequency=1400, max_frequency=1600)


def process_samples(raw_samples: np.ndarray, sample_rate: int, notch_freq=2600, notch_q=1) -> np.ndarray:
    """data_packet.py:408-465 ``__process_samples``: the packet's own notch (``iirnotch`` at the PACKET's sample
    rate + ``filtfilt``), ``abs(hilbert)``, ``medfilt(., 3)`` (zero padded), then the 0.5 / 99.5 percentile
    stretch ``rint(255 * (am - low) / (delta + 0.000001))`` clipped to 0..255, as int."""
    from oracle import wefax_oracle as O
    x = np.asarray(raw_samples, dtype=np.float64)
    b, a = O.notch_coefficients(notch_freq, notch_q, sample_rate)
    am = np.abs(O.hilbert(O.filtfilt(b, a, x)))
    padded = np.concatenate([[0.0], am, [0.0]])
    am = np.median(np.stack([padded[:-2], padded[1:-1], padded[2:]]), axis=0)
    low, high = O.percentile_linear(am, 0.5), O.percentile_linear(am, 99.5)
    d = np.rint(255 * (am - low) / ((high - low) + 0.000001))
    return np.clip(d, 0, 255).astype(np.int64)


def packet_pattern_search(samples: np.ndarray, sample_rate: int) -> list:
    """data_packet.py:313-331: template ``[255] + [0]*samples(0.025) + [255]`` and signal both shifted by -128,
    greedy picker seeded with the sentinel ``(-mindistance, 0)`` (which the first positive correlation within
    ``mindistance`` REPLACES), no peak limit; the first entry of the list is dropped on return."""
    n = len(samples)
    length = n / sample_rate

    def nsamp(x):
        return int((x / length) * n)

    k = nsamp(0.025)
    mindistance = nsamp(0.4)
    L = k + 2
    d = np.asarray(samples, dtype=np.int64) - 128
    csum = np.concatenate([[0], np.cumsum(d)])
    m = n - L
    if m <= 0:
        return []
    i = np.arange(m)
    corr = 127 * d[i] + 127 * d[i + k + 1] - 128 * (csum[i + k + 1] - csum[i + 1])
    peaks = [(-mindistance, 0)]
    for pos in range(m):
        c = int(corr[pos])
        if pos - peaks[-1][0] > mindistance:
            peaks.append((pos, c))
        elif c > peaks[-1][1]:
            peaks[-1] = (pos, c)
    return [p[0] for p in peaks][1:]


def find_sync_pulse(raw_samples: np.ndarray, sample_rate: int, settings=None, notch_freq=2600, notch_q=1) -> dict:
    """data_packet.py:301-343: exactly one spectral peak (height 0.5, prominence 0.2, NO distance filter), inside
    1400..1600 Hz, and at least one pulse position from the template search on the packet's own grey levels."""
    s = dict(DEFAULT_SYNC_SETTINGS, **(settings or {}))
    freq, amp = fourier_transform(raw_samples, sample_rate)
    peaks = find_peaks(amp, 1, s["height"], s["prominence"])
    in_range = all(s["min_frequency"] <= f <= s["max_frequency"] for f in freq[peaks])
    pulses = packet_pattern_search(process_samples(raw_samples, sample_rate, notch_freq, notch_q), sample_rate)
    freq_found = bool(in_range and len(peaks) == 1)
    return dict(frequency_peak_found=freq_found, samples_peak_found=bool(pulses),
                pulse_found=bool(freq_found and pulses), n_fft_peaks=int(len(peaks)),
                peaks_samples=[int(p) for p in pulses])


def state_machine(pcm: np.ndarray, sample_rate: int, packet_seconds: float = 1.0, settings=None, sync_settings=None):
    """wefax_live.py:175-200 run over the consecutive packets of a recording: start tone for >= 4 s of consecutive
    packets, THEN the first packet whose sync pulse is found starts the picture at its last pulse
    (``data_points = packet.samples[peaks_samples[-1]:]``), THEN the stop tone for >= 4 s ends it.  Returns the
    list of ``(start_packet, image_start_sample, stop_packet)`` (entries are None where the recording ends first)."""
    plen = int(sample_rate * packet_seconds)
    npk = pcm.shape[0] // plen
    out = []
    start_found = phasing_found = False
    n_start = n_stop = 0
    cur = None
    for k in range(npk):
        seg = pcm[k * plen:(k + 1) * plen]
        if not start_found:
            n_start = n_start + 1 if contain_start_tone(seg, sample_rate, settings) else 0
            if n_start * packet_seconds >= 4:
                start_found = True
                cur = [k, None, None]
        if start_found and not phasing_found:
            info = find_sync_pulse(seg, sample_rate, sync_settings)
            if info["pulse_found"]:
                phasing_found = True
                cur[1] = k * plen + info["peaks_samples"][-1]
        if start_found and phasing_found:
            n_stop = n_stop + 1 if contain_stop_tone(seg, sample_rate, settings) else 0
            if n_stop * packet_seconds >= 4:
                cur[2] = k
                out.append(tuple(cur))
                # the live decoder ends its session here; a file scan looks for the next transmission
                start_found = phasing_found = False
                n_start = n_stop = 0
                cur = None
    if cur is not None:
        out.append(tuple(cur))
    return out


def scan(pcm: np.ndarray, sample_rate: int, packet_seconds: float = 1.0, settings=None):
    """Start / stop flags of consecutive packets of a recording (what the live state machine,
    wefax_live.py:175-200, evaluates packet by packet)."""
    plen = int(sample_rate * packet_seconds)
    npk = pcm.shape[0] // plen
    start = np.zeros(npk, dtype=bool)
    stop = np.zeros(npk, dtype=bool)
    for k in range(npk):
        seg = pcm[k * plen: (k + 1) * plen]
        start[k] = contain_start_tone(seg, sample_rate, settings)
        stop[k] = contain_stop_tone(seg, sample_rate, settings)
    return start, stop
