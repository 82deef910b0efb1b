"""TEST INFRASTRUCTURE — CPU restatement (numpy, float64) of the reference decoder.

This is the parity oracle for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it;
the product package ``wefax_b200`` never does.

It restates, stage by stage, what ``/root/reference/wefax.py`` computes in
``Demodulator.process()`` (``wefax.py:46-93``).  The reference delegates the
arithmetic to third-party packages that are not under ``/root/reference``
(pins from ``requirements.txt``: numpy 1.24.2, scipy 1.10.0, Pillow 9.4.0); the
published algorithm of each call is restated here in plain numpy with
``numpy.fft`` as the only transform primitive and ``scipy.signal.lfilter`` as the
only recursion primitive (the same routine the reference's ``filtfilt`` runs).

PARITY PIN: the reference's own tests hold no numeric vectors for this path
(SURVEY.md §4), so this oracle is pinned against outputs of the reference itself
run in the build container: ``tests/golden/make_golden.py`` runs the unmodified
reference (``oracle/ref_runner.py``) on the shipped 1-s fixtures and on seeded
synthetics and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against them (float stages to
1e-9 of peak, integer stages bit-exact).  Versions the vectors were made with are
stored in each fixture.
"""
from __future__ import annotations

import math

import numpy as np

TARGET_RATE = 11025


# --------------------------------------------------------------------------- #
# per-LPM integer constants                                                    #
# --------------------------------------------------------------------------- #
def line_constants(lpm, sample_rate: int = TARGET_RATE) -> dict:
    """Integer/float constants the reference derives from LPM.

    ``time_for_one_frame = 1 / (lpm / 60)`` (wefax.py:33,44); ``samples(x) =
    int(x * frame_len * sample_rate)`` (wefax.py:223); template
    ``[1]*samples(0.005) + [0]*samples(0.001) + [1]*samples(0.005)``
    (wefax.py:225); ``mindistance = int(frame_len * sample_rate * 0.8)``
    (wefax.py:229); regular spacing window ``frame_len*sample_rate -/+ 500``
    exclusive (wefax.py:264-267); ``frame_width = int(frame_len * sample_rate)``
    (wefax.py:298).  Python float semantics (left-to-right products) are kept.
    """
    frame_len = 1 / (lpm / 60)
    n1 = int(0.005 * frame_len * sample_rate)
    n0 = int(0.001 * frame_len * sample_rate)
    return dict(frame_len=frame_len, n1=n1, n0=n0, template_len=2 * n1 + n0,
                mindistance=int(frame_len * sample_rate * 0.8),
                dev_min=frame_len * sample_rate - 500,
                dev_max=frame_len * sample_rate + 500,
                width=int(frame_len * sample_rate))


# --------------------------------------------------------------------------- #
# ingest                                                                        #
# --------------------------------------------------------------------------- #
def merge_channels(frames: np.ndarray) -> np.ndarray:
    """wefax.py:360-373: ``np.divide(np.add(L, R), 2)`` per frame.

    ``np.add`` on two numpy scalars of the stored dtype keeps that dtype, so for
    int16 PCM the sum WRAPS before the (float64) division.
    """
    frames = np.asarray(frames)
    s = np.add(frames[:, 0], frames[:, 1])          # same dtype: wraps for ints
    return np.divide(s, 2).astype(np.float64)


def resample(x: np.ndarray, num: int) -> np.ndarray:
    """wefax.py:384 ``scipy.signal.resample(x, num)`` (real input, no window).

    rfft; keep the first ``m//2+1`` bins, ``m = min(num, n)``; if ``m`` is even and
    ``num != n`` the unpaired bin ``m//2`` is doubled (down-sampling) or halved
    (up-sampling); scale by ``num/n``; irfft to ``num`` samples.
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    m = min(num, n)
    m2 = m // 2 + 1
    X = np.fft.rfft(x)[:m2].copy()
    if m % 2 == 0 and num != n:
        X[m // 2] *= 2 if num < n else 0.5
    return np.fft.irfft(X / (n / num), n=num)


def resampled_length(n_samples: int, sample_rate: int) -> int:
    """wefax.py:357,384: ``int(11025 * (len(data) / sample_rate))``."""
    return int(TARGET_RATE * (n_samples / sample_rate))


# --------------------------------------------------------------------------- #
# notch (zero-phase)                                                            #
# --------------------------------------------------------------------------- #
def notch_coefficients(freq, q, fs):
    """wefax.py:68-70 ``scipy.signal.iirnotch(w0, Q, fs)``.

    Second-order notch (Orfanidis eq. 11.3.4-11.3.7) with a -3 dB bandwidth
    ``w0/Q``: ``beta = tan(bw/2)``, ``gain = 1/(1+beta)``,
    ``b = gain*[1, -2cos(w0), 1]``, ``a = [1, -2*gain*cos(w0), 2*gain-1]``.
    """
    w0 = 2 * float(freq) / fs
    if w0 > 1.0 or w0 < 0.0:
        raise ValueError("w0 should be such that 0 < w0 < 1")
    bw = w0 / float(q) * math.pi
    w0 = w0 * math.pi
    beta = math.tan(bw / 2.0)
    gain = 1.0 / (1.0 + beta)
    b = gain * np.array([1.0, -2.0 * math.cos(w0), 1.0])
    a = np.array([1.0, -2.0 * gain * math.cos(w0), 2.0 * gain - 1.0])
    return b, a


def _lfilter_zi(b, a):
    """Steady-state initial state of a transposed direct-form-II section for a
    unit step (``scipy.signal.lfilter_zi``): solve ``(I - A^T) zi = B``."""
    n = max(len(a), len(b))
    a = np.r_[a, np.zeros(n - len(a))] / a[0]
    b = np.r_[b, np.zeros(n - len(b))] / a[0]
    companion_t = np.zeros((n - 1, n - 1))
    companion_t[:, 0] = -a[1:]
    companion_t[np.arange(n - 2), np.arange(1, n - 1)] = 1.0
    return np.linalg.solve(np.eye(n - 1) - companion_t, b[1:] - a[1:] * b[0])


def filtfilt(b, a, x) -> np.ndarray:
    """wefax.py:72 ``scipy.signal.filtfilt(b, a, x)`` with its defaults.

    ``padtype='odd'``, ``padlen = 3*max(len(a), len(b))``: the signal is extended by
    ``padlen`` odd-reflected samples at each end, filtered forwards with the initial
    state ``zi*ext[0]``, reversed, filtered again with ``zi*y[-1]``, reversed and the
    padding stripped.
    """
    from scipy.signal import lfilter  # the recursion primitive (direct form II transposed)
    x = np.asarray(x, dtype=np.float64)
    padlen = 3 * max(len(a), len(b))
    if x.shape[0] <= padlen:
        raise ValueError("The length of the input vector x must be greater than padlen, "
                         "which is %d." % padlen)
    left = 2 * x[0] - x[padlen:0:-1]
    right = 2 * x[-1] - x[-2:-(padlen + 2):-1]
    ext = np.concatenate((left, x, right))
    zi = _lfilter_zi(b, a)
    y, _ = lfilter(b, a, ext, zi=zi * ext[0])
    y, _ = lfilter(b, a, y[::-1], zi=zi * y[-1])
    return y[::-1][padlen:-padlen]


# --------------------------------------------------------------------------- #
# demodulate                                                                    #
# --------------------------------------------------------------------------- #
def hilbert(x: np.ndarray) -> np.ndarray:
    """wefax.py:174 ``scipy.signal.hilbert(x)``: one length-N FFT, bins
    ``1..ceil(N/2)-1`` doubled, negative frequencies zeroed (DC and, for even N, the
    Nyquist bin kept), inverse FFT."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    X = np.fft.fft(x)
    h = np.zeros(n)
    if n % 2 == 0:
        h[0] = h[n // 2] = 1
        h[1:n // 2] = 2
    else:
        h[0] = 1
        h[1:(n + 1) // 2] = 2
    return np.fft.ifft(X * h)


def medfilt5(v: np.ndarray) -> np.ndarray:
    """wefax.py:175 ``scipy.signal.medfilt(v, 5)``: running median of 5 with the
    two ends zero-padded."""
    v = np.asarray(v, dtype=np.float64)
    p = np.concatenate((np.zeros(2), v, np.zeros(2)))
    win = np.stack([p[i:i + v.shape[0]] for i in range(5)])
    win.sort(axis=0)
    return win[2].copy()


def demodulate(x: np.ndarray) -> np.ndarray:
    """wefax.py:166-183."""
    return medfilt5(np.abs(hilbert(x)))


# --------------------------------------------------------------------------- #
# digitalize                                                                    #
# --------------------------------------------------------------------------- #
def percentile_linear(v: np.ndarray, q: float) -> float:
    """``numpy.percentile(v, q)`` with the default 'linear' method: virtual index
    ``(N-1)*(q/100)``, neighbours ``floor`` and ``floor+1`` of the sorted data and
    numpy's two-sided lerp (``a+(b-a)*t`` for ``t<0.5``, ``b-(b-a)*(1-t)`` otherwise)."""
    n = v.shape[0]
    virt = (n - 1) * (q / 100)
    lo = int(math.floor(virt))
    hi = min(lo + 1, n - 1)
    part = np.partition(v, [lo, hi])
    a, b = float(part[lo]), float(part[hi])
    t = virt - lo
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def digitalize(env: np.ndarray):
    """wefax.py:185-216: global 0.5/99.5 percentile stretch to 0..255,
    ``numpy.round`` (half to even), clip, int.  Returns ``(values, low, high)``."""
    env = np.asarray(env, dtype=np.float64)
    low = percentile_linear(env, 0.5)
    high = percentile_linear(env, 99.5)
    delta = high - low
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.round(255 * (env - low) / delta)
    d[d < 0] = 0
    d[d > 255] = 255
    return d.astype(np.int64), low, high


# --------------------------------------------------------------------------- #
# phasing / sync search                                                         #
# --------------------------------------------------------------------------- #
def sync_correlation(dig: np.ndarray, n1: int, n0: int) -> np.ndarray:
    """wefax.py:225,232-236: ``corr[i] = dot(sync-128, (data-128)[i:i+L])`` with
    ``sync = [1]*n1+[0]*n0+[1]*n1`` — i.e. weights -127, -128, -127.  Exact int64."""
    L = 2 * n1 + n0
    s = np.asarray(dig, dtype=np.int64) - 128
    cs = np.concatenate(([0], np.cumsum(s)))
    m = s.shape[0] - L
    if m <= 0:
        return np.zeros(0, dtype=np.int64)
    i = np.arange(m)
    sa = cs[i + n1] - cs[i]
    sb = cs[i + n1 + n0] - cs[i + n1]
    sc = cs[i + L] - cs[i + n1 + n0]
    return -127 * sa - 128 * sb - 127 * sc


def pattern_search(dig: np.ndarray, consts: dict, max_peaks: int = 100):
    """wefax.py:221-261: greedy peak picker over the correlation.

    ``peaks=[(0,0)]``; a sample further than ``mindistance`` from the last peak
    opens a new peak, otherwise a strictly larger correlation replaces the last
    peak (and moves its position); stop as soon as there are 100 peaks.
    """
    n1, n0 = consts["n1"], consts["n0"]
    L = 2 * n1 + n0
    mind = consts["mindistance"]
    dig = np.asarray(dig)
    m = dig.shape[0] - L                      # range(len(data) - len(sync))
    pos, val = [0], [0]
    chunk = 1 << 20                           # the scan normally ends within ~100 lines
    for base in range(0, max(m, 0), chunk):
        stop = min(base + chunk, m)
        corr_l = sync_correlation(dig[base:stop + L], n1, n0).tolist()
        for i, c in enumerate(corr_l, start=base):
            if i - pos[-1] > mind:
                pos.append(i)
                val.append(c)
            elif c > val[-1]:
                pos[-1] = i
                val[-1] = c
            if len(pos) == max_peaks:
                return pos
    return pos


def find_phasing(peaks, consts: dict):
    """wefax.py:263-294 including its indexing quirks.

    ``clear`` keeps ``peaks[i]`` (``1 <= i <= len-2``) whose distance to the previous
    peak is regular; the grouping loop then runs ``i`` over ``range(1, len(clear)-1)``
    but indexes ``peaks`` (not ``clear``), closes a group only on an irregular gap and
    never appends the trailing group; the result is the first longest group.
    Raises ``ValueError`` when there is no group (``max([])``).
    """
    def regular(x):
        return consts["dev_max"] > x > consts["dev_min"]

    clear = [peaks[i] for i in range(1, len(peaks) - 1) if regular(peaks[i] - peaks[i - 1])]
    groups, group = [], []
    for i in range(1, len(clear) - 1):
        if regular(peaks[i] - peaks[i - 1]):
            group.append(peaks[i])
        else:
            groups.append(group)
            group = []
    if not groups:
        raise ValueError("max() iterable argument is empty")
    return max(groups, key=len)


# --------------------------------------------------------------------------- #
# image                                                                         #
# --------------------------------------------------------------------------- #
def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bicubic_rows(in_size: int, out_size: int):
    """Pillow's ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for one axis
    (``src/libImaging/Resample.c``, BICUBIC support 2.0, PRECISION_BITS = 22).
    Returns ``(xmin[out], count[out], k[out, ksize])`` with integer coefficients."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, dtype=np.int64)
    cnt_a = np.zeros(out_size, dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        xmin_a[xx] = xmin
        cnt_a[xx] = xmax
    return xmin_a, cnt_a, kk


def resize_rows_x4(img: np.ndarray) -> np.ndarray:
    """wefax.py:325 ``image.resize((w, 4*h))`` for mode 'L': Pillow's default filter
    is BICUBIC; the horizontal pass is skipped (same width); the vertical pass
    accumulates ``(1<<21) + sum(px*k)`` in int32, shifts by 22 and clips to 0..255."""
    h, w = img.shape
    if h == 0:
        return np.zeros((0, w), dtype=np.uint8)
    xmin, cnt, kk = bicubic_rows(h, 4 * h)
    src = img.astype(np.int64)
    out = np.empty((4 * h, w), dtype=np.uint8)
    ksize = kk.shape[1]
    acc = np.full((4 * h, w), 1 << 21, dtype=np.int64)
    for t in range(ksize):
        rows = np.minimum(xmin + t, h - 1)
        acc += src[rows] * np.where(t < cnt, kk[:, t], 0)[:, None]
    out[:] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
    return out


def convert_to_image(dig: np.ndarray, width: int) -> np.ndarray:
    """wefax.py:296-327: ``h = len(data)//w`` full lines, luminance ``255 - value``,
    then the x4 vertical bicubic.  Fewer than one line of data makes the reference's
    first ``putpixel`` fail with ``IndexError('image index out of range')``."""
    dig = np.asarray(dig)
    h = dig.shape[0] // width
    if h == 0 and dig.shape[0] > 0:
        raise IndexError("image index out of range")
    lum = (255 - dig[: h * width]).astype(np.uint8).reshape(h, width)
    return resize_rows_x4(lum)


# --------------------------------------------------------------------------- #
# whole path                                                                    #
# --------------------------------------------------------------------------- #
def decode(pcm: np.ndarray, sample_rate: int, lpm=120,
           notch_freq=2600, notch_q=1, stop_after: str | None = None) -> dict:
    """``Demodulator.process()`` (wefax.py:46-93) on in-memory PCM.

    ``pcm`` is what ``scipy.io.wavfile.read`` returned: ``(n,)`` mono or
    ``(n, 2)`` stereo, in its stored dtype.  Returns a dict with the reference's
    post-``process()`` attributes.  An exception the reference would raise in the
    sync search / image stage is stored under ``"error"`` as ``(type name, message)``.
    """
    out: dict = {"error": None}
    pcm = np.asarray(pcm)
    data = merge_channels(pcm) if pcm.ndim == 2 else pcm
    length = data.shape[0] / sample_rate
    if sample_rate != TARGET_RATE:
        data = resample(data, int(TARGET_RATE * length))
        sample_rate = TARGET_RATE
        length = data.shape[0] / sample_rate
    out["sample_rate"], out["length"] = sample_rate, length
    b, a = notch_coefficients(int(notch_freq), notch_q, sample_rate)
    out["audio_data"] = audio = filtfilt(b, a, data)
    if stop_after == "audio":
        return out
    out["demodulated_data"] = env = demodulate(audio)
    if stop_after == "demodulated":
        return out
    dig, low, high = digitalize(env)
    out["digitalized_data"], out["low"], out["high"] = dig, low, high
    consts = line_constants(lpm, sample_rate)
    out["peaks"] = peaks = pattern_search(dig, consts)
    try:
        out["phasing_signals"] = ph = find_phasing(peaks, consts)
        out["start_frame"] = sf = ph[-1] if ph else 0
        out["output_image"] = convert_to_image(dig[sf:], consts["width"])
    except (ValueError, IndexError) as exc:
        out["error"] = (type(exc).__name__, str(exc))
    return out
