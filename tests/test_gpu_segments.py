"""Segment mode on the GPU (SURVEY.md 8(e), configs[2]) through the C-ABI: several native contexts on one
device stand in for several GPUs (the per-rank code is the same; tests/test_segments.py covers the
two-process exchange over gloo).

Tolerance of segment mode against the whole-recording decode, halo 65536 samples:
envelope within 2e-3 of its peak, percentiles within 1e-3 of the peak, start_frame identical, >= 99.5 % of
grey levels and pixels within +-1.  With ONE segment there is no halo and the result is the plain decode's."""
import numpy as np
import pytest

from oracle import wefax_oracle as O
from wefax_b200 import segments as S
from wefax_b200 import synth
from wefax_b200.decoder import Decoder

from segment_fakes import OracleSegmentWorker

pytestmark = pytest.mark.gpu


def _decoders(k):
    return [Decoder(0) for _ in range(k)]


def _close(ds):
    for d in ds:
        d.close()


def _cat(parts):
    return np.concatenate([parts[k] for k in sorted(parts)])


def test_one_segment_equals_the_plain_decode_bit_for_bit():
    pcm = synth.synth_recording(30.0, lpm=120, seed=7, noise_sigma=0.03)
    with Decoder(0) as dec:
        whole = dec.decode(pcm, 11025, 120, want=("demodulated", "digitalized", "raster"))
        res = S.decode_segmented(pcm, 11025, 120, [dec], want=("raster", "digitalized", "demodulated"))
    assert res.error() is None and whole.error(0) is None
    assert (res.low, res.high) == (whole.low_high[0, 0], whole.low_high[0, 1])
    assert res.peaks == whole.peaks[0] and res.phasing_signals == whole.phasing_signals[0]
    assert res.start_frame == int(whole.start_frame[0])
    assert np.array_equal(res.digitalized[0], whole.digitalized[0])
    assert np.array_equal(res.demodulated[0], whole.demodulated[0])
    assert np.array_equal(res.image, whole.image(0))


@pytest.mark.parametrize("level", [0, 1, 2])
def test_histogram_kernel_counts_the_core(level):
    pcm = synth.synth_recording(12.0, lpm=120, seed=3, noise_sigma=0.05)
    n = pcm.shape[0]
    lo, hi = 1003, n - 2001                                   # unaligned core inside the segment
    with Decoder(0) as dec:
        dec.segment_envelope(pcm, 11025, lo, hi)
        med = dec.segment_quantise(1.0, 2.0, want=("demodulated",))["demodulated"]
        keys = med.view(np.uint32).astype(np.int64)
        # prefixes that exist in the data: the digits of the median and maximum values
        srt = np.sort(keys)
        picks = [srt[len(srt) // 2], srt[-1], srt[0], srt[len(srt) // 3]]
        prefix = [0] * 4 if level == 0 else [int(k >> (21 if level == 1 else 10)) for k in picks]
        got = dec.segment_histogram(level, prefix)
    fake = OracleSegmentWorker()
    fake.med, fake.core = med, (0, med.shape[0])
    want = fake.segment_histogram(level, prefix)
    assert np.array_equal(got, want)
    assert got.sum() > 0


# (segments, rate, seconds, lpm, seed, noise): the seeds 21 / 5 / 3 give start_frame 20541 / 283057 / 24436 (20468 at
# 48 kHz), i.e. image lines that do not start on the segment cuts.  (48 kHz, seed 13 is the known sensitive case of
# tests/test_segments.py::test_start_frame_is_exact_only_given_the_grey_levels.)
@pytest.mark.parametrize("G,rate,seconds,lpm,seed,noise", [
    (2, 11025, 130.0, 120, 21, 0.03), (4, 11025, 300.0, 120, 21, 0.03), (3, 11025, 100.0, 180, 5, 0.04),
    (2, 11025, 80.0, 240, 3, 0.02), (2, 48000, 150.0, 120, 5, 0.03), (2, 11025, 260.0, 60, 13, 0.03)])
def test_segments_match_the_whole_decode(G, rate, seconds, lpm, seed, noise):
    pcm = synth.synth_recording(seconds, sample_rate=rate, lpm=lpm, seed=seed, noise_sigma=noise,
                                carrier_offset_hz=30.0 if lpm == 240 else 0.0)
    ds = _decoders(G)
    try:
        whole = ds[0].decode(pcm, rate, lpm, want=("demodulated", "digitalized", "raster"))
        res = S.decode_segmented(pcm, rate, lpm, ds, want=("raster", "digitalized", "demodulated"))
    finally:
        _close(ds)
    assert res.error() is None and whole.error(0) is None
    dem, dig = _cat(res.demodulated), _cat(res.digitalized)
    ref_dem, ref_dig = whole.demodulated[0], whole.digitalized[0]
    assert dem.shape == ref_dem.shape and dig.shape == ref_dig.shape
    peak = float(np.abs(ref_dem).max())
    assert np.abs(dem - ref_dem).max() / peak < 2e-3
    assert abs(res.low - whole.low_high[0, 0]) / peak < 1e-3 and abs(res.high - whole.low_high[0, 1]) / peak < 1e-3
    assert (np.abs(dig.astype(int) - ref_dig.astype(int)) <= 1).mean() >= 0.995
    # the search is exact GIVEN the grey levels; a +-1 grey level can move a peak inside a flat correlation
    # maximum, the phasing group and start_frame are what the image depends on
    assert res.start_frame == int(whole.start_frame[0]) and len(res.peaks) == len(whole.peaks[0])
    assert res.phasing_signals == whole.phasing_signals[0]
    img = whole.image(0)
    assert res.image.shape == img.shape
    assert (np.abs(res.image.astype(int) - img.astype(int)) <= 1).mean() >= 0.995


def test_segments_against_the_oracle_and_the_cpu_protocol():
    """Same protocol, GPU workers vs oracle-backed workers: same plan, same start_frame, images within +-1."""
    pcm = synth.synth_recording(130.0, lpm=120, seed=21, noise_sigma=0.03)
    ds = _decoders(2)
    try:
        gpu = S.decode_segmented(pcm, 11025, 120, ds, halo=30000)
    finally:
        _close(ds)
    cpu = S.decode_segmented(pcm, 11025, 120, [OracleSegmentWorker(), OracleSegmentWorker()], halo=30000)
    ref = O.decode(pcm, 11025, 120)
    assert gpu.start_frame == cpu.start_frame == ref["start_frame"] == 20541
    assert gpu.phasing_signals == cpu.phasing_signals == list(ref["phasing_signals"])
    peak = np.abs(ref["demodulated_data"]).max()
    assert abs(gpu.low - cpu.low) / peak < 1e-5 and abs(gpu.high - cpu.high) / peak < 1e-5
    assert (np.abs(gpu.image.astype(int) - cpu.image.astype(int)) <= 1).mean() >= 0.999
    assert (np.abs(gpu.image.astype(int) - ref["output_image"].astype(int)) <= 1).mean() >= 0.995


def test_stereo_and_device_resident_segments():
    import torch
    mono = synth.synth_recording(140.0, lpm=120, seed=4, noise_sigma=0.02)
    stereo = np.stack([mono, mono], axis=1)
    ds = _decoders(2)
    try:
        a = S.decode_segmented(mono, 11025, 120, ds)
        b = S.decode_segmented(stereo, 11025, 120, ds)
        dev = torch.from_numpy(mono).cuda()
        c = S.decode_segmented(None, 11025, 120, ds, n_frames=mono.shape[0],
                               segment_pcm=lambda sg: S.segment_frames(dev, sg).contiguous())
        # image rows left on the GPU and assembled there (the multi-GPU path sends them to rank 0 the same way)
        d = S.decode_segmented(mono, 11025, 120, ds, rows_on_device=True, exchange=S.HostExchange(device=0))
        assert all(v.is_cuda for v in d.rows.values())
        assert np.array_equal(a.image, d.image)
    finally:
        _close(ds)
    # (L + R) / 2 of identical channels wraps for |x| >= 16384 (wefax.py:372); this recording stays below
    assert np.abs(mono).max() < 16384
    assert np.array_equal(a.image, b.image) and np.array_equal(a.image, c.image)
    assert (a.low, a.high, a.start_frame) == (c.low, c.high, c.start_frame)


def test_call_order_and_arguments_are_checked():
    pcm = synth.synth_recording(10.0, lpm=120, seed=1)
    with Decoder(0) as dec:
        with pytest.raises(ValueError, match="wefax_segment_envelope has not run"):
            dec.segment_histogram(0)
        with pytest.raises(ValueError, match="outside the extended segment"):
            dec.segment_envelope(pcm, 11025, 0, pcm.shape[0] + 1)
        with pytest.raises(ValueError, match="straddles the seam"):
            dec.segment_envelope(pcm, 11025, 0, pcm.shape[0], seam=1024)
        dec.segment_envelope(pcm, 11025, 0, pcm.shape[0])
        with pytest.raises(ValueError, match="wefax_segment_quantise has not run"):
            dec.segment_raster(120, 0, 4, 0, 4)
        dec.segment_quantise(100.0, 9000.0)
        with pytest.raises(ValueError, match="outside the extended segment"):
            dec.segment_raster(120, 0, 10 ** 6, 0, 4)
        assert dec.segment_raster(120, 0, 6, 1, 4).shape == (16, 5512)


@pytest.mark.slow
@pytest.mark.parametrize("G", [2, 8])
def test_configs2_full_size_in_segments(G):
    """BASELINE configs[2]: 20 min at 48 kHz (57.6 M frames -> 13.23 M samples) in G segments."""
    pcm = synth.synth_recording(1200.0, sample_rate=48000, lpm=120, seed=1)
    ds = _decoders(G)
    try:
        whole = ds[0].decode(pcm, 48000, 120, want=("digitalized", "raster"))
        res = S.decode_segmented(pcm, 48000, 120, ds, want=("raster", "digitalized"))
    finally:
        _close(ds)
    assert res.error() is None
    dig = _cat(res.digitalized)
    assert (np.abs(dig.astype(int) - whole.digitalized[0].astype(int)) <= 1).mean() >= 0.995
    assert res.start_frame == int(whole.start_frame[0])
    img = whole.image(0)
    assert res.image.shape == img.shape == (4 * ((13_230_000 - res.start_frame) // 5512), 5512)
    assert (np.abs(res.image.astype(int) - img.astype(int)) <= 1).mean() >= 0.995


def test_one_process_drives_two_gpus():
    """One host thread per GPU, no process group (skipped on single-GPU boxes)."""
    from wefax_b200 import _native as N
    if N.load().wefax_device_count() < 2:
        pytest.skip("needs two GPUs")
    pcm = synth.synth_recording(130.0, lpm=120, seed=21, noise_sigma=0.03)
    one = [Decoder(0), Decoder(0)]
    two = [Decoder(0), Decoder(1)]
    try:
        a = S.decode_segmented(pcm, 11025, 120, one, want=("raster", "digitalized"))
        b = S.decode_segmented(pcm, 11025, 120, two, want=("raster", "digitalized"))
        c = S.decode_segmented(pcm, 11025, 120, two, rows_on_device=True, exchange=S.HostExchange(device=0))
    finally:
        _close(one + two)
    assert (a.low, a.high, a.start_frame) == (b.low, b.high, b.start_frame) == (c.low, c.high, c.start_frame)
    assert a.start_frame == 20541
    assert np.array_equal(a.image, b.image) and np.array_equal(a.image, c.image)
    assert all(np.array_equal(a.digitalized[k], b.digitalized[k]) for k in a.digitalized)
    assert sorted(v.device.index for v in c.rows.values()) == [0, 1]
