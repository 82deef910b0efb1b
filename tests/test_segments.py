"""Segment mode on CPU (SURVEY.md 8(e), configs[2]): segment planning, the exact global percentiles from
summed digit histograms, line ownership / assembly, and the two-rank exchange over gloo.  The GPU worker
is replaced by tests/segment_fakes.py (oracle-backed); tests/test_gpu_segments.py runs the real one."""
import os
import socket

import numpy as np
import pytest

from oracle import wefax_oracle as O
from wefax_b200 import _native as N
from wefax_b200 import segments as S
from wefax_b200 import synth

from segment_fakes import OracleSegmentWorker


@pytest.mark.parametrize("n_frames,rate", [(57_600_000, 48000), (39_690_000, 11025), (1_234_567, 11025),
                                           (960_000, 8000), (4_410_000, 44100)])
@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_plan_tiles_the_recording(n_frames, rate, G):
    halo = 20000
    segs = S.plan_segments(n_frames, rate, G, halo=halo)
    n_total = n_frames if rate == 11025 else N.resampled_length(n_frames, rate)
    assert [s.index for s in segs] == list(range(G))
    assert segs[0].core_begin == 0 and segs[-1].core_end == n_total
    for a, b in zip(segs, segs[1:]):
        assert a.core_end == b.core_begin                       # cores tile [0, n_total) without overlap
    for s in segs:
        assert s.out_begin <= s.core_begin < s.core_end <= s.out_end
        assert s.n_out < n_total or G == 1
        if G > 1:                                               # halo on BOTH sides: circular at the ends
            assert s.core_begin - s.out_begin >= halo and s.out_end - s.core_end >= halo
            # exact ratio: the extended segment resamples to the planned number of samples, from a whole sample
            assert s.n_out * rate == (s.in_end - s.in_begin) * 11025
            assert s.out_begin * rate == s.in_begin * 11025
        # the seam is where the wrapped halo meets the recording's own samples
        if G > 1 and s.index == 0:
            assert s.in_begin < 0 and s.seam == -s.out_begin and s.seam % 16 == 0
        elif G > 1 and s.index == G - 1:
            assert s.in_end > n_frames and s.seam == n_total - s.out_begin
        else:
            assert s.seam == 0 and 0 <= s.in_begin and s.in_end <= n_frames
        frames = S.segment_frames(np.arange(n_frames, dtype=np.int32), s)
        assert frames.shape[0] == s.in_end - s.in_begin
        assert frames[0] == s.in_begin % n_frames and frames[-1] == (s.in_end - 1) % n_frames
        if s.seam:
            k = s.seam * rate // 11025
            assert frames[k] == 0 and frames[k - 1] == n_frames - 1


def test_ragged_recordings_are_refused_in_segment_mode():
    assert len(S.plan_segments(2_000_003, 44100, 1)) == 1
    with pytest.raises(ValueError, match="whole number"):
        S.plan_segments(2_000_003, 44100, 2)
    with pytest.raises(ValueError, match="too short"):
        S.plan_segments(48000, 48000, 4)
    with pytest.raises(ValueError):
        S.plan_segments(1000, 11025, 0)


def test_plan_head_extends_the_first_segment_only():
    plain = S.plan_segments(13_230_000, 11025, 8, halo=20000, plannable=None)
    long_head = S.plan_segments(13_230_000, 11025, 8, halo=20000, plannable=None, head=3_000_000)
    assert long_head[0].out_end >= 3_000_000 > plain[0].out_end
    assert long_head[0].core_end == plain[0].core_end and long_head[1:] == plain[1:]
    whole = S.plan_segments(661_500, 11025, 3, halo=20000, plannable=None, head=10**9)
    assert (whole[0].out_begin, whole[0].out_end, whole[0].seam) == (0, 661_500, 0)


def test_plan_prefers_directly_plannable_lengths():
    for G in (2, 4, 8):
        for s in S.plan_segments(57_600_000, 48000, G):
            assert N.fft_plan_describe(s.n_out // 2)[1] == 0
            assert N.fft_plan_describe((s.in_end - s.in_begin) // 2)[1] == 0


@pytest.mark.parametrize("n,parts", [(10, 1), (1001, 3), (200_000, 4), (77_777, 8)])
def test_percentiles_from_summed_histograms_are_numpy_exact(n, parts):
    rng = np.random.default_rng(n)
    v = np.abs(rng.normal(8000, 900, n)).astype(np.float32)
    v[rng.integers(0, n, n // 50)] = 0.0                         # ties and zeros
    v[: n // 10] = np.float32(8192.0)
    cuts = np.linspace(0, n, parts + 1).astype(int)

    def histogram(level, prefix):
        total = np.zeros((4, 2048), dtype=np.int64)
        for a, b in zip(cuts, cuts[1:]):
            wk = OracleSegmentWorker()
            wk.med, wk.core = v, (int(a), int(b))
            total += wk.segment_histogram(level, prefix)
        return total

    ranks, fracs = S.percentile_targets(n)
    got = S.select_order_statistics(histogram, ranks)
    srt = np.sort(v)
    assert got == [float(srt[r]) for r in ranks]
    low, high = S._lerp(got[0], got[1], fracs[0]), S._lerp(got[2], got[3], fracs[1])
    want = np.percentile(v.astype(np.float64), (0.5, 99.5))
    assert low == want[0] and high == want[1]


def test_owned_lines_partition_the_image():
    segs = S.plan_segments(3_000_000, 11025, 5, halo=20000, plannable=None)
    for start in (0, 1, 5511, 5512, 123_456, 551_200):
        w = 5512
        n_lines = (3_000_000 - start) // w
        spans = [S.owned_lines(s, start, w, n_lines) for s in segs]
        assert spans[0][0] == 0 and spans[-1][1] == n_lines
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
        for s, (r0, r1) in zip(segs, spans):
            for r in (r0, r1 - 1):
                if r1 > r0:
                    assert s.core_begin <= start + r * w < s.core_end


def _recording(seconds=40.0, lpm=120, seed=5, rate=11025):
    return synth.synth_recording(seconds, lpm=lpm, seed=seed, noise_sigma=0.03, sample_rate=rate)


def test_one_segment_is_the_plain_decode():
    """G = 1: no halo, no approximation -> every output equals the oracle's whole decode."""
    pcm = _recording(20.0)
    ref = O.decode(pcm, 11025, 120)
    res = S.decode_segmented(pcm, 11025, 120, [OracleSegmentWorker()], want=("raster", "digitalized", "demodulated"),
                             plannable=None)
    assert res.error() is None and ref["error"] is None
    assert res.start_frame == ref["start_frame"] and res.peaks == list(ref["peaks"])
    assert res.phasing_signals == list(ref["phasing_signals"])
    # percentiles of the float32 envelope vs the float64 one
    assert abs(res.low - ref["low"]) <= 1e-6 * abs(ref["high"]) and abs(res.high - ref["high"]) <= 1e-6 * abs(ref["high"])
    assert (np.abs(res.digitalized[0].astype(int) - ref["digitalized_data"]) <= 1).all()
    assert res.image.shape == ref["output_image"].shape
    assert (np.abs(res.image.astype(int) - ref["output_image"].astype(int)) <= 1).mean() >= 0.999


@pytest.mark.parametrize("G,rate,seconds,seed", [(3, 11025, 60.0, 5), (2, 48000, 30.0, 5), (2, 11025, 130.0, 21)])
def test_segments_match_the_whole_decode_within_the_stated_tolerance(G, rate, seconds, seed):
    pcm = _recording(seconds, rate=rate, seed=seed)       # seed 21: start_frame 20541 (lines off the segment cuts)
    ref = O.decode(pcm, rate, 120)
    halo = 20000
    workers = [OracleSegmentWorker() for _ in range(G)]
    res = S.decode_segmented(pcm, rate, 120, workers, halo=halo, want=("raster", "digitalized", "demodulated"),
                             plannable=None)
    assert res.error() is None
    peak = np.abs(ref["demodulated_data"]).max()
    dem = np.concatenate([res.demodulated[k] for k in sorted(res.demodulated)])
    dig = np.concatenate([res.digitalized[k] for k in sorted(res.digitalized)])
    assert dem.shape == ref["demodulated_data"].shape
    assert np.abs(dem - ref["demodulated_data"]).max() / peak < 5e-3        # halo truncation of the transforms
    assert abs(res.low - ref["low"]) / peak < 1e-3 and abs(res.high - ref["high"]) / peak < 1e-3
    assert (np.abs(dig.astype(int) - ref["digitalized_data"]) <= 1).mean() >= 0.995
    assert res.start_frame == ref["start_frame"] == (20541 if seed == 21 else 0)
    assert res.image.shape == ref["output_image"].shape
    assert (np.abs(res.image.astype(int) - ref["output_image"].astype(int)) <= 1).mean() >= 0.995


def test_start_frame_is_exact_only_given_the_grey_levels():
    """The stated limit of segment mode: the phasing search is the reference's, run on segment 0's grey levels.  Those
    differ from the whole decode's by +-1 in ~1 % of the samples (halo truncation), and the reference's greedy peak
    picker / grouping (wefax.py:239-294) is discontinuous in them.  On this recording one flipped peak changes the
    longest group: whole decode 50309, two segments 0, even in float64.  Everything else stays within tolerance."""
    pcm = synth.synth_recording(150.0, sample_rate=48000, lpm=120, seed=13, noise_sigma=0.03)
    ref = O.decode(pcm, 48000, 120)
    res = S.decode_segmented(pcm, 48000, 120, [OracleSegmentWorker(), OracleSegmentWorker()], plannable=None,
                             want=("raster", "digitalized"))
    dig = np.concatenate([res.digitalized[k] for k in sorted(res.digitalized)])
    assert (np.abs(dig.astype(int) - ref["digitalized_data"]) <= 1).all()
    assert (dig != ref["digitalized_data"]).mean() < 0.02
    assert (ref["start_frame"], res.start_frame) == (50309, 0)
    # given segment 0's own grey levels the search result IS the reference's
    head = S.plan_decode(pcm.shape[0], 48000, 120, 2, plannable=None)[0]
    consts = O.line_constants(120, 11025)
    own = np.concatenate([res.digitalized[0], np.zeros(0, np.uint8)])
    peaks = O.pattern_search(own.astype(np.int64), consts)
    assert len(peaks) == 100 and head.out_end > peaks[-1]
    assert res.peaks == list(peaks) and res.phasing_signals == list(O.find_phasing(peaks, consts))


def test_short_first_segment_is_refused():
    pcm = _recording(60.0)
    with pytest.raises(ValueError, match="first segment too short"):
        S.decode_segmented(pcm, 11025, 120, [OracleSegmentWorker() for _ in range(6)], halo=20000, plannable=None,
                           head=0)
    with pytest.raises(ValueError, match="halo"):
        S.decode_segmented(pcm, 11025, 60, [OracleSegmentWorker() for _ in range(2)], halo=20000, plannable=None)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, queue):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pcm = _recording(60.0)
    res = S.decode_segmented(pcm, 11025, 120, [OracleSegmentWorker()], halo=20000, plannable=None)
    dist.barrier()
    if rank == 0:
        queue.put((res.low, res.high, res.start_frame, res.image))
    else:
        assert res.image is None and res.rows
    dist.destroy_process_group()


def test_two_ranks_over_gloo_equal_two_local_segments():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    low, high, start, image = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pcm = _recording(60.0)
    local = S.decode_segmented(pcm, 11025, 120, [OracleSegmentWorker(), OracleSegmentWorker()], halo=20000,
                               plannable=None)
    assert (low, high, start) == (local.low, local.high, local.start_frame)
    assert np.array_equal(image, local.image)


def test_plan_invariants_hold_for_arbitrary_recordings():
    """Property test: any (length, rate, segment count, halo) either plans a valid tiling or is refused."""
    import math
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(rate=st.sampled_from([8000, 11025, 16000, 22050, 44100, 48000]), units=st.integers(1, 400_000),
           G=st.integers(1, 8), halo=st.integers(1, 70_000), head=st.integers(0, 2_000_000))
    def check(rate, units, G, halo, head):
        g = math.gcd(rate, 11025)
        u_in, u_out = rate // g, 11025 // g
        n_frames = units * u_in
        n_total = n_frames if rate == 11025 else N.resampled_length(n_frames, rate)
        try:
            segs = S.plan_segments(n_frames, rate, G, halo=halo, plannable=None, head=head)
        except ValueError:
            assert G > 1            # one segment is always possible
            return
        assert len(segs) == G and segs[0].core_begin == 0 and segs[-1].core_end == n_total
        for a, b in zip(segs, segs[1:]):
            assert a.core_end == b.core_begin
        for s in segs:
            assert s.core_begin < s.core_end and s.out_begin <= s.core_begin and s.core_end <= s.out_end
            assert s.n_out <= n_total                                  # never laps itself
            if G > 1:
                assert s.n_out * rate == (s.in_end - s.in_begin) * 11025
                whole = s.n_out == n_total and s.seam == 0             # stretched to the whole recording
                if not whole:
                    assert s.core_begin - s.out_begin >= halo and s.out_end - s.core_end >= halo
                if s.seam:
                    assert s.seam <= s.core_begin - s.out_begin or s.core_end - s.out_begin <= s.seam
            frames = S.segment_frames(np.empty(n_frames, dtype=np.int8), s)
            assert frames.shape[0] == s.in_end - s.in_begin
        if G > 1 and segs[0].n_out < n_total:
            assert segs[0].out_end >= min(head, n_total - 1) or segs[0].out_end >= n_total - u_out

    check()
