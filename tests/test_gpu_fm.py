"""GPU tests of the FM-discriminator extension (csrc/fm.cu).

There is NO reference parity for this mode (the reference only describes it, README.md:85-101).
Two checks instead:
  * CUDA (fp32) against oracle/fm_oracle.py (numpy float64 of the same definition): per-sample grey
    within 2e-3 of the 0..1 range, line start within +-1 sample, >= 99.9 % of pixels within +-1;
  * both against the GROUND TRUTH of the synthetic generator (the grey levels that were transmitted,
    box-averaged over the same pixel spans): mean absolute error and fraction within +-16 levels,
    bounds set ~1.5x above what the float64 oracle itself achieves (a discriminator behind a
    1.4 kHz band-pass cannot follow 64-sample steps instantly).
"""
import numpy as np
import pytest

from oracle import fm_oracle as F
from wefax_b200 import synth

pytestmark = pytest.mark.gpu

SR = 11025


@pytest.fixture(scope="module")
def dec():
    from wefax_b200.decoder import Decoder
    d = Decoder(0)
    yield d
    d.close()


def _compare(res, ref, img_frac=0.999):
    g, rg = res.grey, ref["grey"]
    core = slice(200, len(rg) - 200)
    assert np.abs(g[core] - rg[core]).max() < 2e-3
    assert abs(res.line_start - ref["line_start"]) <= 1
    if res.line_start == ref["line_start"]:
        assert res.image.shape == ref["image"].shape
        d = np.abs(res.image.astype(int) - ref["image"].astype(int))
        assert (d <= 1).mean() >= img_frac, float((d <= 1).mean())


@pytest.mark.parametrize("lpm,ioc,block,noise", [(120, 576, 64, 0.0), (60, 288, 64, 0.0), (240, 576, 32, 0.02)])
def test_fm_decode_matches_oracle_and_ground_truth(dec, lpm, ioc, block, noise):
    from wefax_b200.fm import decode_fm
    pcm, truth = synth.synth_recording(60.0, lpm=lpm, ioc=ioc, seed=7, block=block, noise_sigma=noise, return_grey=True)
    search_from, image_end = 5 * SR, len(pcm) - 15 * SR          # phasing starts at 5 s, stop tone 15 s before the end
    res = decode_fm(dec, pcm, SR, lpm=lpm, ioc=ioc, search_from=search_from, image_end=image_end, want_grey=True)
    ref = F.decode(pcm, lpm, ioc, search_from=search_from, fold_lines=20, image_end=image_end)
    _compare(res, ref)
    assert res.width == int(round(np.pi * ioc)) and res.image.shape[1] == res.width
    # ground truth: the transmitted grey, same pixel geometry from the TRUE line start
    assert abs(res.line_start - search_from) <= 3
    timg = F.image(truth, search_from, lpm, ioc, image_end)
    rows = min(timg.shape[0], res.image.shape[0])
    err = np.abs(res.image[:rows].astype(int) - timg[:rows].astype(int))
    oerr = np.abs(ref["image"][:rows].astype(int) - timg[:rows].astype(int))
    assert err.mean() <= 1.5 * oerr.mean() + 0.5
    assert err.mean() < (6.0 if noise == 0 else 14.0), float(err.mean())
    assert (err <= 16).mean() > (0.93 if noise == 0 else 0.75), float((err <= 16).mean())


def test_fm_decode_finds_the_picture_from_the_tones(dec):
    """search_from / image_end from the start / stop tone scan (N2 feeding N4)."""
    from wefax_b200.fm import decode_fm
    pcm, truth = synth.synth_recording(60.0, seed=3, block=64, return_grey=True)
    res = decode_fm(dec, pcm, SR)
    assert res.search_from == 5 * SR and res.image_end == 45 * SR
    assert abs(res.line_start - 5 * SR) <= 3
    assert res.image.shape in ((79, 1810), (80, 1810))     # one sample late costs the last row
    # phasing lines: 5 % white then black
    ph = res.image[2:28]
    assert ph[:, 5:80].mean() > 200 and ph[:, 150:].mean() < 30


def test_fm_decode_of_a_48k_recording(dec):
    """Resampled input: same picture as decoding the 11025 Hz rendering of the same transmission (+-16 levels)."""
    from wefax_b200.fm import decode_fm
    pcm48 = synth.synth_recording(40.0, sample_rate=48000, seed=5, block=64)
    pcm11 = synth.synth_recording(40.0, sample_rate=SR, seed=5, block=64)
    a = decode_fm(dec, pcm48, 48000, search_from=5 * SR, image_end=25 * SR)
    b = decode_fm(dec, pcm11, SR, search_from=5 * SR, image_end=25 * SR)
    assert abs(a.line_start - b.line_start) <= 2
    rows = min(a.image.shape[0], b.image.shape[0])
    assert rows > 10
    # the two renderings draw different random greys per block grid; compare the deterministic part: phasing pulse
    assert a.image[:rows, 5:80].mean() > 150 and b.image[:rows, 5:80].mean() > 150


def test_fm_rejects_bad_parameters(dec):
    from wefax_b200.fm import decode_fm
    pcm = synth.synth_recording(20.0, seed=1)
    with pytest.raises(ValueError):
        decode_fm(dec, pcm, SR, band=(2600.0, 1200.0), search_from=0, image_end=len(pcm))
    with pytest.raises(ValueError):
        decode_fm(dec, pcm, SR, lpm=-5, search_from=0, image_end=len(pcm))
