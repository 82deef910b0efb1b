"""Generate the golden vectors that pin ``oracle/wefax_oracle.py``.

Runs the UNMODIFIED reference decoder (``/root/reference/wefax.py``) in-process
through ``oracle/ref_runner.py`` — so it only works in the build container where
``/root/reference`` is mounted — and writes

* ``tests/golden/full_<case>.npz``   : input PCM + every post-``process()``
  attribute of the reference, full arrays (small cases);
* ``tests/golden/digests.json``      : for longer seeded synthetics, SHA-256 of
  the integer outputs and a strided float64 sample of the float stages.

Usage (build container):  python tests/golden/make_golden.py            # everything
                          python tests/golden/make_golden.py NAME ...   # (re)make only these digest cases
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_runner  # noqa: E402
from wefax_b200 import synth  # noqa: E402

FIXTURE_DIR = os.path.join(ref_runner.REFERENCE_ROOT, "test_files", "parts")

#: small cases stored in full: name -> (kind, args)
FULL_CASES = {
    "fixture_image": ("wav", "image.wav", 120),
    "fixture_stop_tone": ("wav", "stop_tone.wav", 120),
    "fixture_start_tone": ("wav", "start_tone.wav", 120),
    "fixture_start_tone_noisy": ("wav", "start_tone_noisy.wav", 120),
    "fixture_start_tone_start": ("wav", "start_tone_start.wav", 120),
    "synth_12s_240": ("synth", dict(duration_s=12.0, lpm=240, seed=11, noise_sigma=0.03), 240),
    "synth_10s_48k_120": ("synth", dict(duration_s=10.0, sample_rate=48000, lpm=120, seed=12,
                                        noise_sigma=0.02), 120),
    "synth_stereo_8s_240": ("stereo", dict(duration_s=8.0, lpm=240, seed=13), 240),
}

#: longer seeded synthetics pinned by digest: name -> (synth kwargs, lpm)
DIGEST_CASES = {
    "clean_30s_120": (dict(duration_s=30.0, lpm=120, seed=3), 120),
    "noisy_40s_60": (dict(duration_s=40.0, lpm=60, seed=3, noise_sigma=0.05), 60),
    "offset_25s_240": (dict(duration_s=25.0, lpm=240, seed=3, noise_sigma=0.02,
                            carrier_offset_hz=30.0), 240),
    "drift_33s_90": (dict(duration_s=33.3, lpm=90, seed=3, noise_sigma=0.1, drift_ppm=5.0), 90),
    "resamp_20s_48k_120": (dict(duration_s=20.0, sample_rate=48000, lpm=120, seed=3,
                                noise_sigma=0.03), 120),
    "noisy_45s_100": (dict(duration_s=45.0, lpm=100, seed=4, noise_sigma=0.04), 100),
    "noisy_30s_180": (dict(duration_s=30.0, lpm=180, seed=5, noise_sigma=0.04), 180),
    "odd_len_prime_240": (dict(duration_s=100003 / 11025, lpm=240, seed=6, noise_sigma=0.02), 240),
    "resamp_16s_44k1_120": (dict(duration_s=16.0, sample_rate=44100, lpm=120, seed=7,
                                 noise_sigma=0.02), 120),
    "upsamp_20s_8k_120": (dict(duration_s=20.0, sample_rate=8000, lpm=120, seed=8,
                               noise_sigma=0.02), 120),
    # long enough for pattern_search to reach its 100-peak break (wefax.py:251); non-zero start_frame
    "long_130s_120": (dict(duration_s=130.0, lpm=120, seed=21, noise_sigma=0.03), 120),
    "long_100s_180": (dict(duration_s=100.0, lpm=180, seed=5, noise_sigma=0.04), 180),
    "long_80s_240_offset": (dict(duration_s=80.0, lpm=240, seed=3, noise_sigma=0.02, carrier_offset_hz=30.0), 240),
}

FLOAT_SAMPLE = 4096


def stereo_case(**kw) -> np.ndarray:
    """Two channels whose sum overflows int16 in places (wefax.py:372 wraps)."""
    left = synth.synth_recording(amplitude=0.8, **kw).astype(np.int32)
    right = np.roll(left, 1)
    return np.stack([left, right], axis=1).astype(np.int16)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def float_sample(a: np.ndarray) -> list:
    idx = np.linspace(0, a.shape[0] - 1, FLOAT_SAMPLE).astype(np.int64)
    return [float(v) for v in a[idx]]


def versions() -> dict:
    import PIL
    import scipy
    return dict(numpy=np.__version__, scipy=scipy.__version__, pillow=PIL.__version__,
                python=sys.version.split()[0])


def main() -> None:
    from scipy.io import wavfile
    tmp = tempfile.mkdtemp()
    ver = versions()
    only = sys.argv[1:]

    for name, (kind, arg, lpm) in ({} if only else FULL_CASES).items():
        if kind == "wav":
            path = os.path.join(FIXTURE_DIR, arg)
            sr, pcm = wavfile.read(path)
        else:
            pcm = stereo_case(**arg) if kind == "stereo" else synth.synth_recording(**arg)
            sr = arg.get("sample_rate", 11025)
            path = os.path.join(tmp, name + ".wav")
            synth.write_wav(path, pcm, sr)
        r = ref_runner.run_reference(path, lpm)
        blob = dict(pcm=pcm, sample_rate_in=sr, lpm=lpm,
                    audio_data=r["audio_data"], demodulated_data=r["demodulated_data"],
                    digitalized_data=r["digitalized_data"].astype(np.uint8),
                    error=json.dumps(r["error"]), versions=json.dumps(ver),
                    progress_titles=json.dumps(r["progress_titles"]))
        if r["error"] is None:
            blob.update(phasing_signals=np.asarray(r["phasing_signals"], dtype=np.int64),
                        start_frame=r["start_frame"], output_image=r["output_image"])
        np.savez_compressed(os.path.join(HERE, f"full_{name}.npz"), **blob)
        print(name, "error=", r["error"], "start_frame=", r.get("start_frame"))

    digests = {"versions": ver, "cases": {}}
    if only:
        with open(os.path.join(HERE, "digests.json")) as fh:
            digests = json.load(fh)
        if digests["versions"] != ver:
            raise SystemExit(f"installed versions {ver} differ from the file's {digests['versions']}: regenerate all")
    for name, (kw, lpm) in DIGEST_CASES.items():
        if only and name not in only:
            continue
        pcm = synth.synth_recording(**kw)
        sr = kw.get("sample_rate", 11025)
        path = os.path.join(tmp, name + ".wav")
        synth.write_wav(path, pcm, sr)
        r = ref_runner.run_reference(path, lpm)
        entry = dict(synth=kw, lpm=lpm, pcm_sha256=sha(pcm), n_out=int(r["audio_data"].shape[0]),
                     error=r["error"],
                     audio_sample=float_sample(r["audio_data"]),
                     audio_absmax=float(np.abs(r["audio_data"]).max()),
                     demod_sample=float_sample(r["demodulated_data"]),
                     demod_absmax=float(np.abs(r["demodulated_data"]).max()),
                     digitalized_sha256=sha(r["digitalized_data"].astype(np.uint8)))
        if r["error"] is None:
            entry.update(phasing_signals=r["phasing_signals"], start_frame=int(r["start_frame"]),
                         image_shape=list(r["output_image"].shape),
                         image_sha256=sha(r["output_image"]))
        digests["cases"][name] = entry
        print(name, "error=", r["error"], "start_frame=", r.get("start_frame"),
              "phasing=", r.get("phasing_signals"))
    with open(os.path.join(HERE, "digests.json"), "w") as fh:
        json.dump(digests, fh, indent=1)


#: recordings whose 1-s packets are run through the reference's tone test
TONE_CASES = {
    "clean_30s_ioc576": dict(duration_s=30.0, lpm=120, ioc=576, seed=3),
    "clean_30s_ioc288": dict(duration_s=30.0, lpm=120, ioc=288, seed=4),
    "noisy_30s_ioc576": dict(duration_s=30.0, lpm=120, ioc=576, seed=5, noise_sigma=0.05),
    "offset_30s_ioc576": dict(duration_s=30.0, lpm=60, ioc=576, seed=6, noise_sigma=0.02, carrier_offset_hz=40.0),
}


def sync_entry(sp: dict) -> dict:
    """DataPacket.find_sync_pulse() of the reference for one packet + a digest of its processed samples."""
    return dict(pulse_found=sp["pulse_found"], frequency_peak_found=sp["frequency_peak_found"],
                samples_peak_found=sp["samples_peak_found"], n_fft_peaks=sp["n_fft_peaks"],
                peaks_samples=sp["peaks_samples"], samples_sha256=sha(sp["samples"].astype(np.uint8)))


def make_tones() -> None:
    """tests/golden/tones.json: contain_start_tone / contain_stop_tone of the reference's DataPacket."""
    from scipy.io import wavfile
    out = {"versions": versions(), "fixtures": {}, "synthetic": {}}
    for name in ("image.wav", "stop_tone.wav", "start_tone.wav", "start_tone_noisy.wav", "start_tone_start.wav"):
        sr, pcm = wavfile.read(os.path.join(FIXTURE_DIR, name))
        start, stop = ref_runner.run_reference_tones(pcm, sr)
        sp = ref_runner.run_reference_sync_pulse(pcm, sr)
        out["fixtures"][name] = dict(sample_rate=int(sr), start=start, stop=stop, sync_pulse=sync_entry(sp))
        print(name, sr, start, stop, sp["pulse_found"], sp["peaks_samples"])
    for name, kw in TONE_CASES.items():
        pcm = synth.synth_recording(**kw)
        flags = [ref_runner.run_reference_tones(pcm[k * 11025:(k + 1) * 11025], 11025, kw["lpm"])
                 for k in range(pcm.shape[0] // 11025)]
        pulses = [ref_runner.run_reference_sync_pulse(pcm[k * 11025:(k + 1) * 11025], 11025, kw["lpm"])
                  for k in range(pcm.shape[0] // 11025)]
        out["synthetic"][name] = dict(synth=kw, pcm_sha256=sha(pcm), start=[f[0] for f in flags],
                                      stop=[f[1] for f in flags], sync_pulse=[sync_entry(p) for p in pulses])
        print(name, "start:", "".join("1" if f[0] else "." for f in flags), "stop:",
              "".join("1" if f[1] else "." for f in flags))
    with open(os.path.join(HERE, "tones.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    if "--tones-only" not in sys.argv:
        main()
    make_tones()
