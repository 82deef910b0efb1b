"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck): fused grey map + raster
(aligned, unaligned and generic widths, edge tiles, sequential phasing scan with lazy grey levels), symmetric and
recursive notch, percentile brackets, tone and sync-pulse scans, FM decode, lanes, segment mode."""
import os
import numpy as np, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wefax_b200 import synth
from wefax_b200.decoder import Decoder
from wefax_b200.tones import scan_tones
from wefax_b200.fm import decode_fm
pcm = synth.synth_recording(16.0, seed=1, noise_sigma=0.02)      # n = 176400 -> half 88200 = 225 x 392
with Decoder(0) as dec:
    r = dec.decode(pcm, 11025, 120, want=("audio", "demodulated", "digitalized", "raster"))
    print("decode", r.status, r.start_frame, r.height)
    b = dec.decode(np.stack([pcm, pcm[::-1].copy(), pcm]), 11025, [120, 60, 240])
    print("batch", b.status)
    print("tones", [a.sum() for a in scan_tones(dec, pcm, 11025)[:2]])
    f = decode_fm(dec, pcm, 11025, search_from=11025, image_end=len(pcm))
    print("fm", f.image.shape, f.line_start)
    # round 2: fused grey map + raster on every width class, odd lengths, raster without digitalized
    for lpm in (60, 90, 120, 240):
        q = synth.synth_recording(30.0, lpm=lpm, seed=3, noise_sigma=0.03)
        r = dec.decode(np.stack([q[:-1], q[1:]]), 11025, lpm, want=("raster",))
        print("fused", lpm, r.status, r.height)
    from wefax_b200.tones import scan_recording, scan_sync_pulses
    print("state machine", scan_recording(dec, synth.synth_recording(30.0, seed=3), 11025))
    q48 = synth.synth_recording(2.0, sample_rate=48000, seed=2)
    print("sync pulses 48k (recursive notch)", scan_sync_pulses(dec, q48, 48000)["peaks_samples"])
    print("high-Q notch", float(np.abs(dec.filtfilt(pcm.astype(np.float32), 2600, 30)).max()))
    big = synth.synth_recording(100.0, seed=4, noise_sigma=0.03)             # n >= 2^20: bracketed percentiles
    r = dec.decode(big, 11025, 120)
    print("bracketed percentiles", r.status, r.low_high)
os.environ["WEFAX_SYNC_FORCE_SCAN"] = "1"
with Decoder(0) as dec:
    print("sequential scan, lazy grey", dec.decode(pcm, 11025, 120).start_frame)
del os.environ["WEFAX_SYNC_FORCE_SCAN"]
os.environ["WEFAX_DEPTH_FIRST"] = "1"
with Decoder(0) as dec:
    print("lanes", dec.decode(np.stack([pcm] * 5), 11025, [120, 60, 240, 90, 120]).start_frame)
del os.environ["WEFAX_DEPTH_FIRST"]
# segment mode: three contexts on one GPU (circular halo with a seam at both ends, histogram exchange, raster margins),
# at 11025 Hz and through the resampler
from wefax_b200 import segments as S
long_pcm = synth.synth_recording(150.0, seed=21, noise_sigma=0.03)
ds = [Decoder(0) for _ in range(3)]
s = S.decode_segmented(long_pcm, 11025, 120, ds, halo=20000, want=("raster", "digitalized", "demodulated"))
print("segments", s.status, s.start_frame, s.image.shape, sorted(s.rows))
pcm48 = synth.synth_recording(150.0, sample_rate=48000, seed=5, noise_sigma=0.03)
s = S.decode_segmented(pcm48, 48000, 120, ds[:2], halo=20000)
print("segments 48k", s.status, s.start_frame, s.image.shape)
for d in ds:
    d.close()
