import numpy as np, sys
sys.path.insert(0, "/root/repo")
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
