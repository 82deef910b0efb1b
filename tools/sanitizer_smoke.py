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
